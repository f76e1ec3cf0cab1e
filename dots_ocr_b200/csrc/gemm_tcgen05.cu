// bf16 GEMM on the 5th-gen tensor cores:  D[M,N] = A[M,K] * B[N,K]^T  (both operands K-major).
//
//   TMA (cp.async.bulk.tensor, 128-B swizzle) -> shared-memory ring -> tcgen05.mma (one elected
//   thread) -> fp32 accumulators in TMEM (double buffered) -> tcgen05.ld -> fused epilogue -> HBM.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM
// allocator, warps 4-7 = epilogue (one TMEM lane quarter each).  One CTA per SM.
//
// Used for every dense contraction on the dots.ocr hot path (SURVEY.md §8a rows a5, a8, a11-a13,
// a16, a20-a22).  Epilogues reproduce the HF eager rounding points (each nn.Linear rounds to bf16
// before the next elementwise op).
#include "common.h"
#include "ptx.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;        // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int ACC_STAGES = 2;
constexpr int GEMM_THREADS = 384;      // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: two epilogue groups
constexpr int EPI_GROUPS = 2;          // each group = 4 warps (one per TMEM lane quarter) handling half of the tile's columns

struct GemmParams {
    int M, N, K;                   // rows of A, rows of B, reduction length
    int m_blocks, n_blocks, splits, kb_per_split, num_k_blocks;
    int static_is_b;               // which operand holds weights (never written by an in-flight kernel): 1 = B, 0 = A (swap-AB)
    void* out;
    long long ldo;
    const bf16* bias;
    const bf16* res;
    long long ldr;
    // decode (swap-AB) only: operands pre-tiled in HBM and fetched with 1-D bulk copies instead of tensor-map boxes
    const uint8_t* a_tiled;        // ops.tile_weight(W): [m_blocks][num_k_blocks] blobs of 16 KB; nullptr: tmap_a
    const uint8_t* b_tiled;        // activations, k-block-tiled with BLOCK_N rows per tile (ops.tile_rows); nullptr: tmap_b
    int out_tiled;                 // SWIGLU_T: act is written k-block-tiled with BLOCK_N rows per tile (B operand of down_proj)
    // DOTS_EPI_ROPE: 2-D rotary embedding of the ViT fused into the q|k|v projection (columns < rope_cols are q and k heads)
    const float* rope_cos;         // [M, 64] fp32
    const float* rope_sin;
    int rope_cols;
    unsigned long long* trace;     // timeline instrumentation (nullptr unless armed)
    int stages;                    // ring depth of this launch (0: the template's STAGES)
};

constexpr bool epi_is_swap_ab(int epi) { return epi == DOTS_EPI_F32_PARTIAL_T || epi == DOTS_EPI_BF16_T || epi == DOTS_EPI_SWIGLU_T; }

template <int BLOCK_N, int EPI>
struct GemmSmem {
    static constexpr bool SWAP = epi_is_swap_ab(EPI);
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    // Decode (swap-AB) kernels keep the ring under half an SM's shared memory so that two CTAs -- usually of two
    // consecutive kernels of the decode step, overlapped by programmatic dependent launch -- stream weights at once.
#ifndef GEMM_SWAP_STAGES
#define GEMM_SWAP_STAGES 6
#endif
    static constexpr int STAGES = SWAP ? (BLOCK_N >= 256 ? 2 : (BLOCK_N >= 128 ? 4 : GEMM_SWAP_STAGES)) : (BLOCK_N >= 256) ? 4 : (BLOCK_N >= 128 ? 6 : 8);
    // STAGES is the maximum ring depth of the decode kernels; a launch picks its own depth (GemmParams::stages, <= STAGES) and asks
    // for that much shared memory only, so that short kernels (q|k|v: 3 k-blocks per CTA, o_proj: 2) leave room for their neighbours
    // on the SM and long ones (gate|up: 24) keep more bytes in flight.  Two CTAs per SM by registers for the narrow batch tiles.
    static constexpr int MIN_CTAS_ = (SWAP && BLOCK_N <= 64) ? 2 : 1;
    static constexpr int MIN_CTAS = MIN_CTAS_;
    static constexpr int BAR_BYTES = 1024;
    // SWIGLU_T: the up-projection warps hand their bf16-rounded values to the gate warps through shared memory
    static constexpr int XCH_BYTES = (EPI == DOTS_EPI_SWIGLU_T) ? BLOCK_N * 64 * 2 : 0;
    static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + XCH_BYTES + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = (ACC_STAGES * BLOCK_N <= 32) ? 32 : (ACC_STAGES * BLOCK_N <= 64) ? 64
                                   : (ACC_STAGES * BLOCK_N <= 128) ? 128 : (ACC_STAGES * BLOCK_N <= 256) ? 256 : 512;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ void tile_coords(const GemmParams& p, int t, int& m_blk, int& n_blk, int& split) {
    split = t % p.splits;
    int r = t / p.splits;
    n_blk = r % p.n_blocks;
    m_blk = r / p.n_blocks;
}

// Epilogue of one accumulator tile, executed by one epilogue thread: `row` = output row of this thread (TMEM lane),
// `t_row` = TMEM address of its accumulator row, `eg` = epilogue group (which half of the tile's columns), `m_blk` = tile row
// index in units of BLOCK_M (swap-AB epilogues only).  Shared by the 1-CTA and the 2-CTA kernels.
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, int row, int m_blk, int n_blk, int split, uint32_t t_row, int eg,
                                                   int wq, int lane, bf16* xch) {
    if constexpr (EPI == DOTS_EPI_SWIGLU) {
        // B rows are interleaved per 128-block: [64 gate rows | 64 up rows]; out has N/2 columns.
        static_assert(EPI != DOTS_EPI_SWIGLU || BLOCK_N == 256, "swiglu epilogue needs BLOCK_N=256");
        bf16* out = reinterpret_cast<bf16*>(p.out);
#pragma unroll 1
        for (int c = eg * 2; c < eg * 2 + 2; ++c) {
            uint32_t g[32], u[32];
            tmem_ld_32x32b_x32(t_row + (c >> 1) * 128 + (c & 1) * 32, g);
            tmem_ld_32x32b_x32(t_row + (c >> 1) * 128 + 64 + (c & 1) * 32, u);
            tmem_ld_wait();
            const int col = n_blk * 128 + c * 32;
            if (row < p.M) {
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float g0 = bf16_round(__uint_as_float(g[2 * j])), g1 = bf16_round(__uint_as_float(g[2 * j + 1]));
                    float u0 = bf16_round(__uint_as_float(u[2 * j])), u1 = bf16_round(__uint_as_float(u[2 * j + 1]));
                    float a0 = bf16_round(silu_f(g0)), a1 = bf16_round(silu_f(g1));
                    o[j] = pack_bf16x2(a0 * u0, a1 * u1);
                }
                bf16* dst = out + (long long)row * p.ldo + col;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (col + q * 8 + 8 <= p.N / 2)
                        *reinterpret_cast<uint4*>(dst + q * 8) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            }
        }
    } else if constexpr (EPI == DOTS_EPI_ROPE) {
        // ViT q|k|v projection with the 2-D rotary embedding fused ([V]:287 + :295-302): an epilogue group owns one 128-column head;
        // q = bf16(acc) first (the Linear's output dtype), then NeoX rotate-half in fp32 with one rounding -- the arithmetic of
        // vit_rope_apply_kernel, so the result is bit-identical to GEMM + that kernel, without the extra read-modify-write pass.
        static_assert(EPI != DOTS_EPI_ROPE || BLOCK_N == 256, "rope epilogue needs 256-wide tiles (one head per epilogue group)");
        bf16* out = reinterpret_cast<bf16*>(p.out);
        const int col_head = n_blk * BLOCK_N + eg * 128;
        const bool rot = col_head < p.rope_cols;
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t a[32], b[32];
            tmem_ld_32x32b_x32(t_row + eg * 128 + half * 32, a);
            tmem_ld_32x32b_x32(t_row + eg * 128 + 64 + half * 32, b);
            tmem_ld_wait();
            if (!row_ok || col_head + 128 > p.N) continue;
            bf16* dst = out + (long long)row * p.ldo + col_head + half * 32;
            const float4* cp = reinterpret_cast<const float4*>(p.rope_cos + (long long)row * 64 + half * 32);
            const float4* sp = reinterpret_cast<const float4*>(p.rope_sin + (long long)row * 64 + half * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float x1[8], x2[8], o1[8], o2[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { x1[j] = bf16_round(__uint_as_float(a[q * 8 + j])); x2[j] = bf16_round(__uint_as_float(b[q * 8 + j])); }
                if (rot) {
                    float c[8], sn[8];
                    *reinterpret_cast<float4*>(c) = __ldg(cp + 2 * q); *reinterpret_cast<float4*>(c + 4) = __ldg(cp + 2 * q + 1);
                    *reinterpret_cast<float4*>(sn) = __ldg(sp + 2 * q); *reinterpret_cast<float4*>(sn + 4) = __ldg(sp + 2 * q + 1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o1[j] = __fsub_rn(__fmul_rn(x1[j], c[j]), __fmul_rn(x2[j], sn[j]));
                        o2[j] = __fadd_rn(__fmul_rn(x2[j], c[j]), __fmul_rn(x1[j], sn[j]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { o1[j] = x1[j]; o2[j] = x2[j]; }
                }
                *reinterpret_cast<uint4*>(dst + q * 8) = make_uint4(pack_bf16x2(o1[0], o1[1]), pack_bf16x2(o1[2], o1[3]), pack_bf16x2(o1[4], o1[5]), pack_bf16x2(o1[6], o1[7]));
                *reinterpret_cast<uint4*>(dst + 64 + q * 8) = make_uint4(pack_bf16x2(o2[0], o2[1]), pack_bf16x2(o2[2], o2[3]), pack_bf16x2(o2[4], o2[5]), pack_bf16x2(o2[6], o2[7]));
            }
        }
    } else if constexpr (EPI == DOTS_EPI_SWIGLU_T) {
        // swap-AB decode GEMM over the gate|up weight: tile rows 0-63 are gate features, rows 64-127 the matching up
        // features (warps 0,1 / 2,3 of this group).  Up warps publish bf16(up) through shared memory, gate warps
        // finish  act[b][f] = bf16( bf16(silu(bf16 g)) * bf16 u )  -- the same rounding points as the prefill epilogue.
        // Each epilogue group (4 warps, one per TMEM lane quarter) owns half of the batch columns when the tile is wide enough;
        // inside a group all four warps share the SiLU work: gate warps publish bf16(g) of the upper half of the group's columns,
        // up warps publish bf16(u) of the lower half; then gate warps finish the lower half and up warps the upper half.
        constexpr int GROUPS = (BLOCK_N >= 64) ? EPI_GROUPS : 1;       // 16-column TMEM chunks: a group needs >= 32 columns
        if (eg < GROUPS) {
            bf16* out = reinterpret_cast<bf16*>(p.out);
            const int fl = (wq & 1) * 32 + lane;                 // feature within the 64-block
            const bool is_up = wq >= 2;
            constexpr int SPAN = BLOCK_N / GROUPS;               // batch columns of this group
            constexpr int HALF = SPAN / 2;
            static_assert(HALF % 16 == 0, "batch tile halves are read in 16-column TMEM chunks");
            const int g0 = eg * SPAN;
            const int f = m_blk * 64 + fl;
#pragma unroll 1
            for (int h0 = 0; h0 < HALF; h0 += 32) {              // 32 batch columns of each half per pass (register budget)
                constexpr int W = HALF < 32 ? HALF : 32;
                float mine[W];
                const int pub0 = g0 + (is_up ? 0 : HALF) + h0, keep0 = g0 + (is_up ? HALF : 0) + h0;
#pragma unroll
                for (int c = 0; c < W / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_row + pub0 + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) xch[(pub0 + c * 16 + j) * 64 + fl] = __float2bfloat16_rn(__uint_as_float(v[j]));
                }
#pragma unroll
                for (int c = 0; c < W / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(t_row + keep0 + c * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) mine[c * 16 + j] = __uint_as_float(v[j]);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");      // the four warps of this group only
                if (f < p.M / 2) {
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const int b = n_blk * BLOCK_N + keep0 + j;
                        if (b < p.N) {
                            const float other = __bfloat162float(xch[(keep0 + j) * 64 + fl]);
                            const float gv = is_up ? other : bf16_round(mine[j]);
                            const float uv = is_up ? bf16_round(mine[j]) : other;
                            const long long o = p.out_tiled ? tiled_row_off(b, f, BLOCK_N) : (long long)b * p.ldo + f;
                            out[o] = __float2bfloat16_rn(bf16_round(silu_f(gv)) * uv);
                        }
                    }
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + eg) : "memory");      // xch rows of this group are free for the next tile
        }
    } else if constexpr (EPI == DOTS_EPI_F32_PARTIAL_T) {
        // swap-AB decode GEMM: A rows are output features, B rows are batch rows.
        // partial[split][b][feature] fp32; lanes write consecutive features (coalesced).
        // (this epilogue is the exposed tail of a 3-12 k-block kernel: one base pointer, pointer bumps, no per-element bound
        // arithmetic on full chunks)
        float* out = reinterpret_cast<float*>(p.out);
#pragma unroll 1
        for (int c = eg; c < BLOCK_N / 32; c += EPI_GROUPS) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + c * 32, v);
            const int b0 = n_blk * BLOCK_N + c * 32;
            float* dst = out + ((long long)split * p.N + b0) * p.ldo + row;
            const int nb = min(32, p.N - b0);
            tmem_ld_wait();
            if (row < p.M) {
                if (nb == 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { *dst = __uint_as_float(v[j]); dst += p.ldo; }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { if (j < nb) *dst = __uint_as_float(v[j]); dst += p.ldo; }
                }
            }
        }
    } else if constexpr (EPI == DOTS_EPI_BF16_T) {
        // swap-AB, no split: out[b][feature] = bf16(acc + bias[feature]); lanes = consecutive features.
        bf16* out = reinterpret_cast<bf16*>(p.out);
        const float bias_v = (p.bias != nullptr && row < p.M) ? __bfloat162float(p.bias[row]) : 0.f;
#pragma unroll 1
        for (int c = eg; c < BLOCK_N / 32; c += EPI_GROUPS) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + c * 32, v);
            tmem_ld_wait();
            if (row < p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int b = n_blk * BLOCK_N + c * 32 + j;
                    if (b < p.N) out[(long long)b * p.ldo + row] = __float2bfloat16_rn(__uint_as_float(v[j]) + bias_v);
                }
            }
        }
    } else {
        // Two 32-column chunks per iteration: both TMEM loads and (for the residual epilogue) all eight 16-byte
        // residual loads are issued before the first use, so one iteration pays one memory latency, not two.
        static_assert(BLOCK_N % 64 == 0, "generic epilogue walks the tile in 64-column steps");
        bf16* out = reinterpret_cast<bf16*>(p.out);
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int c = eg * (BLOCK_N / 64); c < (eg + 1) * (BLOCK_N / 64); c += 2) {
            uint32_t v[2][32];
            tmem_ld_32x32b_x32(t_row + c * 32, v[0]);
            tmem_ld_32x32b_x32(t_row + (c + 1) * 32, v[1]);
            const int col0 = n_blk * BLOCK_N + c * 32;
            uint4 rres[2][4];
            if constexpr (EPI == DOTS_EPI_RESIDUAL) {
                const bf16* rsrc = p.res + (long long)row * p.ldr + col0;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        rres[h][q] = (row_ok && col0 + h * 32 + q * 8 + 8 <= p.N)
                                         ? *reinterpret_cast<const uint4*>(rsrc + h * 32 + q * 8) : make_uint4(0, 0, 0, 0);
            }
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int col = col0 + h * 32;
                if (!(row_ok && col < p.N)) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[h][j]);
                if constexpr (EPI == DOTS_EPI_BIAS || EPI == DOTS_EPI_BIAS_GELU) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (col + q * 8 + 8 <= p.N) {
                            uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col + q * 8));
                            f[q * 8 + 0] += bf16_lo(b.x); f[q * 8 + 1] += bf16_hi(b.x);
                            f[q * 8 + 2] += bf16_lo(b.y); f[q * 8 + 3] += bf16_hi(b.y);
                            f[q * 8 + 4] += bf16_lo(b.z); f[q * 8 + 5] += bf16_hi(b.z);
                            f[q * 8 + 6] += bf16_lo(b.w); f[q * 8 + 7] += bf16_hi(b.w);
                        }
                    }
                }
                if constexpr (EPI == DOTS_EPI_BIAS_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf_f(bf16_round(f[j]));
                }
                if constexpr (EPI == DOTS_EPI_RESIDUAL) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 r = rres[h][q];
                        f[q * 8 + 0] = bf16_round(f[q * 8 + 0]) + bf16_lo(r.x);
                        f[q * 8 + 1] = bf16_round(f[q * 8 + 1]) + bf16_hi(r.x);
                        f[q * 8 + 2] = bf16_round(f[q * 8 + 2]) + bf16_lo(r.y);
                        f[q * 8 + 3] = bf16_round(f[q * 8 + 3]) + bf16_hi(r.y);
                        f[q * 8 + 4] = bf16_round(f[q * 8 + 4]) + bf16_lo(r.z);
                        f[q * 8 + 5] = bf16_round(f[q * 8 + 5]) + bf16_hi(r.z);
                        f[q * 8 + 6] = bf16_round(f[q * 8 + 6]) + bf16_lo(r.w);
                        f[q * 8 + 7] = bf16_round(f[q * 8 + 7]) + bf16_hi(r.w);
                    }
                }
                bf16* dst = out + (long long)row * p.ldo + col;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (col + q * 8 + 8 <= p.N) {
                        uint4 o;
                        o.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                        o.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                        o.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                        o.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                        *reinterpret_cast<uint4*>(dst + q * 8) = o;
                    }
                }
            }
        }
    }
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, GemmSmem<BLOCK_N, EPI>::MIN_CTAS)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p) {
    using S = GemmSmem<BLOCK_N, EPI>;
    constexpr int MAX_STAGES = S::STAGES;
    const int STAGES = (p.stages > 0 && p.stages < MAX_STAGES) ? p.stages : MAX_STAGES;      // ring depth of this launch
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * S::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [MAX_STAGES]
    uint64_t* empty_bar = bars + MAX_STAGES;         // [MAX_STAGES]
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;     // [ACC_STAGES]
    uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [ACC_STAGES]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
    bf16* xch = reinterpret_cast<bf16*>(smem + STAGES * S::STAGE_BYTES + S::BAR_BYTES);   // [BLOCK_N][64] (SWIGLU_T only)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_blocks * p.n_blocks * p.splits;

    if (warp == 0 && lane == 0) {
        if (!p.a_tiled) prefetch_tensormap(&tmap_a);
        if (!p.b_tiled) prefetch_tensormap(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < ACC_STAGES; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4 * EPI_GROUPS);   // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);      // warp-uniform for the compiler
    pdl_launch_dependents();        // the next kernel of the stream may begin its own prologue / weight prefetch
    constexpr int TRACE_KID = 100 + EPI;
    if (threadIdx.x == 0) trace_point(p.trace, TRACE_KID, 0);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            // operand fetch of one ring stage: a tensor-map box, or (decode, pre-tiled operands) one contiguous bulk copy
            auto load_a = [&](int st, int kb, int m_blk) {
                if (p.a_tiled) bulk_load(smem_a + st * S::A_BYTES, p.a_tiled + ((size_t)m_blk * p.num_k_blocks + kb) * S::A_BYTES, S::A_BYTES, &full_bar[st]);
                else tma_load_2d(smem_a + st * S::A_BYTES, &tmap_a, kb * BLOCK_K, m_blk * BLOCK_M, &full_bar[st]);
            };
            auto load_b = [&](int st, int kb, int n_blk) {
                if (p.b_tiled) bulk_load(smem_b + st * S::B_BYTES, p.b_tiled + (size_t)kb * S::B_BYTES, S::B_BYTES, &full_bar[st]);
                else tma_load_2d(smem_b + st * S::B_BYTES, &tmap_b, kb * BLOCK_K, n_blk * BLOCK_N, &full_bar[st]);
            };
            // Weights never depend on a predecessor kernel: put the first ring-full of weight tiles in flight BEFORE
            // waiting for the dependency, so HBM keeps streaming across the kernel boundary.
            int pre = 0;
            if ((int)blockIdx.x < num_tiles) {
                int m_blk, n_blk, split;
                tile_coords(p, blockIdx.x, m_blk, n_blk, split);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1 && pre < STAGES; ++kb, ++pre) {
                    mbar_expect_tx(&full_bar[pre], S::STAGE_BYTES);
                    if (p.static_is_b) tma_load_2d(smem_b + pre * S::B_BYTES, &tmap_b, kb * BLOCK_K, n_blk * BLOCK_N, &full_bar[pre]);
                    else load_a(pre, kb, m_blk);
                }
            }
            pdl_wait();
            trace_point(p.trace, TRACE_KID, 1);
            int issued = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m_blk, n_blk, split;
                tile_coords(p, t, m_blk, n_blk, split);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++issued) {
                    if (issued < pre) {
                        // static half already in flight on this stage's barrier: add the dependent half
                        if (p.static_is_b) load_a(stage, kb, m_blk);
                        else load_b(stage, kb, n_blk);
                    } else {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                        load_a(stage, kb, m_blk);
                        load_b(stage, kb, n_blk);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp runs the loop convergently (uniform registers for descriptors / TMEM addresses); only the
        // tcgen05 instructions sit under the elected-lane predicate.
        {
            const bool leader = elect_one();
            constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N);
            const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem_a));
            const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b));
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m_blk, n_blk, split;
                tile_coords(p, t, m_blk, n_blk, split);
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (leader && kb == kb0 && t == (int)blockIdx.x) trace_point(p.trace, TRACE_KID, 2);
                    const uint64_t da = da0 + (uint64_t)(stage * (S::A_BYTES >> 4));
                    const uint64_t db = db0 + (uint64_t)(stage * (S::B_BYTES >> 4));
                    const uint32_t first = (kb > kb0) ? 1u : 0u;
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            // advance 32 bytes (16 bf16) along K inside the swizzle row: +2 in the >>4 address field
                            umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : first);
                        }
                        umma_commit(&empty_bar[stage]);          // smem slot reusable once these MMAs retire
                        if (kb + 1 == kb1) umma_commit(&tmem_full[acc]);   // accumulator complete -> epilogue
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
            if (leader) trace_point(p.trace, TRACE_KID, 3);
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        pdl_wait();                                          // residual / bias reads and all output writes follow the dependency
        const int wq = warp & 3;                             // TMEM lane quarter this warp may read
        const int eg = (warp - 4) >> 2;                      // epilogue group: which half of the tile's columns
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int m_blk, n_blk, split;
            tile_coords(p, t, m_blk, n_blk, split);
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * BLOCK_M + wq * 32 + lane;          // row of A this thread owns
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BLOCK_N;

            gemm_epilogue_tile<BLOCK_N, EPI>(p, row, m_blk, n_blk, split, t_row, eg, wq, lane, xch);
            // release this accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
        if (threadIdx.x == 128) trace_point(p.trace, TRACE_KID, 4);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

extern int g_gemm_pair;

template <int BLOCK_N, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    using S = GemmSmem<BLOCK_N, EPI>;
    auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, EPI>;
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    }
    const int tiles = p.m_blocks * p.n_blocks * p.splits;
    const int slots = num_sms() * S::MIN_CTAS;
    const int grid = tiles < slots ? tiles : slots;
    GemmParams pt = p;
    pt.trace = g_trace;
    const int stages = (p.stages > 0 && p.stages < S::STAGES) ? p.stages : S::STAGES;
    const size_t smem = (size_t)stages * S::STAGE_BYTES + S::BAR_BYTES + S::XCH_BYTES + 1024;
    DOTS_CHECK_CUDA(launch_ex(kern, dim3(grid), dim3(GEMM_THREADS), smem, stream, true, ta, tb, pt));
    return 0;
}


// ----------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per cluster of two CTAs.  Each CTA stages its own 128 rows of A
// and HALF of the B tile (128 of its 256 rows) per k-block; the pair's tensor cores read both halves, so per CTA the B
// traffic into shared memory and out of it is halved (stage = 16 + 16 KB -> six stages), which is what keeps a
// 128-wide-M kernel off the shared-memory roofline.  The leader CTA (cluster rank 0) issues the MMAs and owns the `full`
// barriers (both CTAs' TMA bytes are credited to them) and the `tmem_empty` barriers (both CTAs' epilogue warps arrive);
// `empty` and `tmem_full` are signalled in both CTAs by multicast commits.
constexpr int G2_STAGES = 6;
constexpr int G2_HALF_N = 128;
constexpr int G2_BLOCK_N = 256;
constexpr int G2_A_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KB
constexpr int G2_B_BYTES = G2_HALF_N * BLOCK_K * 2;        // 16 KB
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_SMEM = G2_STAGES * G2_STAGE_BYTES + 1024 /*barriers*/ + 1024 /*align*/;

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + G2_STAGES * G2_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]   (leader's copy is the live one)
    uint64_t* empty_bar = bars + G2_STAGES;          // [STAGES]   per CTA
    uint64_t* tmem_full = bars + 2 * G2_STAGES;      // [2]        per CTA
    uint64_t* tmem_empty = tmem_full + ACC_STAGES;   // [2]        (leader's copy is the live one)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int num_tiles = p.m_blocks * p.n_blocks;   // m_blocks counts 256-row tiles here

    if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_a); prefetch_tensormap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < G2_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * 4 * EPI_GROUPS); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_2sm(tmem_ptr, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // the peer's barriers exist before anything remote touches them
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_launch_dependents();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            pdl_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += n_pairs) {
                const int n_blk = t % p.n_blocks, m_pair = t / p.n_blocks;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);      // both CTAs' tiles
                    tma_load_2d_2sm(smem_a + stage * G2_A_BYTES, &tmap_a, kb * BLOCK_K, m_pair * 256 + (int)rank * BLOCK_M, full_leader);
                    tma_load_2d_2sm(smem_b + stage * G2_B_BYTES, &tmap_b, kb * BLOCK_K, n_blk * G2_BLOCK_N + (int)rank * G2_HALF_N, full_leader);
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader) {
            const bool elected = elect_one();
            constexpr uint32_t idesc = umma_idesc_bf16(256, G2_BLOCK_N);
            const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem_a));
            const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b));
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = pair; t < num_tiles; t += n_pairs) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * G2_BLOCK_N;
                for (int kb = 0; kb < p.num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = da0 + (uint64_t)(stage * (G2_A_BYTES >> 4));
                    const uint64_t db = db0 + (uint64_t)(stage * (G2_B_BYTES >> 4));
                    const uint32_t first = (kb > 0) ? 1u : 0u;
                    if (elected) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : first);
                        umma_commit_2sm(&empty_bar[stage], 3);                                   // slot free in both CTAs
                        if (kb + 1 == p.num_k_blocks) umma_commit_2sm(&tmem_full[acc], 3);       // accumulators ready in both CTAs
                    }
                    __syncwarp();
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs: each owns 128 rows of the 256-row tile) =====================
        pdl_wait();
        const int wq = warp & 3;
        const int eg = (warp - 4) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = pair; t < num_tiles; t += n_pairs) {
            const int n_blk = t % p.n_blocks, m_pair = t / p.n_blocks;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_pair * 256 + (int)rank * BLOCK_M + wq * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * G2_BLOCK_N;
            gemm_epilogue_tile<G2_BLOCK_N, EPI>(p, row, 0, n_blk, 0, t_row, eg, wq, lane, nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));     // the leader's barrier counts both CTAs
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // nobody leaves while the peer may still signal or read this CTA
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

template <int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    auto kern = gemm2_bf16_tcgen05_kernel<EPI>;
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM));
    }
    const int tiles = p.m_blocks * p.n_blocks;
    int pairs = num_sms() / 2;
    if (tiles < pairs) pairs = tiles;
    DOTS_CHECK_CUDA(launch_ex_cluster(kern, dim3(2 * pairs), dim3(GEMM_THREADS), (size_t)G2_SMEM, stream, true, 2u, ta, tb, p));
    return 0;
}

}  // namespace dots

using namespace dots;

namespace dots { int g_gemm_pair = 1; }      // CTA-pair (cta_group::2) kernel for large prefill GEMMs (default on); dots_set_gemm_pair()

extern "C" int dots_set_gemm_pair(int enable) {
    dots::g_gemm_pair = enable ? 1 : 0;
    return 0;
}

extern "C" int dots_vit_rope_apply(void* qkv, long long ld, int S, int heads, int head_dim, const float* cos_t, const float* sin_t, void* stream);

static int gemm_bf16_impl(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                          int M, int N, int K, int epilogue, const void* bias, const void* residual,
                          long long ldr, const float* rope_cos, const float* rope_sin, int rope_cols, void* stream) {
    DOTS_REQUIRE(M > 0 && N > 0 && K > 0, "dots_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
    DOTS_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && ldo % 8 == 0 && N % 8 == 0,
                 "dots_gemm_bf16: K, N and all pitches must be multiples of 8 (16-byte TMA/vector alignment)");
    DOTS_REQUIRE((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out)) % 16 == 0,
                 "dots_gemm_bf16: pointers must be 16-byte aligned");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.out = out; p.ldo = ldo;
    p.bias = reinterpret_cast<const bf16*>(bias);
    p.res = reinterpret_cast<const bf16*>(residual);
    p.ldr = ldr;
    p.splits = 1;
    p.static_is_b = 1;             // W holds weights
    p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    p.kb_per_split = p.num_k_blocks;
    p.m_blocks = (M + BLOCK_M - 1) / BLOCK_M;
    p.rope_cos = rope_cos; p.rope_sin = rope_sin; p.rope_cols = rope_cols;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ta, tb;
    if (make_tmap_2d_bf16(&ta, A, M, K, lda, BLOCK_M)) return -4;

    // CTA-pair kernel (256 x 256 tiles) for problems with at least two waves of such tiles
    const bool pair_ok = g_gemm_pair && (N % 256 == 0) && ((long long)((M + 255) / 256) * (N / 256) >= num_sms());
    if (epilogue == DOTS_EPI_ROPE) {
        DOTS_REQUIRE(rope_cos && rope_sin && rope_cols % 256 == 0 && rope_cols <= N && N % 128 == 0,
                     "dots_gemm_bf16_rope: need cos/sin tables, rope_cols %% 256 == 0 (q and k heads of 128), N %% 128 == 0");
        const long long tiles256 = (long long)p.m_blocks * ((N + 255) / 256);
        if (!(N % 256 == 0 && (pair_ok || tiles256 >= 2LL * num_sms()))) {
            // small problems run on 128-wide tiles, where a head straddles two epilogue groups: plain store, then the rotary pass
            if (int rc = gemm_bf16_impl(A, lda, W, ldw, out, ldo, M, N, K, DOTS_EPI_STORE, nullptr, nullptr, 0, nullptr, nullptr, 0, stream)) return rc;
            return dots_vit_rope_apply(out, ldo, M, rope_cols / 256, 128, rope_cos, rope_sin, stream);
        }
    }
    if (pair_ok && epilogue != DOTS_EPI_F32_PARTIAL_T && epilogue != DOTS_EPI_BF16_T && epilogue != DOTS_EPI_SWIGLU_T) {
        if ((epilogue == DOTS_EPI_BIAS || epilogue == DOTS_EPI_BIAS_GELU) && !bias) {
            set_error("dots_gemm_bf16: bias epilogue without bias pointer");
            return -1;
        }
        if (epilogue == DOTS_EPI_RESIDUAL) {
            DOTS_REQUIRE(residual && ldr % 8 == 0, "dots_gemm_bf16: residual epilogue needs residual pointer, ldr %% 8 == 0");
        }
        p.m_blocks = (M + 255) / 256;
        p.n_blocks = N / 256;
        CUtensorMap tb2;
        if (make_tmap_2d_bf16(&tb2, W, N, K, ldw, G2_HALF_N)) return -4;
        switch (epilogue) {
            case DOTS_EPI_STORE: return launch_gemm2<DOTS_EPI_STORE>(ta, tb2, p, st);
            case DOTS_EPI_BIAS: return launch_gemm2<DOTS_EPI_BIAS>(ta, tb2, p, st);
            case DOTS_EPI_BIAS_GELU: return launch_gemm2<DOTS_EPI_BIAS_GELU>(ta, tb2, p, st);
            case DOTS_EPI_RESIDUAL: return launch_gemm2<DOTS_EPI_RESIDUAL>(ta, tb2, p, st);
            case DOTS_EPI_SWIGLU: return launch_gemm2<DOTS_EPI_SWIGLU>(ta, tb2, p, st);
            case DOTS_EPI_ROPE: return launch_gemm2<DOTS_EPI_ROPE>(ta, tb2, p, st);
        }
    }

    if (epilogue == DOTS_EPI_SWIGLU) {
        DOTS_REQUIRE(N % 256 == 0, "dots_gemm_bf16: SWIGLU epilogue needs N %% 256 == 0 (got %d)", N);
        p.n_blocks = N / 256;
        if (make_tmap_2d_bf16(&tb, W, N, K, ldw, 256)) return -4;
        return launch_gemm<256, DOTS_EPI_SWIGLU>(ta, tb, p, st);
    }
    if ((epilogue == DOTS_EPI_BIAS || epilogue == DOTS_EPI_BIAS_GELU) && !bias) {
        set_error("dots_gemm_bf16: bias epilogue without bias pointer");
        return -1;
    }
    if (epilogue == DOTS_EPI_RESIDUAL) {
        DOTS_REQUIRE(residual && ldr % 8 == 0, "dots_gemm_bf16: residual epilogue needs residual pointer, ldr %% 8 == 0");
    }
    // Tile width: 256 when it divides the work well, else 128 (more tiles for small problems).
    const long long tiles256 = (long long)p.m_blocks * ((N + 255) / 256);
    const bool use256 = (N % 256 == 0 || N > 4096) && tiles256 >= 2LL * num_sms();
    if (use256) {
        p.n_blocks = (N + 255) / 256;
        if (make_tmap_2d_bf16(&tb, W, N, K, ldw, 256)) return -4;
        switch (epilogue) {
            case DOTS_EPI_STORE: return launch_gemm<256, DOTS_EPI_STORE>(ta, tb, p, st);
            case DOTS_EPI_BIAS: return launch_gemm<256, DOTS_EPI_BIAS>(ta, tb, p, st);
            case DOTS_EPI_BIAS_GELU: return launch_gemm<256, DOTS_EPI_BIAS_GELU>(ta, tb, p, st);
            case DOTS_EPI_RESIDUAL: return launch_gemm<256, DOTS_EPI_RESIDUAL>(ta, tb, p, st);
            case DOTS_EPI_ROPE: return launch_gemm<256, DOTS_EPI_ROPE>(ta, tb, p, st);
        }
    } else {
        p.n_blocks = (N + 127) / 128;
        if (make_tmap_2d_bf16(&tb, W, N, K, ldw, 128)) return -4;
        switch (epilogue) {
            case DOTS_EPI_STORE: return launch_gemm<128, DOTS_EPI_STORE>(ta, tb, p, st);
            case DOTS_EPI_BIAS: return launch_gemm<128, DOTS_EPI_BIAS>(ta, tb, p, st);
            case DOTS_EPI_BIAS_GELU: return launch_gemm<128, DOTS_EPI_BIAS_GELU>(ta, tb, p, st);
            case DOTS_EPI_RESIDUAL: return launch_gemm<128, DOTS_EPI_RESIDUAL>(ta, tb, p, st);
        }
    }
    set_error("dots_gemm_bf16: unknown epilogue %d", epilogue);
    return -1;
}

extern "C" int dots_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo,
                              int M, int N, int K, int epilogue, const void* bias, const void* residual,
                              long long ldr, void* stream) {
    DOTS_REQUIRE(epilogue != DOTS_EPI_ROPE, "dots_gemm_bf16: the rotary epilogue is dots_gemm_bf16_rope");
    return gemm_bf16_impl(A, lda, W, ldw, out, ldo, M, N, K, epilogue, bias, residual, ldr, nullptr, nullptr, 0, stream);
}

// ViT q|k|v projection with the 2-D rotary embedding applied to the first rope_cols columns (q heads then k heads, 128 wide) in the
// GEMM epilogue: out = [rope(bf16(A W_q^T)) | rope(bf16(A W_k^T)) | bf16(A W_v^T)]; cos/sin [M, 64] fp32 (dots_vit_rope_table).
extern "C" int dots_gemm_bf16_rope(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldo, int M, int N, int K,
                                   const float* cos_t, const float* sin_t, int rope_cols, void* stream) {
    return gemm_bf16_impl(A, lda, W, ldw, out, ldo, M, N, K, DOTS_EPI_ROPE, nullptr, nullptr, 0, cos_t, sin_t, rope_cols, stream);
}

// Decode-time (skinny) GEMM, swap-AB + split-K:  partial[s][b][n] = sum_{k in split s} X[b,k] * W[n,k].
// The weight matrix is the 128-row tcgen05 M operand so that a batch of <= 256 rows still runs on
// the tensor cores while the weights stream once from HBM.
extern "C" int dots_gemm_skinny_bf16(const void* X, long long ldx, const void* W, long long ldw, float* partial,
                                     void* out_bf16, long long ldo, const void* bias, int batch, int N, int K, int splits,
                                     void* stream) {
    DOTS_REQUIRE(batch > 0 && batch <= 256 && N > 0 && K > 0, "dots_gemm_skinny_bf16: bad shape batch=%d N=%d K=%d", batch, N, K);
    DOTS_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "dots_gemm_skinny_bf16: K and pitches must be multiples of 8");
    DOTS_REQUIRE((partial != nullptr) != (out_bf16 != nullptr), "dots_gemm_skinny_bf16: pass exactly one of partial / out_bf16");
    GemmParams p{};
    p.static_is_b = 0;     // swap-AB: the A operand holds the weights
    p.M = N;               // A operand = weights: rows are output features
    p.N = batch;           // B operand = activations
    p.K = K;
    p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    if (splits < 1) splits = 1;
    if (splits > p.num_k_blocks) splits = p.num_k_blocks;
    p.kb_per_split = (p.num_k_blocks + splits - 1) / splits;
    p.splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;   // every split non-empty
    DOTS_REQUIRE(p.splits == splits, "dots_gemm_skinny_bf16: splits=%d does not tile %d k-blocks (would use %d)", splits,
                 p.num_k_blocks, p.splits);
    const bool to_bf16 = out_bf16 != nullptr;
    if (to_bf16) {
        DOTS_REQUIRE(splits == 1, "dots_gemm_skinny_bf16: bf16 output needs splits == 1");
        p.out = out_bf16; p.ldo = ldo; p.bias = reinterpret_cast<const bf16*>(bias);
    } else {
        p.out = partial; p.ldo = N;
    }
    p.m_blocks = (N + BLOCK_M - 1) / BLOCK_M;
    p.n_blocks = 1;
    p.stages = 4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ta, tb;
    if (make_tmap_2d_bf16(&ta, W, N, K, ldw, BLOCK_M)) return -4;
    const int bn = batch <= 32 ? 32 : batch <= 64 ? 64 : batch <= 128 ? 128 : 256;
    if (make_tmap_2d_bf16(&tb, X, batch, K, ldx, bn)) return -4;
    if (to_bf16) {
        switch (bn) {
            case 32: return launch_gemm<32, DOTS_EPI_BF16_T>(ta, tb, p, st);
            case 64: return launch_gemm<64, DOTS_EPI_BF16_T>(ta, tb, p, st);
            case 128: return launch_gemm<128, DOTS_EPI_BF16_T>(ta, tb, p, st);
            default: return launch_gemm<256, DOTS_EPI_BF16_T>(ta, tb, p, st);
        }
    }
    switch (bn) {
        case 32: return launch_gemm<32, DOTS_EPI_F32_PARTIAL_T>(ta, tb, p, st);
        case 64: return launch_gemm<64, DOTS_EPI_F32_PARTIAL_T>(ta, tb, p, st);
        case 128: return launch_gemm<128, DOTS_EPI_F32_PARTIAL_T>(ta, tb, p, st);
        default: return launch_gemm<256, DOTS_EPI_F32_PARTIAL_T>(ta, tb, p, st);
    }
}

// Decode gate|up GEMM with the SwiGLU fused into the epilogue (no split-K: 2I/128 tiles already cover the SMs).
// W is the interleaved gate|up weight [2I, K] ([64 gate | 64 up] per 128 rows); act [batch, I] bf16.
extern "C" int dots_gemm_skinny_swiglu_bf16(const void* X, long long ldx, const void* W, long long ldw, void* act, long long ld_act,
                                            int batch, int two_i, int K, void* stream) {
    DOTS_REQUIRE(batch > 0 && batch <= 256 && two_i > 0 && two_i % 128 == 0 && K > 0, "dots_gemm_skinny_swiglu_bf16: bad shape batch=%d 2I=%d K=%d",
                 batch, two_i, K);
    DOTS_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0, "dots_gemm_skinny_swiglu_bf16: K and pitches must be multiples of 8");
    GemmParams p{};
    p.static_is_b = 0;
    p.M = two_i; p.N = batch; p.K = K;
    p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    p.kb_per_split = p.num_k_blocks;
    p.splits = 1;
    p.out = act; p.ldo = ld_act;
    p.m_blocks = two_i / BLOCK_M;
    p.n_blocks = 1;
    p.stages = 4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ta, tb;
    if (make_tmap_2d_bf16(&ta, W, two_i, K, ldw, BLOCK_M)) return -4;
    const int bn = batch <= 32 ? 32 : batch <= 64 ? 64 : batch <= 128 ? 128 : 256;
    if (make_tmap_2d_bf16(&tb, X, batch, K, ldx, bn)) return -4;
    switch (bn) {
        case 32: return launch_gemm<32, DOTS_EPI_SWIGLU_T>(ta, tb, p, st);
        case 64: return launch_gemm<64, DOTS_EPI_SWIGLU_T>(ta, tb, p, st);
        case 128: return launch_gemm<128, DOTS_EPI_SWIGLU_T>(ta, tb, p, st);
        default: return launch_gemm<256, DOTS_EPI_SWIGLU_T>(ta, tb, p, st);
    }
}

// Ring depth per decode GEMM family (dots_set_decode_stages): [0] split-K partial GEMMs (capped by the k-blocks a CTA owns),
// [1] gate|up + SwiGLU, [2] lm_head; 16 KB of weights + 8 KB of activations per stage at batch 64.  Measured at batch 64 (ablate_r2j_*):
// 4/4/4 2.059 ms per step, 4/5/4 2.022, 4/6/4 2.019, 3/5/4 2.035, 6/5/4 1.991 (default), 4/5/6 2.024.
namespace dots { int g_dec_stages[3] = {6, 5, 4}; }

extern "C" int dots_set_decode_stages(int partial, int swiglu, int head) {
    DOTS_REQUIRE(partial >= 2 && partial <= 8 && swiglu >= 2 && swiglu <= 8 && head >= 2 && head <= 8, "dots_set_decode_stages: depths must be 2..8");
    dots::g_dec_stages[0] = partial; dots::g_dec_stages[1] = swiglu; dots::g_dec_stages[2] = head;
    return 0;
}

// ---- decode GEMMs over pre-tiled operands (1-D bulk copies; see ops.tile_weight / ops.tile_rows) ----
// gate|up projection + SwiGLU of one decode step, batch <= 64.  Wt = tile_weight(interleaved gate|up weight [2I, K]); Xt = normalised
// activations, k-block-tiled with 32 / 64 rows per tile; act_t = bf16(bf16(silu(bf16 g)) * bf16 u), written k-block-tiled (same rows
// per tile): the B operand of down_proj.
extern "C" int dots_decode_gemm_swiglu(const void* Xt, const void* Wt, void* act_t, int batch, int two_i, int K, void* stream) {
    DOTS_REQUIRE(Xt && Wt && act_t && batch > 0 && batch <= 64 && two_i > 0 && two_i % 128 == 0 && K > 0,
                 "dots_decode_gemm_swiglu: bad arguments batch=%d 2I=%d K=%d", batch, two_i, K);
    DOTS_REQUIRE((two_i / 2) % 64 == 0, "dots_decode_gemm_swiglu: I must be a multiple of 64 (tiled output)");
    GemmParams p{};
    p.static_is_b = 0;
    p.M = two_i; p.N = batch; p.K = K;
    p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    p.kb_per_split = p.num_k_blocks;
    p.splits = 1;
    p.out = act_t; p.ldo = two_i / 2;
    p.m_blocks = two_i / BLOCK_M;
    p.n_blocks = 1;
    p.a_tiled = reinterpret_cast<const uint8_t*>(Wt);
    p.b_tiled = reinterpret_cast<const uint8_t*>(Xt);
    p.out_tiled = 1;
    p.stages = g_dec_stages[1] < p.num_k_blocks ? g_dec_stages[1] : p.num_k_blocks;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ta{}, tb{};            // unused with tiled operands
    return batch <= 32 ? launch_gemm<32, DOTS_EPI_SWIGLU_T>(ta, tb, p, st) : launch_gemm<64, DOTS_EPI_SWIGLU_T>(ta, tb, p, st);
}

// lm_head of one decode step, batch <= 64: out[b, n] = bf16(X[b, :] . W[n, :]).  Wt = tile_weight(W).  The activations are either
// k-block-tiled (x_tile_rows = 32 / 64: Xt) or row-major (x_tile_rows = 0: X with pitch ldx, fetched through a tensor map).
extern "C" int dots_decode_gemm_head(const void* X, long long ldx, int x_tile_rows, const void* Wt, void* out_bf16, long long ldo, int batch,
                                     int N, int K, void* stream) {
    DOTS_REQUIRE(X && Wt && out_bf16 && batch > 0 && batch <= 64 && N > 0 && K > 0, "dots_decode_gemm_head: bad arguments batch=%d N=%d K=%d", batch, N, K);
    const int bn = batch <= 32 ? 32 : 64;
    DOTS_REQUIRE(x_tile_rows == 0 || x_tile_rows == bn, "dots_decode_gemm_head: x_tile_rows must be 0 or %d for batch %d", bn, batch);
    GemmParams p{};
    p.static_is_b = 0;
    p.M = N; p.N = batch; p.K = K;
    p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    p.kb_per_split = p.num_k_blocks;
    p.splits = 1;
    p.out = out_bf16; p.ldo = ldo;
    p.m_blocks = (N + BLOCK_M - 1) / BLOCK_M;
    p.n_blocks = 1;
    p.a_tiled = reinterpret_cast<const uint8_t*>(Wt);
    p.stages = g_dec_stages[2];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ta{}, tb{};
    if (x_tile_rows) p.b_tiled = reinterpret_cast<const uint8_t*>(X);
    else {
        DOTS_REQUIRE(K % 8 == 0 && ldx % 8 == 0, "dots_decode_gemm_head: K and ldx must be multiples of 8");
        if (make_tmap_2d_bf16(&tb, X, batch, K, ldx, bn)) return -4;
    }
    return bn == 32 ? launch_gemm<32, DOTS_EPI_BF16_T>(ta, tb, p, st) : launch_gemm<64, DOTS_EPI_BF16_T>(ta, tb, p, st);
}

// Split-K partials over tiled operands: partial[s][b][n] (fp32) = sum over the s-th K slice of X[b, k] * W[n, k]; the 7-kernel decode
// layer with bulk-copied operands (dots_gemm_skinny_bf16 is the tensor-map version for batches of 65..256).
extern "C" int dots_decode_gemm_partial(const void* Xt, const void* Wt, float* partial, int batch, int N, int K, int splits, void* stream) {
    DOTS_REQUIRE(Xt && Wt && partial && batch > 0 && batch <= 64 && N > 0 && K > 0, "dots_decode_gemm_partial: bad arguments batch=%d N=%d K=%d", batch, N, K);
    GemmParams p{};
    p.static_is_b = 0;
    p.M = N; p.N = batch; p.K = K;
    p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
    if (splits < 1) splits = 1;
    if (splits > p.num_k_blocks) splits = p.num_k_blocks;
    p.kb_per_split = (p.num_k_blocks + splits - 1) / splits;
    p.splits = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    DOTS_REQUIRE(p.splits == splits, "dots_decode_gemm_partial: splits=%d does not tile %d k-blocks (would use %d)", splits, p.num_k_blocks, p.splits);
    p.out = partial; p.ldo = N;
    p.m_blocks = (N + BLOCK_M - 1) / BLOCK_M;
    p.n_blocks = 1;
    p.a_tiled = reinterpret_cast<const uint8_t*>(Wt);
    p.b_tiled = reinterpret_cast<const uint8_t*>(Xt);
    p.stages = p.kb_per_split < g_dec_stages[0] ? (p.kb_per_split < 2 ? 2 : p.kb_per_split) : g_dec_stages[0];
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ta{}, tb{};
    return batch <= 32 ? launch_gemm<32, DOTS_EPI_F32_PARTIAL_T>(ta, tb, p, st) : launch_gemm<64, DOTS_EPI_F32_PARTIAL_T>(ta, tb, p, st);
}
