// One persistent kernel for the GEMM chain of a decoder layer during decode (batch <= 64):
//
//   P0  o_proj        partial[s][b][:]   = attn[b] . W_o^T            (swap-AB tcgen05, split-K)
//   P1  finalize      resid += bf16(sum partial);  normed = RMSNorm(resid) * ln_mid
//   P2  gate|up       act[b] = bf16( bf16(silu(bf16 g)) * bf16 u )    (SwiGLU in the epilogue)
//   P3  down_proj     partial[s][b][:]   = act[b] . W_down^T          (split-K)
//   P4  finalize      resid += bf16(sum partial);  normed = RMSNorm(resid) * ln_next
//   P5  qkv (layer+1) partial[s][b][:]   = normed[b] . W_qkv^T        (split-K; consumed by the attention kernel)
//
// Why one kernel: each of these GEMMs gives every SM a single tile, so as separate launches every SM pays the pipeline
// fill / epilogue / drain of six kernels per layer with HBM idle in between.  Here the weight stream never stops: warp 0
// keeps requesting weight tiles of LATER phases (they depend on nothing) into an 8-stage shared-memory ring while the
// other warps are still in the epilogue or waiting at a phase boundary; only the small activation operand of a stage
// waits for the device-wide phase barrier (a counter in global memory; 148 co-resident CTAs, one per SM).
//
// Arithmetic and rounding points are exactly those of dots_gemm_skinny_bf16 / dots_gemm_skinny_swiglu_bf16 /
// dots_decode_residual_rmsnorm (same tiles, same k-ranges, same reduction order): results are bit-identical.
// SURVEY.md §8a rows a20, a21, a15, a16 (decode half).
#include "../common.h"
#include "../ptx.cuh"
#include "experiments.h"

namespace dots {

constexpr int CH_BLOCK_M = 128;
constexpr int CH_BLOCK_K = 64;
#ifndef CH_STAGES_N
#define CH_STAGES_N 4            // 4 stages = 106 KB of shared memory: the next attention kernel's CTA (101 KB) fits beside it
#endif
constexpr int CH_STAGES = CH_STAGES_N;
constexpr int CH_THREADS = 256;
constexpr int CH_A_BYTES = CH_BLOCK_M * CH_BLOCK_K * 2;      // 16 KB weight tile

struct ChainGemm {
    int m_tiles, splits, kb_per_split, num_kb;   // work items = m_tiles * splits (<= gridDim.x), k-blocks per item
    int n_out;                                   // rows of the weight matrix (output features; 2I for gate|up)
};

struct ChainParams {
    ChainGemm g[4];                // 0: o_proj, 1: gate|up, 2: down_proj, 3: qkv of the next layer
    int n_gemm;                    // 3 (last layer: no P5) or 4
    int batch, hidden, inter;
    float* partial;                // fp32 split-K scratch shared by P0 / P3 / P5
    bf16* resid;                   // [batch, hidden] residual stream (updated in place)
    bf16* normed;                  // [batch, hidden]
    bf16* act;                     // [batch, inter]
    const bf16* ln_mid;            // post-attention RMSNorm weight
    const bf16* ln_next;           // next layer's input RMSNorm weight (or the final norm)
    float eps;
    unsigned int* counters;        // [0..4] monotonically increasing device-wide phase counters, [7] launch epoch (all zero initially);
                                   // [16..] optional timing stamps (see CH_TIMING)
};

// Optional phase timing (debug builds: -DCH_TIMING): CTA 0 and CTA `batch` record %globaltimer at each phase boundary
// into counters[16 + cta_slot * 32 + 2 * i] (64-bit, ns).
#ifdef CH_TIMING
__device__ __forceinline__ void ch_stamp(const ChainParams& p, int cta, int i) {
    if (cta != 0 && cta != p.batch) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    reinterpret_cast<unsigned long long*>(p.counters + 16)[(cta == 0 ? 0 : 16) + i] = t;
}
#else
__device__ __forceinline__ void ch_stamp(const ChainParams&, int, int) {}
#endif

template <int BLOCK_N>
struct ChainSmem {
    static constexpr int B_BYTES = BLOCK_N * CH_BLOCK_K * 2;
    static constexpr int STAGE_BYTES = CH_A_BYTES + B_BYTES;
    static constexpr int XCH_BYTES = BLOCK_N * 64 * 2;
    static constexpr int TOTAL = CH_STAGES * STAGE_BYTES + XCH_BYTES + 1024 /*barriers*/ + 1024 /*align*/;
};

__device__ __forceinline__ float ch_silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Device-wide phase barrier, arrive side (one thread per CTA, after a CTA-local barrier over the writers).
__device__ __forceinline__ void phase_arrive(unsigned int* ctr) {
    asm volatile("fence.proxy.async;" ::: "memory");     // the readers may use TMA (async proxy)
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");      // release: cumulative over the CTA barrier before it
}
// Wait side: spin until every CTA of the grid has arrived (bounded: a scheduling bug traps instead of hanging the box).
__device__ __forceinline__ void phase_wait(const unsigned int* ctr, unsigned int target) {
    long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
        if (clock64() - t0 > 8000000000LL) {
            printf("dots: decode-chain phase watchdog block %d (counter %u / %u)\n", (int)blockIdx.x, ld_acquire_u32(ctr), target);
            __trap();
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
}

// x = bf16(sum_s partial[s][b]); resid[b] = bf16(resid[b] + x); normed[b] = RMSNorm(resid[b]) * w   -- one sequence per CTA,
// executed by the 128 epilogue threads (same arithmetic, split order and rounding as decode_residual_rmsnorm_kernel except
// for the order of the fp32 sum of squares, which is a fixed tree here as there).
__device__ __forceinline__ void chain_finalize_row(const ChainParams& p, int b, int splits, const bf16* __restrict__ w, float* s_red,
                                                   int et /*0..127*/) {
    const int H = p.hidden;
    const int nchunks = H >> 3;
    float ss = 0.f;
    float x[2][8];
    int nmine = 0;
    for (int c = et; c < nchunks; c += 128, ++nmine) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* p0 = p.partial + (long long)b * H + c * 8;
        const long long sstride = (long long)p.batch * H;
        // eight splits' worth of 32-byte loads in flight per round trip; adds stay in split order
        for (int s0 = 0; s0 < splits; s0 += 8) {
            float4 a[8], d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (s0 + u < splits) {
                    const float4* ps = reinterpret_cast<const float4*>(p0 + (s0 + u) * sstride);
                    a[u] = __ldcg(ps); d[u] = __ldcg(ps + 1);      // L2 reads: another phase of this kernel read these addresses earlier
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (s0 + u < splits) {
                    acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
                    acc[4] += d[u].x; acc[5] += d[u].y; acc[6] += d[u].z; acc[7] += d[u].w;
                }
            }
        }
        const uint4 rr = __ldcg(reinterpret_cast<const uint4*>(p.resid + (long long)b * H) + c);
        const float r[8] = {bf16_lo(rr.x), bf16_hi(rr.x), bf16_lo(rr.y), bf16_hi(rr.y), bf16_lo(rr.z), bf16_hi(rr.z), bf16_lo(rr.w), bf16_hi(rr.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float v = bf16_round(bf16_round(acc[j]) + r[j]);
            x[nmine][j] = v;
            ss += v * v;
        }
        reinterpret_cast<uint4*>(p.resid + (long long)b * H)[c] =
            make_uint4(pack_bf16x2(x[nmine][0], x[nmine][1]), pack_bf16x2(x[nmine][2], x[nmine][3]),
                       pack_bf16x2(x[nmine][4], x[nmine][5]), pack_bf16x2(x[nmine][6], x[nmine][7]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((et & 31) == 0) s_red[et >> 5] = ss;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float tot = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    const float rinv = rsqrtf(tot / (float)H + p.eps);
    nmine = 0;
    for (int c = et; c < nchunks; c += 128, ++nmine) {
        const uint4 ww = __ldg(reinterpret_cast<const uint4*>(w) + c);
        const float g[8] = {bf16_lo(ww.x), bf16_hi(ww.x), bf16_lo(ww.y), bf16_hi(ww.y), bf16_lo(ww.z), bf16_hi(ww.z), bf16_lo(ww.w), bf16_hi(ww.w)};
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = bf16_round(x[nmine][j] * rinv) * g[j];
        reinterpret_cast<uint4*>(p.normed + (long long)b * H)[c] =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");       // s_red reusable; all writes of this CTA issued
}

// tma_a[g] = weight of GEMM g (box 128 x 64); tma_b[g] = its activation operand (box BLOCK_N x 64):
// attn, normed, act, normed.
struct ChainMaps {
    CUtensorMap a[4];
    CUtensorMap b[4];
};

template <int BLOCK_N>
__global__ void __maxnreg__(160)      // 256 threads; the cap leaves registers for a co-resident attention CTA
decode_chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
    using S = ChainSmem<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + CH_STAGES * CH_A_BYTES;
    bf16* xch = reinterpret_cast<bf16*>(smem + CH_STAGES * S::STAGE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CH_STAGES * S::STAGE_BYTES + S::XCH_BYTES);
    uint64_t* full_bar = bars;                        // [STAGES]
    uint64_t* empty_bar = bars + CH_STAGES;           // [STAGES]
    uint64_t* tmem_full = bars + 2 * CH_STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;             // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_red = reinterpret_cast<float*>(tmem_ptr + 4);   // [4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned int n_ctas = gridDim.x;
    const int cta = blockIdx.x;

    if (warp == 0 && lane == 0) {
        for (int g = 0; g < p.n_gemm; ++g) { prefetch_tensormap(&maps.a[g]); prefetch_tensormap(&maps.b[g]); }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < CH_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_launch_dependents();

    // work item of this CTA in GEMM g (at most one): m tile and k range
    auto item = [&](int g, int& m_blk, int& kb0, int& kb1, int& split) -> bool {
        const ChainGemm& G = p.g[g];
        if (cta >= G.m_tiles * G.splits) return false;
        split = cta % G.splits;
        m_blk = cta / G.splits;
        kb0 = split * G.kb_per_split;
        kb1 = min(G.num_kb, kb0 + G.kb_per_split);
        return true;
    };
    // device-wide counter a GEMM's activation operand waits for: g0 <- predecessor kernel (PDL), g1 <- P1, g2 <- P2, g3 <- P4
    // counters: [0]=P0 done, [1]=P1, [2]=P2, [3]=P3, [4]=P4

    if (warp == 0) {
        // ===================== weight producer: runs ahead of every phase barrier, bounded only by the ring =====================
        if (lane == 0) {
            int n = 0;
            for (int g = 0; g < p.n_gemm; ++g) {
                int m_blk, kb0, kb1, split;
                if (!item(g, m_blk, kb0, kb1, split)) continue;
                for (int kb = kb0; kb < kb1; ++kb, ++n) {
                    const int st = n % CH_STAGES;
                    if (n >= CH_STAGES) mbar_wait(&empty_bar[st], ((n / CH_STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full_bar[st], S::STAGE_BYTES);        // covers the weight tile and the activation tile of the stage
                    tma_load_2d(smem_a + st * CH_A_BYTES, &maps.a[g], kb * CH_BLOCK_K, m_blk * CH_BLOCK_M, &full_bar[st]);
                }
            }
        }
    } else if (warp == 3) {
        // ===================== activation producer: follows the phase barriers =====================
        // device-wide counter a GEMM's activation operand waits for: g0 <- predecessor kernel, g1 <- P1, g2 <- P2, g3 <- P4
        if (lane == 0) {
            pdl_wait();
            const unsigned int target = (ld_acquire_u32(p.counters + 7) + 1u) * n_ctas;      // counters are monotonic across launches
            int n = 0;
            for (int g = 0; g < p.n_gemm; ++g) {
                int m_blk, kb0, kb1, split;
                if (!item(g, m_blk, kb0, kb1, split)) continue;
                if (g > 0) phase_wait(p.counters + (g == 1 ? 1 : g == 2 ? 2 : 4), target);
                for (int kb = kb0; kb < kb1; ++kb, ++n) {
                    const int st = n % CH_STAGES;
                    if (n >= CH_STAGES) mbar_wait(&empty_bar[st], ((n / CH_STAGES) & 1) ^ 1);   // the stage's previous tenant has been consumed
                    tma_load_2d(smem_b + st * S::B_BYTES, &maps.b[g], kb * CH_BLOCK_K, 0, &full_bar[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_bf16(CH_BLOCK_M, BLOCK_N);
        const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem_a));
        const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b));
        int n = 0, acc = 0;
        uint32_t acc_phase = 0;
        for (int g = 0; g < p.n_gemm; ++g) {
            int m_blk, kb0, kb1, split;
            if (!item(g, m_blk, kb0, kb1, split) || kb1 <= kb0) continue;
            mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = kb0; kb < kb1; ++kb, ++n) {
                const int st = n % CH_STAGES;
                mbar_wait(&full_bar[st], (n / CH_STAGES) & 1);
                tc_fence_after();
                if (g == 1 && lane == 0) { if (kb == kb0) ch_stamp(p, cta, 12); if (kb == kb0 + 8) ch_stamp(p, cta, 13); if (kb + 1 == kb1) ch_stamp(p, cta, 14); }
                const uint64_t da = da0 + (uint64_t)(st * (CH_A_BYTES >> 4));
                const uint64_t db = db0 + (uint64_t)(st * (S::B_BYTES >> 4));
                const uint32_t first = (kb > kb0) ? 1u : 0u;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < CH_BLOCK_K / 16; ++k) umma_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, k > 0 ? 1u : first);
                    umma_commit(&empty_bar[st]);
                    if (kb + 1 == kb1) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue / finalize warps =====================
        pdl_wait();
        const unsigned int epoch = ld_acquire_u32(p.counters + 7);
        const unsigned int target = (epoch + 1u) * n_ctas;
        const int wq = warp & 3;
        const int et = threadIdx.x - 128;
        int acc = 0;
        uint32_t acc_phase = 0;
        auto arrive = [&](int ctr) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) phase_arrive(p.counters + ctr);
        };
        auto gemm_epilogue = [&](int g) {
            int m_blk, kb0, kb1, split;
            if (!item(g, m_blk, kb0, kb1, split) || kb1 <= kb0) return;
            const ChainGemm& G = p.g[g];
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            if (g == 1 && et == 0) ch_stamp(p, cta, 15);
            const int row = m_blk * CH_BLOCK_M + wq * 32 + lane;
            const uint32_t t_row = tmem_base + ((uint32_t)(wq * 32) << 16) + acc * BLOCK_N;
            if (g == 1) {
                // gate|up: rows 0-63 gate, 64-127 up (see DOTS_EPI_SWIGLU_T in gemm_tcgen05.cu).  All four warps share the
                // SiLU work: gate warps publish bf16(g) of the upper half of the batch, up warps publish bf16(u) of the lower
                // half; then gate warps finish batch columns [0, BN/2) and up warps [BN/2, BN).
                const int fl = (wq & 1) * 32 + lane;
                const bool is_up = wq >= 2;
                constexpr int HALF = BLOCK_N / 2;
                static_assert(HALF % 16 == 0, "batch tile halves are read in 16-column TMEM chunks");
                float mine[HALF];                       // my accumulator values for the half I finish
                {
                    // publish the half I do NOT finish
                    const int pub0 = is_up ? 0 : HALF;
#pragma unroll
                    for (int c = 0; c < HALF / 16; ++c) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(t_row + pub0 + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) xch[(pub0 + c * 16 + j) * 64 + fl] = __float2bfloat16_rn(__uint_as_float(v[j]));
                    }
                    const int keep0 = is_up ? HALF : 0;
#pragma unroll
                    for (int c = 0; c < HALF / 16; ++c) {
                        uint32_t v[16];
                        tmem_ld_32x32b_x16(t_row + keep0 + c * 16, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) mine[c * 16 + j] = __uint_as_float(v[j]);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                {
                    const int f = m_blk * 64 + fl;
                    const int b0 = is_up ? HALF : 0;
                    if (f < G.n_out / 2) {
#pragma unroll
                        for (int j = 0; j < HALF; ++j) {
                            const int b = b0 + j;
                            if (b < p.batch) {
                                const float other = __bfloat162float(xch[b * 64 + fl]);
                                const float gv = is_up ? other : bf16_round(mine[j]);
                                const float uv = is_up ? bf16_round(mine[j]) : other;
                                p.act[(long long)b * p.inter + f] = __float2bfloat16_rn(bf16_round(ch_silu(gv)) * uv);
                            }
                        }
                    }
                }
            } else {
                // split-K partial store: partial[split][b][feature], lanes write consecutive features
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(t_row + c * 32, v);
                    tmem_ld_wait();
                    if (row < G.n_out) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int b = c * 32 + j;
                            if (b < p.batch) p.partial[((long long)split * p.batch + b) * G.n_out + row] = __uint_as_float(v[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        };
        auto finalize = [&](int wait_ctr, int splits, const bf16* w) {
            if (cta < p.batch) {
                if (et == 0) phase_wait(p.counters + wait_ctr, target);
                asm volatile("bar.sync 1, 128;" ::: "memory");
                chain_finalize_row(p, cta, splits, w, s_red, et);
            }
        };
        const bool stamper = (et == 0);
        if (stamper) ch_stamp(p, cta, 0);
        gemm_epilogue(0); if (stamper) ch_stamp(p, cta, 1); arrive(0);                 // P0 o_proj
        if (stamper) ch_stamp(p, cta, 2);
        finalize(0, p.g[0].splits, p.ln_mid); if (stamper) ch_stamp(p, cta, 3); arrive(1);      // P1
        if (stamper) ch_stamp(p, cta, 4);
        gemm_epilogue(1); if (stamper) ch_stamp(p, cta, 5); arrive(2);                 // P2 gate|up + SwiGLU
        if (stamper) ch_stamp(p, cta, 6);
        gemm_epilogue(2); if (stamper) ch_stamp(p, cta, 7); arrive(3);                 // P3 down_proj
        if (stamper) ch_stamp(p, cta, 8);
        finalize(3, p.g[2].splits, p.ln_next); if (stamper) ch_stamp(p, cta, 9); arrive(4);     // P4
        if (stamper) ch_stamp(p, cta, 10);
        if (p.n_gemm > 3) gemm_epilogue(3);                             // P5 qkv of the next layer (kernel end publishes it)
        if (stamper) ch_stamp(p, cta, 11);
        if (cta == 0 && et == 0) {
            // every CTA has read the epoch before it arrived at P4: once P4 is complete the next launch's epoch may be published
            phase_wait(p.counters + 4, target);
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.counters + 7), "r"(epoch + 1u) : "memory");
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N);
    }
}

}  // namespace dots

using namespace dots;

extern "C" int dots_decode_chain(const void* attn, const void* w_o, const void* w_gu, const void* w_down, const void* w_qkv_next,
                                 float* partial, void* resid, void* normed, void* act, const void* ln_mid, const void* ln_next,
                                 unsigned int* counters, int batch, int hidden, int inter, int qkv_n, int attn_dim, int splits_o,
                                 int splits_down, int splits_qkv, float eps, void* stream) {
    DOTS_REQUIRE(batch > 0 && batch <= 64, "dots_decode_chain: batch must be 1..64 (got %d)", batch);
    DOTS_REQUIRE(hidden % 128 == 0 && inter % 64 == 0 && attn_dim % 64 == 0 && hidden % 64 == 0, "dots_decode_chain: bad dims");
    DOTS_REQUIRE(hidden / 8 <= 256, "dots_decode_chain: hidden <= 2048");
    const int sms = num_sms();
    ChainParams p{};
    auto setg = [&](int i, int n_out, int K, int splits) -> int {
        ChainGemm& G = p.g[i];
        G.n_out = n_out;
        G.m_tiles = (n_out + CH_BLOCK_M - 1) / CH_BLOCK_M;
        G.num_kb = (K + CH_BLOCK_K - 1) / CH_BLOCK_K;
        if (splits < 1) splits = 1;
        G.kb_per_split = (G.num_kb + splits - 1) / splits;
        G.splits = (G.num_kb + G.kb_per_split - 1) / G.kb_per_split;
        return (G.splits == splits && G.m_tiles * G.splits <= sms) ? 0 : -1;
    };
    DOTS_REQUIRE(setg(0, hidden, attn_dim, splits_o) == 0, "dots_decode_chain: o_proj tiles x splits must tile K and fit the SM count");
    DOTS_REQUIRE(setg(1, 2 * inter, hidden, 1) == 0, "dots_decode_chain: gate|up tiles must fit the SM count");
    DOTS_REQUIRE(setg(2, hidden, inter, splits_down) == 0, "dots_decode_chain: down_proj tiles x splits must tile K and fit the SM count");
    p.n_gemm = 3;
    if (w_qkv_next != nullptr) {
        DOTS_REQUIRE(setg(3, qkv_n, hidden, splits_qkv) == 0, "dots_decode_chain: qkv tiles x splits must tile K and fit the SM count");
        p.n_gemm = 4;
    }
    p.batch = batch; p.hidden = hidden; p.inter = inter;
    p.partial = partial; p.resid = (bf16*)resid; p.normed = (bf16*)normed; p.act = (bf16*)act;
    p.ln_mid = (const bf16*)ln_mid; p.ln_next = (const bf16*)ln_next; p.eps = eps; p.counters = counters;
    const int bn = batch <= 32 ? 32 : 64;
    ChainMaps maps;
    if (make_tmap_2d_bf16(&maps.a[0], w_o, hidden, attn_dim, attn_dim, CH_BLOCK_M)) return -4;
    if (make_tmap_2d_bf16(&maps.a[1], w_gu, 2 * inter, hidden, hidden, CH_BLOCK_M)) return -4;
    if (make_tmap_2d_bf16(&maps.a[2], w_down, hidden, inter, inter, CH_BLOCK_M)) return -4;
    if (make_tmap_2d_bf16(&maps.b[0], attn, batch, attn_dim, attn_dim, bn)) return -4;
    if (make_tmap_2d_bf16(&maps.b[1], normed, batch, hidden, hidden, bn)) return -4;
    if (make_tmap_2d_bf16(&maps.b[2], act, batch, inter, inter, bn)) return -4;
    if (p.n_gemm > 3) {
        if (make_tmap_2d_bf16(&maps.a[3], w_qkv_next, qkv_n, hidden, hidden, CH_BLOCK_M)) return -4;
        if (make_tmap_2d_bf16(&maps.b[3], normed, batch, hidden, hidden, bn)) return -4;
    } else {
        maps.a[3] = maps.a[0]; maps.b[3] = maps.b[0];
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(decode_chain_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmem<32>::TOTAL));
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(decode_chain_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainSmem<64>::TOTAL));
    }
    // Every CTA must be resident at once (device-wide phase barriers): one per SM.  Programmatic dependent launch lets the
    // CTAs start streaming weights while the attention kernel before them drains; all mutable state is read after pdl_wait().
    if (bn == 32) DOTS_CHECK_CUDA(launch_ex(decode_chain_kernel<32>, dim3(sms), dim3(CH_THREADS), (size_t)ChainSmem<32>::TOTAL, st, true, maps, p));
    else DOTS_CHECK_CUDA(launch_ex(decode_chain_kernel<64>, dim3(sms), dim3(CH_THREADS), (size_t)ChainSmem<64>::TOTAL, st, true, maps, p));
    return 0;
}
