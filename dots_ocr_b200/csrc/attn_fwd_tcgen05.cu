// Variable-length fused attention forward on the 5th-gen tensor cores (head_dim 128).
//
// One CTA owns TWO 128-row query tiles (A and B) of one (sequence, head) and ping-pongs them so the
// tensor pipe works on one tile while the other tile's softmax runs (FA4-style schedule):
//
//   warp 0        TMA producer: Q_A, Q_B once; K_j / V_j tiles (128 keys) through mbarrier rings
//   warp 1        tcgen05.mma issuer:  S_t = Q_t K_j^T  (SS: both operands in shared memory)
//                                      O_t += P_t V_j   (TS: P read from TMEM, V MN-major in shared memory)
//   warps 4-7     softmax warpgroup for tile A  (thread = query row = TMEM lane)
//   warps 8-11    softmax warpgroup for tile B
//
// TMEM (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,384)  O_B [384,512); P_t (bf16 pairs) overwrites
// the first 64 columns of S_t.  Softmax is fp32 with exp2 and a pre-scaled log2(e); the running max is only
// advanced when it grows by more than 2^8 (lazy rescale), so O in TMEM is rarely touched by the softmax warps.
// P is rounded to bf16 before P*V, the row sum is accumulated from the unrounded fp32 values
// (flash_attention_2's rounding points).
//
// SURVEY.md §8a rows a10 (ViT, bidirectional, one segment per image) and a19 (LLM prefill, causal GQA).
#include "common.h"
#include "ptx.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int FA_D = 128;
constexpr int FA_BM = 128;          // rows per query tile (two tiles per CTA)
constexpr int FA_BN = 128;          // keys per KV tile
constexpr int FA_THREADS = 384;
constexpr int FA_TILE_BYTES = 128 * 128 * 2;        // 32 KB: one [128 x 128] bf16 tile = two 16-KB swizzle boxes
constexpr int FA_KSTAGES = 3;
constexpr int FA_VSTAGES = 2;
constexpr int FA_SMEM = (2 + FA_KSTAGES + FA_VSTAGES) * FA_TILE_BYTES + 1024 /*barriers*/ + 1024 /*align*/;

struct FaParams {
    const int* cu;
    bf16* o;
    long long os;
    int n_q_heads, group;
    float scale_log2;
};

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                         // [2 tiles][2 boxes][128 rows][128 B]
    uint8_t* sK = sQ + 2 * FA_TILE_BYTES;                       // [KSTAGES]
    uint8_t* sV = sK + FA_KSTAGES * FA_TILE_BYTES;              // [VSTAGES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FA_VSTAGES * FA_TILE_BYTES);
    uint64_t* q_full = bars;                    // [1]
    uint64_t* k_full = bars + 1;                // [KSTAGES]
    uint64_t* k_empty = k_full + FA_KSTAGES;    // [KSTAGES]
    uint64_t* v_full = k_empty + FA_KSTAGES;    // [VSTAGES]
    uint64_t* v_empty = v_full + FA_VSTAGES;    // [VSTAGES]
    uint64_t* s_full = v_empty + FA_VSTAGES;    // [2]   MMA -> softmax: S_t ready
    uint64_t* p_full = s_full + 2;              // [2]   softmax -> MMA: P_t written (and O_t rescaled)
    uint64_t* o_done = p_full + 2;              // [2]   MMA -> softmax: last P*V of tile t retired
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

    const int seq = blockIdx.z, head = blockIdx.y;
    const int tok0 = p.cu[seq];
    const int L = p.cu[seq + 1] - tok0;
    const int q0 = blockIdx.x * (2 * FA_BM);
    if (q0 >= L) return;
    const int kvh = head / p.group;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int n_kv = (L + FA_BN - 1) / FA_BN;
    if (CAUSAL) n_kv = min(n_kv, (q0 + 2 * FA_BM + FA_BN - 1) / FA_BN);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_q); prefetch_tensormap(&tm_k); prefetch_tensormap(&tm_v);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < FA_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
        for (int i = 0; i < FA_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_done[i], 1); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");   // warpgroup 0 hands its registers to the softmax warpgroups
    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * FA_TILE_BYTES);
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sQ + t * FA_TILE_BYTES + h * (FA_TILE_BYTES / 2), &tm_q, head * FA_D + h * 64,
                                tok0 + q0 + t * FA_BM, q_full);
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&k_empty[ks], kph ^ 1);
                mbar_expect_tx(&k_full[ks], FA_TILE_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sK + ks * FA_TILE_BYTES + h * (FA_TILE_BYTES / 2), &tm_k, kvh * FA_D + h * 64, tok0 + j * FA_BN,
                                &k_full[ks]);
                if (++ks == FA_KSTAGES) { ks = 0; kph ^= 1; }
                mbar_wait(&v_empty[vs], vph ^ 1);
                mbar_expect_tx(&v_full[vs], FA_TILE_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sV + vs * FA_TILE_BYTES + h * (FA_TILE_BYTES / 2), &tm_v, kvh * FA_D + h * 64, tok0 + j * FA_BN,
                                &v_full[vs]);
                if (++vs == FA_VSTAGES) { vs = 0; vph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);     // S = Q K^T : A, B K-major
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 0, 1);     // O += P V : A from TMEM, B (V) MN-major
            const uint32_t tS[2] = {tmem_base + 0, tmem_base + 128};
            const uint32_t tO[2] = {tmem_base + 256, tmem_base + 384};

            auto issue_s = [&](int t, int ks) {
                const uint32_t qa = smem_u32(sQ + t * FA_TILE_BYTES);
                const uint32_t ka = smem_u32(sK + ks * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = (k >> 2) * (FA_TILE_BYTES / 2) + (k & 3) * 32;      // box, then 32 B per k-step
                    umma_bf16_ss(tS[t], umma_desc_k_sw128(qa + off), umma_desc_k_sw128(ka + off), idesc_s, k > 0 ? 1u : 0u);
                }
                umma_commit(&s_full[t]);
            };
            auto issue_pv = [&](int t, int vs, bool first) {
                const uint32_t va = smem_u32(sV + vs * FA_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // 16 keys per step: 16 rows x 128 B = 2048 B into the tile; d halves are 16 KB apart (LBO), 8-row groups 1 KB (SBO)
                    const uint64_t vd = umma_desc_mn_sw128(va + k * 2048, FA_TILE_BYTES / 2, 1024);
                    umma_bf16_ts(tO[t], tS[t] + k * 8, vd, idesc_o, (first && k == 0) ? 0u : 1u);
                }
            };

            mbar_wait(q_full, 0);
            int ks = 0, vs = 0;
            uint32_t kph = 0, vph = 0, pph = 0;
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0);
            issue_s(1, 0);
            umma_commit(&k_empty[0]);                       // K_0 free once both S MMAs retire
            int ks_next = 1 % FA_KSTAGES;
            uint32_t kph_next = (FA_KSTAGES == 1) ? 1u : 0u;
            (void)ks; (void)kph;
            for (int j = 0; j < n_kv; ++j) {
                const bool more = (j + 1 < n_kv);
                mbar_wait(&v_full[vs], vph);
                if (more) mbar_wait(&k_full[ks_next], kph_next);
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    mbar_wait(&p_full[t], pph);
                    tc_fence_after();
                    issue_pv(t, vs, j == 0);
                    if (more) issue_s(t, ks_next);          // overwrites S_t / P_t: ordered after P_t V_j on the tensor pipe
                    else umma_commit(&o_done[t]);
                }
                umma_commit(&v_empty[vs]);
                if (more) umma_commit(&k_empty[ks_next]);
                pph ^= 1;
                if (++vs == FA_VSTAGES) { vs = 0; vph ^= 1; }
                if (more) { if (++ks_next == FA_KSTAGES) { ks_next = 0; kph_next ^= 1; } }
            }
        }
    } else if (warp >= 4) {
        // =============================== softmax warpgroups ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");       // one query row = 128 fp32 scores in registers
        const int t = (warp - 4) >> 2;                   // 0: tile A, 1: tile B
        const int wq = warp & 3;                         // TMEM lane quarter
        const int row = wq * 32 + lane;                  // row within the tile == TMEM lane
        const int qi = q0 + t * FA_BM + row;             // query index within the sequence
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + t * 128;
        const uint32_t tO = tmem_base + lane_addr + 256 + t * 128;
        float m_run = -INFINITY;                         // running max in raw score units
        float l_run = 0.f;
        uint32_t sph = 0;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[t], sph);
            sph ^= 1;
            tc_fence_after();
            float s[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
            tmem_ld_wait();
            const int k0 = j * FA_BN;
            const bool need_mask = (k0 + FA_BN > L) || (CAUSAL && (k0 + FA_BN - 1 > q0 + t * FA_BM + wq * 32));
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < 128; ++i) {
                    const int kj = k0 + i;
                    if (kj >= L || (CAUSAL && kj > qi)) s[i] = -INFINITY;
                }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 128; ++i) mx = fmaxf(mx, s[i]);
            // lazy rescale: only move the reference max when it would grow by more than 2^8 in the exp2 domain
            const bool need = (mx > m_run) && ((mx - m_run) * p.scale_log2 > 8.0f);
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_run;
                const float alpha = (m_run == -INFINITY) ? 0.f : ex2f((m_run - m_new) * p.scale_log2);
                if (j > 0) {
                    // S_t(j) complete implies P_t V_{j-1} retired (in-order tensor pipe): O_t is stable here
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tO + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32b_x32(tO + c * 32, v);
                    }
                }
                l_run *= alpha;
                m_run = m_new;
            }
            const float mneg = (m_run == -INFINITY) ? 0.f : -m_run * p.scale_log2;
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float p0 = ex2f(fmaf(s[c * 64 + 2 * i], p.scale_log2, mneg));
                    const float p1 = ex2f(fmaf(s[c * 64 + 2 * i + 1], p.scale_log2, mneg));
                    lsum += p0 + p1;
                    pk[i] = pack_bf16x2(p0, p1);
                }
                tmem_st_32x32b_x32(tS + c * 32, pk);
            }
            l_run += lsum;
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[t]);
        }
        // ---- epilogue: O_t / l -> bf16 -> global ------------------------------------------
        mbar_wait(&o_done[t], 0);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* dst = p.o + (long long)(tok0 + qi) * p.os + head * FA_D;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tO + c * 32, v);
            tmem_ld_wait();
            if (qi < L) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 o4;
                    o4.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv, __uint_as_float(v[q * 8 + 1]) * inv);
                    o4.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv, __uint_as_float(v[q * 8 + 3]) * inv);
                    o4.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv, __uint_as_float(v[q * 8 + 5]) * inv);
                    o4.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv, __uint_as_float(v[q * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = o4;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace dots

using namespace dots;

extern "C" int dots_attn_varlen_fwd_tc(const void* q, long long q_stride, const void* k, long long k_stride, const void* v,
                                       long long v_stride, void* out, long long o_stride, const int* cu_seqlens, int n_seqs,
                                       int max_seqlen, long long total_tokens, int n_q_heads, int n_kv_heads, int head_dim,
                                       int causal, float softmax_scale, void* stream) {
    DOTS_REQUIRE(head_dim == FA_D, "dots_attn_varlen_fwd_tc: head_dim must be 128 (got %d)", head_dim);
    DOTS_REQUIRE(n_seqs > 0 && max_seqlen > 0 && total_tokens > 0 && n_q_heads > 0 && n_kv_heads > 0 && n_q_heads % n_kv_heads == 0,
                 "dots_attn_varlen_fwd_tc: bad shape");
    DOTS_REQUIRE(q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && o_stride % 8 == 0,
                 "dots_attn_varlen_fwd_tc: token strides must be multiples of 8 elements");
    DOTS_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "dots_attn_varlen_fwd_tc: 16-byte aligned pointers");
    CUtensorMap tq, tk, tv;
    if (make_tmap_2d_bf16(&tq, q, total_tokens, (uint64_t)n_q_heads * FA_D, q_stride, 128, 64)) return -4;
    if (make_tmap_2d_bf16(&tk, k, total_tokens, (uint64_t)n_kv_heads * FA_D, k_stride, 128, 64)) return -4;
    if (make_tmap_2d_bf16(&tv, v, total_tokens, (uint64_t)n_kv_heads * FA_D, v_stride, 128, 64)) return -4;
    FaParams p;
    p.cu = cu_seqlens; p.o = (bf16*)out; p.os = o_stride;
    p.n_q_heads = n_q_heads; p.group = n_q_heads / n_kv_heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    dim3 grid((max_seqlen + 2 * FA_BM - 1) / (2 * FA_BM), n_q_heads, n_seqs);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured = false;
    if (!configured) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        configured = true;
    }
    if (causal) attn_fwd_tcgen05_kernel<true><<<grid, FA_THREADS, FA_SMEM, st>>>(tq, tk, tv, p);
    else attn_fwd_tcgen05_kernel<false><<<grid, FA_THREADS, FA_SMEM, st>>>(tq, tk, tv, p);
    DOTS_LAUNCH_CHECK();
    return 0;
}
