// Variable-length fused attention forward on the 5th-gen tensor cores (head_dim 128).
//
// One CTA owns TWO 128-row query tiles ("chains" 0 and 1) of one (sequence, head); the chains are independent dependency
// chains (S -> softmax -> P -> O), so the tensor pipe works on one while the other is in its softmax:
//
//   warp 0        TMA producer: Q_0, Q_1 once; K_j / V_j tiles through mbarrier rings (128-B swizzle boxes)
//   warp 1 / 2    tcgen05.mma issuers, one per chain (the warp runs the control flow convergently, one elected lane issues):
//                     S_t(j) = Q_t K_j^T   (SS: both operands in shared memory)
//                     O_t  += P_t(j) V_j   (TS: P read from TMEM, V MN-major in shared memory)
//   warps 4-7     softmax warpgroup of chain 0  (thread = query row = TMEM lane)
//   warps 8-11    softmax warpgroup of chain 1
//
// Two tilings (FA_BN_KEYS): 128-key tiles with one S buffer per chain (default: S(j+1) is issued right behind P(j) V_j), or
// 64-key tiles with S double-buffered in TMEM (S(j+1) computed during softmax(j)); measured 1107/1235 vs 1042/1150 TFLOP/s at
// L = 5476 / 19600.  TMEM (512 columns): chain t at t*256: S [0,128) (or two 64-column buffers), O [128,256).  P(j) (bf16
// pairs) overwrites the first half of the S buffer it came from.  Softmax is fp32 in the exp2 domain with packed fp32x2
// arithmetic; the running max only advances when it grows by more than 2^8 (lazy rescale), so O in TMEM is rarely touched
// by the softmax warps.  P is rounded to bf16 before P*V, the row sum accumulates the unrounded fp32 values
// (flash_attention_2's rounding points).
//
// SURVEY.md §8a rows a10 (ViT, bidirectional, one segment per image) and a19 (LLM prefill, causal GQA).
#include "common.h"
#include "ptx.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int FA_D = 128;
constexpr int FA_BM = 128;          // rows per query tile (two tiles per CTA)
#ifndef FA_BN_KEYS
#define FA_BN_KEYS 128
#endif
constexpr int FA_BN = FA_BN_KEYS;   // keys per KV tile: 64 (S double-buffered in TMEM) or 128 (one S buffer per chain)
constexpr int FA_SBUF = (FA_BN == 64) ? 2 : 1;      // S buffers per chain (TMEM: 2 x 64 or 1 x 128 columns)
static_assert(FA_BN == 64 || FA_BN == 128, "FA_BN_KEYS must be 64 or 128");
#ifndef FA_SPLIT_N
#define FA_SPLIT_N 1
#endif
constexpr int FA_SPLIT = FA_SPLIT_N;    // softmax warps per TMEM lane quarter and chain: 2 = each thread handles half of a row's key columns
static_assert(FA_SPLIT == 1 || (FA_SPLIT == 2 && FA_BN_KEYS == 128), "the split softmax is written for 128-key tiles");
constexpr int FA_THREADS = 128 + 256 * FA_SPLIT;
constexpr int FA_QTILE_BYTES = FA_BM * FA_D * 2;     // 32 KB = two [128 x 64] swizzle boxes
constexpr int FA_KVTILE_BYTES = FA_BN * FA_D * 2;    // 16 KB = two [64 x 64] swizzle boxes
constexpr int FA_KSTAGES = (FA_BN == 64) ? 4 : 3;
constexpr int FA_VSTAGES = (FA_BN == 64) ? 4 : 2;
// FA_SPLIT == 2 adds a 2 KB row-max / row-sum exchange buffer between the two column halves; with 128-key tiles that only fits the
// 227 KB limit without alignment slack, so that build requires (and checks) a 1024-byte aligned dynamic shared memory base.
constexpr int FA_SMEM = (FA_SPLIT_N == 2)
    ? 2 * FA_QTILE_BYTES + (FA_KSTAGES + FA_VSTAGES) * FA_KVTILE_BYTES + 512 /*barriers*/ + 2048 /*exchange*/ + 512
    : 2 * FA_QTILE_BYTES + (FA_KSTAGES + FA_VSTAGES) * FA_KVTILE_BYTES + 1024 /*barriers*/ + 1024 /*align*/;
static_assert(FA_SMEM <= 232448, "shared memory budget");

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

// exp2 of two packed fp32 values on the FMA/ALU pipes (no MUFU): round-to-nearest range reduction with the 1.5*2^23
// magic constant, a degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, far below the
// bf16 rounding applied to P), and the integer part added straight into the exponent field.  The softmax warps run
// one pair in four through this path so the MUFU pipe (16 ex2 / clk / SM) stops being the co-bottleneck of the
// tensor pipe -- the FlashAttention-4 trick.
__device__ __forceinline__ void ex2_poly_f32x2(uint64_t x2, float& p0, float& p1) {
    float x0, x1;
    unpack_f32x2(x2, x0, x1);
    const uint64_t xc = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
    const uint64_t t = fadd_f32x2(xc, pack_f32x2(12582912.f, 12582912.f));
    const uint64_t n = fadd_f32x2(t, pack_f32x2(-12582912.f, -12582912.f));
    const uint64_t f = ffma_f32x2(n, pack_f32x2(-1.f, -1.f), xc);
    uint64_t q = ffma_f32x2(f, pack_f32x2(0.055171459913253784f, 0.055171459913253784f), pack_f32x2(0.2426108568906784f, 0.2426108568906784f));
    q = ffma_f32x2(q, f, pack_f32x2(0.6932609677314758f, 0.6932609677314758f));
    q = ffma_f32x2(q, f, pack_f32x2(0.9999281167984009f, 0.9999281167984009f));
    float t0, t1, q0, q1;
    unpack_f32x2(t, t0, t1);
    unpack_f32x2(q, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

#ifndef FA_EMU_MASK
#define FA_EMU_MASK 0u               // bit i set: packed pair i of the 32 pairs of a row takes the polynomial path
#endif

// S buffer (and its barrier) used by key tile j, and the parity of that barrier's phase for tile j
__device__ __forceinline__ constexpr int fa_buf(int j) { return FA_SBUF == 2 ? (j & 1) : 0; }
__device__ __forceinline__ constexpr uint32_t fa_par(int j) { return (uint32_t)(FA_SBUF == 2 ? ((j >> 1) & 1) : (j & 1)); }

template <bool CAUSAL>
__global__ void __launch_bounds__(FA_THREADS, 1)
attn_fwd_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const FaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                         // [2 chains][2 boxes][128 rows][128 B]
    uint8_t* sK = sQ + 2 * FA_QTILE_BYTES;                      // [KSTAGES][2 boxes][64 rows][128 B]
    uint8_t* sV = sK + FA_KSTAGES * FA_KVTILE_BYTES;            // [VSTAGES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FA_VSTAGES * FA_KVTILE_BYTES);
    uint64_t* q_full = bars;                    // [1]
    uint64_t* k_full = bars + 1;                // [KSTAGES]
    uint64_t* k_empty = k_full + FA_KSTAGES;    // [KSTAGES]
    uint64_t* v_full = k_empty + FA_KSTAGES;    // [VSTAGES]
    uint64_t* v_empty = v_full + FA_VSTAGES;    // [VSTAGES]
    uint64_t* s_full = v_empty + FA_VSTAGES;    // [2 chains][2 buffers]  MMA -> softmax: S_t(j) ready in buffer j&1
    uint64_t* p_full = s_full + 4;              // [2][2]                 softmax -> MMA: P_t(j) written over buffer j&1
    uint64_t* pv_done = p_full + 4;             // [2][2]                 MMA -> softmax: O_t += P_t(j) V_j retired
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 4);
    float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 512);      // [2 chains][2 halves][128 rows] (FA_SPLIT == 2)
    if (FA_SPLIT == 2 && (smem_u32(smem_raw) & 1023u) != 0u) __trap();

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
        for (int i = 0; i < FA_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }   // released by both issuing warps
        for (int i = 0; i < FA_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
        for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], FA_SPLIT == 1 ? 128 : 8); mbar_init(&pv_done[i], 1); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // The CTA allocates all 512 TMEM columns, so the allocation starts at column 0 / lane 0.  Using the literal keeps every
    // tcgen05.mma operand in uniform registers: the single issuing thread spends ~10 instructions per MMA instead of ~45
    // (per-MMA broadcast loops on a non-uniform TMEM address), which matters when one 128x64x16 MMA is only 32 clocks.
    if (*tmem_ptr != 0u) __trap();
    constexpr uint32_t tmem_base = 0u;

    // register hand-over: 128 x (168 - dec) released >= 256 x (inc - 168) acquired
    if (FA_SPLIT == 1 && warp < 4) {
        if (FA_BN == 64) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    }
    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * FA_QTILE_BYTES);
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sQ + t * FA_QTILE_BYTES + h * (FA_QTILE_BYTES / 2), &tm_q, head * FA_D + h * 64,
                                tok0 + q0 + t * FA_BM, q_full);
            for (int j = 0; j < n_kv; ++j) {
                const int ks = j % FA_KSTAGES, vs = j % FA_VSTAGES;
                mbar_wait(&k_empty[ks], ((j / FA_KSTAGES) & 1) ^ 1);
                mbar_expect_tx(&k_full[ks], FA_KVTILE_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sK + ks * FA_KVTILE_BYTES + h * (FA_KVTILE_BYTES / 2), &tm_k, kvh * FA_D + h * 64, tok0 + j * FA_BN,
                                &k_full[ks]);
                mbar_wait(&v_empty[vs], ((j / FA_VSTAGES) & 1) ^ 1);
                mbar_expect_tx(&v_full[vs], FA_KVTILE_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(sV + vs * FA_KVTILE_BYTES + h * (FA_KVTILE_BYTES / 2), &tm_v, kvh * FA_D + h * 64, tok0 + j * FA_BN,
                                &v_full[vs]);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =============================== MMA issuers: warp 1 drives chain 0, warp 2 drives chain 1 ===============================
        // Two issuing warps so that one chain's softmax hand-off never waits behind the other chain's.
        // The whole warp runs the control flow (waits, descriptor arithmetic) convergently so the compiler keeps every
        // tcgen05 operand in uniform registers; only the instructions themselves sit under the elected-lane predicate.
        {
            const int n_kv_u = __shfl_sync(0xffffffffu, n_kv, 0);       // warp-uniform trip count
            const int t = __shfl_sync(0xffffffffu, warp - 1, 0);        // this warp's chain
            const bool leader = elect_one();
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, FA_BN, 0, 0);   // S = Q K^T : A, B K-major
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 128, 0, 1);     // O += P V : A from TMEM, B (V) MN-major
            // Descriptors are base + (byte offset >> 4): the address field is the low 14 bits in 16-byte units and shared
            // memory tops out below 2^18 bytes, so plain 64-bit adds never carry out of the field.
            const uint64_t dq0 = umma_desc_k_sw128(smem_u32(sQ));
            const uint64_t dk0 = umma_desc_k_sw128(smem_u32(sK));
            const uint64_t dv0 = umma_desc_mn_sw128(smem_u32(sV), FA_KVTILE_BYTES / 2, 1024);
            auto issue_s = [&](int t, int j) {
                const uint32_t tS = tmem_base + t * 256 + fa_buf(j) * 64;
                const uint64_t dq = dq0 + (uint64_t)(t * (FA_QTILE_BYTES >> 4));
                const uint64_t dk = dk0 + (uint64_t)((j % FA_KSTAGES) * (FA_KVTILE_BYTES >> 4));
                uint64_t* bar = &s_full[t * 2 + fa_buf(j)];
                if (leader) {
#ifndef FA_DBG_NOS
#pragma unroll
                    for (int k = 0; k < 8; ++k) {      // 16 head-dim elements per step: box k>>2, 32 B inside the swizzle row
                        const uint64_t qoff = (uint64_t)(((k >> 2) * (FA_QTILE_BYTES / 2) + (k & 3) * 32) >> 4);
                        const uint64_t koff = (uint64_t)(((k >> 2) * (FA_KVTILE_BYTES / 2) + (k & 3) * 32) >> 4);
                        umma_bf16_ss(tS, dq + qoff, dk + koff, idesc_s, k > 0 ? 1u : 0u);
                    }
#endif
                    umma_commit(bar);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int t, int j) {
                const uint32_t tP = tmem_base + t * 256 + fa_buf(j) * 64;
                const uint32_t tO = tmem_base + t * 256 + 128;
                const uint64_t dv = dv0 + (uint64_t)((j % FA_VSTAGES) * (FA_KVTILE_BYTES >> 4));
                const uint32_t acc0 = (j == 0) ? 0u : 1u;
                uint64_t* bar = &pv_done[t * 2 + fa_buf(j)];
                if (leader) {
#ifndef FA_DBG_NOPV
#pragma unroll
                    for (int k = 0; k < FA_BN / 16; ++k) {
                        // 16 keys per step: 16 rows x 128 B = 2048 B into the box; d halves are 8 KB apart (LBO), 8-row groups 1 KB (SBO)
                        umma_bf16_ts(tO, tP + k * 8, dv + (uint64_t)((k * 2048) >> 4), idesc_o, k == 0 ? acc0 : 1u);
                    }
#endif
                    umma_commit(bar);
                }
                __syncwarp();
            };
            auto commit = [&](uint64_t* bar) {
                if (leader) umma_commit(bar);
                __syncwarp();
            };

            mbar_wait(q_full, 0);
            for (int jj = 0; jj < FA_SBUF && jj < n_kv_u; ++jj) {
                mbar_wait(&k_full[jj % FA_KSTAGES], 0);
                tc_fence_after();
                issue_s(t, jj);
                commit(&k_empty[jj % FA_KSTAGES]);
            }
            for (int j = 0; j < n_kv_u; ++j) {
                const int jn = j + FA_SBUF;                 // the S tile that reuses the buffer P_t(j) lives in
                const bool more = (jn < n_kv_u);
                mbar_wait(&v_full[j % FA_VSTAGES], (j / FA_VSTAGES) & 1);
                if (more) mbar_wait(&k_full[jn % FA_KSTAGES], (jn / FA_KSTAGES) & 1);
                mbar_wait(&p_full[t * 2 + fa_buf(j)], fa_par(j));
                tc_fence_after();
                issue_pv(t, j);
                if (more) issue_s(t, jn);                   // overwrites that buffer: ordered after P_t(j) V_j (same issuing thread)
                commit(&v_empty[j % FA_VSTAGES]);
                if (more) commit(&k_empty[jn % FA_KSTAGES]);
            }
        }
    } else if (FA_SPLIT == 2 && warp >= 4) {
#if FA_SPLIT_N == 2
        // =============================== softmax, two warps per lane quarter: each thread owns HALF of a row's key columns ===============================
        // Halving the per-thread work halves the S -> P latency of a chain (the kernel is bound by that hand-off chain, not by
        // MUFU or issue slots); the two halves agree on the row max through shared memory, keep partial row sums and each
        // finishes 64 of the 128 output dims.
        const int t = (warp - 4) >> 3;                   // chain
        const int hh = ((warp - 4) >> 2) & 1;            // column half
        const int wq = warp & 3;                         // TMEM lane quarter
        const int row = wq * 32 + lane;
        const int qi = q0 + t * FA_BM + row;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + t * 256;
        const uint32_t tSh = tS + hh * 64;               // my 64 score columns
        const uint32_t tPh = tS + hh * 32;               // my 32 packed-P columns
        const uint32_t tOh = tS + 128 + hh * 64;         // my 64 output dims
        float m_run = -INFINITY;
        float l_run = 0.f;                               // partial row sum over my columns
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[t * 2], j & 1);
            tc_fence_after();
            const int k0 = j * FA_BN + hh * 64;
            const bool need_mask = (k0 + 64 > L) || (CAUSAL && (k0 + 63 > q0 + t * FA_BM + wq * 32));
            // ---- pass 1: row max of my half ----
            float mxl = -INFINITY;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float s[32];
                tmem_ld_32x32b_x32(tSh + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
                tmem_ld_wait();
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kj = k0 + c * 32 + i;
                        if (kj >= L || (CAUSAL && kj > qi)) s[i] = -INFINITY;
                    }
                }
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) m4[q] = fmaxf(m4[q], fmaxf(s[i + 2 * q], s[i + 2 * q + 1]));
                }
                mxl = fmaxf(mxl, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
            }
            float* xb = xch + t * 256;      // single buffer: the second barrier of a step orders its reads before the next step's writes
            xb[hh * 128 + row] = mxl;
            asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
            const float mx = fmaxf(mxl, xb[(hh ^ 1) * 128 + row]);
            // lazy rescale: both halves see the same (mx, m_run) and take the same decision
            const bool need = (mx > m_run) && ((mx - m_run) * p.scale_log2 > 8.0f);
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_run;
                const float alpha = (m_run == -INFINITY) ? 0.f : ex2f((m_run - m_new) * p.scale_log2);
                if (j > 0) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tOh + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32b_x32(tOh + c * 32, v);
                    }
                }
                l_run *= alpha;
                m_run = m_new;
            }
            const float mneg = (m_run == -INFINITY) ? 0.f : -m_run * p.scale_log2;
            const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), mn2 = pack_f32x2(mneg, mneg);
            uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
            // ---- pass 2: p = exp2(s * scale - m * scale), bf16 P over my packed columns ----
            uint32_t pk[32];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float s[32];
                tmem_ld_32x32b_x32(tSh + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
                tmem_ld_wait();
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kj = k0 + c * 32 + i;
                        if (kj >= L || (CAUSAL && kj > qi)) s[i] = -INFINITY;
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint64_t x2 = ffma_f32x2(pack_f32x2(s[2 * i], s[2 * i + 1]), sc2, mn2);
                    float x0, x1;
                    unpack_f32x2(x2, x0, x1);
                    const float p0 = ex2f(x0), p1 = ex2f(x1);
                    acc2[i & 3] = fadd_f32x2(acc2[i & 3], pack_f32x2(p0, p1));
                    pk[c * 16 + i] = pack_bf16x2(p0, p1);
                }
            }
            // every thread of the chain must have finished READING S before anyone overwrites it with P: my partner's columns
            // [0, 32) of S are my partner's P target and vice versa
            tc_fence_before();
            asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
            tc_fence_after();
            tmem_st_32x32b_x32(tPh, pk);
            {
                float a0, a1, b0, b1;
                unpack_f32x2(fadd_f32x2(acc2[0], acc2[1]), a0, a1);
                unpack_f32x2(fadd_f32x2(acc2[2], acc2[3]), b0, b1);
                l_run += (a0 + a1) + (b0 + b1);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t * 2]);
        }
        // ---- epilogue: combine the partial row sums, then my 64 dims of O_t / l -> bf16 -> global ----
        float* xb = xch + t * 256;
        xb[hh * 128 + row] = l_run;
        asm volatile("bar.sync %0, 256;" ::"r"(1 + t) : "memory");
        const float l_tot = l_run + xb[(hh ^ 1) * 128 + row];
        mbar_wait(&pv_done[t * 2], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
        bf16* dst = p.o + (long long)(tok0 + qi) * p.os + head * FA_D + hh * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tOh + c * 32, v);
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
#endif
    } else if (FA_SPLIT == 1 && warp >= 4) {
        // =============================== softmax warpgroups ===============================
        if (FA_BN == 64) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int t = (warp - 4) >> 2;                   // chain
        const int wq = warp & 3;                         // TMEM lane quarter
        const int row = wq * 32 + lane;                  // row within the tile == TMEM lane
        const int qi = q0 + t * FA_BM + row;             // query index within the sequence
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tSbase = tmem_base + lane_addr + t * 256;
        const uint32_t tO = tSbase + 128;
        float m_run = -INFINITY;                         // running max in raw score units
        float l_run = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            const uint32_t tS = tSbase + fa_buf(j) * 64;
            mbar_wait(&s_full[t * 2 + fa_buf(j)], fa_par(j));
            tc_fence_after();
#ifdef FA_DBG_NOSOFTMAX
            tc_fence_before();
            mbar_arrive(&p_full[t * 2 + fa_buf(j)]);
            continue;
#endif
            float s[FA_BN];
#pragma unroll
            for (int c = 0; c < FA_BN / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
            tmem_ld_wait();
            const int k0 = j * FA_BN;
            const bool need_mask = (k0 + FA_BN > L) || (CAUSAL && (k0 + FA_BN - 1 > q0 + t * FA_BM + wq * 32));
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < FA_BN; ++i) {
                    const int kj = k0 + i;
                    if (kj >= L || (CAUSAL && kj > qi)) s[i] = -INFINITY;
                }
            }
            // row max: four independent chains of 3-input max (FMNMX3)
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < FA_BN; i += 8) {
#pragma unroll
                for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(mx4[c], fmaxf(s[i + 2 * c], s[i + 2 * c + 1]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // lazy rescale: only move the reference max when it would grow by more than 2^8 in the exp2 domain
            const bool need = (mx > m_run) && ((mx - m_run) * p.scale_log2 > 8.0f);
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_run;
                const float alpha = (m_run == -INFINITY) ? 0.f : ex2f((m_run - m_new) * p.scale_log2);
                if (j > 0) {
                    // O_t must hold every product issued so far: S_t(j) ready only implies P(j-2) V retired, so wait for
                    // P(j-1) V explicitly.  Its barrier is in one of two states (that phase pending / complete): skipping
                    // this wait on other iterations cannot alias the parity.
                    // (with a single S buffer S_t(j) was issued behind P(j-1) V, so it has retired already)
                    if (FA_SBUF == 2) mbar_wait(&pv_done[t * 2 + fa_buf(j - 1)], fa_par(j - 1));
                    tc_fence_after();
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
            // p = exp2(s * scale - m * scale): packed fp32x2 FMA / add (FFMA2, FADD2); four independent packed row-sum accumulators
            const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), mn2 = pack_f32x2(mneg, mneg);
            uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int c = 0; c < FA_BN / 64; ++c) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint64_t x2 = ffma_f32x2(pack_f32x2(s[c * 64 + 2 * i], s[c * 64 + 2 * i + 1]), sc2, mn2);
                    float p0, p1;
                    if ((FA_EMU_MASK >> i) & 1u) {
                        ex2_poly_f32x2(x2, p0, p1);
                    } else {
                        float x0, x1;
                        unpack_f32x2(x2, x0, x1);
                        p0 = ex2f(x0); p1 = ex2f(x1);
                    }
                    acc2[i & 3] = fadd_f32x2(acc2[i & 3], pack_f32x2(p0, p1));
                    pk[i] = pack_bf16x2(p0, p1);
                }
                tmem_st_32x32b_x32(tS + c * 32, pk);
            }
            {
                float a0, a1, b0, b1;
                unpack_f32x2(fadd_f32x2(acc2[0], acc2[1]), a0, a1);
                unpack_f32x2(fadd_f32x2(acc2[2], acc2[3]), b0, b1);
                l_run += (a0 + a1) + (b0 + b1);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[t * 2 + fa_buf(j)]);
        }
        // ---- epilogue: O_t / l -> bf16 -> global ------------------------------------------
        mbar_wait(&pv_done[t * 2 + fa_buf(n_kv - 1)], fa_par(n_kv - 1));
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
    if (make_tmap_2d_bf16(&tk, k, total_tokens, (uint64_t)n_kv_heads * FA_D, k_stride, FA_BN, 64)) return -4;
    if (make_tmap_2d_bf16(&tv, v, total_tokens, (uint64_t)n_kv_heads * FA_D, v_stride, FA_BN, 64)) return -4;
    FaParams p;
    p.cu = cu_seqlens; p.o = (bf16*)out; p.os = o_stride;
    p.n_q_heads = n_q_heads; p.group = n_q_heads / n_kv_heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    dim3 grid((max_seqlen + 2 * FA_BM - 1) / (2 * FA_BM), n_q_heads, n_seqs);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
    }
    if (causal) attn_fwd_tcgen05_kernel<true><<<grid, FA_THREADS, FA_SMEM, st>>>(tq, tk, tv, p);
    else attn_fwd_tcgen05_kernel<false><<<grid, FA_THREADS, FA_SMEM, st>>>(tq, tk, tv, p);
    DOTS_LAUNCH_CHECK();
    return 0;
}
