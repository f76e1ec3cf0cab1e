// Variable-length fused attention forward on CTA PAIRS (tcgen05 cta_group::2), head_dim 128.
//
// Same algorithm and rounding points as attn_fwd_tcgen05.cu (two 128-row query tiles = "chains" per CTA, 128-key tiles, one S
// buffer per chain, softmax warpgroups with packed fp32x2 arithmetic and lazy rescale), but a cluster of two CTAs shares
// every K/V tile: each CTA stages only HALF of it (K: 64 of the 128 keys; V: 64 of the 128 head dims) and the pair's tensor
// cores read both halves (M = 256 MMAs: rows 0-127 = this chain in CTA 0, rows 128-255 = the same chain in CTA 1).
// Per CTA that halves the TMA writes into shared memory and the B-operand reads out of it -- the single-CTA kernel needs
// ~128 B/clk of shared-memory bandwidth at full tensor rate (the whole SM budget), this one ~80 B/clk.
//
//   both CTAs   warp 0        TMA producer: own Q tiles, own half of K_j / V_j; bytes are credited to the LEADER's barriers
//   leader      warps 1, 2    tcgen05.mma.cta_group::2 issuers, one per chain:  S_t = Q_t K_j^T (SS),  O_t += P_t V_j (TS)
//   both CTAs   warps 4-7 / 8-11   softmax warpgroups of chain 0 / 1 for this CTA's 128 rows
//
// Barriers: q/k/v `full` live in the leader (TMA of both CTAs completes on them); k/v `empty`, `s_full`, `pv_done` exist in
// both CTAs and are signalled by multicast commits; `p_full` lives in the leader and collects one arrival per softmax warp
// of both CTAs.  SURVEY.md §8a rows a10, a19.
//
// STATUS: experiment.  Passes the parity tests, but measures 650 vs 1080 TFLOP/s for the single-CTA kernel: attention is bound
// by the per-chain hand-off latency, which cluster-scope barriers lengthen, not by shared-memory bandwidth.
#include "../common.h"
#include "../ptx.cuh"
#include "experiments.h"

namespace dots {

constexpr int F2_D = 128;
constexpr int F2_BM = 128;                      // rows per chain per CTA
constexpr int F2_BN = 128;                      // keys per tile
constexpr int F2_THREADS = 384;
constexpr int F2_QTILE_BYTES = F2_BM * F2_D * 2;            // 32 KB (two [128 x 64] boxes)
constexpr int F2_KHALF_BYTES = (F2_BN / 2) * F2_D * 2;      // 16 KB: 64 keys x 128 dims (two [64 x 64] boxes)
constexpr int F2_VHALF_BYTES = F2_BN * (F2_D / 2) * 2;      // 16 KB: 128 keys x 64 dims (one [128 x 64] box)
constexpr int F2_KSTAGES = 4;
constexpr int F2_VSTAGES = 3;
constexpr int F2_SMEM = 2 * F2_QTILE_BYTES + F2_KSTAGES * F2_KHALF_BYTES + F2_VSTAGES * F2_VHALF_BYTES + 1024 + 1024;

struct Fa2Params {
    const int* cu;
    bf16* o;
    long long os;
    int n_q_heads, group;
    float scale_log2;
};

__device__ __forceinline__ void f2_tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void f2_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem, both CTAs] (+)= A[tmem, each CTA's rows] * B[smem, N/2 from each CTA]
__device__ __forceinline__ void umma_bf16_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float f2_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(F2_THREADS, 1)
attn_fwd_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                             const __grid_constant__ CUtensorMap tm_v, const Fa2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                          // [2 chains][2 boxes][128 rows][128 B]
    uint8_t* sK = sQ + 2 * F2_QTILE_BYTES;                       // [KSTAGES][2 boxes][64 keys][128 B]   this CTA's 64 keys
    uint8_t* sV = sK + F2_KSTAGES * F2_KHALF_BYTES;              // [VSTAGES][128 keys][128 B]           this CTA's 64 dims
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + F2_VSTAGES * F2_VHALF_BYTES);
    uint64_t* q_full = bars;                     // [1]         leader
    uint64_t* k_full = bars + 1;                 // [KSTAGES]   leader
    uint64_t* k_empty = k_full + F2_KSTAGES;     // [KSTAGES]   both
    uint64_t* v_full = k_empty + F2_KSTAGES;     // [VSTAGES]   leader
    uint64_t* v_empty = v_full + F2_VSTAGES;     // [VSTAGES]   both
    uint64_t* s_full = v_empty + F2_VSTAGES;     // [2 chains]  both
    uint64_t* p_full = s_full + 2;               // [2 chains]  leader (8 warp arrivals: 4 softmax warps x 2 CTAs)
    uint64_t* pv_done = p_full + 2;              // [2 chains]  both
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int seq = blockIdx.z, head = blockIdx.y;
    const int tok0 = p.cu[seq];
    const int L = p.cu[seq + 1] - tok0;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int q0_pair = (blockIdx.x >> 1) * (4 * F2_BM);          // 512 query rows per CTA pair
    if (q0_pair >= L) return;                                     // both CTAs of the pair leave together
    const int q0 = q0_pair + (int)rank * (2 * F2_BM);             // this CTA's 256 rows (possibly all past L: still needed as half of the pair)
    const int kvh = head / p.group;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int n_kv = (L + F2_BN - 1) / F2_BN;
    if (CAUSAL) n_kv = min(n_kv, (q0_pair + 4 * F2_BM + F2_BN - 1) / F2_BN);

    if (warp == 0 && lane == 0) {
        prefetch_tensormap(&tm_q); prefetch_tensormap(&tm_k); prefetch_tensormap(&tm_v);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < F2_KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 2); }    // two issuing warps
        for (int i = 0; i < F2_VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 8); mbar_init(&pv_done[i], 1); }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_2sm(tmem_ptr, 512);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    if (*tmem_ptr != 0u) __trap();               // the pair owns all 512 columns: the allocation starts at column 0
    constexpr uint32_t tmem_base = 0u;

    if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
        // =============================== TMA producer (both CTAs) ===============================
        if (lane == 0) {
            const uint32_t qf = mapa_shared(smem_u32(q_full), 0);
            if (leader) mbar_expect_tx(q_full, 2 * 2 * F2_QTILE_BYTES);              // both CTAs' Q tiles
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d_2sm(sQ + t * F2_QTILE_BYTES + h * (F2_QTILE_BYTES / 2), &tm_q, head * F2_D + h * 64, tok0 + q0 + t * F2_BM, qf);
            for (int j = 0; j < n_kv; ++j) {
                const int ks = j % F2_KSTAGES, vs = j % F2_VSTAGES;
                mbar_wait(&k_empty[ks], ((j / F2_KSTAGES) & 1) ^ 1);
                const uint32_t kf = mapa_shared(smem_u32(&k_full[ks]), 0);
                if (leader) mbar_expect_tx(&k_full[ks], 2 * F2_KHALF_BYTES);
#pragma unroll
                for (int h = 0; h < 2; ++h)       // this CTA's 64 keys of the tile, dims [64h, 64h + 64)
                    tma_load_2d_2sm(sK + ks * F2_KHALF_BYTES + h * (F2_KHALF_BYTES / 2), &tm_k, kvh * F2_D + h * 64,
                                    tok0 + j * F2_BN + (int)rank * (F2_BN / 2), kf);
                mbar_wait(&v_empty[vs], ((j / F2_VSTAGES) & 1) ^ 1);
                const uint32_t vf = mapa_shared(smem_u32(&v_full[vs]), 0);
                if (leader) mbar_expect_tx(&v_full[vs], 2 * F2_VHALF_BYTES);
                // all 128 keys of the tile, this CTA's 64 dims
                tma_load_2d_2sm(sV + vs * F2_VHALF_BYTES, &tm_v, kvh * F2_D + (int)rank * 64, tok0 + j * F2_BN, vf);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // =============================== MMA issuers (leader CTA): warp 1 drives chain 0, warp 2 chain 1 ===============================
        if (leader) {
            const int n_kv_u = __shfl_sync(0xffffffffu, n_kv, 0);
            const int t = __shfl_sync(0xffffffffu, warp - 1, 0);
            const bool elected = elect_one();
            constexpr uint32_t idesc_s = umma_idesc_bf16(256, F2_BN, 0, 0);           // S = Q K^T : A, B K-major, M = 256 over the pair
            constexpr uint32_t idesc_o = umma_idesc_bf16(256, 128, 0, 1);             // O += P V : A from TMEM, B (V) MN-major
            const uint64_t dq0 = umma_desc_k_sw128(smem_u32(sQ));
            const uint64_t dk0 = umma_desc_k_sw128(smem_u32(sK));
            const uint64_t dv0 = umma_desc_mn_sw128(smem_u32(sV), F2_VHALF_BYTES, 1024);   // one 64-wide MN chunk per CTA (LBO unused)
            const uint32_t tS = tmem_base + t * 256;
            const uint32_t tO = tmem_base + t * 256 + 128;
            auto issue_s = [&](int j) {
                const uint64_t dq = dq0 + (uint64_t)(t * (F2_QTILE_BYTES >> 4));
                const uint64_t dk = dk0 + (uint64_t)((j % F2_KSTAGES) * (F2_KHALF_BYTES >> 4));
                if (elected) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {      // 16 head dims per step: box k>>2, 32 B inside the swizzle row
                        const uint64_t qoff = (uint64_t)(((k >> 2) * (F2_QTILE_BYTES / 2) + (k & 3) * 32) >> 4);
                        const uint64_t koff = (uint64_t)(((k >> 2) * (F2_KHALF_BYTES / 2) + (k & 3) * 32) >> 4);
                        umma_bf16_ss_2sm(tS, dq + qoff, dk + koff, idesc_s, k > 0 ? 1u : 0u);
                    }
                    umma_commit_2sm(&s_full[t], 3);
                }
                __syncwarp();
            };
            auto issue_pv = [&](int j) {
                const uint64_t dv = dv0 + (uint64_t)((j % F2_VSTAGES) * (F2_VHALF_BYTES >> 4));
                const uint32_t acc0 = (j == 0) ? 0u : 1u;
                if (elected) {
#pragma unroll
                    for (int k = 0; k < F2_BN / 16; ++k)      // 16 keys per step: 16 rows x 128 B = 2048 B into the box
                        umma_bf16_ts_2sm(tO, tS + k * 8, dv + (uint64_t)((k * 2048) >> 4), idesc_o, k == 0 ? acc0 : 1u);
                    umma_commit_2sm(&pv_done[t], 3);
                }
                __syncwarp();
            };
            auto commit2 = [&](uint64_t* bar) {
                if (elected) umma_commit_2sm(bar, 3);
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0);
            commit2(&k_empty[0]);
            for (int j = 0; j < n_kv_u; ++j) {
                const int jn = j + 1;
                const bool more = (jn < n_kv_u);
                mbar_wait(&v_full[j % F2_VSTAGES], (j / F2_VSTAGES) & 1);
                if (more) mbar_wait(&k_full[jn % F2_KSTAGES], (jn / F2_KSTAGES) & 1);
                mbar_wait(&p_full[t], j & 1);
                tc_fence_after();
                issue_pv(j);
                if (more) issue_s(jn);                        // overwrites S/P of tile j: ordered behind P(j) V_j on the tensor pipe
                commit2(&v_empty[j % F2_VSTAGES]);
                if (more) commit2(&k_empty[jn % F2_KSTAGES]);
            }
        }
    } else if (warp >= 4) {
        // =============================== softmax warpgroups (both CTAs) ===============================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int t = (warp - 4) >> 2;
        const int wq = warp & 3;
        const int row = wq * 32 + lane;
        const int qi = q0 + t * F2_BM + row;
        const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
        const uint32_t tS = tmem_base + lane_addr + t * 256;
        const uint32_t tO = tS + 128;
        const uint32_t p_full_leader = mapa_shared(smem_u32(&p_full[t]), 0);
        float m_run = -INFINITY;
        float l_run = 0.f;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[t], j & 1);
            tc_fence_after();
            float s[F2_BN];
#pragma unroll
            for (int c = 0; c < F2_BN / 32; ++c) tmem_ld_32x32b_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
            tmem_ld_wait();
            const int k0 = j * F2_BN;
            const bool need_mask = (k0 + F2_BN > L) || (CAUSAL && (k0 + F2_BN - 1 > q0 + t * F2_BM + wq * 32));
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < F2_BN; ++i) {
                    const int kj = k0 + i;
                    if (kj >= L || (CAUSAL && kj > qi)) s[i] = -INFINITY;
                }
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < F2_BN; i += 8) {
#pragma unroll
                for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(mx4[c], fmaxf(s[i + 2 * c], s[i + 2 * c + 1]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // lazy rescale (S(j) was issued behind P(j-1) V, so O already holds every product issued so far)
            const bool need = (mx > m_run) && ((mx - m_run) * p.scale_log2 > 8.0f);
            if (__any_sync(0xffffffffu, need)) {
                const float m_new = need ? mx : m_run;
                const float alpha = (m_run == -INFINITY) ? 0.f : f2_ex2((m_run - m_new) * p.scale_log2);
                if (j > 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tO + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        f2_tmem_st_32x32b_x32(tO + c * 32, v);
                    }
                }
                l_run *= alpha;
                m_run = m_new;
            }
            const float mneg = (m_run == -INFINITY) ? 0.f : -m_run * p.scale_log2;
            const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2), mn2 = pack_f32x2(mneg, mneg);
            uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int c = 0; c < F2_BN / 64; ++c) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint64_t x2 = ffma_f32x2(pack_f32x2(s[c * 64 + 2 * i], s[c * 64 + 2 * i + 1]), sc2, mn2);
                    float x0, x1;
                    unpack_f32x2(x2, x0, x1);
                    const float p0 = f2_ex2(x0), p1 = f2_ex2(x1);
                    acc2[i & 3] = fadd_f32x2(acc2[i & 3], pack_f32x2(p0, p1));
                    pk[i] = pack_bf16x2(p0, p1);
                }
                f2_tmem_st_32x32b_x32(tS + c * 32, pk);
            }
            {
                float a0, a1, b0, b1;
                unpack_f32x2(fadd_f32x2(acc2[0], acc2[1]), a0, a1);
                unpack_f32x2(fadd_f32x2(acc2[2], acc2[3]), b0, b1);
                l_run += (a0 + a1) + (b0 + b1);
            }
            f2_tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(p_full_leader);        // one arrival per softmax warp of either CTA
        }
        // ---- epilogue: O_t / l -> bf16 -> global ------------------------------------------
        mbar_wait(&pv_done[t], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        bf16* dst = p.o + (long long)(tok0 + qi) * p.os + head * F2_D;
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
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 512);
    }
}

}  // namespace dots

using namespace dots;

extern "C" int dots_attn_varlen_fwd_pair(const void* q, long long q_stride, const void* k, long long k_stride, const void* v,
                                         long long v_stride, void* out, long long o_stride, const int* cu_seqlens, int n_seqs,
                                         int max_seqlen, long long total_tokens, int n_q_heads, int n_kv_heads, int head_dim,
                                         int causal, float softmax_scale, void* stream) {
    DOTS_REQUIRE(head_dim == F2_D, "dots_attn_varlen_fwd_pair: head_dim must be 128 (got %d)", head_dim);
    DOTS_REQUIRE(n_seqs > 0 && max_seqlen > 0 && total_tokens > 0 && n_q_heads > 0 && n_kv_heads > 0 && n_q_heads % n_kv_heads == 0,
                 "dots_attn_varlen_fwd_pair: bad shape");
    DOTS_REQUIRE(q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && o_stride % 8 == 0,
                 "dots_attn_varlen_fwd_pair: token strides must be multiples of 8 elements");
    DOTS_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) % 16 == 0, "dots_attn_varlen_fwd_pair: 16-byte aligned pointers");
    CUtensorMap tq, tk, tv;
    if (make_tmap_2d_bf16(&tq, q, total_tokens, (uint64_t)n_q_heads * F2_D, q_stride, 128, 64)) return -4;
    if (make_tmap_2d_bf16(&tk, k, total_tokens, (uint64_t)n_kv_heads * F2_D, k_stride, F2_BN / 2, 64)) return -4;
    if (make_tmap_2d_bf16(&tv, v, total_tokens, (uint64_t)n_kv_heads * F2_D, v_stride, F2_BN, 64)) return -4;
    Fa2Params p;
    p.cu = cu_seqlens; p.o = (bf16*)out; p.os = o_stride;
    p.n_q_heads = n_q_heads; p.group = n_q_heads / n_kv_heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    dim3 grid(2 * ((max_seqlen + 4 * F2_BM - 1) / (4 * F2_BM)), n_q_heads, n_seqs);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tcgen05_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    }
    if (causal) DOTS_CHECK_CUDA(launch_ex_cluster(attn_fwd_tcgen05_pair_kernel<true>, grid, dim3(F2_THREADS), (size_t)F2_SMEM, st, false, 2u, tq, tk, tv, p));
    else DOTS_CHECK_CUDA(launch_ex_cluster(attn_fwd_tcgen05_pair_kernel<false>, grid, dim3(F2_THREADS), (size_t)F2_SMEM, st, false, 2u, tq, tk, tv, p));
    return 0;
}
