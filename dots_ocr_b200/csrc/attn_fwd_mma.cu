// Variable-length fused attention forward (flash-attention style), head_dim = 128.
//
//   * ViT: bidirectional, one segment per image (cu_seqlens)      -- SURVEY.md §8a row a10
//   * LLM prefill: causal, grouped-query (q head h reads kv head h / group) -- row a19
//
// v1 of this kernel runs on the legacy warp-level tensor path (mma.sync m16n8k16, ldmatrix,
// cp.async double buffering).  Layout: token-major, q/k/v may live inside a fused qkv buffer
// (token stride passed in elements).  Softmax in fp32 with exp2 and a pre-scaled log2(e),
// P rounded to bf16 before P*V (the flash_attention_2 rounding points).
#include "common.h"
#include "mma_sm80.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int ATT_D = 128;
constexpr int ATT_BM = 64;
constexpr int ATT_BN = 64;
constexpr int ATT_THREADS = 128;
constexpr int ATT_SMEM = (ATT_BM + 4 * ATT_BN) * ATT_D * 2;   // Q + 2xK + 2xV = 80 KB

struct AttnParams {
    const bf16* q; const bf16* k; const bf16* v; bf16* o;
    long long qs, ks, vs, os;          // token strides (elements)
    const int* cu;                     // [n_seqs + 1]
    int n_q_heads, group;              // group = q heads per kv head
    float scale_log2;
};

__device__ __forceinline__ void load_tile_async(uint8_t* smem_tile, const bf16* gbase, long long stride, int row0,
                                                int rows_valid, int tid) {
    // 64 rows x 16 chunks of 16 B; 128 threads -> 8 chunks each
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int idx = tid + j * ATT_THREADS;
        const int r = idx >> 4, c = idx & 15;
        const bool ok = (row0 + r) < rows_valid;
        const bf16* src = gbase + (long long)(ok ? (row0 + r) : 0) * stride + c * 8;
        cp_async_16(smem_tile + swz128(r, c), src, ok);
    }
}

template <bool CAUSAL>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_mma_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + ATT_BM * ATT_D * 2;
    uint8_t* sV = sK + 2 * ATT_BN * ATT_D * 2;

    const int seq = blockIdx.z;
    const int head = blockIdx.y;
    const int tok0 = p.cu[seq];
    const int L = p.cu[seq + 1] - tok0;
    const int q0 = blockIdx.x * ATT_BM;
    if (q0 >= L) return;
    const int kvh = head / p.group;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    const bf16* qg = p.q + (long long)tok0 * p.qs + head * ATT_D;
    const bf16* kg = p.k + (long long)tok0 * p.ks + kvh * ATT_D;
    const bf16* vg = p.v + (long long)tok0 * p.vs + kvh * ATT_D;

    int n_tiles = (L + ATT_BN - 1) / ATT_BN;
    if (CAUSAL) n_tiles = min(n_tiles, (q0 + ATT_BM + ATT_BN - 1) / ATT_BN);

    load_tile_async(sQ, qg, p.qs, q0, L, tid);
    load_tile_async(sK, kg, p.ks, 0, L, tid);
    cp_async_commit();
    load_tile_async(sV, vg, p.vs, 0, L, tid);
    cp_async_commit();

    cp_async_wait<1>();          // Q and K0 landed
    __syncthreads();

    // Q fragments for this warp's 16 rows, all 8 k-steps
    uint32_t qf[8][4];
    {
        const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) ldmatrix_x4(qf[kk], smem_u32(sQ) + swz128(r, kk * 2 + (lane >> 4)));
    }

    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int qrow0 = q0 + warp * 16 + g;          // rows qrow0 and qrow0 + 8

    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j & 1;
        const uint32_t kb = smem_u32(sK) + buf * (ATT_BN * ATT_D * 2);
        const uint32_t vb = smem_u32(sV) + buf * (ATT_BN * ATT_D * 2);
        // K_j is resident here (waited at the end of the previous iteration / prologue)

        // ---- S = Q K^T --------------------------------------------------------------
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
            for (int pr = 0; pr < 4; ++pr) {
                uint32_t bfr[4];
                const int r = pr * 16 + (lane & 7) + 8 * (lane >> 4);
                ldmatrix_x4(bfr, kb + swz128(r, kk * 2 + ((lane >> 3) & 1)));
                mma_bf16_16816(s[2 * pr], qf[kk], bfr[0], bfr[1]);
                mma_bf16_16816(s[2 * pr + 1], qf[kk], bfr[2], bfr[3]);
            }
        }
        // prefetch K_{j+1} into the other buffer (its previous reader, iteration j-1, is done:
        // every warp passed the __syncthreads at the end of iteration j-1)
        if (j + 1 < n_tiles) load_tile_async(sK + (buf ^ 1) * (ATT_BN * ATT_D * 2), kg, p.ks, (j + 1) * ATT_BN, L, tid);
        cp_async_commit();

        // ---- mask + online softmax ---------------------------------------------------
        const int kbase = j * ATT_BN + 2 * t;
        const bool need_mask = (j * ATT_BN + ATT_BN > L) || (CAUSAL && (j * ATT_BN + ATT_BN - 1 > q0 + warp * 16));
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float val = s[nb][e] * p.scale_log2;
                if (need_mask) {
                    const int kj = kbase + nb * 8 + (e & 1);
                    const int qi = qrow0 + ((e >> 1) << 3);
                    if (kj >= L || (CAUSAL && kj > qi)) val = -INFINITY;
                }
                s[nb][e] = val;
                mx[e >> 1] = fmaxf(mx[e >> 1], val);
            }
        }
        float alpha[2], msafe[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);
            msafe[h] = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[h] = fast_exp2(m_run[h] - msafe[h]);          // exp2(-inf) = 0 on the first tile
            m_run[h] = m_new;
            l_run[h] *= alpha[h];
        }
        uint32_t pf[4][4];                 // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = fast_exp2(s[nb][0] - msafe[0]);
            const float p1 = fast_exp2(s[nb][1] - msafe[0]);
            const float p2 = fast_exp2(s[nb][2] - msafe[1]);
            const float p3 = fast_exp2(s[nb][3] - msafe[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
            o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
        }

        // ---- O += P V ------------------------------------------------------------------
        cp_async_wait<1>();             // V_j landed (only K_{j+1} may still be in flight)
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 8; ++dp) {
                uint32_t bfr[4];
                const int r = kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
                ldmatrix_x4_trans(bfr, vb + swz128(r, dp * 2 + (lane >> 4)));
                mma_bf16_16816(o[2 * dp], pf[kk], bfr[0], bfr[1]);
                mma_bf16_16816(o[2 * dp + 1], pf[kk], bfr[2], bfr[3]);
            }
        }
        __syncthreads();                // everyone done with K_j (S phase) and V_j
        if (j + 1 < n_tiles) load_tile_async(sV + (buf ^ 1) * (ATT_BN * ATT_D * 2), vg, p.vs, (j + 1) * ATT_BN, L, tid);
        cp_async_commit();
        cp_async_wait<1>();             // K_{j+1} landed
        __syncthreads();
    }

    // ---- finalize -------------------------------------------------------------------------
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    bf16* og = p.o + (long long)tok0 * p.os + head * ATT_D;
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
        const int col = nb * 8 + 2 * t;
        if (qrow0 < L)
            *reinterpret_cast<uint32_t*>(og + (long long)qrow0 * p.os + col) = pack_bf16x2(o[nb][0] * inv0, o[nb][1] * inv0);
        if (qrow0 + 8 < L)
            *reinterpret_cast<uint32_t*>(og + (long long)(qrow0 + 8) * p.os + col) = pack_bf16x2(o[nb][2] * inv1, o[nb][3] * inv1);
    }
}

}  // namespace dots

using namespace dots;

extern "C" int dots_attn_varlen_fwd(const void* q, long long q_stride, const void* k, long long k_stride, const void* v,
                                    long long v_stride, void* out, long long o_stride, const int* cu_seqlens,
                                    int n_seqs, int max_seqlen, int n_q_heads, int n_kv_heads, int head_dim,
                                    int causal, float softmax_scale, void* stream) {
    DOTS_REQUIRE(head_dim == ATT_D, "dots_attn_varlen_fwd: head_dim must be 128 (got %d)", head_dim);
    DOTS_REQUIRE(n_seqs > 0 && max_seqlen > 0 && n_q_heads > 0 && n_kv_heads > 0 && n_q_heads % n_kv_heads == 0,
                 "dots_attn_varlen_fwd: bad shape n_seqs=%d max_seqlen=%d heads=%d/%d", n_seqs, max_seqlen, n_q_heads, n_kv_heads);
    DOTS_REQUIRE(q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && o_stride % 2 == 0,
                 "dots_attn_varlen_fwd: token strides must keep 16-byte alignment");
    AttnParams p;
    p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (bf16*)out;
    p.qs = q_stride; p.ks = k_stride; p.vs = v_stride; p.os = o_stride;
    p.cu = cu_seqlens;
    p.n_q_heads = n_q_heads;
    p.group = n_q_heads / n_kv_heads;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    dim3 grid((max_seqlen + ATT_BM - 1) / ATT_BM, n_q_heads, n_seqs);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    }
    if (causal) attn_fwd_mma_kernel<true><<<grid, ATT_THREADS, ATT_SMEM, st>>>(p);
    else attn_fwd_mma_kernel<false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(p);
    DOTS_LAUNCH_CHECK();
    return 0;
}
