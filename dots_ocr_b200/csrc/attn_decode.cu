// Decode-step attention over the in-place KV cache (decode half).
//
// One query token per sequence, grouped-query: the G q-heads that share a kv-head are packed into
// the 16-row M dimension of mma.sync so K and V are read from HBM exactly once per step.
// Keys are split (a) across CTAs (grid.x = n_splits, "flash decoding") and (b) across the 4 warps
// of a CTA, each warp streaming its own 16-key tiles through a private cp.async double buffer.
// Partial (m, l, O) triples are merged in shared memory, then across splits by a small combine
// kernel.  This kernel is HBM-bound: bytes = 2 * ctx * 256 B per (sequence, kv head).
#include "common.h"
#include "ptx.cuh"
#include "mma_sm80.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int DEC_D = 128;
constexpr int DEC_TILE = 16;                 // keys per warp tile
constexpr int DEC_WARPS = 4;
constexpr int DEC_THREADS = DEC_WARPS * 32;
constexpr int DEC_TILE_BYTES = DEC_TILE * DEC_D * 2;                       // 4 KB
constexpr int DEC_SMEM = 4096 /*Q*/ + DEC_WARPS * 4 * DEC_TILE_BYTES;     // Q + per-warp 2x(K,V) = 68 KB

struct DecParams {
    const bf16* q;            // [B, n_q_heads * 128]
    const bf16* kc;           // [B, n_kv_heads, ctx_max, 128]
    const bf16* vc;
    const int* ctx_len;       // [B] keys visible to the current token (its own key included)
    bf16* out;                // [B, n_q_heads * 128]
    float* part_o;            // [B, n_q_heads, n_splits, 128]
    float* part_ml;           // [B, n_q_heads, n_splits, 2]
    long long ctx_max;
    int n_q_heads, n_kv_heads, group, n_splits;
    float scale_log2;
    // fused QKV finalize (dots_attn_decode_fused): q/k/v of the current token arrive as split-K fp32 partials
    // [qkv_splits][B][(nq + 2 nkv) * 128] of the QKV GEMM; this kernel adds the bias, applies RoPE, appends k, v to the
    // cache and keeps q in shared memory.  qkv_partial == nullptr: q is read from p.q (plain dots_attn_decode).
    const float* qkv_partial;
    int qkv_splits;
    const bf16* qkv_bias;
    const int* pos;
    const float* inv_freq;
    bf16* kc_w;
    bf16* vc_w;
};

// HF Qwen2 RoPE rounding points (modeling_qwen2.py:102-146): cos/sin are bf16, every product and the sum round to bf16.
__device__ __forceinline__ void dec_rope_bf16_4(const float (&x1)[4], const float (&x2)[4], int pos, const float* __restrict__ inv_freq,
                                                int i0, float (&o1)[4], float (&o2)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float ang = __fmul_rn((float)pos, inv_freq[i0 + j]);
        const float c = bf16_round(cosf(ang)), sn = bf16_round(sinf(ang));
        o1[j] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x1[j], c)), bf16_round(__fmul_rn(-x2[j], sn))));
        o2[j] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x2[j], c)), bf16_round(__fmul_rn(x1[j], sn))));
    }
}

__device__ __forceinline__ void dec_load_tile(uint8_t* dst, const bf16* gsrc_rows, int key0, int key_end, int lane) {
    // 16 rows x 16 chunks; 32 lanes -> 8 chunks each
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int idx = lane + j * 32;
        const int r = idx >> 4, c = idx & 15;
        const bool ok = (key0 + r) < key_end;
        const bf16* src = gsrc_rows + (long long)(ok ? (key0 + r) : 0) * DEC_D + c * 8;
        cp_async_16(dst + swz128(r, c), src, ok);
    }
}

__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const DecParams p) {
    pdl_wait();
    pdl_launch_dependents();
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sQ = smem;
    const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    uint8_t* sK = smem + 4096 + warp * 4 * DEC_TILE_BYTES;      // [2][16][128]
    uint8_t* sV = sK + 2 * DEC_TILE_BYTES;

    const int ctx = p.ctx_len[b];
    int chunk = (ctx + p.n_splits - 1) / p.n_splits;
    chunk = (chunk + DEC_TILE * DEC_WARPS - 1) / (DEC_TILE * DEC_WARPS) * (DEC_TILE * DEC_WARPS);
    const int k_begin = split * chunk;
    const int k_end = min(ctx, k_begin + chunk);

    const bf16* kbase = p.kc + ((long long)b * p.n_kv_heads + kvh) * p.ctx_max * DEC_D;
    const bf16* vbase = p.vc + ((long long)b * p.n_kv_heads + kvh) * p.ctx_max * DEC_D;

    // Q tile: rows 0..G-1 = the group's q heads, rows G..15 zero
    if (p.qkv_partial == nullptr) {
        const bf16* qg = p.q + (long long)b * p.n_q_heads * DEC_D + (long long)kvh * p.group * DEC_D;
        for (int idx = tid; idx < 16 * 16; idx += DEC_THREADS) {
            const int r = idx >> 4, c = idx & 15;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (r < p.group) val = *reinterpret_cast<const uint4*>(qg + r * DEC_D + c * 8);
            *reinterpret_cast<uint4*>(sQ + swz128(r, c)) = val;
        }
    } else {
        // ---- fused QKV finalize: split-K reduce (fixed order) + bias + RoPE; q -> sQ, k/v -> cache row `pos` ----
        for (int idx = tid; idx < (16 - p.group) * 16; idx += DEC_THREADS) {
            const int r = p.group + (idx >> 4), c = idx & 15;
            *reinterpret_cast<uint4*>(sQ + swz128(r, c)) = make_uint4(0, 0, 0, 0);
        }
        const int posb = p.pos[b];
        const bool owns_new = (posb >= k_begin) && (posb < k_end);        // the split whose key range holds the new token
        const int N = (p.n_q_heads + 2 * p.n_kv_heads) * DEC_D;
        const long long sstride = (long long)(gridDim.z) * N;
        const int n_units = (p.group + 2) * 16;                            // (head, 4-column pair chunk)
        for (int u = tid; u < n_units; u += DEC_THREADS) {
            const int hl = u >> 4, c4 = u & 15;
            if (hl >= p.group && !owns_new) continue;
            const int col0 = (hl < p.group ? (kvh * p.group + hl)
                                           : (hl == p.group ? p.n_q_heads + kvh : p.n_q_heads + p.n_kv_heads + kvh)) * DEC_D + c4 * 4;
            const float* src = p.qkv_partial + (long long)b * N + col0;
            float x1[4] = {0.f, 0.f, 0.f, 0.f}, x2[4] = {0.f, 0.f, 0.f, 0.f};
            int sidx = 0;
            for (; sidx + 4 <= p.qkv_splits; sidx += 4) {
                float4 a[4], d[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    a[w] = *reinterpret_cast<const float4*>(src + (sidx + w) * sstride);
                    d[w] = *reinterpret_cast<const float4*>(src + (sidx + w) * sstride + 64);
                }
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    x1[0] += a[w].x; x1[1] += a[w].y; x1[2] += a[w].z; x1[3] += a[w].w;
                    x2[0] += d[w].x; x2[1] += d[w].y; x2[2] += d[w].z; x2[3] += d[w].w;
                }
            }
            for (; sidx < p.qkv_splits; ++sidx) {
                const float4 a = *reinterpret_cast<const float4*>(src + sidx * sstride);
                const float4 d = *reinterpret_cast<const float4*>(src + sidx * sstride + 64);
                x1[0] += a.x; x1[1] += a.y; x1[2] += a.z; x1[3] += a.w;
                x2[0] += d.x; x2[1] += d.y; x2[2] += d.z; x2[3] += d.w;
            }
            const uint2 b1 = *reinterpret_cast<const uint2*>(p.qkv_bias + col0);
            const uint2 b2 = *reinterpret_cast<const uint2*>(p.qkv_bias + col0 + 64);
            x1[0] = bf16_round(x1[0] + bf16_lo(b1.x)); x1[1] = bf16_round(x1[1] + bf16_hi(b1.x));
            x1[2] = bf16_round(x1[2] + bf16_lo(b1.y)); x1[3] = bf16_round(x1[3] + bf16_hi(b1.y));
            x2[0] = bf16_round(x2[0] + bf16_lo(b2.x)); x2[1] = bf16_round(x2[1] + bf16_hi(b2.x));
            x2[2] = bf16_round(x2[2] + bf16_lo(b2.y)); x2[3] = bf16_round(x2[3] + bf16_hi(b2.y));
            float o1[4], o2[4];
            if (hl <= p.group) {
                dec_rope_bf16_4(x1, x2, posb, p.inv_freq, c4 * 4, o1, o2);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { o1[j] = x1[j]; o2[j] = x2[j]; }
            }
            const uint2 r1 = make_uint2(pack_bf16x2(o1[0], o1[1]), pack_bf16x2(o1[2], o1[3]));
            const uint2 r2 = make_uint2(pack_bf16x2(o2[0], o2[1]), pack_bf16x2(o2[2], o2[3]));
            if (hl < p.group) {
                *reinterpret_cast<uint2*>(sQ + swz128(hl, c4 >> 1) + (c4 & 1) * 8) = r1;
                *reinterpret_cast<uint2*>(sQ + swz128(hl, 8 + (c4 >> 1)) + (c4 & 1) * 8) = r2;
            } else {
                bf16* dst = (hl == p.group ? p.kc_w : p.vc_w) + (((long long)b * p.n_kv_heads + kvh) * p.ctx_max + posb) * DEC_D + c4 * 4;
                *reinterpret_cast<uint2*>(dst) = r1;
                *reinterpret_cast<uint2*>(dst + 64) = r2;
            }
        }
    }
    __syncthreads();
    uint32_t qf[8][4];
    {
        const int r = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) ldmatrix_x4(qf[kk], smem_u32(sQ) + swz128(r, kk * 2 + (lane >> 4)));
    }

    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    const int n_tiles_cta = (k_end > k_begin) ? (k_end - k_begin + DEC_TILE - 1) / DEC_TILE : 0;
    const int my_tiles = (n_tiles_cta > warp) ? (n_tiles_cta - warp + DEC_WARPS - 1) / DEC_WARPS : 0;

    if (my_tiles > 0) {
        dec_load_tile(sK, kbase, k_begin + warp * DEC_TILE, k_end, lane);
        dec_load_tile(sV, vbase, k_begin + warp * DEC_TILE, k_end, lane);
    }
    cp_async_commit();
    for (int i = 0; i < my_tiles; ++i) {
        const int buf = i & 1;
        const int key0 = k_begin + (warp + i * DEC_WARPS) * DEC_TILE;
        if (i + 1 < my_tiles) {
            const int nk = key0 + DEC_WARPS * DEC_TILE;
            dec_load_tile(sK + (buf ^ 1) * DEC_TILE_BYTES, kbase, nk, k_end, lane);
            dec_load_tile(sV + (buf ^ 1) * DEC_TILE_BYTES, vbase, nk, k_end, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const uint32_t kb = smem_u32(sK) + buf * DEC_TILE_BYTES;
        const uint32_t vb = smem_u32(sV) + buf * DEC_TILE_BYTES;

        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t bfr[4];
            const int r = (lane & 7) + 8 * (lane >> 4);
            ldmatrix_x4(bfr, kb + swz128(r, kk * 2 + ((lane >> 3) & 1)));
            mma_bf16_16816(s[0], qf[kk], bfr[0], bfr[1]);
            mma_bf16_16816(s[1], qf[kk], bfr[2], bfr[3]);
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float val = s[nb][e] * p.scale_log2;
                const int kj = key0 + nb * 8 + 2 * t + (e & 1);
                if (kj >= k_end) val = -INFINITY;
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
            alpha[h] = fast_exp2(m_run[h] - msafe[h]);
            m_run[h] = m_new;
            l_run[h] *= alpha[h];
        }
        uint32_t pf[4];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const float p0 = fast_exp2(s[nb][0] - msafe[0]);
            const float p1 = fast_exp2(s[nb][1] - msafe[0]);
            const float p2 = fast_exp2(s[nb][2] - msafe[1]);
            const float p3 = fast_exp2(s[nb][3] - msafe[1]);
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pf[nb * 2 + 0] = pack_bf16x2(p0, p1);
            pf[nb * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
            o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
        }
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
            uint32_t bfr[4];
            const int r = (lane & 7) + 8 * ((lane >> 3) & 1);
            ldmatrix_x4_trans(bfr, vb + swz128(r, dp * 2 + (lane >> 4)));
            mma_bf16_16816(o[2 * dp], pf, bfr[0], bfr[1]);
            mma_bf16_16816(o[2 * dp + 1], pf, bfr[2], bfr[3]);
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 1);
    l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 2);

    // ---- merge the 4 warps (rows g < group only; rows 8..15 are padding) -----------------
    __syncthreads();                                  // all warps done with their K/V buffers
    float* sO = reinterpret_cast<float*>(smem + 4096);                  // [4 warps][8 rows][128]
    float* sML = sO + DEC_WARPS * 8 * DEC_D;                            // [4 warps][8 rows][2]
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
        sO[(warp * 8 + g) * DEC_D + nb * 8 + 2 * t] = o[nb][0];
        sO[(warp * 8 + g) * DEC_D + nb * 8 + 2 * t + 1] = o[nb][1];
    }
    if (t == 0) {
        sML[(warp * 8 + g) * 2] = m_run[0];
        sML[(warp * 8 + g) * 2 + 1] = l_run[0];
    }
    __syncthreads();
    for (int idx = tid; idx < p.group * DEC_D; idx += DEC_THREADS) {
        const int r = idx / DEC_D, c = idx % DEC_D;
        float m = -INFINITY;
#pragma unroll
        for (int w = 0; w < DEC_WARPS; ++w) m = fmaxf(m, sML[(w * 8 + r) * 2]);
        float acc = 0.f, l = 0.f;
#pragma unroll
        for (int w = 0; w < DEC_WARPS; ++w) {
            const float mw = sML[(w * 8 + r) * 2];
            const float wgt = (mw == -INFINITY) ? 0.f : fast_exp2(mw - m);
            acc += wgt * sO[(w * 8 + r) * DEC_D + c];
            l += wgt * sML[(w * 8 + r) * 2 + 1];
        }
        const int head = kvh * p.group + r;
        if (p.n_splits == 1) {
            p.out[((long long)b * p.n_q_heads + head) * DEC_D + c] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
        } else {
            const long long pi = ((long long)b * p.n_q_heads + head) * p.n_splits + split;
            p.part_o[pi * DEC_D + c] = acc;
            if (c == 0) { p.part_ml[pi * 2] = m; p.part_ml[pi * 2 + 1] = l; }
        }
    }
}

__global__ void __launch_bounds__(DEC_D)
attn_decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, bf16* __restrict__ out,
                           int n_splits) {
    pdl_wait();
    pdl_launch_dependents();
    const long long bh = blockIdx.x;          // b * n_q_heads + head
    const int c = threadIdx.x;
    float m = -INFINITY;
    for (int s = 0; s < n_splits; ++s) m = fmaxf(m, part_ml[(bh * n_splits + s) * 2]);
    float acc = 0.f, l = 0.f;
    for (int s = 0; s < n_splits; ++s) {
        const float ms = part_ml[(bh * n_splits + s) * 2];
        const float w = (ms == -INFINITY) ? 0.f : fast_exp2(ms - m);
        acc += w * part_o[(bh * n_splits + s) * DEC_D + c];
        l += w * part_ml[(bh * n_splits + s) * 2 + 1];
    }
    out[bh * DEC_D + c] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
}

}  // namespace dots

using namespace dots;

static int launch_attn_decode(DecParams& p, int batch, int n_q_heads, int n_kv_heads, int head_dim, int n_splits, float softmax_scale,
                              void* stream, const char* who) {
    DOTS_REQUIRE(head_dim == DEC_D, "%s: head_dim must be 128", who);
    DOTS_REQUIRE(batch > 0 && n_q_heads % n_kv_heads == 0 && n_q_heads / n_kv_heads <= 8,
                 "%s: bad heads %d/%d (group must be <= 8)", who, n_q_heads, n_kv_heads);
    DOTS_REQUIRE(n_splits >= 1 && (n_splits == 1 || (p.part_o && p.part_ml)), "%s: n_splits>1 needs partial buffers", who);
    p.n_q_heads = n_q_heads; p.n_kv_heads = n_kv_heads; p.group = n_q_heads / n_kv_heads; p.n_splits = n_splits;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured = false;
    if (!configured) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
        configured = true;
    }
    dim3 grid(n_splits, n_kv_heads, batch);
    DOTS_CHECK_CUDA(launch_ex(attn_decode_kernel, dim3(grid), dim3(DEC_THREADS), (size_t)(DEC_SMEM), st, true, p));
    if (n_splits > 1) {
        DOTS_CHECK_CUDA(launch_ex(attn_decode_combine_kernel, dim3(batch * n_q_heads), dim3(DEC_D), (size_t)(0), st, true, p.part_o, p.part_ml, p.out, n_splits));
    }
    return 0;
}

extern "C" int dots_attn_decode(const void* q, const void* k_cache, const void* v_cache, const int* ctx_len, void* out,
                                float* part_o, float* part_ml, int batch, int n_q_heads, int n_kv_heads, int head_dim,
                                long long ctx_max, int n_splits, float softmax_scale, void* stream) {
    DecParams p{};
    p.q = (const bf16*)q; p.kc = (const bf16*)k_cache; p.vc = (const bf16*)v_cache; p.ctx_len = ctx_len;
    p.out = (bf16*)out; p.part_o = part_o; p.part_ml = part_ml; p.ctx_max = ctx_max;
    return launch_attn_decode(p, batch, n_q_heads, n_kv_heads, head_dim, n_splits, softmax_scale, stream, "dots_attn_decode");
}

extern "C" int dots_attn_decode_fused(const float* qkv_partial, int qkv_splits, const void* qkv_bias, const int* pos, const float* inv_freq,
                                      void* k_cache, void* v_cache, const int* ctx_len, void* out, float* part_o, float* part_ml,
                                      int batch, int n_q_heads, int n_kv_heads, int head_dim, long long ctx_max, int n_splits,
                                      float softmax_scale, void* stream) {
    DOTS_REQUIRE(qkv_partial && qkv_splits >= 1 && qkv_bias && pos && inv_freq, "dots_attn_decode_fused: missing QKV inputs");
    DecParams p{};
    p.q = nullptr; p.kc = (const bf16*)k_cache; p.vc = (const bf16*)v_cache; p.ctx_len = ctx_len;
    p.out = (bf16*)out; p.part_o = part_o; p.part_ml = part_ml; p.ctx_max = ctx_max;
    p.qkv_partial = qkv_partial; p.qkv_splits = qkv_splits; p.qkv_bias = (const bf16*)qkv_bias; p.pos = pos; p.inv_freq = inv_freq;
    p.kc_w = (bf16*)k_cache; p.vc_w = (bf16*)v_cache;
    return launch_attn_decode(p, batch, n_q_heads, n_kv_heads, head_dim, n_splits, softmax_scale, stream, "dots_attn_decode_fused");
}
