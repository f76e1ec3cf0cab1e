// HBM-bound kernels of the dots.ocr hot path: casts, RMSNorm / LayerNorm, 2-D and 1-D RoPE,
// KV-cache append, token embedding + image scatter, greedy argmax, and the decode-step
// "finalize" kernels that fold split-K reduction, bias, residual, norm, RoPE and cache append
// into one pass each.  All are coalesced 16-byte-vector kernels with warp-shuffle reductions;
// rounding points follow the HF eager path (fp32 inside a norm / rope / softmax, one bf16 rounding
// where HF materialises a bf16 tensor).
#include "common.h"
#include "ptx.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ------------------------------------------------------------------ cast + pad
// pixel_values [rows, cols] (fp32 or bf16) -> bf16 [rows, ldo], columns >= cols zero-filled
// (TMA needs a 16-byte row pitch; 588 * 2 B is not).  HF: hidden_states.to(bf16).
__global__ void cast_pad_kernel(const void* __restrict__ in, int in_is_bf16, long long rows, int cols, bf16* __restrict__ out,
                                int ldo) {
    pdl_wait();
    pdl_launch_dependents();
    const long long total = rows * (ldo / 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (ldo / 2);
        const int c = (int)(i % (ldo / 2)) * 2;
        float a = 0.f, b = 0.f;
        if (in_is_bf16) {
            const bf16* src = reinterpret_cast<const bf16*>(in) + r * cols;
            if (c < cols) a = __bfloat162float(src[c]);
            if (c + 1 < cols) b = __bfloat162float(src[c + 1]);
        } else {
            const float* src = reinterpret_cast<const float*>(in) + r * cols;
            if (c < cols) a = src[c];
            if (c + 1 < cols) b = src[c + 1];
        }
        *reinterpret_cast<uint32_t*>(out + r * ldo + c) = pack_bf16x2(a, b);
    }
}

// uint8 RGB page [H, W, 3] (already smart-resized, H and W multiples of patch * merge) -> normalised, patchified bf16 rows
// [gh * gw, ldo]: row order = 2x2 merge blocks contiguous, row layout (c, py, px), columns >= 3 * patch^2 zero.
// Same arithmetic as the host processor + cast: bf16( (float(u8) - 255 * mean_c) / (255 * std_c) )   (fp32, IEEE divide)
// (transformers/models/qwen2_vl/image_processing_qwen2_vl.py:148-232).
__global__ void patchify_u8_kernel(const uint8_t* __restrict__ img, int H, int W, int patch, int merge, float m0, float m1, float m2,
                                   float s0, float s1, float s2, bf16* __restrict__ out, int ldo) {
    pdl_wait();
    pdl_launch_dependents();
    const int gh = H / patch, gw = W / patch;
    const int pd = 3 * patch * patch;
    const long long total = (long long)gh * gw * ldo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / ldo), col = (int)(i % ldo);
        float v = 0.f;
        if (col < pd) {
            // r = ((bh * (gw / merge) + bw) * merge + ih) * merge + iw
            const int iw = r % merge, ih = (r / merge) % merge;
            const int blk = r / (merge * merge);
            const int bw = blk % (gw / merge), bh = blk / (gw / merge);
            const int c = col / (patch * patch), py = (col / patch) % patch, px = col % patch;
            const int y = (bh * merge + ih) * patch + py, x = (bw * merge + iw) * patch + px;
            const float u = (float)img[((long long)y * W + x) * 3 + c];
            const float m = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
            v = __fdiv_rn(__fsub_rn(u, m), sd);
        }
        out[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------ RMSNorm (one warp per row)
// out = bf16( bf16(x * rsqrt(mean(x^2) + eps)) * w )        (Qwen2RMSNorm, modeling_qwen2.py:258-263)
constexpr int NORM_MAX_CHUNKS = 8;     // per lane; cols <= 8 * 32 * 8 = 2048 on the register path

__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w,
                                                      bf16* __restrict__ out, long long ldo, long long rows, int cols, float eps) {
    pdl_wait();
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = blockIdx.x * 8LL + warp;
    if (row >= rows) return;
    const int nchunks = cols >> 3;
    const uint4* src = reinterpret_cast<const uint4*>(x + row * ldx);
    uint4 v[NORM_MAX_CHUNKS];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            v[i] = src[c];
            float f[8];
            unpack8(v[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss / (float)cols + eps);
    const uint4* wsrc = reinterpret_cast<const uint4*>(w);
    uint4* dst = reinterpret_cast<uint4*>(out + row * ldo);
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            float f[8], g[8];
            unpack8(v[i], f);
            unpack8(__ldg(wsrc + c), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j] * r) * g[j];
            dst[c] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------ LayerNorm (PatchMerger.ln_q)
// fp32 statistics, out = bf16((x - mean) * rstd * w + b)       (torch.nn.LayerNorm, eps 1e-6)
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ w,
                                                        const bf16* __restrict__ b, bf16* __restrict__ out, long long ldo,
                                                        long long rows, int cols, float eps) {
    pdl_wait();
    pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = blockIdx.x * 8LL + warp;
    if (row >= rows) return;
    const int nchunks = cols >> 3;
    const uint4* src = reinterpret_cast<const uint4*>(x + row * ldx);
    uint4 v[NORM_MAX_CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            v[i] = src[c];
            float f[8];
            unpack8(v[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[j];
        }
    }
    const float mean = warp_sum(s) / (float)cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            float f[8];
            unpack8(v[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; sq += d * d; }
        }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)cols + eps);
    const uint4* wsrc = reinterpret_cast<const uint4*>(w);
    const uint4* bsrc = reinterpret_cast<const uint4*>(b);
    uint4* dst = reinterpret_cast<uint4*>(out + row * ldo);
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            float f[8], g[8], h[8];
            unpack8(v[i], f);
            unpack8(__ldg(wsrc + c), g);
            unpack8(__ldg(bsrc + c), h);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * g[j] + h[j];
            dst[c] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------ ViT 2-D RoPE
// cos/sin table [S, 64] fp32: angle[s, j] = (j < 32 ? h(s) : w(s)) * inv_freq[j % 32]
// token order = processor order: 2x2 merge blocks contiguous (dots_ocr.py:536-568).
__global__ void vit_rope_table_kernel(const int* __restrict__ cu, const int* __restrict__ grid_hw, int n_img,
                                      const float* __restrict__ inv_freq, int half, int merge, float* __restrict__ cos_t,
                                      float* __restrict__ sin_t, int total_tokens) {
    pdl_wait();
    pdl_launch_dependents();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dim = 2 * half;
    if (idx >= total_tokens * dim) return;
    const int s = idx / dim, j = idx % dim;
    int img = 0;
    while (img + 1 < n_img && s >= cu[img + 1]) ++img;
    const int n = s - cu[img];
    const int W = grid_hw[2 * img + 1];
    const int per_frame = grid_hw[2 * img] * W;
    const int nn = n % per_frame;
    const int blk = nn / (merge * merge), in = nn % (merge * merge);
    const int bw = W / merge;
    const int hpos = (blk / bw) * merge + in / merge;
    const int wpos = (blk % bw) * merge + in % merge;
    const float ang = __fmul_rn((float)(j < half ? hpos : wpos), inv_freq[j % half]);
    cos_t[idx] = cosf(ang);
    sin_t[idx] = sinf(ang);
}

// In-place NeoX rotate-half on the q and k thirds of a fused qkv buffer [S, 3 * heads * 128], fp32 math,
// one rounding (ApplyRotaryEmb, enable_fp32_compute=True).  One thread = 8 (x1, x2) pairs.
__global__ void vit_rope_apply_kernel(bf16* __restrict__ qkv, long long ld, int S, int heads, const float* __restrict__ cos_t,
                                      const float* __restrict__ sin_t) {
    pdl_wait();
    pdl_launch_dependents();
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = (long long)S * heads * 2 * 8;
    if (idx >= total) return;
    const int c8 = (int)(idx & 7);
    const int hh = (int)((idx >> 3) % (heads * 2));
    const long long s = idx / (8LL * heads * 2);
    bf16* base = qkv + s * ld + (long long)hh * 128 + c8 * 8;     // q heads then k heads are contiguous
    uint4 u1 = *reinterpret_cast<uint4*>(base), u2 = *reinterpret_cast<uint4*>(base + 64);
    float x1[8], x2[8], o1[8], o2[8];
    unpack8(u1, x1);
    unpack8(u2, x2);
    const float4* cp = reinterpret_cast<const float4*>(cos_t + s * 64 + c8 * 8);
    const float4* sp = reinterpret_cast<const float4*>(sin_t + s * 64 + c8 * 8);
    float c[8], sn[8];
    *reinterpret_cast<float4*>(c) = cp[0]; *reinterpret_cast<float4*>(c + 4) = cp[1];
    *reinterpret_cast<float4*>(sn) = sp[0]; *reinterpret_cast<float4*>(sn + 4) = sp[1];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        o1[j] = __fsub_rn(__fmul_rn(x1[j], c[j]), __fmul_rn(x2[j], sn[j]));
        o2[j] = __fadd_rn(__fmul_rn(x2[j], c[j]), __fmul_rn(x1[j], sn[j]));
    }
    *reinterpret_cast<uint4*>(base) = pack8(o1);
    *reinterpret_cast<uint4*>(base + 64) = pack8(o2);
}

// ------------------------------------------------------------------ LLM 1-D RoPE (+ KV append)
// HF Qwen2: cos/sin are bf16 tensors; q_embed = q * cos + rotate_half(q) * sin evaluated in bf16
// (every product and the sum round to bf16)   (modeling_qwen2.py:102-146).
__device__ __forceinline__ void rope_bf16_8(const float (&x1)[8], const float (&x2)[8], int pos, const float* inv_freq, int i0,
                                            float (&o1)[8], float (&o2)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float ang = __fmul_rn((float)pos, inv_freq[i0 + j]);
        const float c = bf16_round(cosf(ang)), s = bf16_round(sinf(ang));
        o1[j] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x1[j], c)), bf16_round(__fmul_rn(-x2[j], s))));
        o2[j] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x2[j], c)), bf16_round(__fmul_rn(x1[j], s))));
    }
}

// Prefill: qkv [T, (nq + 2 nkv) * 128] (bias already added by the GEMM epilogue).  Rotates q and k
// in place (the prefill attention reads them from this buffer) and appends k, v to the cache.
__global__ void llm_rope_append_kernel(bf16* __restrict__ qkv, long long ld, int T, int nq, int nkv, const int* __restrict__ pos,
                                       const int* __restrict__ seq_of_tok, const float* __restrict__ inv_freq,
                                       bf16* __restrict__ kc, bf16* __restrict__ vc, long long ctx_max) {
    pdl_wait();
    pdl_launch_dependents();
    const int units = (nq + 2 * nkv) * 8;           // 8 threads per head (each: 8 low + 8 high dims)
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)T * units) return;
    const int tkn = (int)(idx / units);
    const int u = (int)(idx % units);
    const int head = u >> 3, c8 = u & 7;
    bf16* base = qkv + (long long)tkn * ld + head * 128 + c8 * 8;
    const int p = pos[tkn];
    if (head < nq + nkv) {
        float x1[8], x2[8], o1[8], o2[8];
        unpack8(*reinterpret_cast<uint4*>(base), x1);
        unpack8(*reinterpret_cast<uint4*>(base + 64), x2);
        rope_bf16_8(x1, x2, p, inv_freq, c8 * 8, o1, o2);
        const uint4 r1 = pack8(o1), r2 = pack8(o2);
        *reinterpret_cast<uint4*>(base) = r1;
        *reinterpret_cast<uint4*>(base + 64) = r2;
        if (head >= nq) {
            // cache stripes are stored in 64-key tiles with the shared-memory image of the decode kernel (kv_tiled_off)
            bf16* stripe = kc + ((long long)seq_of_tok[tkn] * nkv + (head - nq)) * ctx_max * 128;
            *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, c8 * 8)) = r1;
            *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, 64 + c8 * 8)) = r2;
        }
    } else {
        bf16* stripe = vc + ((long long)seq_of_tok[tkn] * nkv + (head - nq - nkv)) * ctx_max * 128;
        *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, c8 * 8)) = *reinterpret_cast<uint4*>(base);
        *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, 64 + c8 * 8)) = *reinterpret_cast<uint4*>(base + 64);
    }
}

// ------------------------------------------------------------------ page resize (uint8 bicubic + antialias, integer arithmetic)
// One separable pass of the resize the stock image processor runs on the CPU (Pillow's ImagingResample as ported to ATen for
// uint8 tensors; torchvision resize(BICUBIC, antialias=True), image_processing_qwen2_vl.py:148-232):
//     out[.., i, ..] = clip((2^(p-1) + sum_j in[.., xmin[i] + j, ..] * w[i][j]) >> p, 0, 255)        (int32, int16 taps)
// The tap tables come from the host (dots_ocr_b200/resize.py), so the result equals the CPU resize bit for bit.
// inner = bytes of one step along the resampled axis's inner neighbours (3 for the horizontal pass over HWC pixels, W * 3 for the
// vertical pass); a thread produces one output byte, consecutive threads consecutive bytes.
__global__ void resample_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_outer, int in_size, int out_size,
                                   long long inner, const int* __restrict__ xmin, const int* __restrict__ xsize,
                                   const short* __restrict__ w, int ksize, int prec) {
    pdl_wait();
    pdl_launch_dependents();
    const long long total = n_outer * out_size * inner;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long in_ = idx % inner;
        const int i = (int)((idx / inner) % out_size);
        const long long o = idx / (inner * out_size);
        const uint8_t* src = in + (o * in_size + xmin[i]) * inner + in_;
        const short* wi = w + (long long)i * ksize;
        int acc = 1 << (prec - 1);
        const int n = xsize[i];
        for (int j = 0; j < n; ++j) acc += (int)src[(long long)j * inner] * (int)wi[j];
        acc >>= prec;
        out[idx] = (uint8_t)min(max(acc, 0), 255);
    }
}

// ------------------------------------------------------------------ embedding + image scatter
// slots[t] = rank of token t among image-pad tokens (masked_scatter order), or -1 for text tokens.
__global__ void __launch_bounds__(1024) image_slots_kernel(const long long* __restrict__ ids, int T, long long image_token,
                                                           int* __restrict__ slots, int* __restrict__ count_out) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ int warp_excl[32];
    __shared__ int chunk_total;
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < T; base += 1024) {
        const int tkn = base + threadIdx.x;
        const int flag = (tkn < T && ids[tkn] == image_token) ? 1 : 0;
        int incl = flag;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_excl[warp] = incl;          // per-warp totals
        __syncthreads();
        if (warp == 0) {
            const int wv = warp_excl[lane];
            int wi = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += n;
            }
            warp_excl[lane] = wi - wv;                   // exclusive prefix over warps
            if (lane == 31) chunk_total = wi;
        }
        __syncthreads();
        if (tkn < T) slots[tkn] = flag ? (carry + warp_excl[warp] + incl - flag) : -1;
        __syncthreads();                                 // everyone has read carry
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && count_out) *count_out = carry;
}

__global__ void embed_scatter_kernel(const long long* __restrict__ ids, const int* __restrict__ slots, const bf16* __restrict__ table,
                                     const bf16* __restrict__ img, bf16* __restrict__ out, int T, int H, long long vocab) {
    pdl_wait();
    pdl_launch_dependents();
    const int chunks = H >> 3;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)T * chunks) return;
    const int tkn = (int)(idx / chunks), c = (int)(idx % chunks);
    const int slot = slots ? slots[tkn] : -1;
    const bf16* src;
    if (slot >= 0) src = img + (long long)slot * H;
    else {
        long long id = ids[tkn];
        if (id < 0 || id >= vocab) id = 0;
        src = table + id * H;
    }
    reinterpret_cast<uint4*>(out + (long long)tkn * H)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
}

__global__ void gather_rows_kernel(const bf16* __restrict__ src, long long lds, const int* __restrict__ rows, bf16* __restrict__ out,
                                   long long ldo, int n, int cols) {
    pdl_wait();
    pdl_launch_dependents();
    const int chunks = cols >> 3;
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= (long long)n * chunks) return;
    const int r = (int)(idx / chunks), c = (int)(idx % chunks);
    reinterpret_cast<uint4*>(out + (long long)r * ldo)[c] = reinterpret_cast<const uint4*>(src + (long long)rows[r] * lds)[c];
}

// ------------------------------------------------------------------ greedy argmax + sequence state advance
// logits bf16 -> fp32 -> argmax, lowest index wins ties (torch.argmax; generation/utils.py:2762,2793).
// Then the HF bookkeeping of one greedy step: finished rows emit pad, EOS marks a row finished
// (utils.py:2796-2805; any id of generation_config.eos_token_id stops a row), the token is appended, position / context
// length advance.
struct StopIds { long long id[DOTS_MAX_STOP_IDS]; int n; };

__global__ void __launch_bounds__(1024) argmax_advance_kernel(const bf16* __restrict__ logits, long long ldl, int V,
                                                              long long* __restrict__ next_ids, long long* __restrict__ out_ids,
                                                              long long out_ld, int* __restrict__ step, int* __restrict__ pos,
                                                              int* __restrict__ ctx_len, int* __restrict__ finished,
                                                              const StopIds stops, long long pad_id, const long long* __restrict__ forced,
                                                              long long forced_ld) {
    pdl_wait();
    pdl_launch_dependents();
    const int b = blockIdx.x;
    const bf16* row = logits + (long long)b * ldl;
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    const int nchunks = V >> 3;
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(row) + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (f[j] > best) { best = f[j]; bidx = c * 8 + j; }       // strictly greater: first index kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
    }
    __shared__ float sb[32];
    __shared__ int si[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sb[warp] = best; si[warp] = bidx; }
    __syncthreads();
    if (warp == 0) {
        best = (lane < (blockDim.x >> 5)) ? sb[lane] : -INFINITY;
        bidx = (lane < (blockDim.x >> 5)) ? si[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if (lane == 0) {
            long long tok = bidx;
            const int st = step ? step[b] : 0;
            if (forced) tok = forced[(long long)b * forced_ld + st];   // teacher forcing (parity tests)
            if (finished) {
                if (finished[b]) tok = pad_id;
                else {
                    bool hit = false;
#pragma unroll
                    for (int i = 0; i < DOTS_MAX_STOP_IDS; ++i) hit |= (i < stops.n) && (tok == stops.id[i]);
                    if (hit) finished[b] = 1;
                }
            }
            next_ids[b] = tok;
            if (out_ids) out_ids[(long long)b * out_ld + st] = tok;
            if (step) step[b] = st + 1;
            if (pos) pos[b] += 1;
            if (ctx_len) ctx_len[b] += 1;
        }
    }
}

// ------------------------------------------------------------------ decode-step finalize kernels
// acc[0..7] += sum over splits of 8 consecutive fp32 at p0 + s * sstride, in split order (deterministic), with four
// independent 32-byte loads in flight per thread.
__device__ __forceinline__ void sum_splits8(const float* __restrict__ p0, long long sstride, int splits, float (&acc)[8]) {
    // up to 16 splits: every 32-byte load is issued before the first add (one L2 round trip); adds stay in split order
    for (int s0 = 0; s0 < splits; s0 += 16) {
        float4 a[16], d[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (s0 + u < splits) {
                const float4* ps = reinterpret_cast<const float4*>(p0 + (s0 + u) * sstride);
                a[u] = ps[0]; d[u] = ps[1];
            }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (s0 + u < splits) {
                acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
                acc[4] += d[u].x; acc[5] += d[u].y; acc[6] += d[u].z; acc[7] += d[u].w;
            }
        }
    }
}

// x = embed[last_id]; resid = x; normed = RMSNorm(x) * w          (one warp per sequence)
__global__ void __launch_bounds__(256) decode_embed_rmsnorm_kernel(const long long* __restrict__ ids, const bf16* __restrict__ table,
                                                                   long long vocab, const bf16* __restrict__ w, bf16* __restrict__ resid,
                                                                   bf16* __restrict__ normed, int B, int H, float eps,
                                                                   unsigned* __restrict__ counters, int n_counters, int tile_rows) {
    pdl_wait();
    pdl_launch_dependents();
    // first kernel of a decode step: re-arm the rendezvous counters of this step's dots_decode_gemm_resnorm launches
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < n_counters; i += blockDim.x) counters[i] = 0u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    long long id = ids[b];
    if (id < 0 || id >= vocab) id = 0;
    const uint4* src = reinterpret_cast<const uint4*>(table + id * H);
    const int nchunks = H >> 3;
    uint4 v[NORM_MAX_CHUNKS];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            v[i] = __ldg(src + c);
            float f[8];
            unpack8(v[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
            reinterpret_cast<uint4*>(resid + (long long)b * H)[c] = v[i];
        }
    }
    const float r = rsqrtf(warp_sum(ss) / (float)H + eps);
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = lane + i * 32;
        if (c < nchunks) {
            float f[8], g[8];
            unpack8(v[i], f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(w) + c), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j] * r) * g[j];
            // tile_rows > 0: `normed` is the k-block-tiled B operand of the cluster GEMMs (tiled_row_off), else row-major
            bf16* dst = tile_rows > 0 ? normed + tiled_row_off(b, c * 8, tile_rows) : normed + (long long)b * H + c * 8;
            *reinterpret_cast<uint4*>(dst) = pack8(f);
        }
    }
}

// x = bf16(sum partial); resid = bf16(resid + x); normed = RMSNorm(resid) * w
// One CTA per sequence, one 8-column chunk per thread (H / 8 <= 256 threads): the split-K partials of a row are
// [splits] strided reads of 32 B per thread, all independent, so the whole reduction is one round of loads.
__global__ void __launch_bounds__(256) decode_residual_rmsnorm_kernel(const float* __restrict__ partial, int splits,
                                                                      bf16* __restrict__ resid, const bf16* __restrict__ w,
                                                                      bf16* __restrict__ normed, int B, int H, float eps, int tile_rows,
                                                                      unsigned long long* trace) {
    if (threadIdx.x == 0) trace_point(trace, 30, 0);
    pdl_wait();
    pdl_launch_dependents();        // (releasing the dependents ahead of the wait was measured: 1.845 against 1.832 ms per step)
    if (threadIdx.x == 0) trace_point(trace, 30, 1);
    __shared__ float s_part[8];
    const int b = blockIdx.x;
    const int c = threadIdx.x;
    const int nchunks = H >> 3;
    const bool live = c < nchunks;
    float x[8];
    float ss = 0.f;
    uint4 rv = make_uint4(0, 0, 0, 0), gv = make_uint4(0, 0, 0, 0);
    if (live) {
        // the residual row and the norm weight do not depend on the partials: request them first so that the whole kernel is one
        // round of L2 latency (plus the block reduction)
        rv = reinterpret_cast<const uint4*>(resid + (long long)b * H)[c];
        gv = __ldg(reinterpret_cast<const uint4*>(w) + c);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        sum_splits8(partial + (long long)b * H + c * 8, (long long)B * H, splits, acc);
        float rr[8];
        unpack8(rv, rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x[j] = bf16_round(bf16_round(acc[j]) + rr[j]);
            ss += x[j] * x[j];
        }
        reinterpret_cast<uint4*>(resid + (long long)b * H)[c] = pack8(x);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    const int nw = (blockDim.x + 31) >> 5;
    for (int i = 0; i < nw; ++i) tot += s_part[i];
    const float r = rsqrtf(tot / (float)H + eps);
    if (live) {
        float g[8], f[8];
        unpack8(gv, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = bf16_round(x[j] * r) * g[j];
        bf16* dst = tile_rows > 0 ? normed + tiled_row_off(b, c * 8, tile_rows) : normed + (long long)b * H + c * 8;
        *reinterpret_cast<uint4*>(dst) = pack8(f);
    }
    if (threadIdx.x == 0) trace_point(trace, 30, 4);
}

// q,k,v = bf16(sum partial + bias); RoPE(q, k) at pos[b]; q -> q_out, k,v -> cache[b, :, pos[b]]
__global__ void decode_qkv_rope_append_kernel(const float* __restrict__ partial, int splits, const bf16* __restrict__ bias,
                                              const int* __restrict__ pos, const float* __restrict__ inv_freq, bf16* __restrict__ q_out,
                                              bf16* __restrict__ kc, bf16* __restrict__ vc, long long ctx_max, int B, int nq, int nkv) {
    pdl_wait();
    pdl_launch_dependents();
    const int units = (nq + 2 * nkv) * 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * units) return;
    const int b = idx / units, u = idx % units;
    const int head = u >> 3, c8 = u & 7;
    const int N = (nq + 2 * nkv) * 128;
    const int col = head * 128 + c8 * 8;
    float x1[8], x2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { x1[j] = 0.f; x2[j] = 0.f; }
    sum_splits8(partial + (long long)b * N + col, (long long)B * N, splits, x1);
    sum_splits8(partial + (long long)b * N + col + 64, (long long)B * N, splits, x2);
    float b1[8], b2[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(bias + col)), b1);
    unpack8(__ldg(reinterpret_cast<const uint4*>(bias + col + 64)), b2);
#pragma unroll
    for (int j = 0; j < 8; ++j) { x1[j] = bf16_round(x1[j] + b1[j]); x2[j] = bf16_round(x2[j] + b2[j]); }
    const int p = pos[b];
    if (head < nq + nkv) {
        float o1[8], o2[8];
        rope_bf16_8(x1, x2, p, inv_freq, c8 * 8, o1, o2);
        if (head < nq) {
            bf16* dst = q_out + ((long long)b * nq + head) * 128 + c8 * 8;
            *reinterpret_cast<uint4*>(dst) = pack8(o1);
            *reinterpret_cast<uint4*>(dst + 64) = pack8(o2);
        } else {
            bf16* stripe = kc + ((long long)b * nkv + (head - nq)) * ctx_max * 128;
            *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, c8 * 8)) = pack8(o1);
            *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, 64 + c8 * 8)) = pack8(o2);
        }
    } else {
        bf16* stripe = vc + ((long long)b * nkv + (head - nq - nkv)) * ctx_max * 128;
        *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, c8 * 8)) = pack8(x1);
        *reinterpret_cast<uint4*>(stripe + kv_tiled_off(p, 64 + c8 * 8)) = pack8(x2);
    }
}

// act = bf16( bf16(silu(bf16 gate)) * bf16 up ), gate/up interleaved per 128-column block
// ([64 gate | 64 up], the layout the SWIGLU epilogues use), partial [splits][B][2I].
__global__ void decode_swiglu_kernel(const float* __restrict__ partial, int splits, bf16* __restrict__ act, int B, int I) {
    pdl_wait();
    pdl_launch_dependents();
    const int chunks = I >> 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * chunks) return;
    const int b = idx / chunks, c = idx % chunks;
    const int col = c * 8;
    const int gcol = (col >> 6) * 128 + (col & 63);
    float gsum[8], usum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { gsum[j] = 0.f; usum[j] = 0.f; }
    sum_splits8(partial + (long long)b * (2LL * I) + gcol, (long long)B * 2LL * I, splits, gsum);
    sum_splits8(partial + (long long)b * (2LL * I) + gcol + 64, (long long)B * 2LL * I, splits, usum);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float gv = bf16_round(gsum[j]), uv = bf16_round(usum[j]);
        o[j] = bf16_round(gv / (1.0f + expf(-gv))) * uv;
    }
    reinterpret_cast<uint4*>(act + (long long)b * I)[c] = pack8(o);
}

}  // namespace dots

using namespace dots;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int dots_cast_pad_bf16(const void* in, int in_is_bf16, long long rows, int cols, void* out, int ldo, void* stream) {
    DOTS_REQUIRE(rows > 0 && cols > 0 && ldo >= cols && ldo % 8 == 0, "dots_cast_pad_bf16: bad shape rows=%lld cols=%d ldo=%d", rows, cols, ldo);
    const long long total = rows * (ldo / 2);
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    DOTS_CHECK_CUDA(launch_ex(cast_pad_kernel, dim3(blocks), dim3(256), (size_t)(0), ST(stream), true, in, in_is_bf16, rows, cols, (bf16*)out, ldo));
    return 0;
}

extern "C" int dots_patchify_u8(const void* img_hwc, int H, int W, int patch, int merge, const float* mean255, const float* std255,
                                void* out, int ldo, void* stream) {
    DOTS_REQUIRE(img_hwc && out && mean255 && std255, "dots_patchify_u8: null pointer");
    DOTS_REQUIRE(patch > 0 && merge > 0 && H > 0 && W > 0 && H % (patch * merge) == 0 && W % (patch * merge) == 0,
                 "dots_patchify_u8: H, W must be multiples of patch * merge (got %d x %d)", H, W);
    DOTS_REQUIRE(ldo >= 3 * patch * patch && ldo % 8 == 0, "dots_patchify_u8: ldo must cover 3 * patch^2 and keep 16-byte rows");
    const long long total = (long long)(H / patch) * (W / patch) * ldo;
    const int blocks = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    DOTS_CHECK_CUDA(launch_ex(patchify_u8_kernel, dim3(blocks), dim3(256), (size_t)(0), ST(stream), true, (const uint8_t*)img_hwc, H, W, patch, merge,
                              mean255[0], mean255[1], mean255[2], std255[0], std255[1], std255[2], (bf16*)out, ldo));
    return 0;
}

extern "C" int dots_resize_bicubic_u8(const void* img_hwc, int H, int W, void* tmp, void* out, int rh, int rw, const int* xmin_x, const int* xsize_x,
                                      const short* w_x, int ksize_x, int prec_x, const int* xmin_y, const int* xsize_y, const short* w_y, int ksize_y,
                                      int prec_y, void* stream) {
    DOTS_REQUIRE(img_hwc && out && H > 0 && W > 0 && rh > 0 && rw > 0, "dots_resize_bicubic_u8: bad arguments");
    DOTS_REQUIRE(H != rh || W != rw, "dots_resize_bicubic_u8: nothing to do (same size)");
    DOTS_REQUIRE(W == rw || (xmin_x && xsize_x && w_x && ksize_x > 0 && prec_x > 0), "dots_resize_bicubic_u8: horizontal tap tables missing");
    DOTS_REQUIRE(H == rh || (xmin_y && xsize_y && w_y && ksize_y > 0 && prec_y > 0), "dots_resize_bicubic_u8: vertical tap tables missing");
    DOTS_REQUIRE((W == rw || H == rh) || tmp, "dots_resize_bicubic_u8: both axes change: need the [H, rw, 3] intermediate buffer");
    auto blocks = [](long long total) { return (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16); };
    const void* src = img_hwc;
    if (W != rw) {                                       // horizontal pass first (as Pillow / ATen do), uint8 intermediate
        void* dst = (H != rh) ? tmp : out;
        const long long total = (long long)H * rw * 3;
        DOTS_CHECK_CUDA(launch_ex(resample_u8_kernel, dim3(blocks(total)), dim3(256), (size_t)(0), ST(stream), true, (const uint8_t*)src, (uint8_t*)dst,
                                  (long long)H, W, rw, 3LL, xmin_x, xsize_x, w_x, ksize_x, prec_x));
        src = dst;
    }
    if (H != rh) {
        const long long total = (long long)rh * rw * 3;
        DOTS_CHECK_CUDA(launch_ex(resample_u8_kernel, dim3(blocks(total)), dim3(256), (size_t)(0), ST(stream), true, (const uint8_t*)src, (uint8_t*)out,
                                  1LL, H, rh, (long long)rw * 3, xmin_y, xsize_y, w_y, ksize_y, prec_y));
    }
    return 0;
}

extern "C" int dots_rmsnorm(const void* x, long long ldx, const void* w, void* out, long long ldo, long long rows, int cols,
                            float eps, void* stream) {
    DOTS_REQUIRE(rows > 0 && cols % 8 == 0 && cols <= NORM_MAX_CHUNKS * 256 && ldx % 8 == 0 && ldo % 8 == 0,
                 "dots_rmsnorm: need cols %% 8 == 0, cols <= 2048, 16-byte pitches (cols=%d)", cols);
    DOTS_CHECK_CUDA(launch_ex(rmsnorm_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), ST(stream), true, (const bf16*)x, ldx, (const bf16*)w, (bf16*)out, ldo, rows, cols, eps));
    return 0;
}

extern "C" int dots_layernorm(const void* x, long long ldx, const void* w, const void* b, void* out, long long ldo,
                              long long rows, int cols, float eps, void* stream) {
    DOTS_REQUIRE(rows > 0 && cols % 8 == 0 && cols <= NORM_MAX_CHUNKS * 256 && ldx % 8 == 0 && ldo % 8 == 0,
                 "dots_layernorm: need cols %% 8 == 0, cols <= 2048, 16-byte pitches (cols=%d)", cols);
    DOTS_CHECK_CUDA(launch_ex(layernorm_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), ST(stream), true, (const bf16*)x, ldx, (const bf16*)w, (const bf16*)b, (bf16*)out, ldo, rows, cols, eps));
    return 0;
}

extern "C" int dots_vit_rope_table(const int* cu_seqlens, const int* grid_hw, int n_img, const float* inv_freq, int half,
                                   int merge, float* cos_t, float* sin_t, int total_tokens, void* stream) {
    DOTS_REQUIRE(n_img > 0 && total_tokens > 0 && half == 32 && merge > 0, "dots_vit_rope_table: bad args (half must be 32)");
    const long long n = (long long)total_tokens * 2 * half;
    DOTS_CHECK_CUDA(launch_ex(vit_rope_table_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), ST(stream), true, cu_seqlens, grid_hw, n_img, inv_freq, half, merge, cos_t, sin_t, total_tokens));
    return 0;
}

extern "C" int dots_vit_rope_apply(void* qkv, long long ld, int S, int heads, int head_dim, const float* cos_t, const float* sin_t,
                                   void* stream) {
    DOTS_REQUIRE(S > 0 && heads > 0 && head_dim == 128 && ld % 8 == 0, "dots_vit_rope_apply: head_dim must be 128, pitch %% 8 == 0");
    const long long n = (long long)S * heads * 2 * 8;
    DOTS_CHECK_CUDA(launch_ex(vit_rope_apply_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), ST(stream), true, (bf16*)qkv, ld, S, heads, cos_t, sin_t));
    return 0;
}

extern "C" int dots_llm_rope_kv_append(void* qkv, long long ld, int T, int n_q_heads, int n_kv_heads, int head_dim,
                                       const int* positions, const int* seq_of_tok, const float* inv_freq, void* k_cache,
                                       void* v_cache, long long ctx_max, void* stream) {
    DOTS_REQUIRE(T > 0 && head_dim == 128 && ld % 8 == 0, "dots_llm_rope_kv_append: head_dim must be 128");
    const long long n = (long long)T * (n_q_heads + 2 * n_kv_heads) * 8;
    DOTS_CHECK_CUDA(launch_ex(llm_rope_append_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), ST(stream), true, (bf16*)qkv, ld, T, n_q_heads, n_kv_heads, positions, seq_of_tok, inv_freq, (bf16*)k_cache, (bf16*)v_cache, ctx_max));
    return 0;
}

extern "C" int dots_image_slots(const long long* ids, int T, long long image_token_id, int* slots, int* count_out, void* stream) {
    DOTS_REQUIRE(T > 0, "dots_image_slots: empty input");
    DOTS_CHECK_CUDA(launch_ex(image_slots_kernel, dim3(1), dim3(1024), (size_t)(0), ST(stream), true, ids, T, image_token_id, slots, count_out));
    return 0;
}

extern "C" int dots_embed_scatter(const long long* ids, const int* slots, const void* table, const void* img_embeds, void* out,
                                  int T, int H, long long vocab, void* stream) {
    DOTS_REQUIRE(T > 0 && H % 8 == 0, "dots_embed_scatter: bad shape");
    const long long n = (long long)T * (H / 8);
    DOTS_CHECK_CUDA(launch_ex(embed_scatter_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), ST(stream), true, ids, slots, (const bf16*)table, (const bf16*)img_embeds, (bf16*)out, T, H, vocab));
    return 0;
}

extern "C" int dots_gather_rows(const void* src, long long lds, const int* rows, void* out, long long ldo, int n, int cols,
                                void* stream) {
    DOTS_REQUIRE(n > 0 && cols % 8 == 0 && lds % 8 == 0 && ldo % 8 == 0, "dots_gather_rows: bad shape");
    const long long t = (long long)n * (cols / 8);
    DOTS_CHECK_CUDA(launch_ex(gather_rows_kernel, dim3((unsigned)((t + 255) / 256)), dim3(256), (size_t)(0), ST(stream), true, (const bf16*)src, lds, rows, (bf16*)out, ldo, n, cols));
    return 0;
}

extern "C" int dots_argmax_advance(const void* logits, long long ldl, int batch, int vocab, long long* next_ids, long long* out_ids,
                                   long long out_ld, int* step, int* pos, int* ctx_len, int* finished, const long long* stop_ids,
                                   int n_stops, long long pad_id, const long long* forced_ids, long long forced_ld, void* stream) {
    DOTS_REQUIRE(batch > 0 && vocab % 8 == 0 && ldl % 8 == 0, "dots_argmax_advance: vocab and pitch must be multiples of 8");
    DOTS_REQUIRE(n_stops >= 0 && n_stops <= DOTS_MAX_STOP_IDS && (n_stops == 0 || stop_ids), "dots_argmax_advance: at most %d stop ids (got %d)",
                 DOTS_MAX_STOP_IDS, n_stops);
    StopIds stops{};
    stops.n = n_stops;
    for (int i = 0; i < n_stops; ++i) stops.id[i] = stop_ids[i];
    DOTS_CHECK_CUDA(launch_ex(argmax_advance_kernel, dim3(batch), dim3(1024), (size_t)(0), ST(stream), true, (const bf16*)logits, ldl, vocab, next_ids, out_ids, out_ld, step, pos, ctx_len, finished, stops, pad_id, forced_ids, forced_ld));
    return 0;
}

extern "C" int dots_decode_embed_rmsnorm(const long long* ids, const void* table, long long vocab, const void* w, void* resid,
                                         void* normed, int batch, int H, float eps, unsigned int* counters, int n_counters, int tile_rows,
                                         void* stream) {
    DOTS_REQUIRE(batch > 0 && H % 8 == 0 && H <= NORM_MAX_CHUNKS * 256, "dots_decode_embed_rmsnorm: H %% 8, H <= 2048");
    DOTS_REQUIRE(n_counters >= 0 && (n_counters == 0 || counters), "dots_decode_embed_rmsnorm: n_counters without a counter array");
    DOTS_REQUIRE(tile_rows == 0 || (tile_rows % 8 == 0 && batch <= tile_rows && H % 64 == 0), "dots_decode_embed_rmsnorm: bad tile_rows %d", tile_rows);
    DOTS_CHECK_CUDA(launch_ex(decode_embed_rmsnorm_kernel, dim3((batch + 7) / 8), dim3(256), (size_t)(0), ST(stream), true, ids, (const bf16*)table, vocab, (const bf16*)w, (bf16*)resid, (bf16*)normed, batch, H, eps, counters, n_counters, tile_rows));
    return 0;
}

extern "C" int dots_decode_residual_rmsnorm(const float* partial, int splits, void* resid, const void* w, void* normed, int batch,
                                            int H, float eps, int tile_rows, void* stream) {
    DOTS_REQUIRE(batch > 0 && splits > 0 && H % 8 == 0 && H <= NORM_MAX_CHUNKS * 256, "dots_decode_residual_rmsnorm: bad shape");
    DOTS_REQUIRE(tile_rows == 0 || (tile_rows % 8 == 0 && batch <= tile_rows && H % 64 == 0), "dots_decode_residual_rmsnorm: bad tile_rows %d", tile_rows);
    const int threads = ((H / 8) + 31) / 32 * 32;
    DOTS_CHECK_CUDA(launch_ex(decode_residual_rmsnorm_kernel, dim3(batch), dim3(threads), (size_t)(0), ST(stream), true, partial, splits, (bf16*)resid, (const bf16*)w, (bf16*)normed, batch, H, eps, tile_rows, g_trace));
    return 0;
}

extern "C" int dots_decode_qkv_rope_append(const float* partial, int splits, const void* bias, const int* pos, const float* inv_freq,
                                           void* q_out, void* k_cache, void* v_cache, long long ctx_max, int batch, int n_q_heads,
                                           int n_kv_heads, int head_dim, void* stream) {
    DOTS_REQUIRE(batch > 0 && splits > 0 && head_dim == 128, "dots_decode_qkv_rope_append: head_dim must be 128");
    const int n = batch * (n_q_heads + 2 * n_kv_heads) * 8;
    DOTS_CHECK_CUDA(launch_ex(decode_qkv_rope_append_kernel, dim3((n + 127) / 128), dim3(128), (size_t)(0), ST(stream), true, partial, splits, (const bf16*)bias, pos, inv_freq, (bf16*)q_out, (bf16*)k_cache, (bf16*)v_cache, ctx_max, batch, n_q_heads, n_kv_heads));
    return 0;
}

extern "C" int dots_decode_swiglu(const float* partial, int splits, void* act, int batch, int inter, void* stream) {
    DOTS_REQUIRE(batch > 0 && splits > 0 && inter % 64 == 0, "dots_decode_swiglu: intermediate size must be a multiple of 64");
    const int n = batch * (inter / 8);
    DOTS_CHECK_CUDA(launch_ex(decode_swiglu_kernel, dim3((n + 255) / 256), dim3(256), (size_t)(0), ST(stream), true, partial, splits, (bf16*)act, batch, inter));
    return 0;
}
