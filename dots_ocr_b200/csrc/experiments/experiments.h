/*
 * experiments.h -- entry points of the kernels kept as documented NEGATIVE RESULTS.
 * They are not part of the product library: `DOTS_BUILD_EXPERIMENTS=1 python -m dots_ocr_b200.build --force`
 * adds them to libdots_ocr_b200.so for A/B runs (tests/test_experiments_gpu.py skips when they are absent).
 */
#ifndef DOTS_OCR_B200_EXPERIMENTS_H
#define DOTS_OCR_B200_EXPERIMENTS_H
#include "../../../include/dots_ocr_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Same contract on CTA pairs (tcgen05 cta_group::2): a cluster of two CTAs covers 512 query rows and shares every K/V tile,
 * each CTA staging half of it (halves the shared-memory traffic that bounds the single-CTA kernel). */
DOTS_API int dots_attn_varlen_fwd_pair(const void* q, long long q_stride, const void* k, long long k_stride, const void* v,
                              long long v_stride, void* out, long long o_stride, const int* cu_seqlens, int n_seqs,
                              int max_seqlen, long long total_tokens, int n_q_heads, int n_kv_heads, int head_dim,
                              int causal, float softmax_scale, void* stream);

/* One persistent kernel for the GEMM chain of a decoder layer at decode time (batch <= 64):
 *   o_proj -> residual+RMSNorm(ln_mid) -> gate|up+SwiGLU -> down_proj -> residual+RMSNorm(ln_next) -> [qkv of the next layer].
 * Equivalent (bit for bit) to dots_gemm_skinny_bf16(attn, w_o, splits_o) + dots_decode_residual_rmsnorm +
 * dots_gemm_skinny_swiglu_bf16 + dots_gemm_skinny_bf16(act, w_down, splits_down) + dots_decode_residual_rmsnorm
 * [+ dots_gemm_skinny_bf16(normed, w_qkv_next, splits_qkv) -> partial], but the weights of all phases stream through one
 * shared-memory ring without stopping at kernel boundaries; phases are separated by device-wide counters (`counters`: 8 x
 * uint32, zero-initialised ONCE by the caller and then owned by the kernel: monotonic counts + a launch epoch).  w_qkv_next may be NULL (last layer).  Launches one CTA per SM, all co-resident.  [Q]:243-244,302-308,46-48. */
DOTS_API int dots_decode_chain(const void* attn, const void* w_o, const void* w_gu, const void* w_down, const void* w_qkv_next,
                      float* partial, void* resid, void* normed, void* act, const void* ln_mid, const void* ln_next,
                      unsigned int* counters, int batch, int hidden, int inter, int qkv_n, int attn_dim, int splits_o,
                      int splits_down, int splits_qkv, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif
