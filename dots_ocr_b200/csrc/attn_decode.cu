// Decode-step attention over the in-place KV cache (decode half).
//
// One query token per sequence, grouped-query: the G q-heads that share a kv-head are packed into the 16-row M dimension
// of mma.sync so K and V are read from HBM exactly once per step.  The kernel is HBM-bound (2 * ctx * 256 B per
// (sequence, kv head)); what matters is bytes in flight, so K/V arrive through a TMA-fed ring:
//
//   warp 4 (one lane)  producer: 64-key K and V tiles (16 KB each, 128-B swizzle boxes) into a 3-stage mbarrier ring.
//                      Tiles that cannot contain the token being appended are requested BEFORE the programmatic-dependent-
//                      launch wait, i.e. while the QKV GEMM of this layer is still running.
//   warps 0-3          consumers: warp w owns keys [16w, 16w+16) of every tile (QK^T, online softmax, PV with mma.sync),
//                      fixed-order merge of the four partial (m, l, O) through shared memory at the end.
//
// Keys are additionally split across CTAs (grid.x = n_splits, "flash decoding").  Up to DEC_MAX_CLUSTER splits form a
// thread-block CLUSTER: the non-leader CTAs hand their partial (m, l, O) to the leader through distributed shared memory
// and the leader merges them in split order -- no partials in HBM and no combine launch; more splits than that fall
// back to a small combine kernel.  Two CTAs of 103 KB share an SM, so batch x kv_heads x splits >= 2 x SMs keeps
// ~190 KB of K/V in flight per SM.
// The QKV finalize of the current token (bias + RoPE + KV append) runs in the prologue of this kernel, either from the
// split-K fp32 partials of dots_gemm_skinny_bf16 (qkv_partial) or from the bf16 q|k|v row of dots_decode_gemm_qkv (qkv_bf16).
#include "common.h"
#include "ptx.cuh"
#include "mma_sm80.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int DEC_D = 128;
constexpr int DEC_TILE = 16;                 // keys per warp per ring tile
constexpr int DEC_SLICES = 4;                // 16-key slices of a ring tile, one consumer warp each
#ifndef DEC_GROUPS_OVR
#define DEC_GROUPS_OVR 2
#endif
// Consumer groups: group g works on ring tiles g, g + GROUPS, ... with its own running (m, l, O).  One tile costs a warp a ~900-cycle
// DEPENDENT chain (barrier wake -> ldmatrix -> 8 chained HMMAs -> max/shuffles -> exp2 -> 16 HMMAs -> arrive) at 16 % issue
// utilisation, so with one group the kernel streams at one tile per ~0.9 us per SM (4.6 TB/s) whatever the ring depth; two groups
// keep two tiles in progress per SM (profiles/microbench_r2.md: the same traffic without arithmetic takes 21 us, 5.9 TB/s).
constexpr int DEC_GROUPS = DEC_GROUPS_OVR;
constexpr int DEC_WARPS = DEC_SLICES * DEC_GROUPS;                         // consumer warps
constexpr int DEC_THREADS = (DEC_WARPS + 1) * 32;                          // + producer warp
constexpr int DEC_RING_KEYS = DEC_TILE * DEC_SLICES;                       // 64 keys per ring tile
constexpr int DEC_BOX_BYTES = DEC_RING_KEYS * 128;                         // [64 keys][64 dims] bf16 = 8 KB (one swizzle box)
constexpr int DEC_STAGE_BYTES = 4 * DEC_BOX_BYTES;                         // K lo | K hi | V lo | V hi = 32 KB
#ifndef DEC_STAGES_OVR
#define DEC_STAGES_OVR 4
#endif
constexpr int DEC_STAGES = DEC_STAGES_OVR;
// a ring stage must always serve the same consumer group (a group that skipped a phase of a barrier could not tell it from the next)
static_assert(DEC_STAGES % DEC_GROUPS == 0, "ring depth must be a multiple of the number of consumer groups");
constexpr int DEC_SMEM = 1024 /*align*/ + 4096 /*Q*/ + DEC_STAGES * DEC_STAGE_BYTES + 256 /*barriers*/;
constexpr int DEC_MAX_CLUSTER = 4;                                         // splits merged on chip
constexpr int DEC_MERGE_BYTES = 8 * DEC_D * 4 + 8 * 2 * 4;                 // one peer's partial: O [<= 8 heads][128] + (m, l) [<= 8]

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
    const bf16* qkv_bf16;     // [B][(nq + 2 nkv) * 128], bias already added (dots_decode_gemm_qkv); alternative to qkv_partial
    int qkv_splits;
    unsigned long long* trace; // timeline instrumentation (nullptr unless armed)
    int out_tile_rows;        // > 0: `out` is written in the k-block-tiled activation layout with this many rows per tile
    int cluster_merge;        // 1: the n_splits CTAs of a (sequence, kv head) are one cluster and merge through DSMEM
    int fault;                // test-only fault injection (dots_debug_set_fault): 1 = key tile 0, 2 = every other key tile loses its P*V term
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

// byte offset of 16-byte chunk `chunk` (0..15 over the 128 head dims) of key row `row` inside one K or V tile made of
// two 128-B-swizzled TMA boxes ([64 keys][dims 0-63] then [64 keys][dims 64-127])
__device__ __forceinline__ uint32_t dec_tile_off(int row, int chunk) {
    return (uint32_t)((chunk >> 3) * DEC_BOX_BYTES + row * 128 + (((chunk & 7) ^ (row & 7)) << 4));
}

// element offset of out[b][col]: row-major [B, n_q_heads * 128], or (out_tile_rows > 0) the k-block-tiled B operand of o_proj
__device__ __forceinline__ long long dec_out_off(const DecParams& p, int b, int col) {
    return p.out_tile_rows > 0 ? tiled_row_off(b, col, p.out_tile_rows) : (long long)b * p.n_q_heads * DEC_D + col;
}

__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const DecParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;                                        // [STAGES][K lo | K hi | V lo | V hi], 1024-B aligned boxes
    uint8_t* sQ = smem + DEC_STAGES * DEC_STAGE_BYTES;           // 4 KB
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sQ + 4096); // [STAGES]
    uint64_t* empty_bar = full_bar + DEC_STAGES;                 // [STAGES]
    uint8_t* sMerge = sQ + 4096 + 256;                           // leader only: [n_splits - 1] x DEC_MERGE_BYTES (cluster merge)
    const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
    const bool fused = (p.qkv_partial != nullptr) || (p.qkv_bf16 != nullptr);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    // ctx_len / pos were written by the previous step's argmax kernel, which completed before this step's first kernel
    // could even start, so they may be read ahead of the dependency wait.
    const int ctx = p.ctx_len[b];
    int chunk = (ctx + p.n_splits - 1) / p.n_splits;
    chunk = (chunk + DEC_RING_KEYS - 1) / DEC_RING_KEYS * DEC_RING_KEYS;
    const int k_begin = split * chunk;
    const int k_end = min(ctx, k_begin + chunk);
    const int n_tiles = (k_end > k_begin) ? (k_end - k_begin + DEC_RING_KEYS - 1) / DEC_RING_KEYS : 0;
    // The cache stripe of (sequence, kv head) is stored in 64-key tiles of 16 KB that already hold the shared-memory image this
    // kernel wants ([dims 0-63 | dims 64-127][64 keys], 128-B rows swizzled): one 1-D bulk copy per K tile and per V tile.
    const long long stripe = ((long long)b * p.n_kv_heads + kvh) * p.ctx_max * DEC_D;
    const bf16* k_src = p.kc + stripe + (long long)k_begin * DEC_D;           // k_begin is a multiple of 64: tile aligned
    const bf16* v_src = p.vc + stripe + (long long)k_begin * DEC_D;
    // the split whose key range holds key ctx-1, the token appended this step (its cache row is written by this kernel's
    // fused QKV finalize, or by the predecessor kernel on the unfused path)
    const bool holds_new = (ctx - 1 >= k_begin) && (ctx - 1 < k_end);

    if (tid == 0) {
        for (int i = 0; i < DEC_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], DEC_SLICES); }
        fence_barrier_init();
        trace_point(p.trace, 20, 0);
    }
    __syncthreads();

    if (warp == DEC_WARPS) {
        // =============================== TMA producer ===============================
        // The whole warp follows the control flow (bar.sync must be warp-convergent); lane 0 issues the copies.
        // The last tile of the split that holds key ctx-1 is written by this very kernel (fused QKV finalize) or by the
        // predecessor kernel: it waits for the dependency.  Everything older is immutable by now.
        const int early = holds_new ? n_tiles - 1 : n_tiles;
        auto issue = [&](int i) {
            const int st = i % DEC_STAGES;
            uint8_t* dst = ring + st * DEC_STAGE_BYTES;
            mbar_expect_tx(&full_bar[st], DEC_STAGE_BYTES);
            bulk_load(dst, k_src + (long long)i * DEC_RING_KEYS * DEC_D, 2 * DEC_BOX_BYTES, &full_bar[st]);
            bulk_load(dst + 2 * DEC_BOX_BYTES, v_src + (long long)i * DEC_RING_KEYS * DEC_D, 2 * DEC_BOX_BYTES, &full_bar[st]);
        };
        int i = 0;
        if (lane == 0) {
            for (; i < early && i < DEC_STAGES; ++i) issue(i);          // ring-full of immutable tiles ahead of the wait
        }
        __syncwarp();
        pdl_wait();
        if (lane == 0) trace_point(p.trace, 20, 1);
        if (fused && holds_new) {
            // the appended k/v row is produced by the consumer warps of this CTA: wait until they published it
            asm volatile("bar.sync 2, %0;" ::"n"(DEC_THREADS) : "memory");
        }
        if (lane == 0) {
            for (; i < n_tiles; ++i) {
                if (i >= DEC_STAGES) mbar_wait(&empty_bar[i % DEC_STAGES], ((i / DEC_STAGES) & 1) ^ 1);
                issue(i);
            }
        }
        __syncwarp();
        if (p.cluster_merge) cluster_sync_all();       // every thread of the cluster takes part in the merge barrier
        return;
    }

    // =============================== consumer warps ===============================
    // Fused QKV finalize, one (head, 4-column pair) unit per thread.  Everything that does not depend on this layer's QKV GEMM --
    // the bias, the rotary angle and its bf16-rounded cos / sin (the slow accurate cosf / sinf) -- is fetched and computed BEFORE the
    // dependency wait, i.e. while that GEMM is still running; after the wait the partials of all splits are requested at once, so the
    // prologue is one round of L2 latency instead of four.
    const int N = (p.n_q_heads + 2 * p.n_kv_heads) * DEC_D;
    const int n_units = (p.group + 2) * 16;                                // group <= 8: at most 160 units for 256 consumer threads
    const int hl = tid >> 4, c4 = tid & 15;
    const bool has_unit = fused && tid < n_units && (hl < p.group || holds_new);     // k / v rows only in the split that holds the new key
    const int col0 = (hl < p.group ? (kvh * p.group + hl) : (hl == p.group ? p.n_q_heads + kvh : p.n_q_heads + p.n_kv_heads + kvh)) * DEC_D + c4 * 4;
    int posb = 0;
    uint2 b1 = make_uint2(0, 0), b2 = make_uint2(0, 0);
    float cs[4] = {1.f, 1.f, 1.f, 1.f}, sn[4] = {0.f, 0.f, 0.f, 0.f};
    if (has_unit) {
        posb = p.pos[b];                                                    // requires pos[b] == ctx_len[b] - 1
        if (p.qkv_bf16 == nullptr) {
            b1 = *reinterpret_cast<const uint2*>(p.qkv_bias + col0);
            b2 = *reinterpret_cast<const uint2*>(p.qkv_bias + col0 + 64);
        }
        if (hl <= p.group) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float ang = __fmul_rn((float)posb, p.inv_freq[c4 * 4 + j]);
                cs[j] = bf16_round(cosf(ang));
                sn[j] = bf16_round(sinf(ang));
            }
        }
    }
    pdl_wait();

    // Q tile: rows 0..G-1 = the group's q heads, rows G..15 zero
    if (!fused) {
        const bf16* qg = p.q + (long long)b * p.n_q_heads * DEC_D + (long long)kvh * p.group * DEC_D;
        for (int idx = tid; idx < 16 * 16; idx += DEC_WARPS * 32) {
            const int r = idx >> 4, c = idx & 15;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (r < p.group) val = *reinterpret_cast<const uint4*>(qg + r * DEC_D + c * 8);
            *reinterpret_cast<uint4*>(sQ + swz128(r, c)) = val;
        }
    } else {
        // ---- split-K reduce (fixed order) + bias + RoPE; q -> sQ, k/v -> cache row `pos` ----
        for (int idx = tid; idx < (16 - p.group) * 16; idx += DEC_WARPS * 32) {
            const int r = p.group + (idx >> 4), c = idx & 15;
            *reinterpret_cast<uint4*>(sQ + swz128(r, c)) = make_uint4(0, 0, 0, 0);
        }
        if (has_unit) {
            float x1[4], x2[4];
            if (p.qkv_bf16 != nullptr) {
                // q|k|v row of dots_decode_gemm_qkv: reduced, bias added and rounded to bf16 already (HF's Linear output)
                const uint2 a = *reinterpret_cast<const uint2*>(p.qkv_bf16 + (long long)b * N + col0);
                const uint2 d = *reinterpret_cast<const uint2*>(p.qkv_bf16 + (long long)b * N + col0 + 64);
                x1[0] = bf16_lo(a.x); x1[1] = bf16_hi(a.x); x1[2] = bf16_lo(a.y); x1[3] = bf16_hi(a.y);
                x2[0] = bf16_lo(d.x); x2[1] = bf16_hi(d.x); x2[2] = bf16_lo(d.y); x2[3] = bf16_hi(d.y);
            } else {
                const float* src = p.qkv_partial + (long long)b * N + col0;
                const long long sstride = (long long)(gridDim.z) * N;
                x1[0] = x1[1] = x1[2] = x1[3] = 0.f;
                x2[0] = x2[1] = x2[2] = x2[3] = 0.f;
                for (int s0 = 0; s0 < p.qkv_splits; s0 += 8) {             // 8 splits = 16 loads in flight; adds stay in split order
                    float4 a[8], d[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (s0 + w < p.qkv_splits) {
                            a[w] = *reinterpret_cast<const float4*>(src + (s0 + w) * sstride);
                            d[w] = *reinterpret_cast<const float4*>(src + (s0 + w) * sstride + 64);
                        }
#pragma unroll
                    for (int w = 0; w < 8; ++w)
                        if (s0 + w < p.qkv_splits) {
                            x1[0] += a[w].x; x1[1] += a[w].y; x1[2] += a[w].z; x1[3] += a[w].w;
                            x2[0] += d[w].x; x2[1] += d[w].y; x2[2] += d[w].z; x2[3] += d[w].w;
                        }
                }
                x1[0] = bf16_round(x1[0] + bf16_lo(b1.x)); x1[1] = bf16_round(x1[1] + bf16_hi(b1.x));
                x1[2] = bf16_round(x1[2] + bf16_lo(b1.y)); x1[3] = bf16_round(x1[3] + bf16_hi(b1.y));
                x2[0] = bf16_round(x2[0] + bf16_lo(b2.x)); x2[1] = bf16_round(x2[1] + bf16_hi(b2.x));
                x2[2] = bf16_round(x2[2] + bf16_lo(b2.y)); x2[3] = bf16_round(x2[3] + bf16_hi(b2.y));
            }
            float o1[4], o2[4];
            if (hl <= p.group) {
                // HF Qwen2 RoPE rounding points ([Q]:102-146): cos / sin are bf16, every product and the sum round to bf16
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    o1[j] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x1[j], cs[j])), bf16_round(__fmul_rn(-x2[j], sn[j]))));
                    o2[j] = bf16_round(__fadd_rn(bf16_round(__fmul_rn(x2[j], cs[j])), bf16_round(__fmul_rn(x1[j], sn[j]))));
                }
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
                bf16* dst = (hl == p.group ? p.kc_w : p.vc_w) + stripe;
                *reinterpret_cast<uint2*>(dst + kv_tiled_off(posb, c4 * 4)) = r1;
                *reinterpret_cast<uint2*>(dst + kv_tiled_off(posb, 64 + c4 * 4)) = r2;
            }
        }
    }
    if (fused && holds_new) {
        asm volatile("fence.proxy.async;" ::: "memory");        // generic-proxy cache writes -> visible to the TMA (async proxy) reads
        __threadfence();
        asm volatile("bar.sync 2, %0;" ::"n"(DEC_THREADS) : "memory");      // releases the producer's last tile
    }
    asm volatile("bar.sync 1, %0;" ::"n"(DEC_WARPS * 32) : "memory");       // sQ complete (consumer warps only)
    uint32_t qf[8][4];
    {
        const int r = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) ldmatrix_x4(qf[kk], smem_u32(sQ) + swz128(r, kk * 2 + (lane >> 4)));
    }

    // Only rows 0..7 of the 16-row MMA tile carry q heads (group <= 8); rows 8..15 are zero padding, so the softmax works on the
    // first row half of every fragment and feeds P = 0 for the second (their accumulators o[.][2..3] stay 0).  The kernel is bound
    // by the instruction stream of its four consumer warps (ncu: 20 % issue utilisation with one warp per scheduler, ~4 cycles per
    // dependent instruction), so the loop is kept lean: per-lane ldmatrix offsets hoisted, no masking arithmetic on full tiles,
    // the O rescale skipped while the running maximum does not move, no clock reads in the barrier wait.
    float o[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    const int slice = warp % DEC_SLICES, grp = warp / DEC_SLICES;
    uint32_t off_k[8], off_v[8];
    {
        const int r_k = slice * DEC_TILE + (lane & 7) + 8 * (lane >> 4);
        const int r_v = slice * DEC_TILE + (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            off_k[kk] = dec_tile_off(r_k, kk * 2 + ((lane >> 3) & 1));
            off_v[kk] = 2 * DEC_BOX_BYTES + dec_tile_off(r_v, kk * 2 + (lane >> 4));
        }
    }
    const uint32_t ring_u32 = smem_u32(ring);
    const float sc = p.scale_log2;

    if (tid == 0) trace_point(p.trace, 20, 5);          // prologue (QKV finalize, Q fragments) done
    for (int i = grp; i < n_tiles; i += DEC_GROUPS) {
        const int stg = i % DEC_STAGES;
        {
            uint32_t spins = 0;
            while (!mbar_try_wait(&full_bar[stg], (i / DEC_STAGES) & 1)) {
                if (++spins > (1u << 27)) { printf("dots: attn_decode mbarrier watchdog block %d\n", (int)blockIdx.z); __trap(); }
            }
        }
        if (i == 0 && tid == 0) trace_point(p.trace, 20, 2);
        const int key0 = k_begin + i * DEC_RING_KEYS + slice * DEC_TILE;      // this warp's 16 keys of the tile
        if (key0 >= k_end) {                                                   // warp-uniform: nothing of this slice is visible
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stg]);
            continue;
        }
        const uint32_t sb = ring_u32 + stg * DEC_STAGE_BYTES;
        const bool ragged = key0 + DEC_TILE > k_end;                           // warp-uniform: only the last tile of a sequence
        if (ragged) {
            // rows past the last visible key hold whatever the cache stripe contains (possibly NaN bit patterns): zero the V
            // rows so that P = 0 really contributes 0 (K rows are neutralised by the -inf select below)
            for (int idx = lane; idx < DEC_TILE * 16; idx += 32) {
                const int r = idx >> 4, c = idx & 15;
                if (key0 + r >= k_end)
                    *reinterpret_cast<uint4*>(ring + stg * DEC_STAGE_BYTES + 2 * DEC_BOX_BYTES + dec_tile_off(slice * DEC_TILE + r, c)) = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async_smem();       // these generic-proxy writes precede any later bulk refill of the stage
            __syncwarp();
        }

        float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t bfr[4];
            ldmatrix_x4(bfr, sb + off_k[kk]);
            mma_bf16_16816(s[0], qf[kk], bfr[0], bfr[1]);
            mma_bf16_16816(s[1], qf[kk], bfr[2], bfr[3]);
        }
        // row g of the tile: keys 2t, 2t+1 (s[0][0..1]) and 8+2t, 8+2t+1 (s[1][0..1])
        float v0 = s[0][0] * sc, v1 = s[0][1] * sc, v2 = s[1][0] * sc, v3 = s[1][1] * sc;
        if (ragged) {
            const int kj = key0 + 2 * t;
            if (kj >= k_end) v0 = -INFINITY;
            if (kj + 1 >= k_end) v1 = -INFINITY;
            if (kj + 8 >= k_end) v2 = -INFINITY;
            if (kj + 9 >= k_end) v3 = -INFINITY;
        }
        float mx = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        const float m_new = fmaxf(m_run[0], mx);
        const float msafe = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = fast_exp2(m_run[0] - msafe);
        m_run[0] = m_new;
        const float p0 = fast_exp2(v0 - msafe), p1 = fast_exp2(v1 - msafe), p2 = fast_exp2(v2 - msafe), p3 = fast_exp2(v3 - msafe);
        l_run[0] *= alpha;
        l_run[0] += p0 + p1;
        l_run[0] += p2 + p3;
        uint32_t pf[4] = {pack_bf16x2(p0, p1), 0u, pack_bf16x2(p2, p3), 0u};
        if ((p.fault == 1 && i == 0 && split == 0) || (p.fault == 2 && (i & 1) == 0)) pf[0] = pf[2] = 0u;   // test-only fault
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {          // the running maximum moved for some row of this warp: rescale O
#pragma unroll
            for (int j = 0; j < 16; ++j) { o[j][0] *= alpha; o[j][1] *= alpha; }
        }
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
            uint32_t bfr[4];
            ldmatrix_x4_trans(bfr, sb + off_v[dp]);
            mma_bf16_16816(o[2 * dp], pf, bfr[0], bfr[1]);
            mma_bf16_16816(o[2 * dp + 1], pf, bfr[2], bfr[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stg]);           // this warp is done with its slice of the stage
    }
    if (tid == 0) trace_point(p.trace, 20, 3);
    l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 1);
    l_run[0] += __shfl_xor_sync(0xffffffffu, l_run[0], 2);

    // ---- merge the 4 warps (rows g < group only; rows 8..15 are padding) -----------------
    asm volatile("bar.sync 1, %0;" ::"n"(DEC_WARPS * 32) : "memory");   // all consumer warps done with the ring (every TMA tile has landed)
    float* sO = reinterpret_cast<float*>(ring);                         // [4 warps][8 rows][128]
    float* sML = sO + DEC_WARPS * 8 * DEC_D;                            // [warps][8 rows][2]
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
        sO[(warp * 8 + g) * DEC_D + nb * 8 + 2 * t] = o[nb][0];
        sO[(warp * 8 + g) * DEC_D + nb * 8 + 2 * t + 1] = o[nb][1];
    }
    if (t == 0) {
        sML[(warp * 8 + g) * 2] = m_run[0];
        sML[(warp * 8 + g) * 2 + 1] = l_run[0];
    }
    asm volatile("bar.sync 1, %0;" ::"n"(DEC_WARPS * 32) : "memory");
    // every consumer thread finishes: thread (c = tid % 128, part = tid / 128) owns head dim c of the q heads r with r % PARTS == part
    static_assert(DEC_WARPS * 32 % DEC_D == 0 && DEC_WARPS * 8 * (DEC_D + 2) * 4 <= DEC_STAGES * DEC_STAGE_BYTES, "merge scratch lives in the ring");
    constexpr int PARTS = DEC_WARPS * 32 / DEC_D;
    const int c = tid % DEC_D, part = tid / DEC_D;
    const bool finisher = true;
    float accv[8], mv[8], lv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        accv[r] = 0.f; mv[r] = -INFINITY; lv[r] = 0.f;
        if (r % PARTS == part && r < p.group) {
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
            accv[r] = acc; mv[r] = m; lv[r] = l;
        }
    }
    if (p.n_splits == 1) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (r % PARTS == part && r < p.group)
                p.out[dec_out_off(p, b, (kvh * p.group + r) * DEC_D + c)] = __float2bfloat16_rn(lv[r] > 0.f ? accv[r] / lv[r] : 0.f);
    } else if (p.cluster_merge) {
        // ---- on-chip merge across the cluster: peers write (O, m, l) into the leader's shared memory, the leader combines
        //      in split order (same arithmetic as attn_decode_combine_kernel) ----
        if (split != 0 && finisher) {
            const uint32_t base = mapa_shared(smem_u32(sMerge + (size_t)(split - 1) * DEC_MERGE_BYTES), 0);
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r % PARTS == part && r < p.group) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(base + (uint32_t)(r * DEC_D + c) * 4), "f"(accv[r]) : "memory");
            if (c == 0) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (r % PARTS == part && r < p.group) {
                        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(base + (uint32_t)(8 * DEC_D + r * 2) * 4), "f"(mv[r]) : "memory");
                        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(base + (uint32_t)(8 * DEC_D + r * 2 + 1) * 4), "f"(lv[r]) : "memory");
                    }
            }
        }
        cluster_sync_all();
        if (split == 0 && finisher) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r >= p.group || r % PARTS != part) continue;
                float m = mv[r];
                for (int sp = 1; sp < p.n_splits; ++sp)
                    m = fmaxf(m, reinterpret_cast<const float*>(sMerge + (size_t)(sp - 1) * DEC_MERGE_BYTES)[8 * DEC_D + r * 2]);
                float acc = 0.f, l = 0.f;
                {
                    const float w = (mv[r] == -INFINITY) ? 0.f : fast_exp2(mv[r] - m);
                    acc += w * accv[r];
                    l += w * lv[r];
                }
                for (int sp = 1; sp < p.n_splits; ++sp) {
                    const float* pm = reinterpret_cast<const float*>(sMerge + (size_t)(sp - 1) * DEC_MERGE_BYTES);
                    const float ms = pm[8 * DEC_D + r * 2];
                    const float w = (ms == -INFINITY) ? 0.f : fast_exp2(ms - m);
                    acc += w * pm[r * DEC_D + c];
                    l += w * pm[8 * DEC_D + r * 2 + 1];
                }
                p.out[dec_out_off(p, b, (kvh * p.group + r) * DEC_D + c)] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
            }
        }
    } else if (finisher) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r >= p.group || r % PARTS != part) continue;
            const long long pi = ((long long)b * p.n_q_heads + kvh * p.group + r) * p.n_splits + split;
            p.part_o[pi * DEC_D + c] = accv[r];
            if (c == 0) { p.part_ml[pi * 2] = mv[r]; p.part_ml[pi * 2 + 1] = lv[r]; }
        }
    }
    if (tid == 0) trace_point(p.trace, 20, 4);
}

__global__ void __launch_bounds__(DEC_D)
attn_decode_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml, bf16* __restrict__ out,
                           int n_splits, int n_q_heads, int out_tile_rows) {
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
    const int b = (int)(bh / n_q_heads), col = (int)(bh % n_q_heads) * DEC_D + c;
    out[out_tile_rows > 0 ? tiled_row_off(b, col, out_tile_rows) : bh * DEC_D + c] = __float2bfloat16_rn(l > 0.f ? acc / l : 0.f);
}

}  // namespace dots

using namespace dots;

namespace dots {
int g_dec_cluster = 1;        // dots_set_decode_cluster(): merge <= DEC_MAX_CLUSTER key splits on chip (cluster + DSMEM) instead of a combine kernel
int g_debug_fault = 0;        // dots_debug_set_fault(): test-only fault injection
}

static int launch_attn_decode(DecParams& p, int batch, int n_q_heads, int n_kv_heads, int head_dim, int n_splits, float softmax_scale,
                              void* stream, const char* who) {
    DOTS_REQUIRE(head_dim == DEC_D, "%s: head_dim must be 128", who);
    DOTS_REQUIRE(batch > 0 && n_q_heads % n_kv_heads == 0 && n_q_heads / n_kv_heads <= 8,
                 "%s: bad heads %d/%d (group must be <= 8)", who, n_q_heads, n_kv_heads);
    const bool cluster = n_splits > 1 && n_splits <= DEC_MAX_CLUSTER && g_dec_cluster;
    DOTS_REQUIRE(n_splits >= 1 && (n_splits == 1 || cluster || (p.part_o && p.part_ml)), "%s: n_splits > %d needs partial buffers", who,
                 DEC_MAX_CLUSTER);
    p.cluster_merge = cluster ? 1 : 0;
    p.fault = g_debug_fault;
    p.trace = g_trace;
    p.n_q_heads = n_q_heads; p.n_kv_heads = n_kv_heads; p.group = n_q_heads / n_kv_heads; p.n_splits = n_splits;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             DEC_SMEM + (DEC_MAX_CLUSTER - 1) * DEC_MERGE_BYTES));
    }
    DOTS_REQUIRE(p.ctx_max % DEC_RING_KEYS == 0, "%s: ctx_max must be a multiple of %d (the cache is stored in 64-key tiles)", who, DEC_RING_KEYS);
    DOTS_REQUIRE(p.out_tile_rows == 0 || (p.out_tile_rows % 8 == 0 && batch <= p.out_tile_rows), "%s: bad out_tile_rows %d", who, p.out_tile_rows);
    dim3 grid(n_splits, n_kv_heads, batch);
    if (cluster) {
        DOTS_CHECK_CUDA(launch_ex_cluster(attn_decode_kernel, dim3(grid), dim3(DEC_THREADS), (size_t)(DEC_SMEM + (n_splits - 1) * DEC_MERGE_BYTES), st,
                                          true, (unsigned)n_splits, p));
        return 0;
    }
    DOTS_CHECK_CUDA(launch_ex(attn_decode_kernel, dim3(grid), dim3(DEC_THREADS), (size_t)(DEC_SMEM), st, true, p));
    if (n_splits > 1) {
        DOTS_CHECK_CUDA(launch_ex(attn_decode_combine_kernel, dim3(batch * n_q_heads), dim3(DEC_D), (size_t)(0), st, true, p.part_o, p.part_ml, p.out, n_splits, n_q_heads, p.out_tile_rows));
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
                                      void* k_cache, void* v_cache, const int* ctx_len, void* out, int out_tile_rows, float* part_o, float* part_ml,
                                      int batch, int n_q_heads, int n_kv_heads, int head_dim, long long ctx_max, int n_splits,
                                      float softmax_scale, void* stream) {
    DOTS_REQUIRE(qkv_partial && qkv_splits >= 1 && qkv_bias && pos && inv_freq, "dots_attn_decode_fused: missing QKV inputs");
    DecParams p{};
    p.out_tile_rows = out_tile_rows;
    p.q = nullptr; p.kc = (const bf16*)k_cache; p.vc = (const bf16*)v_cache; p.ctx_len = ctx_len;
    p.out = (bf16*)out; p.part_o = part_o; p.part_ml = part_ml; p.ctx_max = ctx_max;
    p.qkv_partial = qkv_partial; p.qkv_splits = qkv_splits; p.qkv_bias = (const bf16*)qkv_bias; p.pos = pos; p.inv_freq = inv_freq;
    p.kc_w = (bf16*)k_cache; p.vc_w = (bf16*)v_cache;
    return launch_attn_decode(p, batch, n_q_heads, n_kv_heads, head_dim, n_splits, softmax_scale, stream, "dots_attn_decode_fused");
}

// Same as dots_attn_decode_fused, with q|k|v of the current token arriving as the bf16 row [batch][(nq + 2 nkv) * 128] that
// dots_decode_gemm_qkv wrote (bias already added): RoPE + KV append + attention.
extern "C" int dots_attn_decode_qkv(const void* qkv_bf16, const int* pos, const float* inv_freq, void* k_cache, void* v_cache, const int* ctx_len,
                                    void* out, int out_tile_rows, float* part_o, float* part_ml, int batch, int n_q_heads, int n_kv_heads,
                                    int head_dim, long long ctx_max, int n_splits, float softmax_scale, void* stream) {
    DOTS_REQUIRE(qkv_bf16 && pos && inv_freq, "dots_attn_decode_qkv: missing QKV inputs");
    DecParams p{};
    p.out_tile_rows = out_tile_rows;
    p.q = nullptr; p.kc = (const bf16*)k_cache; p.vc = (const bf16*)v_cache; p.ctx_len = ctx_len;
    p.out = (bf16*)out; p.part_o = part_o; p.part_ml = part_ml; p.ctx_max = ctx_max;
    p.qkv_bf16 = (const bf16*)qkv_bf16; p.pos = pos; p.inv_freq = inv_freq;
    p.kc_w = (bf16*)k_cache; p.vc_w = (bf16*)v_cache;
    return launch_attn_decode(p, batch, n_q_heads, n_kv_heads, head_dim, n_splits, softmax_scale, stream, "dots_attn_decode_qkv");
}

extern "C" int dots_set_decode_cluster(int enable) {
    dots::g_dec_cluster = enable ? 1 : 0;
    return 0;
}

extern "C" int dots_debug_set_fault(int code) {
    dots::g_debug_fault = code;
    return 0;
}
