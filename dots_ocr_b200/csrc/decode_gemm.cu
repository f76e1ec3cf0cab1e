// Decode-step projections whose output feeds an elementwise stage that needs the FULL reduction: q|k|v (+bias) and
// o_proj / down_proj (+residual, +RMSNorm of the next block).  SURVEY.md §8a rows a15, a16, a20, a21 (decode half);
// reference arithmetic: [Q] modeling_qwen2.py:217-219 (q/k/v Linear), :243 + :302-308 (o_proj, residual, post-attention
// norm), :46-48 + :308 (down_proj, residual), :258-263 (RMSNorm).
//
// With a batch of <= 64 rows these GEMMs are weight-streaming problems (N*K*2 bytes from HBM, a few MFLOP), and the
// output has only N/128 row tiles of the 128-row tensor-core operand: 12 for N = 1536.  To keep every SM streaming, K is
// split across the CTAs of a thread-block CLUSTER (swap-AB: weights are the M operand, the batch is N), each CTA
// accumulates its K slice in TMEM, and the split-K reduction happens on chip:
//
//   TMEM -> registers -> st.shared::cluster into the peer that owns those batch columns (reduce-scatter over distributed
//   shared memory) -> barrier.cluster -> fixed-order sum of the CS partials (deterministic) -> fused epilogue -> HBM.
//
// There are no fp32 partials in HBM/L2 and no separate "finalize" launch.  Epilogues:
//   MODE_QKV      out[b, n] = bf16(acc + bias[n])                                      (HF: nn.Linear with bias)
//   MODE_RESNORM  x = bf16(bf16(acc) + resid[b, n]);  resid[b, n] = x;
//                 normed[b, n] = bf16(bf16(x * rsqrt(mean_n x^2 + eps)) * w[n])         (HF rounding points)
// RMSNorm needs a row statistic over ALL N features, i.e. over every cluster of the grid.  Each CTA publishes the sum of
// squares of its (tile, batch-column) block, bumps a device counter and waits until all CTAs of the grid have done so
// (one L2 round trip; every CTA of the grid is resident: the host checks cudaOccupancyMaxActiveClusters), then normalises
// the x values it still holds in registers.  The counter is zeroed by the first kernel of each decode step.
#include "common.h"
#include "ptx.cuh"
#include "../../include/dots_ocr_b200.h"

namespace dots {

constexpr int DG_BM = 128;
constexpr int DG_BK = 64;
constexpr int DG_UMMA_K = 16;
constexpr int DG_THREADS = 256;        // warp 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 3 idle, 4-7 epilogue
#ifndef DG_STAGES
#define DG_STAGES 4
#endif
constexpr int DG_MODE_QKV = 0;
constexpr int DG_MODE_RESNORM = 1;

struct DgParams {
    const uint8_t* w_tiled;  // weights as [ceil(N/128)][ceil(K/64)] blobs of 16 KB (ops.tile_weight): one bulk copy per ring stage
    const uint8_t* x_tiled;  // activations as [ceil(K/64)] blobs of BN x 128 B (ops.tile_rows with BN rows per tile)
    int N, K, batch;
    int num_k_blocks, kb_per_split;
    const bf16* bias;        // QKV
    bf16* out;               // QKV: [batch, ldo]
    long long ldo;
    bf16* resid;             // RESNORM: [batch, N] in/out
    const bf16* ln_w;        // RESNORM: weight of the RMSNorm that follows
    bf16* normed;            // RESNORM: normalised activations in the k-block-tiled layout, BN rows per tile (next GEMM's B operand)
    float* stats;            // RESNORM: [n_tiles][64] sums of squares
    unsigned* counter;       // RESNORM: CTAs of this launch that have published their statistics
    float eps;
    unsigned long long* trace;   // timeline instrumentation (nullptr unless armed)
};

template <int BN, int CS>
struct DgSmem {
    static_assert(BN % CS == 0 && (BN / CS) % 4 == 0, "each cluster rank owns a multiple of 4 batch columns");
    static constexpr int COLS = BN / CS;                         // batch columns owned by one rank
    static constexpr int A_BYTES = DG_BM * DG_BK * 2;            // 16 KB of weights
    static constexpr int B_BYTES = BN * DG_BK * 2;               // activations
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int RING_BYTES = DG_STAGES * STAGE_BYTES;
    static constexpr int RECV_BYTES = CS * DG_BM * COLS * 4;     // [source rank][feature][COLS] fp32
    static constexpr int BAR_BYTES = 1024;
    static constexpr int TOTAL = RING_BYTES + RECV_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
    static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BN, int CS, int MODE>
__global__ void __launch_bounds__(DG_THREADS, 1)
decode_gemm_cluster_kernel(const DgParams p) {
    using S = DgSmem<BN, CS>;
    constexpr int COLS = S::COLS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + DG_STAGES * S::A_BYTES;
    float* recv = reinterpret_cast<float*>(smem + S::RING_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::RING_BYTES + S::RECV_BYTES);
    uint64_t* full_bar = bars;                      // [STAGES]
    uint64_t* empty_bar = bars + DG_STAGES;         // [STAGES]
    uint64_t* tmem_full = bars + 2 * DG_STAGES;     // [1]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* s_red = reinterpret_cast<float*>(tmem_ptr + 4);          // [4 warps][COLS] + [COLS] rstd

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();        // == blockIdx.x: the K split of this CTA and the batch columns it finalises
    const int tile = blockIdx.y;                    // 128 output features
    const int kb0 = rank * p.kb_per_split;
    const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_split);
    const int nk = max(0, kb1 - kb0);               // an empty split contributes zeros

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < DG_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, S::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
    pdl_launch_dependents();
    constexpr int TRACE_KID = 40 + MODE;
    if (threadIdx.x == 0) trace_point(p.trace, TRACE_KID, 0);

    float acc[BN];          // epilogue warps: this thread's feature row, all batch columns (partial over this CTA's K slice)
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            // weights never depend on a predecessor kernel: the first ring-full is requested before the dependency wait
            const uint8_t* wsrc = p.w_tiled + ((size_t)tile * p.num_k_blocks + kb0) * S::A_BYTES;      // this CTA's k-blocks are contiguous
            const uint8_t* xsrc = p.x_tiled + (size_t)kb0 * S::B_BYTES;
            const int pre = min(nk, DG_STAGES);
            for (int i = 0; i < pre; ++i) {
                mbar_expect_tx(&full_bar[i], S::STAGE_BYTES);
                bulk_load(smem_a + i * S::A_BYTES, wsrc + (size_t)i * S::A_BYTES, S::A_BYTES, &full_bar[i]);
            }
            pdl_wait();
            trace_point(p.trace, TRACE_KID, 1);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < nk; ++i) {
                if (i >= pre) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                    bulk_load(smem_a + stage * S::A_BYTES, wsrc + (size_t)i * S::A_BYTES, S::A_BYTES, &full_bar[stage]);
                }
                bulk_load(smem_b + stage * S::B_BYTES, xsrc + (size_t)i * S::B_BYTES, S::B_BYTES, &full_bar[stage]);
                if (++stage == DG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp convergent, one elected lane issues) =====================
        const bool leader = elect_one();
        constexpr uint32_t idesc = umma_idesc_bf16(DG_BM, BN);
        const uint64_t da0 = umma_desc_k_sw128(smem_u32(smem_a));
        const uint64_t db0 = umma_desc_k_sw128(smem_u32(smem_b));
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < nk; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = da0 + (uint64_t)(stage * (S::A_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)(stage * (S::B_BYTES >> 4));
            if (leader) {
#pragma unroll
                for (int k = 0; k < DG_BK / DG_UMMA_K; ++k) umma_bf16_ss(tmem_base, da + 2 * k, db + 2 * k, idesc, (i > 0 || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (i + 1 == nk) umma_commit(tmem_full);
            }
            __syncwarp();
            if (++stage == DG_STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== epilogue, part 1: TMEM -> registers -> owning peer (reduce-scatter over DSMEM) =====================
        const int f = (warp - 4) * 32 + lane;                   // feature row of the tile == TMEM lane
        if (nk > 0) {
            mbar_wait(tmem_full, 0);
            tc_fence_after();
            if (threadIdx.x == 128) trace_point(p.trace, TRACE_KID, 3);
            const uint32_t t_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(t_row + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c * 32 + j] = __uint_as_float(v[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < BN; ++j) acc[j] = 0.f;
        }
        const uint32_t my_slot = smem_u32(recv + ((size_t)rank * DG_BM + f) * COLS);
#pragma unroll
        for (int d = 0; d < CS; ++d) {
            const uint32_t dst = mapa_shared(my_slot, (uint32_t)d);
#pragma unroll
            for (int q = 0; q < COLS / 4; ++q)
                st_cluster_v4(dst + q * 16, acc[d * COLS + 4 * q], acc[d * COLS + 4 * q + 1], acc[d * COLS + 4 * q + 2], acc[d * COLS + 4 * q + 3]);
        }
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();             // every partial of this CTA's batch columns has landed in `recv`
    if (threadIdx.x == 128) trace_point(p.trace, TRACE_KID, 5);

    if (warp >= 4) {
        // ===================== epilogue, part 2: fixed-order reduction + fused elementwise stage =====================
        pdl_wait();                 // residual reads and every output write follow the dependency
        const int f = (warp - 4) * 32 + lane;
        const int feat = tile * DG_BM + f;
        const bool feat_ok = feat < p.N;
        float s[COLS];
#pragma unroll
        for (int j = 0; j < COLS; ++j) s[j] = 0.f;
#pragma unroll
        for (int src = 0; src < CS; ++src) {                  // split order: deterministic
            const float4* r4 = reinterpret_cast<const float4*>(recv + ((size_t)src * DG_BM + f) * COLS);
#pragma unroll
            for (int q = 0; q < COLS / 4; ++q) {
                const float4 t = r4[q];
                s[4 * q] += t.x; s[4 * q + 1] += t.y; s[4 * q + 2] += t.z; s[4 * q + 3] += t.w;
            }
        }
        const int b0 = rank * COLS;
        if constexpr (MODE == DG_MODE_QKV) {
            const float bias_v = (p.bias != nullptr && feat_ok) ? __bfloat162float(p.bias[feat]) : 0.f;
#pragma unroll
            for (int j = 0; j < COLS; ++j)
                if (feat_ok && b0 + j < p.batch) p.out[(long long)(b0 + j) * p.ldo + feat] = __float2bfloat16_rn(s[j] + bias_v);
        } else {
            float x[COLS], ssq[COLS];
#pragma unroll
            for (int j = 0; j < COLS; ++j) {
                x[j] = 0.f;
                if (feat_ok && b0 + j < p.batch) {
                    const long long o = (long long)(b0 + j) * p.N + feat;
                    x[j] = bf16_round(bf16_round(s[j]) + __bfloat162float(p.resid[o]));
                    p.resid[o] = __float2bfloat16_rn(x[j]);
                }
                ssq[j] = x[j] * x[j];
            }
#pragma unroll
            for (int j = 0; j < COLS; ++j) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], o);
                if (lane == 0) s_red[(warp - 4) * COLS + j] = ssq[j];
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int et = threadIdx.x - 128;                   // 0..127 within the epilogue warps
            if (et < COLS) {
                const float tot = (s_red[et] + s_red[COLS + et]) + (s_red[2 * COLS + et] + s_red[3 * COLS + et]);
                p.stats[tile * 64 + b0 + et] = tot;
                __threadfence();
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) {
                atomicAdd(p.counter, 1u);
                const unsigned target = gridDim.x * gridDim.y;
                const long long t0 = clock64();
                while (*reinterpret_cast<volatile unsigned*>(p.counter) < target) {
                    if (clock64() - t0 > 4000000000LL) {       // ~2 s: a scheduling bug becomes a launch error, not a hung GPU
                        printf("dots: decode_gemm grid wait watchdog (tile %d rank %d)\n", tile, rank);
                        __trap();
                    }
                }
                __threadfence();
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et < COLS) {
                float tot = 0.f;
                for (int t = 0; t < (int)gridDim.y; ++t) tot += __ldcg(p.stats + t * 64 + b0 + et);     // tile order: deterministic
                s_red[4 * COLS + et] = rsqrtf(tot / (float)p.N + p.eps);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const float w = feat_ok ? __bfloat162float(p.ln_w[feat]) : 0.f;
#pragma unroll
            for (int j = 0; j < COLS; ++j)
                if (feat_ok && b0 + j < p.batch)
                    p.normed[tiled_row_off(b0 + j, feat, BN)] = __float2bfloat16_rn(bf16_round(x[j] * s_red[4 * COLS + j]) * w);
        }
    }

    if (threadIdx.x == 128) trace_point(p.trace, TRACE_KID, 4);
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, S::TMEM_COLS);
    }
}

template <int BN, int CS, int MODE>
static int launch_decode_gemm(const DgParams& p, int n_tiles, cudaStream_t st) {
    using S = DgSmem<BN, CS>;
    auto kern = decode_gemm_cluster_kernel<BN, CS, MODE>;
    static bool configured[64] = {false};
    if (first_use_on_device(configured)) {
        DOTS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    }
    DgParams pt = p;
    pt.trace = g_trace;
    DOTS_CHECK_CUDA(launch_ex_cluster(kern, dim3(CS, n_tiles), dim3(DG_THREADS), (size_t)S::TOTAL, st, true, (unsigned)CS, pt));
    return 0;
}

template <int BN, int CS>
static int max_clusters_resnorm(int* out) {
    using S = DgSmem<BN, CS>;
    auto kern = decode_gemm_cluster_kernel<BN, CS, DG_MODE_RESNORM>;
    DOTS_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS, 64); cfg.blockDim = dim3(DG_THREADS); cfg.dynamicSmemBytes = S::TOTAL;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    DOTS_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    *out = n;
    return 0;
}

constexpr int DG_CS = 8;       // portable cluster size; two clusters of 8 fit one GPC (>= 16 SMs) -> >= 16 co-resident clusters

static int prep(DgParams& p, const void* Xt, const void* Wt, int batch, int N, int K, int& bn, const char* who) {
    DOTS_REQUIRE(batch > 0 && batch <= 64 && N > 0 && K > 0, "%s: batch must be 1..64 (got %d), N=%d K=%d", who, batch, N, K);
    DOTS_REQUIRE((reinterpret_cast<uintptr_t>(Xt) | reinterpret_cast<uintptr_t>(Wt)) % 16 == 0, "%s: tiled operands must be 16-byte aligned", who);
    p.w_tiled = reinterpret_cast<const uint8_t*>(Wt);
    p.x_tiled = reinterpret_cast<const uint8_t*>(Xt);
    p.N = N; p.K = K; p.batch = batch;
    p.num_k_blocks = (K + DG_BK - 1) / DG_BK;
    p.kb_per_split = (p.num_k_blocks + DG_CS - 1) / DG_CS;
    bn = batch <= 32 ? 32 : 64;
    return 0;
}

}  // namespace dots

using namespace dots;

// Operand layouts of both entry points: Wt = ops.tile_weight(W) ([ceil(N/128)][ceil(K/64)] blobs of 16 KB), Xt = activations in the
// k-block-tiled layout with 32 (batch <= 32) or 64 rows per tile (ops.tile_rows / tiled_row_off); both are fetched with 1-D bulk copies.
extern "C" int dots_decode_gemm_qkv(const void* Xt, const void* Wt, const void* bias, void* out, long long ldo, int batch, int N, int K,
                                    void* stream) {
    DOTS_REQUIRE(Xt && Wt && out, "dots_decode_gemm_qkv: null pointer");
    DgParams p{};
    int bn = 0;
    if (int rc = prep(p, Xt, Wt, batch, N, K, bn, "dots_decode_gemm_qkv")) return rc;
    p.bias = reinterpret_cast<const bf16*>(bias);
    p.out = reinterpret_cast<bf16*>(out);
    p.ldo = ldo;
    const int n_tiles = (N + DG_BM - 1) / DG_BM;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return bn == 32 ? launch_decode_gemm<32, DG_CS, DG_MODE_QKV>(p, n_tiles, st) : launch_decode_gemm<64, DG_CS, DG_MODE_QKV>(p, n_tiles, st);
}

extern "C" int dots_decode_gemm_resnorm(const void* Xt, const void* Wt, void* resid, const void* ln_w, void* normed_t, float* stats,
                                        unsigned int* counter, int batch, int N, int K, float eps, void* stream) {
    DOTS_REQUIRE(Xt && Wt && resid && ln_w && normed_t && stats && counter, "dots_decode_gemm_resnorm: null pointer");
    DOTS_REQUIRE(N % 64 == 0, "dots_decode_gemm_resnorm: N must be a multiple of 64 (tiled output)");
    DgParams p{};
    int bn = 0;
    if (int rc = prep(p, Xt, Wt, batch, N, K, bn, "dots_decode_gemm_resnorm")) return rc;
    p.resid = reinterpret_cast<bf16*>(resid);
    p.ln_w = reinterpret_cast<const bf16*>(ln_w);
    p.normed = reinterpret_cast<bf16*>(normed_t);
    p.stats = stats; p.counter = counter; p.eps = eps;
    const int n_tiles = (N + DG_BM - 1) / DG_BM;
    // every cluster of the launch must be resident at once (the epilogue waits on a device-wide counter)
    static int max_cl[2][64] = {{0}};
    int dev = 0;
    DOTS_CHECK_CUDA(cudaGetDevice(&dev));
    int& cached = max_cl[bn == 32 ? 0 : 1][dev & 63];
    if (cached == 0) {
        int n = 0;
        if (int rc = (bn == 32 ? max_clusters_resnorm<32, DG_CS>(&n) : max_clusters_resnorm<64, DG_CS>(&n))) return rc;
        cached = n > 0 ? n : -1;
    }
    DOTS_REQUIRE(cached >= n_tiles, "dots_decode_gemm_resnorm: %d row tiles need %d co-resident clusters of %d, the device holds %d", n_tiles,
                 n_tiles, DG_CS, cached);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return bn == 32 ? launch_decode_gemm<32, DG_CS, DG_MODE_RESNORM>(p, n_tiles, st) : launch_decode_gemm<64, DG_CS, DG_MODE_RESNORM>(p, n_tiles, st);
}

extern "C" int dots_decode_gemm_max_clusters(int batch, int* out) {
    DOTS_REQUIRE(out && batch > 0 && batch <= 64, "dots_decode_gemm_max_clusters: batch must be 1..64");
    return batch <= 32 ? max_clusters_resnorm<32, DG_CS>(out) : max_clusters_resnorm<64, DG_CS>(out);
}
