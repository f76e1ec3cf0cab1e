// Hardware fundamentals that steer the decode-step design.  Stand-alone: compiled and run on the GPU box,
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/microbench tools/microbench.cu -lcuda && /tmp/microbench
// Prints one JSON object per measurement:
//   stream      HBM -> shared memory through TMA (cp.async.bulk 1-D) as a function of bytes in flight per SM and CTAs per SM
//   pdl_chain   time per kernel of a chain of dependent kernels (store -> flush -> dependent load) with / without
//               programmatic dependent launch, inside a CUDA graph: the cost of one kernel boundary
//   grid_bar    one device-wide barrier (atomic counter + polling) inside a persistent kernel
//   flag_hop    release-store by one CTA -> acquire-load seen by another CTA (cross-SM signalling latency through L2)
//   dsmem_red   cluster of 8: every CTA writes a 4 KB slice into each peer (reduce-scatter pattern) + cluster barrier
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ stream
// Each CTA streams its contiguous slab of `bytes_per_cta` through a ring of `stages` x `box` bytes; one producer lane, one
// consumer warp that only touches one word per box (the point is the memory system, not the consumer).
__global__ void stream_kernel(const uint8_t* __restrict__ src, size_t bytes_per_cta, int stages, int box, unsigned* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 32;
    uint8_t* ring = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* base = src + (size_t)blockIdx.x * bytes_per_cta;
    const int n = (int)(bytes_per_cta / box);
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < n; ++i) {
                const int st = i % stages;
                if (i >= stages) mbar_wait(&empty[st], ((i / stages) & 1) ^ 1);
                mbar_expect_tx(&full[st], box);
                bulk_load(ring + (size_t)st * box, base + (size_t)i * box, box, &full[st]);
            }
        }
    } else if (warp == 1) {
        unsigned acc = 0;
        for (int i = 0; i < n; ++i) {
            const int st = i % stages;
            mbar_wait(&full[st], (i / stages) & 1);
            acc += *reinterpret_cast<const volatile unsigned*>(ring + (size_t)st * box + lane * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ stream through 2-D tensor maps
// Same ring, but every box is a cp.async.bulk.tensor.2d of [box_rows x 64 bf16] (128-byte rows, 128-B swizzle) -- the shape every
// weight / KV tile of the decode kernels has.  pitch_elems = 64: the box rows are contiguous in memory (16 KB blob);
// pitch_elems = 1536: rows 3072 B apart (a K-major weight matrix).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__global__ void stream2d_kernel(const __grid_constant__ CUtensorMap tm, int boxes_per_cta, int stages, int box_rows, int kblocks, unsigned* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 32;
    uint8_t* ring = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int box = box_rows * 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < boxes_per_cta; ++i) {
                const int st = i % stages;
                if (i >= stages) mbar_wait(&empty[st], ((i / stages) & 1) ^ 1);
                mbar_expect_tx(&full[st], box);
                const long long g = (long long)blockIdx.x * boxes_per_cta + i;        // global box index: row tile = g / kblocks, k-block = g % kblocks
                tma_load_2d(ring + (size_t)st * box, &tm, (int)(g % kblocks) * 64, (int)(g / kblocks) * box_rows, &full[st]);
            }
        }
    } else if (warp == 1) {
        unsigned acc = 0;
        for (int i = 0; i < boxes_per_cta; ++i) {
            const int st = i % stages;
            mbar_wait(&full[st], (i / stages) & 1);
            acc += *reinterpret_cast<const volatile unsigned*>(ring + (size_t)st * box + lane * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ KV-attention access pattern
// The decode attention kernel's traffic without its arithmetic: CTA i streams the K stripe and the V stripe of one (sequence, kv
// head) -- two sequential streams `stride` bytes apart from its neighbours', 16 KB per tile each -- through a ring whose stage holds
// one K tile + one V tile; a consumer warp waits `spin` clocks per stage before releasing it (0 = memory system only).
__global__ void kv_stream_kernel(const uint8_t* __restrict__ kbase, const uint8_t* __restrict__ vbase, size_t stride, int tiles, int stages,
                                 int spin, unsigned* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 32;
    uint8_t* ring = smem + 1024;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* ks = kbase + (size_t)blockIdx.x * stride;
    const uint8_t* vs = vbase + (size_t)blockIdx.x * stride;
    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < tiles; ++i) {
                const int st = i % stages;
                if (i >= stages) mbar_wait(&empty[st], ((i / stages) & 1) ^ 1);
                mbar_expect_tx(&full[st], 32768);
                bulk_load(ring + (size_t)st * 32768, ks + (size_t)i * 16384, 16384, &full[st]);
                bulk_load(ring + (size_t)st * 32768 + 16384, vs + (size_t)i * 16384, 16384, &full[st]);
            }
        }
    } else if (warp == 1) {
        unsigned acc = 0;
        for (int i = 0; i < tiles; ++i) {
            const int st = i % stages;
            mbar_wait(&full[st], (i / stages) & 1);
            acc += *reinterpret_cast<const volatile unsigned*>(ring + (size_t)st * 32768 + lane * 4);
            if (spin > 0) { const long long t0 = clock64(); while (clock64() - t0 < spin) {} }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[st]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// ------------------------------------------------------------------------------------------------ pdl chain
// Kernel i: wait for kernel i-1, read what it wrote (148 x 128 floats), add 1, write.  A real producer/consumer boundary.
__global__ void chain_kernel(const float* __restrict__ in, float* __restrict__ out, int use_pdl) {
    if (use_pdl) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    out[i] = in[i] + 1.0f;
}

// ------------------------------------------------------------------------------------------------ grid barrier
__global__ void grid_bar_kernel(unsigned* counter, int iters, float* data) {
    const unsigned n = gridDim.x;
    for (int it = 0; it < iters; ++it) {
        data[blockIdx.x * blockDim.x + threadIdx.x] += 1.0f;            // a store the barrier has to publish
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(counter, 1u);
            const unsigned target = (unsigned)(it + 1) * n;
            while (*reinterpret_cast<volatile unsigned*>(counter) < target) {}
            __threadfence();
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ flag hop (ping-pong)
__global__ void flag_hop_kernel(unsigned* flags, int iters, long long* cycles) {
    // CTA 0 and CTA 1 (different SMs) bounce a counter: each hop = st.release -> ld.acquire by the other SM
    unsigned* mine = flags + blockIdx.x * 32;
    unsigned* other = flags + (1 - blockIdx.x) * 32;
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    for (int it = 1; it <= iters; ++it) {
        if (blockIdx.x == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(mine), "r"((unsigned)it) : "memory");
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(other) : "memory"); } while (v < (unsigned)it);
        } else {
            unsigned v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(other) : "memory"); } while (v < (unsigned)it);
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(mine), "r"((unsigned)it) : "memory");
        }
    }
    if (blockIdx.x == 0) cycles[0] = clock64() - t0;
}

// ------------------------------------------------------------------------------------------------ DSMEM reduce-scatter
__global__ void __cluster_dims__(8, 1, 1) dsmem_kernel(int iters, float* out) {
    __shared__ __align__(16) float buf[8][1024];        // [source rank][4 KB]
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    float4 v = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        // every thread (256) writes 16 B into its slot of each peer's buf[rank]
        for (unsigned p = 0; p < 8; ++p) {
            uint32_t local = smem_u32(&buf[rank][threadIdx.x * 4]);
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(p));
            asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
        for (int s = 0; s < 8; ++s) acc += buf[s][threadIdx.x * 4];
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static float time_graph(cudaGraphExec_t g, cudaStream_t st, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) CK(cudaGraphLaunch(g, st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < reps; ++i) CK(cudaGraphLaunch(g, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char** argv) {
    const bool quick = argc > 1;       // any argument: skip the long 1-D sweep
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    // ---- stream: 8 GB buffer (>> L2), each config reads ~4 GB
    if (!quick) {
        const size_t total = (size_t)8 << 30;
        uint8_t* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
        unsigned* sink; CK(cudaMalloc(&sink, 4));
        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        const int boxes[] = {8192, 16384, 32768};
        for (int per_sm = 1; per_sm <= 4; ++per_sm) {
            for (int box : boxes) {
                for (int stages = 2; stages <= 24; stages += (stages < 8 ? 1 : 4)) {
                    const size_t smem = 1024 + (size_t)stages * box;
                    if (smem * per_sm > 220 * 1024) continue;
                    const int grid = sms * per_sm;
                    size_t per_cta = ((size_t)4 << 30) / grid / box * box;
                    // pad the request so that no more than per_sm CTAs fit one SM (even spread over the SMs)
                    size_t smem_launch = smem;
                    const size_t floor_ = (size_t)(227 * 1024) / (per_sm + 1) + 1024;
                    if (smem_launch < floor_) smem_launch = floor_;
                    if (smem_launch > 220 * 1024) smem_launch = 220 * 1024;
                    stream_kernel<<<grid, 64, smem_launch, st>>>(src, per_cta, stages, box, sink);
                    CK(cudaEventRecord(e0, st));
                    stream_kernel<<<grid, 64, smem_launch, st>>>(src, per_cta, stages, box, sink);
                    CK(cudaEventRecord(e1, st));
                    CK(cudaStreamSynchronize(st));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    printf("{\"k\": \"stream\", \"ctas_per_sm\": %d, \"box\": %d, \"stages\": %d, \"inflight_kb_per_sm\": %d, \"gbs\": %.0f}\n", per_sm, box,
                           stages, (int)((size_t)stages * box * per_sm / 1024), (double)per_cta * grid / ms / 1e6);
                    fflush(stdout);
                }
            }
        }
        // fewer SMs at a deep ring: how many SMs saturate HBM?
        for (int g : {32, 64, 96, 128, 148}) {
            const int box = 16384, stages = 12;
            size_t per_cta = ((size_t)2 << 30) / g / box * box;
            stream_kernel<<<g, 64, 1024 + stages * box, st>>>(src, per_cta, stages, box, sink);
            CK(cudaEventRecord(e0, st));
            stream_kernel<<<g, 64, 1024 + stages * box, st>>>(src, per_cta, stages, box, sink);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("{\"k\": \"stream_sms\", \"ctas\": %d, \"box\": %d, \"stages\": %d, \"gbs\": %.0f}\n", g, box, stages, (double)per_cta * g / ms / 1e6);
        }
        CK(cudaFree(src)); CK(cudaFree(sink));
    }

    // ---- decode-attention access pattern: 128 (and 148, 256) CTAs x (K stripe + V stripe), 30 tiles of 16 KB each per stripe
    {
        const size_t stride = (size_t)2240 * 256;                 // one (sequence, kv head) stripe: ctx_max 2240 keys x 256 B
        const size_t layer = stride * 256;                        // room for up to 256 stripes
        const int L = 24;                                         // distinct "layers" so that nothing is served from L2
        uint8_t *kb, *vb; CK(cudaMalloc(&kb, layer * L)); CK(cudaMalloc(&vb, layer * L));
        CK(cudaMemset(kb, 1, layer * L)); CK(cudaMemset(vb, 2, layer * L));
        unsigned* sink; CK(cudaMalloc(&sink, 4));
        CK(cudaFuncSetAttribute(kv_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        for (int ctas : {128, 148, 256}) {
            for (int stages : {3, 4, 6}) {
                for (int spin : {0, 800, 1600}) {
                    size_t smem = 1024 + (size_t)stages * 32768;
                    if (ctas > 148 && smem * 2 > 224 * 1024) continue;
                    kv_stream_kernel<<<ctas, 64, smem, st>>>(kb, vb, stride, 30, stages, spin, sink);
                    CK(cudaEventRecord(e0, st));
                    for (int l = 1; l < L; ++l) kv_stream_kernel<<<ctas, 64, smem, st>>>(kb + layer * l, vb + layer * l, stride, 30, stages, spin, sink);
                    CK(cudaEventRecord(e1, st));
                    CK(cudaStreamSynchronize(st));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    const double bytes = (double)ctas * 30 * 32768 * (L - 1);
                    printf("{\"k\": \"kv_stream\", \"ctas\": %d, \"stages\": %d, \"consumer_spin_clk\": %d, \"us_per_launch\": %.2f, \"gbs\": %.0f}\n", ctas, stages, spin,
                           ms * 1e3 / (L - 1), bytes / ms / 1e6);
                    fflush(stdout);
                }
            }
        }
        CK(cudaFree(kb)); CK(cudaFree(vb)); CK(cudaFree(sink));
    }
    if (argc > 2) return 0;                                       // "quick kv": only the KV pattern

    // ---- 2-D tensor-map streaming: [rows, pitch] bf16, boxes of box_rows x 64 elements (128-B rows)
    {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fp = nullptr; cudaDriverEntryPointQueryResult q;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
        EncodeFn enc = (EncodeFn)fp;
        const size_t total = (size_t)6 << 30;
        uint8_t* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
        unsigned* sink; CK(cudaMalloc(&sink, 4));
        CK(cudaFuncSetAttribute(stream2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        const int pitches[] = {64, 1536, 8960};
        for (int pitch : pitches) {
            for (int box_rows : {64, 128}) {
                for (int swz = 0; swz < 2; ++swz) {
                    const cuuint64_t cols = pitch, rows = total / 2 / pitch;
                    cuuint64_t gdim[2] = {cols, rows}; cuuint64_t gstr[1] = {(cuuint64_t)pitch * 2};
                    cuuint32_t bx[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
                    CUtensorMap tm;
                    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("{\"k\": \"stream2d\", \"error\": %d}\n", (int)r); continue; }
                    const int kblocks = pitch / 64;
                    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
                        for (int stages : {4, 8}) {
                            const int box = box_rows * 128;
                            size_t smem = 1024 + (size_t)stages * box;
                            const size_t floor_ = (size_t)(227 * 1024) / (per_sm + 1) + 1024;
                            if (smem < floor_) smem = floor_;
                            if (smem * per_sm > 224 * 1024) continue;
                            const int grid = sms * per_sm;
                            const int boxes_per_cta = (int)(((size_t)3 << 30) / grid / box);
                            stream2d_kernel<<<grid, 64, smem, st>>>(tm, boxes_per_cta, stages, box_rows, kblocks, sink);
                            CK(cudaEventRecord(e0, st));
                            stream2d_kernel<<<grid, 64, smem, st>>>(tm, boxes_per_cta, stages, box_rows, kblocks, sink);
                            CK(cudaEventRecord(e1, st));
                            CK(cudaStreamSynchronize(st));
                            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                            printf("{\"k\": \"stream2d\", \"pitch_elems\": %d, \"box_rows\": %d, \"swizzle128\": %d, \"ctas_per_sm\": %d, \"stages\": %d, \"gbs\": %.0f}\n",
                                   pitch, box_rows, swz, per_sm, stages, (double)boxes_per_cta * box * grid / ms / 1e6);
                            fflush(stdout);
                        }
                    }
                }
            }
        }
        CK(cudaFree(src)); CK(cudaFree(sink));
    }

    // ---- pdl chain: 200 dependent kernels in a graph
    for (int use_pdl = 0; use_pdl <= 1; ++use_pdl) {
        float *a, *b; CK(cudaMalloc(&a, 148 * 128 * 4)); CK(cudaMalloc(&b, 148 * 128 * 4));
        CK(cudaMemset(a, 0, 148 * 128 * 4));
        const int n = 200;
        cudaGraph_t graph; cudaGraphExec_t exec;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int i = 0; i < n; ++i) {
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.stream = st;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
            CK(cudaLaunchKernelEx(&cfg, chain_kernel, (const float*)((i & 1) ? b : a), (float*)((i & 1) ? a : b), use_pdl));
        }
        CK(cudaStreamEndCapture(st, &graph));
        CK(cudaGraphInstantiate(&exec, graph, 0));
        const float ms = time_graph(exec, st, 20);
        printf("{\"k\": \"pdl_chain\", \"pdl\": %d, \"us_per_kernel\": %.3f}\n", use_pdl, ms * 1e3 / n);
        CK(cudaGraphExecDestroy(exec)); CK(cudaGraphDestroy(graph)); CK(cudaFree(a)); CK(cudaFree(b));
    }

    // ---- grid barrier
    {
        unsigned* ctr; float* data; CK(cudaMalloc(&ctr, 4)); CK(cudaMalloc(&data, sms * 128 * 4));
        CK(cudaMemset(data, 0, sms * 128 * 4));
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemsetAsync(ctr, 0, 4, st));
            CK(cudaEventRecord(e0, st));
            grid_bar_kernel<<<sms, 128, 0, st>>>(ctr, iters, data);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"k\": \"grid_bar\", \"ctas\": %d, \"us_per_barrier\": %.3f}\n", sms, ms * 1e3 / iters);
        CK(cudaFree(ctr)); CK(cudaFree(data));
    }

    // ---- flag hop
    {
        unsigned* flags; long long* cyc; CK(cudaMalloc(&flags, 256)); CK(cudaMalloc(&cyc, 8));
        CK(cudaMemset(flags, 0, 256));
        const int iters = 2000;
        CK(cudaEventRecord(e0, st));
        flag_hop_kernel<<<2, 32, 0, st>>>(flags, iters, cyc);
        CK(cudaEventRecord(e1, st));
        CK(cudaStreamSynchronize(st));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"k\": \"flag_hop\", \"us_per_one_way_hop\": %.3f}\n", ms * 1e3 / iters / 2);
        CK(cudaFree(flags)); CK(cudaFree(cyc));
    }

    // ---- DSMEM reduce-scatter in clusters of 8
    {
        float* out; CK(cudaMalloc(&out, 144 * 256 * 4));
        const int iters = 1000;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0, st));
            dsmem_kernel<<<144, 256, 0, st>>>(iters, out);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
        }
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("{\"k\": \"dsmem_reduce_scatter\", \"cluster\": 8, \"bytes_per_cta\": %d, \"us_per_round\": %.3f}\n", 8 * 4096, ms * 1e3 / iters);
        CK(cudaFree(out));
    }
    return 0;
}
