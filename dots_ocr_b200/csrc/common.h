// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace dots {

typedef __nv_bfloat16 bf16;

// Error plumbing: every C-ABI entry returns 0 or a negative code; text via dots_last_error().
void set_error(const char* fmt, ...);
// SMs that the next launches may use: the device's count, or the size of the SM partition the caller is feeding
// (dots_set_sm_count; persistent kernels size their grids with it).
int num_sms();
extern int g_sm_override[64];
// Per-device one-time setup (cudaFuncSetAttribute is per device): true the first time it is called on the current device
// with this flag array.
bool first_use_on_device(bool (&flags)[64]);

#define DOTS_CHECK_CUDA(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            dots::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return -2;                                                                     \
        }                                                                                  \
    } while (0)

#define DOTS_REQUIRE(cond, ...)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            dots::set_error(__VA_ARGS__);                                                  \
            return -1;                                                                     \
        }                                                                                  \
    } while (0)

#define DOTS_LAUNCH_CHECK()                                                                \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            dots::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return -3;                                                                     \
        }                                                                                  \
    } while (0)

// Programmatic dependent launch (PDL): kernels launched through launch_ex(..., pdl=true) may start while their stream
// predecessor is still running; they call pdl_wait() (ptx.cuh) before touching anything a predecessor writes.
// dots_set_pdl(0) turns the attribute off globally (plain stream serialisation).
extern int g_pdl;
// Timeline instrumentation (dots_debug_set_trace): when non-null, the decode kernels append (kernel id | point | CTA, %globaltimer,
// clock64) records to this device buffer: word 0 is the record counter, word 1 the capacity in records.  nullptr in normal operation.
extern unsigned long long* g_trace;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Same, for kernels launched as clusters of `cluster_x` CTAs (CTA pairs for cta_group::2 tensor-core work).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, unsigned cluster_x,
                                     Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && g_pdl) ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// 2-D bf16 row-major tensor [rows, cols] with row pitch ld (elements) -> TMA map with a
// {64 x box_rows} box and 128-byte swizzle.  Returns 0 on success.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols = 64);
// 3-D variant: [d2, d1, d0] with strides (elements) s2, s1, unit inner stride.
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                      uint64_t s2, uint32_t b0, uint32_t b1, uint32_t b2);

}  // namespace dots
