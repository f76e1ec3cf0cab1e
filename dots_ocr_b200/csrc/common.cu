// Error plumbing, device queries and TMA tensor-map construction for the C-ABI library.
#include "common.h"
#include "../../include/dots_ocr_b200.h"
#include <stdarg.h>
#include <string.h>

namespace dots {

static thread_local char g_err[1024] = "";
int g_pdl = 1;
unsigned long long* g_trace = nullptr;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cache[64] = {0};          // per device: one process may drive several GPUs
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (g_sm_override[dev] > 0) return g_sm_override[dev];      // launches aimed at an SM partition (dots_set_sm_count)
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cache[dev] = n;
    }
    return cache[dev];
}

bool first_use_on_device(bool (&flags)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (flags[dev]) return false;
    flags[dev] = true;
    return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved through the runtime so the library carries no link-time dependency on libcuda
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled not available from the driver");
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -4;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(2d) failed: CUresult %d (base=%p rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
                  base, (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
        return -4;
    }
    return 0;
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1,
                      uint64_t s2, uint32_t b0, uint32_t b1, uint32_t b2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return -4;
    cuuint64_t gdim[3] = {d0, d1, d2};
    cuuint64_t gstride[2] = {s1 * 2, s2 * 2};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed: CUresult %d", (int)r);
        return -4;
    }
    return 0;
}

}  // namespace dots

extern "C" const char* dots_last_error(void) { return dots::g_err; }

extern "C" int dots_abi_version(void) { return DOTS_ABI_VERSION; }

extern "C" int dots_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    DOTS_CHECK_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DOTS_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

extern "C" int dots_debug_set_trace(void* device_buffer) {
    dots::g_trace = reinterpret_cast<unsigned long long*>(device_buffer);
    return 0;
}

extern "C" int dots_set_pdl(int enable) {
    dots::g_pdl = enable ? 1 : 0;
    return 0;
}

// ---- CUDA-graph helpers: capture the launches a caller issues between begin/end on `stream` ----
extern "C" int dots_graph_begin(void* stream) {
    DOTS_CHECK_CUDA(cudaStreamBeginCapture(reinterpret_cast<cudaStream_t>(stream), cudaStreamCaptureModeThreadLocal));
    return 0;
}
extern "C" int dots_graph_end(void* stream, void** graph_exec_out) {
    cudaGraph_t graph = nullptr;
    DOTS_CHECK_CUDA(cudaStreamEndCapture(reinterpret_cast<cudaStream_t>(stream), &graph));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        dots::set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(e));
        return -2;
    }
    *graph_exec_out = exec;
    return 0;
}
extern "C" int dots_graph_launch(void* graph_exec, void* stream) {
    DOTS_CHECK_CUDA(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), reinterpret_cast<cudaStream_t>(stream)));
    return 0;
}
extern "C" int dots_graph_destroy(void* graph_exec) {
    if (graph_exec) DOTS_CHECK_CUDA(cudaGraphExecDestroy(reinterpret_cast<cudaGraphExec_t>(graph_exec)));
    return 0;
}
