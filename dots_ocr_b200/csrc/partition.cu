// Spatial SM partitions (CUDA green contexts) for running two phases of the page pipeline side by side on one GPU:
// batch i+1's ViT encode + prefill (tensor-core bound, power limited) on one set of SMs while batch i decodes (HBM / latency
// bound, tensor pipe idle) on the rest.  Streams of two ordinary CUDA streams would not do: the prefill kernels are persistent
// one-CTA-per-SM grids that own every SM for milliseconds, so a decode kernel (10 us of work) would queue behind each of them.
//
// The driver entry points are resolved through the runtime (cudaGetDriverEntryPoint), like cuTensorMapEncodeTiled in common.cu,
// so the library keeps no link-time dependency on libcuda.
#include "common.h"
#include "../../include/dots_ocr_b200.h"

namespace dots {

int g_sm_override[64] = {0};     // per device; 0 = the device's SM count (see num_sms())

namespace {

struct DriverApi {
    CUresult (*DeviceGet)(CUdevice*, int);
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType);
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int);
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
    CUresult (*GreenCtxDestroy)(CUgreenCtx);
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int);
    CUresult (*GreenCtxGetDevResource)(CUgreenCtx, CUdevResource*, CUdevResourceType);
    CUresult (*StreamDestroy)(CUstream);
    CUresult (*GetErrorString)(CUresult, const char**);
    bool ok;
};

template <typename F>
bool resolve(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        set_error("%s not available from the driver (green contexts need CUDA 12.4+)", name);
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

DriverApi* api() {
    static DriverApi a{};
    static bool tried = false;
    if (!tried) {
        tried = true;
        a.ok = resolve("cuDeviceGet", a.DeviceGet) && resolve("cuDeviceGetDevResource", a.DeviceGetDevResource) &&
               resolve("cuDevSmResourceSplitByCount", a.DevSmResourceSplitByCount) &&
               resolve("cuDevResourceGenerateDesc", a.DevResourceGenerateDesc) && resolve("cuGreenCtxCreate", a.GreenCtxCreate) &&
               resolve("cuGreenCtxDestroy", a.GreenCtxDestroy) && resolve("cuGreenCtxStreamCreate", a.GreenCtxStreamCreate) &&
               resolve("cuGreenCtxGetDevResource", a.GreenCtxGetDevResource) && resolve("cuStreamDestroy", a.StreamDestroy) &&
               resolve("cuGetErrorString", a.GetErrorString);
    }
    return a.ok ? &a : nullptr;
}

struct Partition {
    CUgreenCtx ctx[2];
    CUstream stream[2];
    int sms[2];
    bool live;
};
Partition g_part[64] = {};

}  // namespace

#define DOTS_CHECK_CU(a, expr)                                                             \
    do {                                                                                   \
        CUresult _r = (expr);                                                              \
        if (_r != CUDA_SUCCESS) {                                                          \
            const char* _s = nullptr;                                                      \
            (a)->GetErrorString(_r, &_s);                                                  \
            dots::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, _s ? _s : "?"); \
            return -2;                                                                     \
        }                                                                                  \
    } while (0)

}  // namespace dots

using namespace dots;

extern "C" int dots_partition_create(int sms_first, void** stream_first, void** stream_rest, int* n_first, int* n_rest) {
    DOTS_REQUIRE(stream_first && stream_rest && n_first && n_rest, "dots_partition_create: null output pointer");
    int dev = 0;
    DOTS_CHECK_CUDA(cudaGetDevice(&dev));
    DOTS_REQUIRE(dev >= 0 && dev < 64, "dots_partition_create: device index %d out of range", dev);
    DOTS_CHECK_CUDA(cudaFree(0));                                  // primary context up
    Partition& P = g_part[dev];
    DOTS_REQUIRE(!P.live, "dots_partition_create: device %d is already partitioned (dots_partition_destroy first)", dev);
    DriverApi* a = api();
    if (!a) return -4;
    const int total = num_sms();
    DOTS_REQUIRE(sms_first >= 8 && sms_first % 8 == 0 && sms_first <= total - 8,
                 "dots_partition_create: the first group needs a multiple of 8 SMs in [8, %d], got %d", total - 8, sms_first);
    CUdevice cud;
    DOTS_CHECK_CU(a, a->DeviceGet(&cud, dev));
    CUdevResource all, first, rest;
    DOTS_CHECK_CU(a, a->DeviceGetDevResource(cud, &all, CU_DEV_RESOURCE_TYPE_SM));
    unsigned int groups = 1;
    DOTS_CHECK_CU(a, a->DevSmResourceSplitByCount(&first, &groups, &all, &rest, 0, (unsigned)sms_first));
    DOTS_REQUIRE(groups == 1 && rest.type == CU_DEV_RESOURCE_TYPE_SM && rest.sm.smCount >= 8,
                 "dots_partition_create: the driver could not split %u SMs into %d + rest (got %u group(s), rest %u)", all.sm.smCount,
                 sms_first, groups, rest.type == CU_DEV_RESOURCE_TYPE_SM ? rest.sm.smCount : 0u);
    CUdevResource* parts[2] = {&first, &rest};
    for (int i = 0; i < 2; ++i) {
        CUdevResourceDesc desc;
        DOTS_CHECK_CU(a, a->DevResourceGenerateDesc(&desc, parts[i], 1));
        DOTS_CHECK_CU(a, a->GreenCtxCreate(&P.ctx[i], desc, cud, CU_GREEN_CTX_DEFAULT_STREAM));
        DOTS_CHECK_CU(a, a->GreenCtxStreamCreate(&P.stream[i], P.ctx[i], CU_STREAM_NON_BLOCKING, 0));
        CUdevResource got;
        DOTS_CHECK_CU(a, a->GreenCtxGetDevResource(P.ctx[i], &got, CU_DEV_RESOURCE_TYPE_SM));
        P.sms[i] = (int)got.sm.smCount;
    }
    P.live = true;
    *stream_first = (void*)P.stream[0];
    *stream_rest = (void*)P.stream[1];
    *n_first = P.sms[0];
    *n_rest = P.sms[1];
    return 0;
}

extern "C" int dots_partition_destroy(void) {
    int dev = 0;
    DOTS_CHECK_CUDA(cudaGetDevice(&dev));
    DOTS_REQUIRE(dev >= 0 && dev < 64, "dots_partition_destroy: device index %d out of range", dev);
    Partition& P = g_part[dev];
    if (!P.live) return 0;
    DriverApi* a = api();
    if (!a) return -4;
    DOTS_CHECK_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 2; ++i) {
        DOTS_CHECK_CU(a, a->StreamDestroy(P.stream[i]));
        DOTS_CHECK_CU(a, a->GreenCtxDestroy(P.ctx[i]));
    }
    P.live = false;
    g_sm_override[dev] = 0;
    return 0;
}

extern "C" int dots_set_sm_count(int n) {
    int dev = 0;
    DOTS_CHECK_CUDA(cudaGetDevice(&dev));
    DOTS_REQUIRE(dev >= 0 && dev < 64, "dots_set_sm_count: device index %d out of range", dev);
    DOTS_REQUIRE(n >= 0 && n <= 1024, "dots_set_sm_count: %d", n);
    g_sm_override[dev] = n;
    return 0;
}
