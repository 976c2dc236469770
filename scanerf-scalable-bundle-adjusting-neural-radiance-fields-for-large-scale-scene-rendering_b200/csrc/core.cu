// Library-wide plumbing for libscanerf_b200.so: error text, device queries, version.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void snrf_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int snrf_sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;  // B200
    }
    return cached;
}

// Scratch for the multi-pass paths (render passes, the unfused encode backward) comes from a private stream-ordered pool that
// keeps its memory across synchronisations (the default pool hands it back to the driver at every sync; re-mapping GBs per
// call costs tens of milliseconds).  snrf_infer_release_scratch() trims it.
static cudaMemPool_t g_scratch_pool[64] = {};
cudaError_t snrf_scratch_alloc(void** ptr, size_t bytes, cudaStream_t s)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaMallocAsync(ptr, bytes, s);
    if (!g_scratch_pool[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        e = cudaMemPoolCreate(&g_scratch_pool[dev], &props);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(g_scratch_pool[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return cudaMallocFromPoolAsync(ptr, bytes, g_scratch_pool[dev], s);
}
int snrf_scratch_release()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_scratch_pool[dev]) return 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemPoolTrimTo(g_scratch_pool[dev], 0);
    if (e != cudaSuccess) { snrf_set_error("snrf_infer_release_scratch: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

SNRF_API const char* snrf_last_error(void) { return g_err; }
SNRF_API int snrf_version(void) { return 100; }
SNRF_API int snrf_device_sm_count(void) { return snrf_sm_count(); }

// L2 fetch granularity of the current device (cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes; a hint to the driver).
// The gather / scatter kernels touch one 32-byte sector per hash-table corner at addresses the hash scatters over the
// whole level slice: a larger granularity drags the neighbouring sector(s) in with every miss, which costs DRAM bandwidth
// and L2 capacity for data that is evicted before its own touches arrive.  bytes <= 0 only queries.  Returns the value in
// effect after the call, or a negative CUDA error code.
SNRF_API int snrf_l2_fetch_granularity(int bytes)
{
    if (bytes > 0) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
        if (e != cudaSuccess) { cudaGetLastError(); snrf_set_error("snrf_l2_fetch_granularity(%d): %s", bytes, cudaGetErrorString(e)); return -(int)e; }
    }
    size_t v = 0;
    cudaError_t e = cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
    if (e != cudaSuccess) { cudaGetLastError(); snrf_set_error("snrf_l2_fetch_granularity: %s", cudaGetErrorString(e)); return -(int)e; }
    return (int)v;
}
