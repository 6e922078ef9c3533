// api_common.cu — error reporting, device selection, pinned host memory (C ABI: include/gdslam_cuda.h)
#include "gd_internal.h"

namespace gd {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int select_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libgdslam_cuda has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return GD_ENODEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device index %d out of range (have %d)", device, n);
        return GD_ENODEVICE;
    }
    GD_CUDA(cudaSetDevice(device));
    return GD_OK;
}

}  // namespace gd

extern "C" {

const char* gd_last_error(void) { return gd::g_last_error.c_str(); }
int gd_abi_version(void) { return GD_ABI_VERSION; }

int gd_device_count(int* count)
{
    if (!count) return GD_EINVAL;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        gd::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return GD_ENODEVICE;
    }
    *count = n;
    return GD_OK;
}

int gd_device_info(int device, char* name, int len, int* sm_count, size_t* total_mem)
{
    GD_TRY(gd::select_device(device));
    cudaDeviceProp p;
    GD_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && len > 0) {
        std::strncpy(name, p.name, (size_t)len - 1);
        name[len - 1] = 0;
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return GD_OK;
}

int gd_host_alloc(void** ptr, size_t bytes)
{
    if (!ptr) return GD_EINVAL;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        gd::set_error("gd_host_alloc: no CUDA device");
        return GD_ENODEVICE;
    }
    cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        gd::set_error("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
        return GD_ENOMEM;
    }
    return GD_OK;
}

int gd_host_free(void* ptr)
{
    if (ptr) cudaFreeHost(ptr);
    return GD_OK;
}

// Bare pinned-memory copy rate of one device: `iters` back-to-back cudaMemcpyAsync of `bytes` in the given direction, timed with
// CUDA events on a private stream.  bench.py runs it on every rank at the same time to measure what the box's host side can
// deliver (the ceiling of the end-to-end leg), independently of any kernel.
int gd_probe_copy(int device, size_t bytes, int iters, int to_device, double* gb_per_s)
{
    if (!gb_per_s || bytes == 0 || iters <= 0) return GD_EINVAL;
    GD_TRY(gd::select_device(device));
    gd::PinnedBuf hb;
    gd::DevBuf db;
    GD_TRY(hb.alloc(bytes));
    GD_TRY(db.alloc(bytes));
    std::memset(hb.p, 1, bytes);
    cudaStream_t s = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    GD_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int rc = GD_OK;
    float ms = 0.f;
    for (int pass = 0; pass < 2 && rc == GD_OK; ++pass) {  // pass 0 warms up
        cudaEventRecord(e0, s);
        for (int i = 0; i < (pass ? iters : 1); ++i) {
            const cudaError_t e = to_device ? cudaMemcpyAsync(db.p, hb.p, bytes, cudaMemcpyHostToDevice, s)
                                            : cudaMemcpyAsync(hb.p, db.p, bytes, cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) {
                gd::set_error("gd_probe_copy: %s", cudaGetErrorString(e));
                rc = GD_ECUDA;
                break;
            }
        }
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = GD_ECUDA;
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    if (rc == GD_OK) *gb_per_s = (double)bytes * iters / (ms * 1e-3) / 1e9;
    return rc;
}

}  // extern "C"
