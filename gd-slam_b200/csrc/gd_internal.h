// gd_internal.h — shared host-side plumbing of libgdslam_cuda (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gdslam_cuda.h"

namespace gd {

void set_error(const char* fmt, ...);

#define GD_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            gd::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return GD_ECUDA;                                                                       \
        }                                                                                          \
    } while (0)

#define GD_TRY(expr)                 \
    do {                             \
        int r__ = (expr);            \
        if (r__ != GD_OK) return r__; \
    } while (0)

#define GD_REQUIRE(cond, msg)                               \
    do {                                                    \
        if (!(cond)) {                                      \
            gd::set_error("%s: %s", __func__, msg);         \
            return GD_EINVAL;                               \
        }                                                   \
    } while (0)

int select_device(int device);  // validates + cudaSetDevice; GD_ENODEVICE when there is none

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// RAII device allocation
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t n)
    {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
            return GD_ENOMEM;
        }
        bytes = n;
        return GD_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
    void* p = nullptr;
    size_t bytes = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    int alloc(size_t n)
    {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMallocHost(&p, n);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMallocHost(%zu) failed: %s", n, cudaGetErrorString(e));
            return GD_ENOMEM;
        }
        bytes = n;
        return GD_OK;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// launch accounting / optional per-family event profile (gd_frontend_profile*)
struct LaunchStats {
    long long launches = 0;
    bool profiling = false;
    struct Family {
        const char* name;
        long long launches = 0;
        float ms = 0.f;
    };
    std::vector<Family> fam;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int family_index(const char* name)
    {
        for (size_t i = 0; i < fam.size(); ++i)
            if (fam[i].name == name || std::strcmp(fam[i].name, name) == 0) return (int)i;
        Family f;
        f.name = name;
        fam.push_back(f);
        return (int)fam.size() - 1;
    }
};

// Brackets kernel launches of one family: counts them and, when profiling, times them with events
// (profiling serialises the stream; it is only enabled by the bench's profile pass).
struct LaunchScope {
    LaunchStats* st;
    cudaStream_t s;
    int idx;
    LaunchScope(LaunchStats* st_, cudaStream_t s_, const char* name, int n_launches) : st(st_), s(s_), idx(-1)
    {
        if (!st) return;
        st->launches += n_launches;
        if (st->profiling) {
            idx = st->family_index(name);
            st->fam[idx].launches += n_launches;
            if (!st->e0) {
                cudaEventCreate(&st->e0);
                cudaEventCreate(&st->e1);
            }
            cudaEventRecord(st->e0, s);
        }
    }
    ~LaunchScope()
    {
        if (st && idx >= 0) {
            cudaEventRecord(st->e1, s);
            cudaEventSynchronize(st->e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, st->e0, st->e1);
            st->fam[idx].ms += ms;
        }
    }
};

// Programmatic dependent launch (sm_90+): a kernel launched through launch_pdl() may be scheduled as soon as every CTA of
// the previous kernel in the stream has executed pdl_trigger() (or exited); its own pdl_wait() then blocks until that
// kernel has completed and its writes are visible.  Kernels call pdl_trigger(); pdl_wait(); before their first global
// access, so the launch latency and the tail of the previous kernel overlap with the next kernel's CTAs becoming resident.
// Both instructions are no-ops for a kernel launched the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// GD_PDL: 1 = always, 0 = never, unset = automatic: on for handles with more than one stream.  (Measured on B200: with a
// single 640x480 stream the early-launched grids cost 26 us per frame — 0.311 vs 0.285 ms — from two streams on they pay.)
inline int pdl_mode()
{
    static const int mode = [] {
        const char* e = std::getenv("GD_PDL");
        return e ? (std::atoi(e) != 0 ? 1 : 0) : -1;
    }();
    return mode;
}
inline thread_local bool g_pdl_auto_on = true;  // set by PdlScope around the enqueue calls of a handle
struct PdlScope {
    bool prev;
    explicit PdlScope(int batch) : prev(g_pdl_auto_on) { g_pdl_auto_on = batch > 1; }
    ~PdlScope() { g_pdl_auto_on = prev; }
    PdlScope(const PdlScope&) = delete;
    PdlScope& operator=(const PdlScope&) = delete;
};
inline bool pdl_enabled() { return pdl_mode() < 0 ? g_pdl_auto_on : pdl_mode() == 1; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// The ordinary launch (full stream serialisation): for persistent kernels, whose CTAs would otherwise sit on every SM at
// griddepcontrol.wait while the predecessor drains.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_plain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Replays a static launch sequence as a CUDA graph, one executable per key (small batches are launch bound).
// enqueue() must only enqueue work on `s` (or on streams forked from and joined back into it), must give the same launches
// for the same key, and must not advance host state.  Never used inside another capture: the batched front-end captures
// the whole step itself and leaves the caches of its cores disabled.
struct GraphCache {
    bool enabled = false;
    std::map<unsigned long long, std::pair<cudaGraphExec_t, long long>> execs;
    GraphCache() = default;
    GraphCache(const GraphCache&) = delete;
    GraphCache& operator=(const GraphCache&) = delete;
    ~GraphCache()
    {
        for (auto& kv : execs)
            if (kv.second.first) cudaGraphExecDestroy(kv.second.first);
    }
    static bool env_default(int batch)
    {
        const char* e = std::getenv("GD_GRAPHS");
        return e ? std::atoi(e) != 0 : batch <= 8;
    }
    template <class F>
    int run(unsigned long long key, cudaStream_t s, LaunchStats* st, F&& enqueue)
    {
        if (!enabled || (st && st->profiling)) return enqueue();
        auto it = execs.find(key);
        if (it == execs.end()) {
            const long long l0 = st ? st->launches : 0;
            cudaGraph_t graph = nullptr;
            const cudaError_t be = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
            if (be != cudaSuccess) {
                set_error("cudaStreamBeginCapture failed (%s): this handle runs plain launches from now on", cudaGetErrorString(be));
                cudaGetLastError();
                enabled = false;
                return enqueue();
            }
            const int rc = enqueue();
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            cudaGraphExec_t exec = nullptr;
            if (rc != GD_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
                if (graph) cudaGraphDestroy(graph);
                set_error("CUDA graph capture / instantiation failed (enqueue rc %d, %s): this handle runs plain launches from now on",
                          rc, cudaGetErrorString(ce));
                cudaGetLastError();
                enabled = false;  // plain launches for good
                if (st) st->launches = l0;
                return enqueue();
            }
            cudaGraphDestroy(graph);
            const long long n = st ? st->launches - l0 : 0;
            if (st) st->launches = l0;  // the capture issued nothing; the replay below is what runs
            it = execs.emplace(key, std::make_pair(exec, n)).first;
        }
        GD_CUDA(cudaGraphLaunch(it->second.first, s));
        if (st) st->launches += it->second.second;
        return GD_OK;
    }
};

}  // namespace gd
