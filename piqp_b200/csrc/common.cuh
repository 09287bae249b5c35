// piqp_b200/csrc/common.cuh -- shared device/host helpers for libpiqp_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdio>
#include <cstdint>
#include <atomic>
#include <shared_mutex>
#include <stdexcept>
#include <string>

namespace b200 {

// ---- error handling: CUDA failures become C++ exceptions inside the library and are turned into
// ---- B200_E_CUDA return codes at the C-ABI (capi.cu); nothing throws across the boundary.
struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
        throw CudaError(buf);
    }
}
#define B200_CUDA(x) ::b200::cuda_check((x), #x, __FILE__, __LINE__)

extern std::atomic<unsigned long long> g_launches;  // counted at every kernel launch (b200_kernel_launch_count); handles may be driven from several host threads
// Device timeline (debugging aid, B200_TIMELINE=1; off: one predictable branch per launch).  There is no nsys in the image and ncu
// serialises launches, so the only way to see where a captured IP iteration spends its time is to stamp it on the device: after every
// launch a one-thread kernel appends (%globaltimer, kernel name id) to a device buffer.  Stamps captured into a CUDA graph are replayed
// with it.  b200_timeline_dump() writes the stamps in execution order and starts over.  stamp i - stamp i-1 = kernel i + its launch gap (+ ~1 us stamp).
extern bool g_timeline_on;
void timeline_stamp(const char* name, cudaStream_t stream);
int timeline_dump(const char* path);
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                 \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);                 \
        B200_CUDA(cudaPeekAtLastError());                                           \
        if (::b200::g_timeline_on) ::b200::timeline_stamp(#kernel, (stream));       \
    } while (0)

// NVTX ranges named after the reference's Tracy zones (include/piqp/utils/tracy.hpp:11-25, PIQP_TRACY_ZoneScopedN("piqp::...")): a
// Nsight Systems timeline of the product then reads like a Tracy capture of the reference.  Header-only NVTX v3: a no-op (one
// predictable branch) unless a profiler injected itself.
struct NvtxZone {
    explicit NvtxZone(const char* name) { nvtxRangePushA(name); }
    ~NvtxZone() { nvtxRangePop(); }
    NvtxZone(const NvtxZone&) = delete;
};
#define B200_ZONE_CAT2(a, b) a##b
#define B200_ZONE_CAT(a, b) B200_ZONE_CAT2(a, b)
#define B200_ZONE(name) ::b200::NvtxZone B200_ZONE_CAT(nvtx_zone_, __LINE__)(name)

// A device-wide synchronisation is illegal while ANY thread captures a stream into a CUDA graph (the IP driver captures one iteration
// per handle, ip_solver.cu: ensure_graph), and handles may be driven from several host threads.  Captures hold this mutex exclusively
// (~0.2 ms, once per handle), device-wide synchronisations hold it shared.
inline std::shared_mutex& capture_mutex() { static std::shared_mutex m; return m; }
inline cudaError_t device_synchronize_shared() { std::shared_lock<std::shared_mutex> lk(capture_mutex()); return cudaDeviceSynchronize(); }

// Opt a kernel in to large dynamic shared memory.  cudaFuncAttributeMaxDynamicSharedMemorySize is per FUNCTION (and device), not per handle or
// launch: handles of different shapes are set up and driven from different host threads (bench.py, tools/mm_suite.py), so a per-handle value
// would race (one handle lowers the limit another is about to launch with: "invalid argument").  Every call therefore sets the same value,
// the device's opt-in maximum minus the kernel's static shared memory; `needed` is only checked against it.
template <class Kern>
inline void allow_dynamic_smem(Kern kern, size_t needed) {
    int dev = 0, optin = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes a;
    B200_CUDA(cudaFuncGetAttributes(&a, kern));
    const size_t limit = (size_t)optin - a.sharedSizeBytes;
    if (needed > limit) { char buf[160]; snprintf(buf, sizeof buf, "kernel needs %zu bytes of dynamic shared memory, the device allows %zu", needed, limit); throw CudaError(buf); }
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
}

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// RAII device buffer.  Allocations come from the device's stream-ordered memory pool (cudaMallocAsync) with the release
// threshold raised to "never", so that creating and destroying solvers re-uses HBM instead of paying the driver's
// cudaMalloc / cudaFree (measured: 30-60 ms to allocate and 30-350 ms to free one batched solver's ~100 buffers).
// Semantics stay those of cudaMalloc / cudaFree: memory is usable from any stream after alloc() returns, and release()
// waits for the device before handing the memory back.
inline void pool_setup_once() {
    static thread_local int configured_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == configured_dev) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    configured_dev = dev;
}
// Inside a ReleaseScope (b200qp_cleanup / b200kkt_destroy, after the handle's own stream has been synchronised) DevBuf::release() skips its
// device-wide synchronisation: a handle owns ~100 buffers, and every one of them would otherwise wait for the in-flight work of all the other
// handles of the process (4 pipelined sub-batches in bench.py, 12 handles in flight in tools/mm_suite.py).
struct ReleaseScope {
    ReleaseScope() { depth()++; }
    ~ReleaseScope() { depth()--; }
    ReleaseScope(const ReleaseScope&) = delete;
    static int& depth() { static thread_local int d = 0; return d; }
};
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) {
            pool_setup_once();
            B200_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), n * sizeof(T), cudaStreamPerThread));
            B200_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
        }
    }
    void zero(cudaStream_t s = 0) { if (n) B200_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    void release() {
        if (p) { if (ReleaseScope::depth() == 0) device_synchronize_shared(); cudaFreeAsync(p, cudaStreamPerThread); }
        p = nullptr; n = 0;
    }
    // stream-ordered release: the memory goes back to the pool once `s` has passed this point -- no device-wide synchronisation
    // (release() waits for the whole device, which serialises handles that work concurrently on different streams)
    void release_on(cudaStream_t s) {
        if (p) cudaFreeAsync(p, s);
        p = nullptr; n = 0;
    }
    T* get() const { return p; }
};

#ifdef __CUDACC__

constexpr double kInf = 1e30;

// ---- warp / block reductions (deterministic: fixed tree, independent of scheduling) ----
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide reductions of NV values at once; `red` is shared scratch of >= 32*NV doubles.
// All threads of the block must call; result is returned to all threads.
enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], const int (&op)[NV], double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        v[i] = op[i] == RED_SUM ? warp_sum(v[i]) : (op[i] == RED_MAX ? warp_max(v[i]) : warp_min(v[i]));
    }
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) red[i * 32 + w] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double x = op[i] == RED_SUM ? 0.0 : (op[i] == RED_MAX ? -INFINITY : INFINITY);
        if (lane < nw) x = red[i * 32 + lane];
        v[i] = op[i] == RED_SUM ? warp_sum(x) : (op[i] == RED_MAX ? warp_max(x) : warp_min(x));
    }
}

__device__ __forceinline__ bool is_finite_d(double x) { return isfinite(x); }

#endif  // __CUDACC__

}  // namespace b200
