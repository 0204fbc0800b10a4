// Shared helpers for the xfemm_b200 CUDA sources (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <atomic>
#include "../../include/xfemm_b200.h"

namespace xb {

void set_error(const char *fmt, ...);

// bookkeeping for bench.py: bytes crossing PCIe and kernels launched by this library
extern std::atomic<long long> g_h2d_bytes, g_d2h_bytes, g_launch; // (a host thread may pre-build the solver structure)

#define XB_CUDA(call)                                                                      \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            xb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,             \
                          cudaGetErrorString(e_));                                         \
            return XFEMM_ERR_CUDA;                                                         \
        }                                                                                  \
    } while (0)

#define XB_ARG(cond, msg)                                                                  \
    do {                                                                                   \
        if (!(cond)) { xb::set_error("invalid argument: %s", msg); return XFEMM_ERR_ARG; } \
    } while (0)

// Device memory goes through a small caching allocator (linprob.cu): cudaFree synchronises the
// device and, with the ~250 buffers of a problem plus its multigrid levels, tearing one problem
// down and setting the next one up cost ~1 s in cudaFree/cudaMalloc alone.  Freed blocks are kept
// (keyed by size) and handed out again; repeated solves of same-sized problems allocate nothing.
cudaError_t dev_alloc(void **p, size_t bytes);
void dev_free(void *p);
void dev_cache_flush(); // really frees everything cached

// Copies between big pageable host arrays and the device, staged through a pinned ring by several host threads
// (linprob.cu).  copy_h2d returns once the source has been read; copy_d2h returns with the data in place.
cudaError_t copy_h2d(void *dst, const void *src, size_t bytes, cudaStream_t st);
cudaError_t copy_d2h(void *dst, const void *src, size_t bytes, cudaStream_t st);

// Optional bump arena: while g_arena points at one, DevBuf allocations are carved from it, so a
// group of buffers (the small multigrid levels) ends up in ONE address range that a stream access
// policy window can pin in L2.  The arena's owner frees the range; arena-backed buffers do not.
struct DevArena
{
    char *base = nullptr;
    size_t cap = 0, used = 0;
};
extern thread_local DevArena *g_arena;
// scope guard: allocations of this thread bypass the arena (temporaries of a set-up step)
struct NoArena
{
    DevArena *prev;
    NoArena() : prev(g_arena) { g_arena = nullptr; }
    ~NoArena() { g_arena = prev; }
};

template <typename T>
struct DevBuf
{
    T *p = nullptr;
    size_t n = 0;
    bool in_arena = false;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() { if (p && !in_arena) dev_free(p); p = nullptr; n = 0; in_arena = false; }
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        if (g_arena && g_arena->used + bytes <= g_arena->cap)
        {
            p = (T *)(g_arena->base + g_arena->used);
            g_arena->used += bytes;
            in_arena = true;
            return cudaSuccess;
        }
        return dev_alloc((void **)&p, count * sizeof(T));
    }
    cudaError_t upload(const T *h, size_t count, cudaStream_t s = 0)
    {
        if (count != n) { cudaError_t e = alloc(count); if (e != cudaSuccess) return e; }
        if (count == 0) return cudaSuccess;
        g_h2d_bytes += (long long)(count * sizeof(T));
        return copy_h2d(p, h, count * sizeof(T), s);
    }
    cudaError_t zero(cudaStream_t s = 0) { return n ? cudaMemsetAsync(p, 0, n * sizeof(T), s) : cudaSuccess; }
};

// ---- device-side reductions -------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over a thread block; result valid in every thread.  `smem` needs 33 doubles.
__device__ __forceinline__ double block_sum(double v, double *smem)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads(); // protect smem reuse between consecutive calls
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    if (wid == 0)
    {
        double t = (lane < nw) ? smem[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) smem[32] = t;
    }
    __syncthreads();
    return smem[32];
}

// Deterministic re-reduction of per-block partials by every block (fixed order → every block
// gets the bit-identical value, so data-dependent control flow stays uniform across the grid).
__device__ __forceinline__ double reduce_partials(const double *__restrict__ part, int count, double *smem)
{
    double v = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) v += __ldcg(part + i);
    return block_sum(v, smem);
}

} // namespace xb
