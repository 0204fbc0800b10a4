// xfemm_linprob_* — GPU backend of CBigLinProb (reference: cfemm/libfemm/spars.cpp:80-474).
// Hand-written sm_100a kernels: SELL-32 SpMV, fused vector updates with warp-shuffle + block
// reductions, Jacobi and multicolour-SSOR preconditioned CG as ONE persistent cooperative
// kernel (no host round trips inside the Krylov loop), Dirichlet / periodic boundary kernels.
#include "linprob.cuh"
#include "amg.cuh"
#include "pcg_args.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <functional>
#include <unordered_map>
#include <mutex>
#include <condition_variable>
#include <pthread.h>
#include <malloc.h>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <cstdio>

namespace cg = cooperative_groups;

namespace xb {

static thread_local std::string g_err;
void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char *last_error() { return g_err.c_str(); }
std::atomic<long long> g_h2d_bytes{0}, g_d2h_bytes{0}, g_launch{0};
thread_local DevArena *g_arena = nullptr;

// ---- caching device allocator (see common.cuh) ---------------------------------------------
// Free lists are kept per (device, rounded size): a process may use several devices (xfemm_cuda_init(device)), and a
// block freed on one must never be handed to a problem on another.
namespace {
std::mutex g_cache_mu;
struct CacheKey
{
    int dev;
    size_t bytes;
    bool operator==(const CacheKey &o) const { return dev == o.dev && bytes == o.bytes; }
};
struct CacheKeyHash
{
    size_t operator()(const CacheKey &k) const { return std::hash<size_t>()(k.bytes * 31 + (size_t)k.dev); }
};
std::unordered_map<CacheKey, std::vector<void *>, CacheKeyHash> g_cache_free; // (device, rounded size) -> blocks
std::unordered_map<void *, CacheKey> g_cache_size;                            // live and cached blocks
size_t g_cache_bytes = 0;                                                     // bytes sitting in the free lists
std::unordered_map<int, bool> g_cache_dirty;                                  // per device: frees since its last sync
inline size_t cache_round(size_t b) { return b < (1u << 20) ? ((b + 511) & ~(size_t)511) : ((b + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1)); }
inline size_t cache_cap()
{
    static const size_t cap = [] { const char *cm = getenv("XFEMM_DEVICE_CACHE_MB"); return (size_t)(cm ? atoi(cm) : 16384) << 20; }();
    return cap;
}
inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }
} // namespace

void dev_cache_flush()
{
    std::lock_guard<std::mutex> lk(g_cache_mu);
    const int here = current_device();
    for (auto &kv : g_cache_free)
    {
        if (kv.second.empty()) continue;
        cudaSetDevice(kv.first.dev);
        cudaDeviceSynchronize();
        for (void *q : kv.second) { g_cache_size.erase(q); cudaFree(q); }
        kv.second.clear();
    }
    cudaSetDevice(here);
    g_cache_free.clear();
    g_cache_bytes = 0;
    g_cache_dirty.clear();
}

cudaError_t dev_alloc(void **p, size_t bytes)
{
    const size_t r = cache_round(std::max<size_t>(bytes, 1));
    const int dev = current_device();
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cache_free.find(CacheKey{dev, r});
        if (it != g_cache_free.end() && !it->second.empty())
        {
            // a block freed earlier may still be read by work queued on its old stream (same device)
            if (g_cache_dirty[dev]) { cudaDeviceSynchronize(); g_cache_dirty[dev] = false; }
            *p = it->second.back();
            it->second.pop_back();
            g_cache_bytes -= r;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, r);
    if (e != cudaSuccess)
    {
        (void)cudaGetLastError();
        dev_cache_flush(); // give the cached blocks back and retry once
        e = cudaMalloc(p, r);
        if (e != cudaSuccess) return e;
    }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_size[*p] = CacheKey{dev, r};
    return cudaSuccess;
}

void dev_free(void *p)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache_size.find(p);
    if (it == g_cache_size.end()) { cudaFree(p); return; }
    const CacheKey key = it->second;
    if (g_cache_bytes + key.bytes > cache_cap())
    {
        g_cache_size.erase(it);
        cudaFree(p);
        return;
    }
    g_cache_free[key].push_back(p);
    g_cache_bytes += key.bytes;
    g_cache_dirty[key.dev] = true;
}

// ---- host <-> device copies of big pageable arrays through a pinned ring ---------------------------------------
// A plain cudaMemcpy from pageable memory lets the driver stage the data with one thread: 11 GB/s up, 17 GB/s down on
// the B200 boxes, against 32 / 54 GB/s from pinned memory.  The arrays of a problem (mesh, pattern, tables: several
// hundred MB at 2 M nodes) are std::vectors of the caller, so the copies here stage them themselves: chunks of 8 MB,
// copied into/out of two pinned buffers by a few host threads while the DMA engine moves the previous chunk.
// One ring per device, shared by all threads (mutex: the set-up threads of a problem take turns).
namespace {
struct StageRing
{
    std::mutex mu;
    char *buf[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = false;
    std::atomic<bool> tried{false};
};
constexpr size_t STAGE_CHUNK = 8u << 20;
constexpr size_t STAGE_MIN = 2u << 20; // smaller copies go the plain way
StageRing g_ring[16];

StageRing *stage_ring()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16 || getenv("XFEMM_NO_STAGED_COPIES")) return nullptr;
    StageRing &r = g_ring[dev];
    if (!r.tried)
    {
        std::lock_guard<std::mutex> lk(r.mu);
        if (!r.tried)
        {
            r.ok = cudaHostAlloc((void **)&r.buf[0], STAGE_CHUNK, cudaHostAllocDefault) == cudaSuccess &&
                   cudaHostAlloc((void **)&r.buf[1], STAGE_CHUNK, cudaHostAllocDefault) == cudaSuccess &&
                   cudaEventCreateWithFlags(&r.ev[0], cudaEventDisableTiming) == cudaSuccess &&
                   cudaEventCreateWithFlags(&r.ev[1], cudaEventDisableTiming) == cudaSuccess;
            if (!r.ok) (void)cudaGetLastError();
            r.tried = true;
        }
    }
    return r.ok ? &r : nullptr;
}

// Three helper threads that live as long as the process (created on first use, re-created in a forked child): the
// staging copies are a few hundred microseconds of memcpy each, so starting threads per chunk cost more than it saved.
// The caller takes one share itself.
class CopyPool
{
    static constexpr int HELPERS = 3;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    char *dst = nullptr;
    const char *src = nullptr;
    size_t bytes = 0;
    unsigned long long generation = 0;
    int pending = 0;
    bool started = false;

    static void share(char *d, const char *s, size_t n, int part, int parts)
    {
        const size_t unit = 4096, units = (n + unit - 1) / unit;
        const size_t a = std::min(n, units * part / parts * unit), b = std::min(n, units * (part + 1) / parts * unit);
        if (b > a) memcpy(d + a, s + a, b - a);
    }
    void worker(int id)
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu);
        for (;;)
        {
            cv_work.wait(lk, [&] { return generation != seen; });
            seen = generation;
            char *d = dst; const char *s = src; const size_t n = bytes;
            lk.unlock();
            share(d, s, n, id + 1, HELPERS + 1);
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }

public:
    void copy(char *d, const char *s, size_t n)
    {
        if (n < (1u << 20)) { memcpy(d, s, n); return; }
        {
            std::unique_lock<std::mutex> lk(mu);
            if (!started)
            {
                for (int i = 0; i < HELPERS; i++) std::thread([this, i] { worker(i); }).detach();
                started = true;
            }
            dst = d; src = s; bytes = n; pending = HELPERS; generation++;
        }
        cv_work.notify_all();
        share(d, s, n, 0, HELPERS + 1);
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};
CopyPool *g_copy_pool = nullptr;
std::once_flag g_copy_pool_once;

void host_copy(char *dst, const char *src, size_t bytes)
{
    std::call_once(g_copy_pool_once, [] {
        g_copy_pool = new CopyPool;
        // a forked child has none of the helper threads: it gets a pool of its own (the parent's object is leaked there)
        pthread_atfork(nullptr, nullptr, [] { g_copy_pool = new CopyPool; });
    });
    g_copy_pool->copy(dst, src, bytes); // (callers hold the ring's mutex: one copy at a time)
}

// is this host pointer page-locked (cudaHostAlloc / cudaHostRegister)?  Then the DMA engine can use it directly.
bool host_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
} // namespace

// returns when `src` has been read (the caller may reuse it); the device copy is ordered on `st`
cudaError_t copy_h2d(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    StageRing *r = (bytes >= STAGE_MIN && !host_pinned(src)) ? stage_ring() : nullptr;
    if (!r) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
    std::lock_guard<std::mutex> lk(r->mu);
    cudaError_t e = cudaSuccess;
    int k = 0;
    for (size_t off = 0; off < bytes && e == cudaSuccess; off += STAGE_CHUNK, k ^= 1)
    {
        const size_t len = std::min(STAGE_CHUNK, bytes - off);
        if ((e = cudaEventSynchronize(r->ev[k])) != cudaSuccess) break; // the copy that last used this buffer
        host_copy(r->buf[k], (const char *)src + off, len);
        if ((e = cudaMemcpyAsync((char *)dst + off, r->buf[k], len, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        e = cudaEventRecord(r->ev[k], st);
    }
    return e;
}

// returns with the data in `dst` (everything queued on `st` before the call has completed by then)
cudaError_t copy_d2h(void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    StageRing *r = (bytes >= STAGE_MIN && !host_pinned(dst)) ? stage_ring() : nullptr;
    if (!r)
    {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaStreamSynchronize(st);
    }
    std::lock_guard<std::mutex> lk(r->mu);
    cudaError_t e = cudaSuccess;
    const size_t nchunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    for (size_t c = 0; c <= nchunks && e == cudaSuccess; c++)
    {
        if (c < nchunks)
        {
            const int k = (int)(c & 1);
            const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
            if ((e = cudaEventSynchronize(r->ev[k])) != cudaSuccess) break;
            if ((e = cudaMemcpyAsync(r->buf[k], (const char *)src + off, len, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
            if ((e = cudaEventRecord(r->ev[k], st)) != cudaSuccess) break;
        }
        if (c > 0)
        {
            const int k = (int)((c - 1) & 1);
            const size_t off = (c - 1) * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
            if ((e = cudaEventSynchronize(r->ev[k])) != cudaSuccess) break;
            host_copy((char *)dst + off, r->buf[k], len);
        }
    }
    return e;
}

static inline int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// =====================================================================================
// small utility kernels
// =====================================================================================
__global__ void k_sell_fill(long long len, const int *__restrict__ src, const double *__restrict__ vals,
                            double *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    {
        int s = src[i];
        out[i] = s >= 0 ? vals[s] : 0.0;
    }
}

__global__ void k_diag_fill(int n_pad, const int *__restrict__ diag_src, const double *__restrict__ vals,
                            double *__restrict__ diag, double *__restrict__ dinv, int *singular_row)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int s = diag_src[r];
    double d = s >= 0 ? vals[s] : 0.0;
    diag[r] = d;
    dinv[r] = (d != 0.0) ? 1.0 / d : 0.0;
    if (s >= 0 && d == 0.0) atomicMin(singular_row, r);
}

// caller order -> internal order (padding rows get 0)
__global__ void k_to_internal(int n_pad, const int *__restrict__ i2c, const double *__restrict__ in, double *__restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int i = i2c[r];
    out[r] = i >= 0 ? in[i] : 0.0;
}
__global__ void k_from_internal(int n_pad, const int *__restrict__ i2c, const double *__restrict__ in, double *__restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int i = i2c[r];
    if (i >= 0) out[i] = in[r];
}

// =====================================================================================
// SELL-32 SpMV building block: one warp computes 32 rows of y = A x (lower + upper + diagonal)
// =====================================================================================
__device__ __forceinline__ double sell_row_offdiag(const double *__restrict__ val, const int *__restrict__ col,
                                                    int base, int k0, int k1, int lane, const double *x)
{
    double acc = 0.0;
    const double *v = val + base + lane;
    const int *c = col + base + lane;
#pragma unroll 8
    for (int k = k0; k < k1; k++)
    {
        // matrix stream: read once per pass → evict-first so the vectors stay resident in L2
        double a = __ldcs(v + k * 32);
        int j = __ldcs(c + k * 32);
        acc = fma(a, x[j], acc);
    }
    return acc;
}

__global__ void __launch_bounds__(256)
k_spmv(int nslices, const int *__restrict__ slice_ptr, const int *__restrict__ sell_col,
       const double *__restrict__ sell_val, const double *__restrict__ diag, const double *x, double *y,
       const unsigned *__restrict__ wide_mask)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nslices; s += nwarps)
    {
        const int base = slice_ptr[s];
        const int w = (slice_ptr[s + 1] - base) >> 5;
        const int row = s * 32 + lane;
        double acc = sell_row_offdiag(sell_val, sell_col, base, 0, w, lane, x);
        if (wide_mask && ((wide_mask[s] >> lane) & 1u)) acc += y[row]; // off-diagonal part left there by k_wide_pre
        y[row] = fma(diag[row], x[row], acc);
    }
}

// res2 = sum (b-u)^2, b2 = sum b^2 → two scalars (deterministic two-stage)
__global__ void k_resnorm_partial(int n, const double *__restrict__ b, const double *__restrict__ u, double *part)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        double bb = b[i], d = bb - u[i];
        a = fma(d, d, a);
        c = fma(bb, bb, c);
    }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = a; part[gridDim.x + blockIdx.x] = c; }
}
__global__ void k_reduce2_final(int nblk, const double *part, double *out2)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) { a += part[i]; c += part[nblk + i]; }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { out2[0] = a; out2[1] = c; }
}

// =====================================================================================
// Persistent preconditioned CG (reference recurrence: spars.cpp:238-316)
// =====================================================================================
// dynamic distribution of a block's slices over its warps (shared-memory ticket counter)
__device__ __forceinline__ int next_slice(int *ctr, int lane)
{
    int v = 0;
    if (lane == 0) v = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}

template <int PRECOND>
__device__ __forceinline__ double precond_apply(const PcgArgs &a, const AmgArgs &g, const AmgArgs &sg, GridBar &grid, int r0, int r1,
                                                int warp_g, int nwarps_g, int lane, int *s_next, int *s_wk,
                                                StageSmem *S, bool x0_ready = false)
{
    // returns this thread's partial of r.z ; z is left in a.z (MC-SSOR, AMG) or implied (Jacobi)
    double dot = 0.0;
    if (PRECOND == XFEMM_PRECOND_AMG) return amg_vcycle(g, sg, grid, s_next, s_wk, S, x0_ready);
    if (PRECOND == XFEMM_PRECOND_JACOBI)
    {
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            double ri = a.r[i];
            dot = fma(ri * a.dinv[i], ri, dot);
        }
        return dot;
    }
    // ---- multicolour SSOR:  z = w(2-w) (D+wU)^-1 D (D+wL)^-1 r   (spars.cpp:197-236, colour order)
    const double w = a.omega, c0 = w * (2.0 - w);
    const int m = a.ncolors;
    // start: z = c0*r ; colour 0 has no lower part → finish its forward step here
    {
        const int rc0 = a.color_slice[1] * 32;
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            double v = c0 * a.r[i];
            if (i < rc0) v *= a.dinv[i];
            a.z[i] = v;
        }
    }
    grid.sync();
    // forward sweep, colours 1..m-1
    for (int c = 1; c < m; c++)
    {
        const int sa = a.color_slice[c], sb = a.color_slice[c + 1];
        for (int s = sa + warp_g; s < sb; s += nwarps_g)
        {
            const int base = a.slice_ptr[s], wl = a.slice_wl[s], row = s * 32 + lane;
            double acc = sell_row_offdiag(a.sell_val, a.sell_col, base, 0, wl, lane, a.z);
            double v = (a.z[row] - w * acc) * a.dinv[row];
            a.z[row] = v;
            if (c == m - 1) dot = fma(a.r[row], v, dot); // last colour is final after the forward sweep
        }
        grid.sync();
    }
    // backward sweep, colours m-2..0:  z_i -= w/d_i * sum_upper a_ij z_j
    for (int c = m - 2; c >= 0; c--)
    {
        const int sa = a.color_slice[c], sb = a.color_slice[c + 1];
        for (int s = sa + warp_g; s < sb; s += nwarps_g)
        {
            const int base = a.slice_ptr[s], wl = a.slice_wl[s], row = s * 32 + lane;
            const int wtot = wl + a.slice_wu[s]; // the "extra" block behind it is not part of the sweeps
            double acc = sell_row_offdiag(a.sell_val, a.sell_col, base, wl, wtot, lane, a.z);
            double v = a.z[row] - w * acc * a.dinv[row];
            a.z[row] = v;
            dot = fma(a.r[row], v, dot);
        }
        if (c > 0) grid.sync();
    }
    if (m == 1) // degenerate: diagonal matrix
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) dot = fma(a.r[i], a.z[i], dot);
    return dot;
}

template <int PRECOND>
__global__ void __launch_bounds__(1024, 1) k_pcg(const __grid_constant__ PcgArgs a, const __grid_constant__ AmgArgs g)
{
    __shared__ __align__(16) unsigned char sg_raw[PRECOND == XFEMM_PRECOND_AMG ? sizeof(AmgArgs) : 16];
    const AmgArgs *sgp = &g; // multigrid: the block's shared-memory copy of g for the out-of-line passes (amg_kernel_init)
    __shared__ double sm[33];
    __shared__ int s_next;
    __shared__ int s_wk[AMG_SWK_SIZE];
    GridBar grid{cg::this_grid(), PRECOND == XFEMM_PRECOND_AMG ? g.grid_ctr : nullptr, s_wk + AMG_SWK_EPOCH + 1};
    const int lane = threadIdx.x & 31;
    const int nwarps_b = blockDim.x >> 5, warp_b = threadIdx.x >> 5;
    const int warp_g = blockIdx.x * nwarps_b + warp_b, nwarps_g = gridDim.x * nwarps_b;
    const int G = gridDim.x;
    // contiguous chunk of slices / rows owned by this block in the vector phases
    const int s0 = (int)((long long)a.nslices * blockIdx.x / G);
    const int s1 = (int)((long long)a.nslices * (blockIdx.x + 1) / G);
    const int r0 = s0 * 32, r1 = s1 * 32;
    double *partA = a.part, *partB = a.part + G;
    int wk0 = 0, wk1 = 0; // wide rows of this block's chunk
    if (a.wide.n_rows) wide_range(a.wide, r0, r1, wk0, wk1);
    // shared-memory ring for the matrix stream of the finest level (stage.cuh); no tiles → direct loads
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    StageSmem S = stage_carve(dyn_smem);
    if (a.tiles)
    {
        const int t0 = a.tile_ptr[blockIdx.x], t1 = a.tile_ptr[blockIdx.x + 1];
        S.ntiles = t1 - t0;
        S.tile = a.tiles + t0;
    }
    if (PRECOND == XFEMM_PRECOND_AMG) sgp = &amg_kernel_init<1>(g, sg_raw, s_wk);

    // ---- res_o = b . M^-1 b   (spars.cpp:257-259) ; z0 is computed with r := b
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
    {
        a.r[i] = a.b[i];
        if (!a.warm) a.x[i] = 0.0;
    }
    grid.sync();
    double part = precond_apply<PRECOND>(a, g, *sgp, grid, r0, r1, warp_g, nwarps_g, lane, &s_next, s_wk, &S);
    part = block_sum(part, sm);
    if (threadIdx.x == 0) partA[blockIdx.x] = part;
    grid.sync();
    const double res_o = reduce_partials(partA, G, sm);
    if (res_o == 0.0)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0) { a.mail[0] = 0; a.mail[1] = 0; a.mail[2] = XFEMM_OK; a.mail[3] = 0; a.mail[4] = 0; }
        return;
    }
    // ---- r = b - A x
    if (a.warm)
    {
        wide_rows_block<double>(a.wide, a.wide_val, a.wide_acc, wk0, wk1, a.x, a.u);
        __syncthreads();
        for (int s = s0 + warp_b; s < s1; s += nwarps_b)
        {
            const int base = a.slice_ptr[s], w = (a.slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
            double acc = sell_row_offdiag(a.sell_val, a.sell_col, base, 0, w, lane, a.x);
            if (wk1 > wk0 && ((a.wide.mask[s] >> lane) & 1u)) acc += a.u[row];
            a.r[row] = a.b[row] - fma(a.diag[row], a.x[row], acc);
        }
        grid.sync();
    }
    // ---- z = M^-1 r ; p = z ; res = z.r
    part = precond_apply<PRECOND>(a, g, *sgp, grid, r0, r1, warp_g, nwarps_g, lane, &s_next, s_wk, &S);
    part = block_sum(part, sm);
    if (threadIdx.x == 0) partB[blockIdx.x] = part;
    grid.sync();
    double res = reduce_partials(partB, G, sm);
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        a.p[i] = (PRECOND == XFEMM_PRECOND_JACOBI) ? a.r[i] * a.dinv[i] : a.z[i];
    grid.sync();

    int it = 0;
    int status = XFEMM_OK;
    double er = 0.0;
    while (true)
    {
        // step i: u = A p ; pAp   (warps of a block pull slices of the block's chunk dynamically)
        int ck = AMG_MAX_STAMPS - 8; // XFEMM_AMG_TRACE: the phases of the CG part go to the tail of the stamp buffer
        if (PRECOND == XFEMM_PRECOND_AMG) amg_stamp(g, ck, AMG_MAX_STAMPS);
        part = 0.0;
        if (PRECOND == XFEMM_PRECOND_AMG) amg_block_stamp(g, 0, true);
        if (S.ntiles > 0)
        {
            // matrix stream through the shared-memory ring (stage.cuh)
            staged_slices(S, a.slice_ptr, a.sell_val, a.sell_col, s0, s1, a.p, &s_next,
                          [&](int row) { return make_double2(a.diag[row], a.p[row]); },
                          [&](int s, int row, double acc, double2 dp) {
                              if (a.wide.n_rows && ((a.wide.mask[s] >> lane) & 1u)) return; // a wide row: see `before`
                              acc = fma(dp.x, dp.y, acc);
                              a.u[row] = acc;
                              part = fma(dp.y, acc, part);
                          },
                          [&]() {
                              // the wide rows of the matrix, dealt out over the blocks: a warp or two of each, before
                              // they join the stream
                              if (a.wide.n_rows)
                                  wide_rows_warp(a.wide, a.wide_val, (int)blockIdx.x, a.wide.n_rows, a.p, [&](int row, double off) {
                                      const double pr = a.p[row], acc = fma(a.diag[row], pr, off);
                                      a.u[row] = acc;
                                      part = fma(pr, acc, part);
                                  }, (int)gridDim.x);
                          });
        }
        else
        {
        wide_rows_block<double>(a.wide, a.wide_val, a.wide_acc, wk0, wk1, a.p, a.u);
        if (threadIdx.x == 0) s_next = s0;
        __syncthreads();
        for (int s = next_slice(&s_next, lane); s < s1; s = next_slice(&s_next, lane))
        {
            const int base = a.slice_ptr[s], w = (a.slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
            double acc = sell_row_offdiag(a.sell_val, a.sell_col, base, 0, w, lane, a.p);
            if (wk1 > wk0 && ((a.wide.mask[s] >> lane) & 1u)) acc += a.u[row];
            const double pr = a.p[row];
            acc = fma(a.diag[row], pr, acc);
            a.u[row] = acc;
            part = fma(pr, acc, part);
        }
        }
        if (PRECOND == XFEMM_PRECOND_AMG) amg_block_stamp(g, 0);
        part = block_sum(part, sm);
        if (threadIdx.x == 0) partA[blockIdx.x] = part;
        grid.sync();
        if (PRECOND == XFEMM_PRECOND_AMG) amg_stamp(g, ck, AMG_MAX_STAMPS);
        const double pAp = reduce_partials(partA, G, sm);
        if (!(pAp > 0.0)) { status = XFEMM_ERR_DIVERGED; break; }
        const double del = res / pAp;
        // step iii: R -= del U (Jacobi also fuses z.r).  Step ii, V += del P, rides along with step v below:
        // it needs nothing but P, and keeping x and p out of this phase leaves the vectors the next SpMV
        // phase gathers from (multigrid: the pre-smoothed iterate) resident in L2.
        if (PRECOND == XFEMM_PRECOND_JACOBI)
        {
            part = 0.0;
            for (int i = r0 + threadIdx.x; i < r1; i += 4 * blockDim.x)
            {
                double uv[4], rv[4], dv[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const int k = i + q * blockDim.x;
                    const bool ok = k < r1;
                    uv[q] = ok ? a.u[k] : 0.0;
                    rv[q] = ok ? a.r[k] : 0.0;
                    dv[q] = ok ? a.dinv[k] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const int k = i + q * blockDim.x;
                    if (k < r1)
                    {
                        const double ri = fma(-del, uv[q], rv[q]);
                        a.r[k] = ri;
                        part = fma(ri * dv[q], ri, part);
                    }
                }
            }
        }
        else
        {
            // four rows per thread and step, loads ahead of the stores (the arrays may alias as far as the compiler
            // knows: written row by row, every step would wait for the previous one's stores)
            for (int i = r0 + threadIdx.x; i < r1; i += 4 * blockDim.x)
            {
                double uv[4], rv[4], dv[4];
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const int k = i + q * blockDim.x;
                    const bool ok = k < r1;
                    uv[q] = ok ? a.u[k] : 0.0;
                    rv[q] = ok ? a.r[k] : 0.0;
                    if (PRECOND == XFEMM_PRECOND_AMG) dv[q] = ok ? a.dinv[k] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const int k = i + q * blockDim.x;
                    if (k < r1)
                    {
                        const double ri = fma(-del, uv[q], rv[q]);
                        a.r[k] = ri;
                        // multigrid: the pre-smoothed iterate of the finest level comes for free here
                        if (PRECOND == XFEMM_PRECOND_AMG) g.lev[0].x[k] = g.omega * dv[q] * ri;
                    }
                }
            }
            grid.sync();
            if (PRECOND == XFEMM_PRECOND_AMG) amg_stamp(g, ck, AMG_MAX_STAMPS);
            // step iv: z = M^-1 r ; z.r
            part = precond_apply<PRECOND>(a, g, *sgp, grid, r0, r1, warp_g, nwarps_g, lane, &s_next, s_wk, &S, true);
        }
        part = block_sum(part, sm);
        if (threadIdx.x == 0) partB[blockIdx.x] = part;
        grid.sync();
        if (PRECOND == XFEMM_PRECOND_AMG) amg_stamp(g, ck, AMG_MAX_STAMPS);
        const double res_new = reduce_partials(partB, G, sm);
        const double rho = res_new / res;
        res = res_new;
        // step ii: V += del P, then step v: P = Z + rho P
        for (int i = r0 + threadIdx.x; i < r1; i += 4 * blockDim.x)
        {
            double pv[4], xv[4], zv[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int k = i + q * blockDim.x;
                const bool ok = k < r1;
                pv[q] = ok ? a.p[k] : 0.0;
                xv[q] = ok ? a.x[k] : 0.0;
                if (PRECOND == XFEMM_PRECOND_JACOBI) zv[q] = ok ? a.r[k] * a.dinv[k] : 0.0;
                else zv[q] = ok ? a.z[k] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int k = i + q * blockDim.x;
                if (k < r1)
                {
                    a.x[k] = fma(del, pv[q], xv[q]);
                    a.p[k] = fma(rho, pv[q], zv[q]);
                }
            }
        }
        if (PRECOND == XFEMM_PRECOND_AMG) amg_stamp(g, ck, AMG_MAX_STAMPS);
        it++;
        er = sqrt(res / res_o);
        if (!(er == er)) { status = XFEMM_ERR_DIVERGED; break; }
        if (er <= a.precision) break;
        if (it >= a.max_iter) { status = XFEMM_ERR_DIVERGED; break; }
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        a.mail[0] = it; a.mail[1] = er; a.mail[2] = status; a.mail[3] = res_o; a.mail[4] = res;
    }
}

template <int PRECOND>
__global__ void __launch_bounds__(512, 2) k_precond_only(const __grid_constant__ PcgArgs a, const __grid_constant__ AmgArgs g)
{
    __shared__ __align__(16) unsigned char sg_raw[PRECOND == XFEMM_PRECOND_AMG ? sizeof(AmgArgs) : 16];
    const AmgArgs *sgp = &g;
    __shared__ int s_next;
    __shared__ int s_wk[AMG_SWK_SIZE];
    GridBar grid{cg::this_grid(), PRECOND == XFEMM_PRECOND_AMG ? g.grid_ctr : nullptr, s_wk + AMG_SWK_EPOCH + 1};
    const int lane = threadIdx.x & 31;
    const int nwarps_b = blockDim.x >> 5, warp_b = threadIdx.x >> 5;
    const int warp_g = blockIdx.x * nwarps_b + warp_b, nwarps_g = gridDim.x * nwarps_b;
    const int G = gridDim.x;
    const int s0 = (int)((long long)a.nslices * blockIdx.x / G);
    const int s1 = (int)((long long)a.nslices * (blockIdx.x + 1) / G);
    const int r0 = s0 * 32, r1 = s1 * 32;
    if (PRECOND == XFEMM_PRECOND_AMG) sgp = &amg_kernel_init<1>(g, sg_raw, s_wk);
    precond_apply<PRECOND>(a, g, *sgp, grid, r0, r1, warp_g, nwarps_g, lane, &s_next, s_wk, nullptr);
    if (PRECOND == XFEMM_PRECOND_JACOBI)
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) a.z[i] = a.r[i] * a.dinv[i];
}

// =====================================================================================
// boundary conditions on the CSR representation
// =====================================================================================
__global__ void k_fill_int(int n, int *p, int v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_dir_scatter(int count, const int *__restrict__ nodes, int *dir_idx)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dir_idx[nodes[k]] = k;
}
// SetValue for all Dirichlet nodes at once (reference semantics, spars.cpp:318-346):
//   entry {k,i} is eliminated when i is fixed and i-bdw <= k < i+bdw (whole column if bdw==0);
//   free row k:  b[k] -= A[k,i]*x_first(i) ; fixed row i: b[i] = A[i,i]*x_last(i).
__global__ void k_dirichlet(int n, const int *__restrict__ rowptr, const int *__restrict__ col, double *vals,
                            double *b, const int *__restrict__ dir_idx, const double *__restrict__ xf,
                            const double *__restrict__ xl, int bdw)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int dk = dir_idx[k];
    const int s0 = rowptr[k], s1 = rowptr[k + 1];
    double bk = b[k], dkk = 0.0;
    for (int s = s0; s < s1; s++)
    {
        const int c = col[s];
        if (c == k) { dkk = vals[s]; continue; }
        const int dc = dir_idx[c];
        // eliminated by SetValue(c): c-bdw <= k < c+bdw ; by SetValue(k): k-bdw <= c < k+bdw
        const bool by_c = dc >= 0 && (bdw == 0 || (k >= c - bdw && k < c + bdw));
        const bool by_k = dk >= 0 && (bdw == 0 || (c >= k - bdw && c < k + bdw));
        if (by_c || by_k)
        {
            if (by_c && dk < 0) bk = fma(-vals[s], xf[dc], bk);
            vals[s] = 0.0;
        }
    }
    if (dk >= 0) bk = dkk * xl[dk];
    b[k] = bk;
}

// Periodicity / AntiPeriodicity in list order (spars.cpp:421-474 / 366-419).  The host plan
// groups pairs into dependency levels (pairs of one level touch disjoint entries); one block
// walks the levels, one warp per pair, lanes over the coupled rows k.
//   item (4 ints): slot(k,i) slot(i,k) slot(k,j) slot(j,k)
//   meta (6 ints): i, j, type, slot(i,i), slot(j,j), unused
// lv0 >= 0: only that level, pairs spread over the whole grid (pairs of one level touch disjoint
// entries; the host launches the levels one after the other).  lv0 < 0: all levels in ONE block.
__global__ void __launch_bounds__(1024)
k_periodic(int nlevels, const int *__restrict__ level_ptr, const int *__restrict__ item_ptr,
           const int *__restrict__ meta, const int *__restrict__ items, double *vals, double *b,
           const unsigned char *__restrict__ fill_mask, unsigned char *created, int lv0)
{
    const int lane = threadIdx.x & 31;
    const int warp = (lv0 >= 0 ? blockIdx.x * (blockDim.x >> 5) : 0) + (threadIdx.x >> 5);
    const int nw = (lv0 >= 0 ? gridDim.x : 1) * (blockDim.x >> 5);
    for (int lv = (lv0 >= 0 ? lv0 : 0); lv < (lv0 >= 0 ? lv0 + 1 : nlevels); lv++)
    {
        for (int pr = level_ptr[lv] + warp; pr < level_ptr[lv + 1]; pr += nw)
        {
            const int *m = meta + 6 * pr;
            const double sgn = m[2] ? -1.0 : 1.0;
            for (int it = item_ptr[pr] + lane; it < item_ptr[pr + 1]; it += 32)
            {
                const int *q = items + 4 * it;
                double v1 = __ldcg(vals + q[0]), v2 = __ldcg(vals + q[2]);
                if (v1 != 0.0 || v2 != 0.0)
                {
                    double c = (v1 + sgn * v2) / 2.0;
                    __stcg(vals + q[0], c); __stcg(vals + q[1], c);
                    __stcg(vals + q[2], sgn * c); __stcg(vals + q[3], sgn * c);
                    if (fill_mask)
                    {
#pragma unroll
                        for (int t = 0; t < 4; t++)
                            if (fill_mask[q[t]]) created[q[t]] = 1;
                    }
                }
            }
            if (lane == 0)
            {
                double c = m[2] ? 0.5 * (__ldcg(vals + m[3]) + __ldcg(vals + m[4]))
                                : (__ldcg(vals + m[3]) + __ldcg(vals + m[4])) / 2.0;
                __stcg(vals + m[3], c); __stcg(vals + m[4], c);
                double bi = __ldcg(b + m[0]), bj = __ldcg(b + m[1]);
                double cb = 0.5 * (bi + sgn * bj);
                __stcg(b + m[0], cb); __stcg(b + m[1], sgn * cb);
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// Long dependency chains (neighbouring pairs along a periodic boundary share entries, so a list
// ordered along the boundary is ONE chain): one block, every warp pre-loads the plan of its next pair
// (indices only, independent of the values) and then waits for its turn; the per-pair critical path is
// the value loads + stores + a shared-memory hand-off instead of a chain of five dependent loads and a
// block barrier.  pair_wait[q] = number of pairs (in level-sorted order) that must be finished first.
__global__ void __launch_bounds__(1024)
k_periodic_turn(int npairs, const int *__restrict__ pair_wait, const int *__restrict__ item_ptr,
                const int *__restrict__ meta, const int *__restrict__ items, double *vals, double *b,
                const unsigned char *__restrict__ fill_mask, unsigned char *created)
{
    __shared__ int done;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    for (int pr = warp; pr < npairs; pr += nw)
    {
        // ---- plan of this pair (does not depend on earlier pairs)
        const int *m = meta + 6 * pr;
        const int ni = m[0], nj = m[1], sii = m[3], sjj = m[4];
        const double sgn = m[2] ? -1.0 : 1.0;
        const int it0 = item_ptr[pr], it1 = item_ptr[pr + 1], need = pair_wait[pr];
        int4 q = make_int4(0, 0, 0, 0);
        if (it0 + lane < it1) q = *reinterpret_cast<const int4 *>(items + 4 * (it0 + lane));
        // ---- wait for the pairs this one depends on
        while (*(volatile int *)&done < need) {}
        __syncwarp();
        // all value loads of the pair in one round trip: lane 0 also fetches the two diagonals and the RHS
        double dii = 0, djj = 0, bi = 0, bj = 0;
        if (lane == 0) { dii = __ldcg(vals + sii); djj = __ldcg(vals + sjj); bi = __ldcg(b + ni); bj = __ldcg(b + nj); }
        for (int it = it0 + lane; it < it1; it += 32)
        {
            if (it != it0 + lane) q = *reinterpret_cast<const int4 *>(items + 4 * it);
            double v1 = __ldcg(vals + q.x), v2 = __ldcg(vals + q.z);
            if (v1 != 0.0 || v2 != 0.0)
            {
                double c = (v1 + sgn * v2) / 2.0;
                __stcg(vals + q.x, c); __stcg(vals + q.y, c);
                __stcg(vals + q.z, sgn * c); __stcg(vals + q.w, sgn * c);
                if (fill_mask)
                {
                    if (fill_mask[q.x]) created[q.x] = 1;
                    if (fill_mask[q.y]) created[q.y] = 1;
                    if (fill_mask[q.z]) created[q.z] = 1;
                    if (fill_mask[q.w]) created[q.w] = 1;
                }
            }
        }
        if (lane == 0)
        {
            double c = m[2] ? 0.5 * (dii + djj) : (dii + djj) / 2.0;
            __stcg(vals + sii, c); __stcg(vals + sjj, c);
            double cb = 0.5 * (bi + sgn * bj);
            __stcg(b + ni, cb); __stcg(b + nj, sgn * cb);
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) atomicAdd(&done, 1);
    }
}

// Long dependency chains, worked through in shared memory.  Pairs that share matrix entries must be applied in list
// order (the averaging does not commute), and neighbouring pairs along a periodic boundary do share entries, so a
// boundary is one chain of up to thousands of pairs; through global memory every link of it costs a round trip to L2
// (~0.7 us per pair).  Here the host cuts every chain (= connected component of the sharing relation; different chains
// are independent and get a block each) into segments whose entries fit shared memory: the block loads the entries of
// a segment, ONE warp applies its pairs in order out of shared memory (~50 cycles per pair), the block writes them
// back.  Same arithmetic, same order, same value-dependent fill rule as k_periodic / k_periodic_turn.
constexpr int PBC_SEG_SLOTS = 8192, PBC_SEG_ITEMS = 4096, PBC_SEG_PAIRS = 1024;
constexpr size_t PBC_SMEM_BYTES = (size_t)PBC_SEG_SLOTS * 9 + (size_t)PBC_SEG_ITEMS * 8 + (size_t)PBC_SEG_PAIRS * 16;
__global__ void __launch_bounds__(1024)
k_periodic_chains(const int *__restrict__ chain_seg, const int *__restrict__ seg_slot, const int *__restrict__ seg_gid,
                  const int *__restrict__ seg_pair, const int *__restrict__ seg_item, const int4 *__restrict__ pair_tab,
                  const ushort4 *__restrict__ sitems, long long nnz, double *vals, double *b,
                  const unsigned char *__restrict__ fill_mask, unsigned char *created)
{
    extern __shared__ __align__(16) unsigned char pbc_smem[];
    double *sv = (double *)pbc_smem;                                        // [PBC_SEG_SLOTS] values
    ushort4 *it = (ushort4 *)(pbc_smem + (size_t)PBC_SEG_SLOTS * 8);        // [PBC_SEG_ITEMS] local slots of an item
    int4 *pt = (int4 *)(pbc_smem + (size_t)PBC_SEG_SLOTS * 8 + (size_t)PBC_SEG_ITEMS * 8); // [PBC_SEG_PAIRS]
    unsigned char *cr = pbc_smem + (size_t)PBC_SEG_SLOTS * 8 + (size_t)PBC_SEG_ITEMS * 8 + (size_t)PBC_SEG_PAIRS * 16; // written?
    const int lane = threadIdx.x & 31;
    for (int sg = chain_seg[blockIdx.x]; sg < chain_seg[blockIdx.x + 1]; sg++)
    {
        const int g0 = seg_slot[sg], ns = seg_slot[sg + 1] - g0;
        const int p0 = seg_pair[sg], np = seg_pair[sg + 1] - p0;
        const int i0 = seg_item[sg], ni = seg_item[sg + 1] - i0;
        for (int t = threadIdx.x; t < ns; t += blockDim.x)
        {
            const int gid = seg_gid[g0 + t];
            sv[t] = gid < nnz ? __ldcg(vals + gid) : __ldcg(b + (gid - nnz));
            cr[t] = 0;
        }
        for (int t = threadIdx.x; t < ni; t += blockDim.x) it[t] = sitems[i0 + t];
        for (int t = threadIdx.x; t < np; t += blockDim.x) pt[t] = pair_tab[p0 + t];
        __syncthreads();
        if (threadIdx.x < 32)
        {
            for (int pr = 0; pr < np; pr++)
            {
                const int4 m = pt[pr];
                const int a0 = m.x & 0xffff, a1 = (m.x >> 16) & 0xffff;
                const double sgn = m.w ? -1.0 : 1.0;
                for (int q = a0 + lane; q < a1; q += 32)
                {
                    const ushort4 s4 = it[q];
                    const double v1 = sv[s4.x], v2 = sv[s4.z];
                    if (v1 != 0.0 || v2 != 0.0)
                    {
                        const double c = (v1 + sgn * v2) / 2.0;
                        sv[s4.x] = c; sv[s4.y] = c; sv[s4.z] = sgn * c; sv[s4.w] = sgn * c;
                        cr[s4.x] = 1; cr[s4.y] = 1; cr[s4.z] = 1; cr[s4.w] = 1;
                    }
                }
                if (lane == 0)
                {
                    const int la = m.y & 0xffff, lb = (m.y >> 16) & 0xffff, lbi = m.z & 0xffff, lbj = (m.z >> 16) & 0xffff;
                    const double dii = sv[la], djj = sv[lb];
                    const double c = m.w ? 0.5 * (dii + djj) : (dii + djj) / 2.0;
                    sv[la] = c; sv[lb] = c;
                    const double cb = 0.5 * (sv[lbi] + sgn * sv[lbj]);
                    sv[lbi] = cb; sv[lbj] = sgn * cb;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < ns; t += blockDim.x)
        {
            const int gid = seg_gid[g0 + t];
            if (gid < nnz)
            {
                __stcg(vals + gid, sv[t]);
                if (cr[t] && fill_mask && fill_mask[gid]) created[gid] = 1;
            }
            else __stcg(b + (gid - nnz), sv[t]);
        }
        __threadfence_block();
        __syncthreads(); // the next segment of the chain may load entries this one wrote
    }
}

// Periodic pairs at the granularity of their UNITS.  The averaging of a pair (spars.cpp:366-474) is a set of
// independent small operations: one per coupled row k (two reads, four writes: the entries (k,i), (i,k), (k,j), (j,k)),
// one for the two diagonals, one for the two right-hand-side entries.  Pairs have to be applied in list order only
// where they touch the SAME entries, and an entry is touched by the pairs of its row and of its column — two, a few
// more at corners.  So while neighbouring pairs chain a whole boundary together (thousands of dependency levels pair by
// pair), the units need a handful of levels: a unit's level = 1 + the highest level of an earlier unit that shares an
// entry with it.  One launch per level, one thread per unit, the whole device; same arithmetic and, per entry, the
// same order of operations as the reference's list-order loop.
//   flags: bit 0 anti-periodic pair, bit 1 diagonal unit (slots x,z), bit 2 right-hand-side unit (rows x,z)
__global__ void k_periodic_units(int u0, int u1, const int4 *__restrict__ unit_slots, const int *__restrict__ unit_flags,
                                 double *vals, double *b, const unsigned char *__restrict__ fill_mask, unsigned char *created)
{
    const int u = u0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= u1) return;
    const int4 q = unit_slots[u];
    const int f = unit_flags[u];
    const double sgn = (f & 1) ? -1.0 : 1.0;
    if (f & 4)
    {
        const double cb = 0.5 * (b[q.x] + sgn * b[q.z]);
        b[q.x] = cb; b[q.z] = sgn * cb;
        return;
    }
    if (f & 2)
    {
        const double dii = vals[q.x], djj = vals[q.z];
        const double c = (f & 1) ? 0.5 * (dii + djj) : (dii + djj) / 2.0;
        vals[q.x] = c; vals[q.z] = c;
        return;
    }
    const double v1 = vals[q.x], v2 = vals[q.z];
    if (v1 != 0.0 || v2 != 0.0)
    {
        const double c = (v1 + sgn * v2) / 2.0;
        vals[q.x] = c; vals[q.y] = c; vals[q.z] = sgn * c; vals[q.w] = sgn * c;
        if (fill_mask)
        {
            if (fill_mask[q.x]) created[q.x] = 1;
            if (fill_mask[q.y]) created[q.y] = 1;
            if (fill_mask[q.z]) created[q.z] = 1;
            if (fill_mask[q.w]) created[q.w] = 1;
        }
    }
}

// Newton bookkeeping (static2d.cpp:953-989)
__global__ void k_newton_norms(int n, const double *__restrict__ V, const double *__restrict__ Vold, double *part)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        double v = V[i], d = v - Vold[i];
        a = fma(d, d, a);
        c = fma(v, v, c);
    }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = a; part[gridDim.x + blockIdx.x] = c; }
}
__global__ void k_relax(int n, double *V, const double *__restrict__ Vold, double relax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) V[i] = relax * V[i] + (1.0 - relax) * Vold[i];
}

// =====================================================================================
// host side
// =====================================================================================
int linprob_build_layout_device(xfemm_linprob *p, int sigma, int wide_thr, bool use_coords); // layout_device.cu
int linprob_check_layout_device(xfemm_linprob *p, const SellLayout &H);

int linprob_build_layout(xfemm_linprob *p, int precond)
{
    if (precond == XFEMM_PRECOND_AMG) precond = XFEMM_PRECOND_JACOBI; // same uncoloured layout
    if (p->layout_precond == precond) return XFEMM_OK;
    delete p->amg; // its transfer maps are tied to the internal row order
    p->amg = nullptr;
    p->tile_grid = -1; // and so are the stream tiles
    SellLayout &L = p->lay;
    int sigma = env_int("XFEMM_SELL_SIGMA", 1024);
    const int wide_thr = p->wide_threshold >= 0 ? p->wide_threshold : env_int("XFEMM_WIDE_ROW", 64);
    cudaStream_t st = p->stream;
    // the uncoloured layout is built on the device from the resident CSR pattern (layout_device.cu); the host
    // builder remains for the multicolour one and as the reference of the XFEMM_LAYOUT_CHECK self-test
    const bool on_device = precond != XFEMM_PRECOND_MCSSOR && !env_int("XFEMM_HOST_LAYOUT", 0);
    const bool check = on_device && env_int("XFEMM_LAYOUT_CHECK", 0);
    const bool use_coords = env_int("XFEMM_MORTON", 1) != 0;
    if (on_device)
    {
        int rc = linprob_build_layout_device(p, sigma, wide_thr, use_coords);
        if (rc != XFEMM_OK) return rc;
    }
    if (!on_device || check)
    {
        SellLayout Hc;
        SellLayout &H = check ? Hc : L;
        // with node coordinates the internal order follows a Z-curve through the plane (gather locality)
        std::vector<int> zorder;
        if ((int)p->h_xy.size() == 2 * p->n && use_coords) morton_order(p->n, p->h_xy.data(), zorder);
        const int *base = zorder.empty() ? nullptr : zorder.data();
        if ((int)p->forced_order.size() == p->n) { base = p->forced_order.data(); sigma = 1; } // taken as is: no window sort
        if (precond == XFEMM_PRECOND_MCSSOR)
        {
            std::vector<int> color;
            int nc = greedy_coloring(p->n, p->h_rowptr.data(), p->h_col.data(), env_int("XFEMM_MAX_COLORS", 7), color);
            build_sell_layout(p->n, p->h_rowptr.data(), p->h_col.data(), &color, nc, sigma, H, wide_thr, base);
        }
        else
            build_sell_layout(p->n, p->h_rowptr.data(), p->h_col.data(), nullptr, 0, sigma, H, wide_thr, base);
        if (check)
        {
            int rc = linprob_check_layout_device(p, H);
            if (rc != XFEMM_OK) return rc;
        }
        else
        {
            XB_CUDA(p->i2c.upload(L.i2c.data(), L.i2c.size(), st));
            XB_CUDA(p->c2i.upload(L.c2i.data(), L.c2i.size(), st));
            XB_CUDA(p->slice_ptr.upload(L.slice_ptr.data(), L.slice_ptr.size(), st));
            XB_CUDA(p->slice_wl.upload(L.slice_wl.data(), L.slice_wl.size(), st));
            XB_CUDA(p->slice_wu.upload(L.slice_wu.data(), L.slice_wu.size(), st));
            XB_CUDA(p->sell_col.upload(L.sell_col.data(), L.sell_col.size(), st));
            XB_CUDA(p->sell_src.upload(L.sell_src.data(), L.sell_src.size(), st));
            XB_CUDA(p->diag_src.upload(L.diag_src.data(), L.diag_src.size(), st));
            XB_CUDA(p->color_slice.upload(L.color_slice.data(), L.color_slice.size(), st));
            XB_CUDA(p->wide.upload(L, st));
        }
    }
    XB_CUDA(p->sell_val.alloc(L.sell_len));
    for (auto *v : {&p->diag, &p->dinv, &p->x, &p->r, &p->p, &p->u, &p->z, &p->w, &p->bint})
    {
        XB_CUDA(v->alloc(L.n_pad));
        XB_CUDA(v->zero(st));
    }
    XB_CUDA(cudaStreamSynchronize(st));
    // the big host vectors are no longer needed
    std::vector<int>().swap(L.sell_col);
    std::vector<int>().swap(L.sell_src);
    std::vector<int>().swap(L.diag_src);
    std::vector<int>().swap(L.wide_col);
    std::vector<int>().swap(L.wide_src);
    p->layout_precond = precond;
    return XFEMM_OK;
}

int linprob_refresh_values(xfemm_linprob *p, int *singular_row);
static int refresh_values(xfemm_linprob *p, int *singular_row) { return linprob_refresh_values(p, singular_row); }
int linprob_refresh_values(xfemm_linprob *p, int *singular_row)
{
    const SellLayout &L = p->lay;
    cudaStream_t st = p->stream;
    int *d_flag = (int *)(p->d_mail.p + 8);
    int big = 0x7fffffff;
    XB_CUDA(cudaMemcpyAsync(d_flag, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    int blocks = (int)std::min<long long>((L.sell_len + 255) / 256, (long long)p->num_sms * 16);
    if (blocks < 1) blocks = 1;
    ++xb::g_launch, k_sell_fill<<<blocks, 256, 0, st>>>(L.sell_len, p->sell_src.p, p->vals.p, p->sell_val.p);
    p->wide.fill(p->vals.p, st);
    ++xb::g_launch, k_diag_fill<<<(L.n_pad + 255) / 256, 256, 0, st>>>(L.n_pad, p->diag_src.p, p->vals.p, p->diag.p, p->dinv.p, d_flag);
    int flag = 0;
    XB_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    *singular_row = (flag == big) ? -1 : flag;
    return XFEMM_OK;
}

// the same on a stream of the caller's choice, without the zero-diagonal report and without waiting (multigrid levels:
// amg_refresh queues all of them behind each other)
int linprob_refresh_values_on(xfemm_linprob *p, cudaStream_t st)
{
    const SellLayout &L = p->lay;
    int *d_flag = (int *)(p->d_mail.p + 8);
    int blocks = (int)std::min<long long>((L.sell_len + 255) / 256, (long long)p->num_sms * 16);
    if (blocks < 1) blocks = 1;
    ++xb::g_launch, k_sell_fill<<<blocks, 256, 0, st>>>(L.sell_len, p->sell_src.p, p->vals.p, p->sell_val.p);
    p->wide.fill(p->vals.p, st);
    ++xb::g_launch, k_diag_fill<<<(L.n_pad + 255) / 256, 256, 0, st>>>(L.n_pad, p->diag_src.p, p->vals.p, p->diag.p, p->dinv.p, d_flag);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

template <typename K>
static int coop_grid(xfemm_linprob *p, K kernel, int threads, int *grid_out, size_t dyn_smem = 0)
{
    int per_sm = 0;
    if (dyn_smem) XB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    XB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem));
    if (per_sm < 1) { set_error("cooperative kernel does not fit on an SM"); return XFEMM_ERR_CUDA; }
    per_sm = std::min(per_sm, env_int("XFEMM_PCG_BLOCKS_PER_SM", 2));
    int grid = p->num_sms * per_sm;
    // no point in more warps than slices
    int warps_per_block = threads / 32;
    int need = (p->lay.nslices + warps_per_block - 1) / warps_per_block;
    grid = std::max(1, std::min(grid, need));
    if (grid > p->max_grid)
    {
        XB_CUDA(p->partials.alloc((size_t)4 * grid));
        p->max_grid = grid;
    }
    *grid_out = grid;
    return XFEMM_OK;
}

PcgArgs linprob_make_args(xfemm_linprob *p, double precision, int warm, int max_iter)
{
    PcgArgs a;
    a.n_pad = p->lay.n_pad; a.nslices = p->lay.nslices; a.ncolors = p->lay.ncolors;
    a.slice_ptr = p->slice_ptr.p; a.slice_wl = p->slice_wl.p; a.slice_wu = p->slice_wu.p; a.sell_col = p->sell_col.p;
    a.color_slice = p->color_slice.p;
    a.sell_val = p->sell_val.p; a.diag = p->diag.p; a.dinv = p->dinv.p; a.b = p->bint.p;
    a.x = p->x.p; a.r = p->r.p; a.p = p->p.p; a.u = p->u.p; a.z = p->z.p; a.w = p->w.p;
    a.part = p->partials.p;
    a.precision = precision;
    const char *om = getenv("XFEMM_SSOR_OMEGA");
    a.omega = om ? atof(om) : 1.2;
    a.max_iter = max_iter; a.warm = warm;
    a.mail = p->d_mail.p;
    a.wide = p->wide.view(); a.wide_val = p->wide.val.p; a.wide_acc = p->wide.acc.p;
    a.tiles = nullptr; a.tile_ptr = nullptr;
    return a;
}

// Slice-aligned tiles (<= STAGE_TILE_ENT entries) of every block's chunk of the SELL stream for a
// launch with `grid` blocks (stage.cuh).  Cached per layout and grid size.  Returns false when the
// staging does not apply (a block would need more tiles than its table holds, or it is disabled).
// first / end: the range of slices the launch works on (a rank's share of a partitioned solve, dist_amg.cu); -1: all
static bool ensure_tiles(xfemm_linprob *p, int grid, int first = -1, int end = -1)
{
    if (env_int("XFEMM_STAGED", 1) == 0) return false;
    if (first < 0) { first = 0; end = p->lay.nslices; }
    if (p->tile_grid == grid && p->tile_first == first && p->tile_end == end) return p->tiles.n > 0;
    p->tile_grid = grid; p->tile_first = first; p->tile_end = end;
    p->tiles.release(); p->tile_ptr.release();
    const std::vector<int> &sp = p->lay.slice_ptr;
    const int ns = end - first;
    std::vector<int4> tiles;
    std::vector<int> tptr(grid + 1, 0);
    for (int b = 0; b < grid; b++)
    {
        const int s0 = first + (int)((long long)ns * b / grid), s1 = first + (int)((long long)ns * (b + 1) / grid);
        int s = s0;
        while (s < s1)
        {
            int e = s + 1;
            if (sp[e] - sp[s] > STAGE_TILE_ENT) return false; // a slice wider than a stage (coloured layouts)
            while (e < s1 && sp[e + 1] - sp[s] <= STAGE_TILE_ENT) e++;
            tiles.push_back(make_int4(s, e, sp[s], sp[e]));
            s = e;
        }
        tptr[b + 1] = (int)tiles.size();
    }
    if (tiles.empty()) return false;
    if (p->tiles.upload(tiles.data(), tiles.size(), p->stream) != cudaSuccess) return false;
    if (p->tile_ptr.upload(tptr.data(), tptr.size(), p->stream) != cudaSuccess) { p->tiles.release(); return false; }
    return true;
}

// y = A x on the solver representation (wide rows first, then the SELL slices)
static void launch_spmv(xfemm_linprob *p, const double *x, double *y)
{
    const int sb = std::max(1, std::min((p->lay.nslices + 7) / 8, p->num_sms * 8));
    p->wide.pre(x, y, p->stream);
    ++g_launch, k_spmv<<<sb, 256, 0, p->stream>>>(p->lay.nslices, p->slice_ptr.p, p->sell_col.p, p->sell_val.p, p->diag.p, x, y,
                                                  p->wide.n_rows ? p->wide.mask.p : nullptr);
}

} // namespace xb

using namespace xb;

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

const char *xfemm_last_error(void) { return xb::last_error(); }
const char *xfemm_version(void) { return "xfemm_b200 0.1 (sm_100a)"; }

int xfemm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// A solve of a multi-million-node problem allocates and frees a few hundred MB of host arrays (mesh tables, pattern,
// layouts).  glibc hands allocations of that size to mmap and gives them back on free, so every solve pays the
// first-touch page faults again: ~0.5 us per 4 KB page, tens of milliseconds per solve at 2 M nodes.  Keeping big
// blocks in the heap lets the next solve of the process reuse warm pages.  Process-wide setting: XFEMM_NO_MALLOPT=1
// leaves the allocator alone.
} // extern "C"
namespace xb {
void tune_host_allocator()
{
    static std::once_flag once;
    std::call_once(once, [] {
        if (getenv("XFEMM_NO_MALLOPT")) return;
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, 1 << 30);
        mallopt(M_TOP_PAD, 64 << 20);
    });
}
} // namespace xb
extern "C" {

int xfemm_cuda_init(int device)
{
    xb::tune_host_allocator();
    int n = xfemm_device_count();
    if (n <= 0) { set_error("no CUDA device available: the xfemm_b200 backend has no CPU fallback"); return XFEMM_ERR_CUDA; }
    XB_ARG(device >= 0 && device < n, "device index out of range");
    XB_CUDA(cudaSetDevice(device));
    XB_CUDA(cudaFree(0));
    return XFEMM_OK;
}

int xfemm_linprob_create(int n, const int *rowptr, const int *colidx, xfemm_linprob **out)
{
    XB_ARG(n > 0 && rowptr && colidx && out, "n, rowptr, colidx, out");
    XB_ARG(rowptr[0] == 0, "rowptr[0] must be 0");
    xb::tune_host_allocator();
    if (xfemm_device_count() <= 0) { set_error("no CUDA device available: the xfemm_b200 backend has no CPU fallback"); return XFEMM_ERR_CUDA; }
    const long long nnz = rowptr[n];
    // validate: sorted columns, diagonal present
    {
        std::atomic<int> bad_cols(n), bad_diag(n); // lowest offending row of either kind (host threads)
        auto lower = [](std::atomic<int> &a, int v) { int cur = a.load(); while (v < cur && !a.compare_exchange_weak(cur, v)) {} };
        parallel_for(n, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++)
            {
                if (rowptr[i + 1] < rowptr[i]) { lower(bad_cols, i); return; }
                bool diag = false;
                for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
                {
                    const int c = colidx[s];
                    if (c < 0 || c >= n || (s > rowptr[i] && colidx[s - 1] >= c)) { lower(bad_cols, i); return; }
                    if (c == i) diag = true;
                }
                if (!diag) { lower(bad_diag, i); return; }
            }
        }, 65536);
        const int bc = bad_cols.load(), bd = bad_diag.load();
        if (bc < n && bc <= bd) { set_error("row %d: columns must be sorted, unique and in range", bc); return XFEMM_ERR_ARG; }
        if (bd < n) { set_error("row %d has no diagonal entry", bd); return XFEMM_ERR_ARG; }
    }
    xfemm_linprob *p = new xfemm_linprob;
    p->n = n; p->nnz = nnz;
    int rc = XFEMM_OK;
    auto fail = [&](int code) { delete p; return code; };
    if (cudaGetDevice(&p->device) != cudaSuccess) { set_error("cudaGetDevice failed"); return fail(XFEMM_ERR_CUDA); }
    // attribute queries, not cudaGetDeviceProperties: that call costs tens of milliseconds and every
    // multigrid level is a linear problem of its own
    int coop = 0;
    if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->device) != cudaSuccess ||
        cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess)
    {
        set_error("cannot query the device");
        return fail(XFEMM_ERR_CUDA);
    }
    if (!coop) { set_error("device lacks cooperative launch"); return fail(XFEMM_ERR_CUDA); }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream create failed"); return fail(XFEMM_ERR_CUDA); }
    p->h_rowptr.assign(rowptr, rowptr + n + 1);
    p->h_col.assign(colidx, colidx + nnz);
    linprob_start_aggregation(p); // host thread: first coarsening of a later multigrid set-up, beside the uploads below
    auto body = [&]() -> int {
        XB_CUDA(p->rowptr.upload(rowptr, n + 1, p->stream));
        XB_CUDA(p->col.upload(colidx, nnz, p->stream));
        XB_CUDA(p->vals.alloc(nnz)); XB_CUDA(p->vals.zero(p->stream));
        XB_CUDA(p->b.alloc(n)); XB_CUDA(p->b.zero(p->stream));
        XB_CUDA(p->V.alloc(n)); XB_CUDA(p->V.zero(p->stream));
        XB_CUDA(p->Vold.alloc(n)); XB_CUDA(p->Vold.zero(p->stream));
        XB_CUDA(p->d_mail.alloc(16)); XB_CUDA(p->d_mail.zero(p->stream));
        XB_CUDA(p->dir_idx.alloc(n));
        ++xb::g_launch, k_fill_int<<<(n + 255) / 256, 256, 0, p->stream>>>(n, p->dir_idx.p, -1);
        XB_CUDA(cudaStreamSynchronize(p->stream));
        return XFEMM_OK;
    };
    rc = body();
    if (rc != XFEMM_OK) { xfemm_linprob_destroy(p); return rc; }
    *out = p;
    return XFEMM_OK;
}

int xfemm_linprob_destroy(xfemm_linprob *p)
{
    if (!p) return XFEMM_OK;
    if (p->stream) cudaStreamSynchronize(p->stream);
    delete p->amg;
    p->amg = nullptr;
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return XFEMM_OK;
}

int xfemm_linprob_n(const xfemm_linprob *p) { return p ? p->n : 0; }
int xfemm_linprob_get_pattern(xfemm_linprob *p, int *rowptr, int *colidx)
{
    XB_ARG(p, "null handle");
    if (rowptr) std::copy(p->h_rowptr.begin(), p->h_rowptr.end(), rowptr);
    if (colidx) std::copy(p->h_col.begin(), p->h_col.end(), colidx);
    return XFEMM_OK;
}
long long xfemm_linprob_nnz(const xfemm_linprob *p) { return p ? p->nnz : 0; }

int xfemm_linprob_wipe(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    XB_CUDA(p->vals.zero(p->stream));
    XB_CUDA(p->b.zero(p->stream));
    return XFEMM_OK;
}

// values + right-hand side as they are now (after the boundary conditions) kept on the device / put back
int xfemm_linprob_snapshot(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    XB_CUDA(p->snap_vals.alloc((size_t)p->nnz));
    XB_CUDA(p->snap_b.alloc((size_t)p->n));
    XB_CUDA(cudaMemcpyAsync(p->snap_vals.p, p->vals.p, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    XB_CUDA(cudaMemcpyAsync(p->snap_b.p, p->b.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    return XFEMM_OK;
}
int xfemm_linprob_restore(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    XB_ARG((long long)p->snap_vals.n == p->nnz && (int)p->snap_b.n == p->n, "no snapshot taken on this problem");
    XB_CUDA(cudaMemcpyAsync(p->vals.p, p->snap_vals.p, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    XB_CUDA(cudaMemcpyAsync(p->b.p, p->snap_b.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    return XFEMM_OK;
}

#define XB_COPY_IN(buf, src, cnt)                                                                       \
    XB_ARG(p && (src), "null pointer");                                                                    \
    xb::g_h2d_bytes += (long long)(cnt) * 8;                                                               \
    XB_CUDA(xb::copy_h2d((buf).p, (src), (size_t)(cnt) * sizeof(double), p->stream));                       \
    XB_CUDA(cudaStreamSynchronize(p->stream));                                                             \
    return XFEMM_OK
#define XB_COPY_OUT(buf, dst, cnt)                                                                       \
    XB_ARG(p && (dst), "null pointer");                                                                    \
    xb::g_d2h_bytes += (long long)(cnt) * 8;                                                               \
    XB_CUDA(xb::copy_d2h((dst), (buf).p, (size_t)(cnt) * sizeof(double), p->stream));                       \
    return XFEMM_OK

int xfemm_linprob_set_values(xfemm_linprob *p, const double *v) { XB_COPY_IN(p->vals, v, p->nnz); }
int xfemm_linprob_get_values(xfemm_linprob *p, double *v) { XB_COPY_OUT(p->vals, v, p->nnz); }
int xfemm_linprob_set_rhs(xfemm_linprob *p, const double *b) { XB_COPY_IN(p->b, b, p->n); }
int xfemm_linprob_get_rhs(xfemm_linprob *p, double *b) { XB_COPY_OUT(p->b, b, p->n); }
int xfemm_linprob_set_solution(xfemm_linprob *p, const double *V) { XB_COPY_IN(p->V, V, p->n); }
int xfemm_linprob_get_solution(xfemm_linprob *p, double *V) { XB_COPY_OUT(p->V, V, p->n); }

int xfemm_linprob_dirichlet(xfemm_linprob *p, int count, const int *nodes, const double *x_first,
                            const double *x_last, int bdw)
{
    XB_ARG(p, "null handle");
    if (count <= 0) return XFEMM_OK;
    XB_ARG(nodes && x_first && x_last, "null arrays");
    for (int k = 0; k < count; k++) XB_ARG(nodes[k] >= 0 && nodes[k] < p->n, "Dirichlet node out of range");
    cudaStream_t st = p->stream;
    XB_CUDA(p->dir_nodes.upload(nodes, count, st));
    XB_CUDA(p->dir_first.upload(x_first, count, st));
    XB_CUDA(p->dir_last.upload(x_last, count, st));
    ++xb::g_launch, k_dir_scatter<<<(count + 255) / 256, 256, 0, st>>>(count, p->dir_nodes.p, p->dir_idx.p);
    ++xb::g_launch, k_dirichlet<<<(p->n + 255) / 256, 256, 0, st>>>(p->n, p->rowptr.p, p->col.p, p->vals.p, p->b.p, p->dir_idx.p,
                                                    p->dir_first.p, p->dir_last.p, bdw);
    // restore the scratch map for the next call
    ++xb::g_launch, k_fill_int<<<(p->n + 255) / 256, 256, 0, st>>>(p->n, p->dir_idx.p, -1);
    XB_CUDA(cudaGetLastError());
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

static inline int find_slot(const xfemm_linprob *p, int row, int c)
{
    const int *b = p->h_col.data() + p->h_rowptr[row], *e = p->h_col.data() + p->h_rowptr[row + 1];
    const int *it = std::lower_bound(b, e, c);
    return (it != e && *it == c) ? (int)(it - p->h_col.data()) : -1;
}

int xfemm_linprob_periodic(xfemm_linprob *p, int npairs, const int *ni, const int *nj, const int *type)
{
    XB_ARG(p, "null handle");
    if (npairs <= 0) return XFEMM_OK;
    XB_ARG(ni && nj && type, "null arrays");
    // FNV-1a over the pair list → reuse the plan across Newton steps / sweep angles
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](int v) { h ^= (uint32_t)v; h *= 1099511628211ull; };
    for (int k = 0; k < npairs; k++) { mix(ni[k]); mix(nj[k]); mix(type[k]); }
    if (h != p->pbc_hash || npairs != p->pbc_pairs)
    {
        // ---- build the plan: symbolic replay in list order on the (fixed) pattern
        std::vector<int> item_ptr(npairs + 1, 0), meta((size_t)6 * npairs), items, level(npairs, 0);
        std::unordered_map<long long, int> last_level; // slot / rhs id → last level that touched it
        std::vector<int> ks;
        int nlevels = 0;
        for (int k = 0; k < npairs; k++)
        {
            int i = ni[k], j = nj[k];
            XB_ARG(i >= 0 && i < p->n && j >= 0 && j < p->n && i != j, "periodic pair out of range");
            if (j < i) std::swap(i, j);
            ks.clear();
            for (int s = p->h_rowptr[i]; s < p->h_rowptr[i + 1]; s++) ks.push_back(p->h_col[s]);
            for (int s = p->h_rowptr[j]; s < p->h_rowptr[j + 1]; s++) ks.push_back(p->h_col[s]);
            std::sort(ks.begin(), ks.end());
            ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
            int lv = 0;
            auto touch = [&](long long id) { auto f = last_level.find(id); if (f != last_level.end()) lv = std::max(lv, f->second + 1); };
            size_t first_item = items.size() / 4;
            for (int kk : ks)
            {
                if (kk == i || kk == j) continue;
                int s_ki = find_slot(p, kk, i), s_ik = find_slot(p, i, kk);
                int s_kj = find_slot(p, kk, j), s_jk = find_slot(p, j, kk);
                if (s_ki < 0 || s_ik < 0 || s_kj < 0 || s_jk < 0)
                {
                    set_error("pattern lacks the periodic fill for pair (%d,%d) row %d — build it with xfemm_pattern_build", i, j, kk);
                    return XFEMM_ERR_ARG;
                }
                items.insert(items.end(), {s_ki, s_ik, s_kj, s_jk});
                touch(s_ki); touch(s_ik); touch(s_kj); touch(s_jk);
            }
            int s_ii = find_slot(p, i, i), s_jj = find_slot(p, j, j);
            touch(s_ii); touch(s_jj);
            touch(-1 - (long long)i); touch(-1 - (long long)j);
            for (size_t q = first_item * 4; q < items.size(); q++) last_level[items[q]] = lv;
            last_level[s_ii] = lv; last_level[s_jj] = lv;
            last_level[-1 - (long long)i] = lv; last_level[-1 - (long long)j] = lv;
            level[k] = lv;
            nlevels = std::max(nlevels, lv + 1);
            item_ptr[k + 1] = (int)(items.size() / 4);
            int *m = &meta[(size_t)6 * k];
            m[0] = i; m[1] = j; m[2] = type[k]; m[3] = s_ii; m[4] = s_jj; m[5] = 0;
        }
        // reorder pairs by level (stable → list order inside a level)
        std::vector<int> order(npairs);
        for (int k = 0; k < npairs; k++) order[k] = k;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return level[a] < level[b]; });
        std::vector<int> level_ptr(nlevels + 1, 0), item_ptr2(npairs + 1, 0), meta2((size_t)6 * npairs), items2;
        items2.reserve(items.size());
        for (int q = 0; q < npairs; q++)
        {
            int k = order[q];
            level_ptr[level[k] + 1]++;
            for (int t = item_ptr[k] * 4; t < item_ptr[k + 1] * 4; t++) items2.push_back(items[t]);
            item_ptr2[q + 1] = (int)(items2.size() / 4);
            for (int t = 0; t < 6; t++) meta2[(size_t)6 * q + t] = meta[(size_t)6 * k + t];
        }
        for (int l = 0; l < nlevels; l++) level_ptr[l + 1] += level_ptr[l];
        cudaStream_t st = p->stream;
        XB_CUDA(p->pbc_level_ptr.upload(level_ptr.data(), level_ptr.size(), st));
        XB_CUDA(p->pbc_item_ptr.upload(item_ptr2.data(), item_ptr2.size(), st));
        XB_CUDA(p->pbc_meta.upload(meta2.data(), meta2.size(), st));
        if (items2.empty()) items2.assign(4, 0);
        XB_CUDA(p->pbc_items.upload(items2.data(), items2.size(), st));
        XB_CUDA(cudaStreamSynchronize(st));
        p->pbc_hash = h; p->pbc_pairs = npairs; p->pbc_levels = nlevels;
        p->h_pbc_level_ptr = level_ptr;
        // pairs a pair has to wait for: everything in the earlier levels
        std::vector<int> wait(std::max(1, npairs), 0);
        for (int l = 0; l < nlevels; l++)
            for (int q = level_ptr[l]; q < level_ptr[l + 1]; q++) wait[q] = level_ptr[l];
        XB_CUDA(p->pbc_wait.upload(wait.data(), wait.size(), st));
        XB_CUDA(cudaStreamSynchronize(st));
        // ---- units and their levels (k_periodic_units); list order: k indexes items / meta
        p->pbc_unit_levels = 0;
        if (nlevels > 1 && env_int("XFEMM_PBC_UNITS", 1))
        {
            std::unordered_map<long long, int> last; // entry id (slot, or -1 - row for the right-hand side) -> level of its last unit
            std::vector<int4> us;
            std::vector<int> uf, ul;
            int nul = 0;
            auto unit = [&](int4 slots, int flags, std::initializer_list<long long> ids) {
                int lv = 0;
                for (long long id : ids) { auto f = last.find(id); if (f != last.end()) lv = std::max(lv, f->second + 1); }
                for (long long id : ids) last[id] = lv;
                us.push_back(slots); uf.push_back(flags); ul.push_back(lv);
                nul = std::max(nul, lv + 1);
            };
            for (int k = 0; k < npairs; k++)
            {
                const int *m = &meta[(size_t)6 * k];
                const int anti = m[2] ? 1 : 0;
                for (int t = item_ptr[k]; t < item_ptr[k + 1]; t++)
                {
                    const int *q = &items[(size_t)4 * t];
                    unit(make_int4(q[0], q[1], q[2], q[3]), anti, {q[0], q[1], q[2], q[3]});
                }
                unit(make_int4(m[3], m[3], m[4], m[4]), anti | 2, {m[3], m[4]});
                unit(make_int4(m[0], m[0], m[1], m[1]), anti | 4, {-1 - (long long)m[0], -1 - (long long)m[1]});
            }
            if (nul <= 64)
            {
                std::vector<int> order(us.size());
                for (size_t u = 0; u < us.size(); u++) order[u] = (int)u;
                std::stable_sort(order.begin(), order.end(), [&](int a, int b2) { return ul[a] < ul[b2]; });
                std::vector<int4> us2(us.size());
                std::vector<int> uf2(us.size());
                p->h_pbc_unit_ptr.assign(nul + 1, 0);
                for (size_t u = 0; u < us.size(); u++) { us2[u] = us[order[u]]; uf2[u] = uf[order[u]]; p->h_pbc_unit_ptr[ul[order[u]] + 1]++; }
                for (int l = 0; l < nul; l++) p->h_pbc_unit_ptr[l + 1] += p->h_pbc_unit_ptr[l];
                XB_CUDA(p->pbc_unit_slots.upload(us2.data(), us2.size(), st));
                XB_CUDA(p->pbc_unit_flags.upload(uf2.data(), uf2.size(), st));
                XB_CUDA(cudaStreamSynchronize(st));
                p->pbc_unit_levels = nul;
            }
        }
        p->pbc_chains = 0;
        if (p->pbc_unit_levels == 0 && nlevels > 32 && p->nnz + p->n < (1ll << 31) && env_int("XFEMM_PBC_CHAINS", 1))
        {
            // ---- chains and segments for k_periodic_chains (pairs in LIST order: k indexes ni/nj/type, items, meta)
            auto ids_of = [&](int k, std::vector<long long> &out) {
                out.clear();
                for (int t = item_ptr[k] * 4; t < item_ptr[k + 1] * 4; t++) out.push_back(items[t]);
                const int *m = &meta[(size_t)6 * k];
                out.push_back(m[3]); out.push_back(m[4]);
                out.push_back(p->nnz + m[0]); out.push_back(p->nnz + m[1]); // right-hand side entries
            };
            std::vector<int> parent(npairs);
            for (int k = 0; k < npairs; k++) parent[k] = k;
            std::function<int(int)> root = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
            std::unordered_map<long long, int> first_pair;
            std::vector<long long> ids;
            for (int k = 0; k < npairs; k++)
            {
                ids_of(k, ids);
                for (long long id : ids)
                {
                    auto f = first_pair.find(id);
                    if (f == first_pair.end()) first_pair[id] = k;
                    else { const int a = root(f->second), b2 = root(k); if (a != b2) parent[std::max(a, b2)] = std::min(a, b2); }
                }
            }
            std::vector<int> chain_of(npairs, -1), chain_root;
            std::vector<std::vector<int>> chains;
            for (int k = 0; k < npairs; k++)
            {
                const int r = root(k);
                if (chain_of[r] < 0) { chain_of[r] = (int)chains.size(); chains.emplace_back(); }
                chains[chain_of[r]].push_back(k);
            }
            std::vector<int> chain_seg(1, 0), seg_slot(1, 0), seg_pair(1, 0), seg_item(1, 0), seg_gid;
            std::vector<int4> pair_tab;
            std::vector<ushort4> sitems;
            bool ok = true;
            for (const auto &ch : chains)
            {
                size_t q = 0;
                while (q < ch.size() && ok)
                {
                    std::unordered_map<long long, int> local;
                    std::vector<int> gids;
                    int nitems = 0, npair = 0;
                    auto lid = [&](long long id) {
                        auto f = local.find(id);
                        if (f != local.end()) return f->second;
                        const int v = (int)gids.size();
                        local[id] = v; gids.push_back((int)id);
                        return v;
                    };
                    while (q < ch.size() && npair < PBC_SEG_PAIRS)
                    {
                        const int k = ch[q];
                        const int cnt = item_ptr[k + 1] - item_ptr[k];
                        ids_of(k, ids);
                        size_t fresh = 0;
                        for (long long id : ids) if (!local.count(id)) fresh++;
                        if (npair > 0 && ((int)gids.size() + (int)fresh > PBC_SEG_SLOTS || nitems + cnt > PBC_SEG_ITEMS)) break;
                        if ((int)fresh > PBC_SEG_SLOTS || cnt > PBC_SEG_ITEMS) { ok = false; break; } // one pair does not fit: other kernel
                        const int it0 = nitems;
                        for (int t = item_ptr[k]; t < item_ptr[k + 1]; t++)
                        {
                            ushort4 s4;
                            s4.x = (unsigned short)lid(items[4 * t]); s4.y = (unsigned short)lid(items[4 * t + 1]);
                            s4.z = (unsigned short)lid(items[4 * t + 2]); s4.w = (unsigned short)lid(items[4 * t + 3]);
                            sitems.push_back(s4);
                        }
                        nitems += cnt;
                        const int *m = &meta[(size_t)6 * k];
                        const int la = lid(m[3]), lb = lid(m[4]), lbi = lid(p->nnz + m[0]), lbj = lid(p->nnz + m[1]);
                        pair_tab.push_back(make_int4(it0 | (nitems << 16), la | (lb << 16), lbi | (lbj << 16), m[2]));
                        npair++; q++;
                    }
                    if (!ok) break;
                    seg_gid.insert(seg_gid.end(), gids.begin(), gids.end());
                    seg_slot.push_back((int)seg_gid.size());
                    seg_pair.push_back((int)pair_tab.size());
                    seg_item.push_back((int)sitems.size());
                }
                chain_seg.push_back((int)seg_slot.size() - 1);
            }
            if (ok && !chains.empty())
            {
                if (sitems.empty()) sitems.push_back(ushort4{0, 0, 0, 0});
                XB_CUDA(p->pbc_chain_seg.upload(chain_seg.data(), chain_seg.size(), st));
                XB_CUDA(p->pbc_seg_slot.upload(seg_slot.data(), seg_slot.size(), st));
                XB_CUDA(p->pbc_seg_gid.upload(seg_gid.data(), seg_gid.size(), st));
                XB_CUDA(p->pbc_seg_pair.upload(seg_pair.data(), seg_pair.size(), st));
                XB_CUDA(p->pbc_seg_item.upload(seg_item.data(), seg_item.size(), st));
                XB_CUDA(p->pbc_pair_tab.upload(pair_tab.data(), pair_tab.size(), st));
                XB_CUDA(p->pbc_sitems.upload(sitems.data(), sitems.size(), st));
                XB_CUDA(cudaStreamSynchronize(st));
                XB_CUDA(cudaFuncSetAttribute(k_periodic_chains, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PBC_SMEM_BYTES));
                p->pbc_chains = (int)chains.size();
            }
        }
    }
    if (p->pbc_unit_levels > 0)
    {
        for (int lv = 0; lv < p->pbc_unit_levels; lv++)
        {
            const int u0 = p->h_pbc_unit_ptr[lv], u1 = p->h_pbc_unit_ptr[lv + 1];
            if (u1 <= u0) continue;
            ++xb::g_launch, k_periodic_units<<<(unsigned)((u1 - u0 + 255) / 256), 256, 0, p->stream>>>(
                u0, u1, p->pbc_unit_slots.p, p->pbc_unit_flags.p, p->vals.p, p->b.p, p->fill_mask.n ? p->fill_mask.p : nullptr, p->created.p);
        }
    }
    else if (p->pbc_chains > 0)
        ++xb::g_launch, k_periodic_chains<<<p->pbc_chains, 1024, PBC_SMEM_BYTES, p->stream>>>(
            p->pbc_chain_seg.p, p->pbc_seg_slot.p, p->pbc_seg_gid.p, p->pbc_seg_pair.p, p->pbc_seg_item.p, p->pbc_pair_tab.p,
            p->pbc_sitems.p, p->nnz, p->vals.p, p->b.p, p->fill_mask.n ? p->fill_mask.p : nullptr, p->created.p);
    else if (p->pbc_levels <= 32)
    {
        // few dependency levels (the usual case: neighbouring boundary nodes share an edge): one launch
        // per level, one warp per pair over the whole device instead of a 3 ms latency chain in one block
        for (int lv = 0; lv < p->pbc_levels; lv++)
        {
            const int np = p->h_pbc_level_ptr[lv + 1] - p->h_pbc_level_ptr[lv];
            if (np <= 0) continue;
            ++xb::g_launch, k_periodic<<<(np + 7) / 8, 256, 0, p->stream>>>(p->pbc_levels, p->pbc_level_ptr.p, p->pbc_item_ptr.p,
                                                                          p->pbc_meta.p, p->pbc_items.p, p->vals.p, p->b.p,
                                                                          p->fill_mask.n ? p->fill_mask.p : nullptr, p->created.p, lv);
        }
    }
    else if (env_int("XFEMM_PBC_TURN", 1))
        ++xb::g_launch, k_periodic_turn<<<1, 1024, 0, p->stream>>>(p->pbc_pairs, p->pbc_wait.p, p->pbc_item_ptr.p, p->pbc_meta.p,
                                                                  p->pbc_items.p, p->vals.p, p->b.p,
                                                                  p->fill_mask.n ? p->fill_mask.p : nullptr, p->created.p);
    else
        ++xb::g_launch, k_periodic<<<1, 1024, 0, p->stream>>>(p->pbc_levels, p->pbc_level_ptr.p, p->pbc_item_ptr.p, p->pbc_meta.p,
                                          p->pbc_items.p, p->vals.p, p->b.p,
                                          p->fill_mask.n ? p->fill_mask.p : nullptr, p->created.p, -1);
    XB_CUDA(cudaGetLastError());
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}

// Value-independent part of the solver set-up (SELL layout, stream tiles are cut at launch, multigrid
// hierarchy): depends on the sparsity pattern only, so a caller may run it on a host thread while the
// assembly and boundary-condition kernels work on the same problem (they touch rowptr/col/vals/b/V only).
int xfemm_linprob_prebuild(xfemm_linprob *p, int precond)
{
    XB_ARG(p, "null handle");
    XB_CUDA(cudaSetDevice(p->device)); // a fresh host thread starts on device 0
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    if (precond == XFEMM_PRECOND_AMG && !p->amg && p->layout_precond < 0 && !getenv("XFEMM_NO_OVERLAP"))
    {
        // the layout of the finest level and the first coarsening (aggregation, coarse pattern, coarse level) only
        // share the sparsity pattern: two host threads, joined before the transfer maps need both layouts
        int lrc = XFEMM_OK;
        std::thread lt([&]() {
            lrc = cudaSetDevice(p->device) == cudaSuccess ? linprob_build_layout(p, precond) : XFEMM_ERR_CUDA;
            if (timing) fprintf(stderr, "[xfemm timing]   prebuild: SELL layout n=%d %.3f s (own thread)\n", p->n, now() - t0);
        });
        std::unique_ptr<AmgHierarchy> h;
        int hrc = amg_build(p, h, &lt);
        if (lt.joinable()) lt.join();
        if (timing) fprintf(stderr, "[xfemm timing]   prebuild: layout + multigrid hierarchy %.3f s\n", now() - t0);
        if (lrc != XFEMM_OK) return lrc;
        if (hrc != XFEMM_OK) return hrc;
        p->amg = h.release();
        return XFEMM_OK;
    }
    int rc = linprob_build_layout(p, precond);
    if (timing) fprintf(stderr, "[xfemm timing]   prebuild: SELL layout n=%d %.3f s\n", p->n, now() - t0);
    if (rc != XFEMM_OK || precond != XFEMM_PRECOND_AMG) return rc;
    if (!p->amg)
    {
        const double t1 = now();
        std::unique_ptr<AmgHierarchy> h;
        rc = amg_build(p, h);
        if (rc != XFEMM_OK) return rc;
        p->amg = h.release();
        if (timing) fprintf(stderr, "[xfemm timing]   prebuild: multigrid hierarchy %.3f s\n", now() - t1);
    }
    return XFEMM_OK;
}

static int prepare_solver(xfemm_linprob *p, int precond, int *singular)
{
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = now();
    const bool had_layout = p->layout_precond >= 0, had_amg = p->amg != nullptr;
    int rc = linprob_build_layout(p, precond);
    if (rc != XFEMM_OK) return rc;
    if (timing && !had_layout) fprintf(stderr, "[xfemm timing] SELL layout build n=%d: %.3f s\n", p->n, now() - t0);
    rc = refresh_values(p, singular);
    if (rc != XFEMM_OK || precond != XFEMM_PRECOND_AMG) return rc;
    if (!p->amg)
    {
        std::unique_ptr<AmgHierarchy> h;
        rc = amg_build(p, h);
        if (rc != XFEMM_OK) return rc;
        p->amg = h.release();
        if (timing && !had_amg) fprintf(stderr, "[xfemm timing] AMG hierarchy build n=%d: %.3f s\n", p->n, now() - t0);
    }
    return amg_refresh(p, *p->amg);
}

} // extern "C"  (internal helpers shared with dist.cu follow)

namespace xb {
// layout + SELL refill + diagonal for `precond`; XFEMM_ERR_SINGULAR on a zero diagonal
int linprob_prepare(xfemm_linprob *p, int precond)
{
    int singular = -1;
    int rc = prepare_solver(p, precond, &singular);
    if (rc != XFEMM_OK) return rc;
    if (singular >= 0) { set_error("singular flag tripped at internal row %d of %d", singular, p->n); return XFEMM_ERR_SINGULAR; }
    return XFEMM_OK;
}
// bint = b in internal order ; V = x back in caller order
int linprob_load_rhs_internal(xfemm_linprob *p)
{
    const int n_pad = p->lay.n_pad;
    ++g_launch, k_to_internal<<<(n_pad + 255) / 256, 256, 0, p->stream>>>(n_pad, p->i2c.p, p->b.p, p->bint.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
// z = M^-1 r with the multigrid V-cycle, on the resident vectors of `p` (internal order), asynchronously
// on p's stream.  The hierarchy must have been prepared (linprob_prepare(p, XFEMM_PRECOND_AMG)).
int linprob_apply_amg_resident(xfemm_linprob *p)
{
    if (!p->amg) { set_error("multigrid hierarchy not prepared"); return XFEMM_ERR_ARG; }
    PcgArgs a = linprob_make_args(p, 0, 0, 0);
    AmgArgs g;
    void *kargs[] = {&a, &g};
    if (p->precond_grid <= 0)
    {
        int grid = 0;
        int rc = coop_grid(p, k_precond_only<XFEMM_PRECOND_AMG>, 512, &grid, AMG_TASK_SMEM_BYTES);
        if (rc != XFEMM_OK) return rc;
        p->precond_grid = grid;
    }
    amg_fill_args(p, *p->amg, g, p->precond_grid);
    g.task_smem_offset = 0;
    a.part = p->partials.p;
    ++g_launch;
    XB_CUDA(cudaLaunchCooperativeKernel((void *)k_precond_only<XFEMM_PRECOND_AMG>, dim3(p->precond_grid), dim3(512), kargs, AMG_TASK_SMEM_BYTES, p->stream));
    return XFEMM_OK;
}
// stream tiles of the finest level for a launch with `grid` blocks (stage.cuh); false: use direct loads
bool linprob_ensure_tiles(xfemm_linprob *p, int grid) { return ensure_tiles(p, grid); }
bool linprob_ensure_tiles_range(xfemm_linprob *p, int grid, int first, int end) { return ensure_tiles(p, grid, first, end); }
// the problem's V (caller order) into `dst` (internal order): initial guess of a warm start
int linprob_load_guess_internal(xfemm_linprob *p, double *dst)
{
    const int n_pad = p->lay.n_pad;
    ++g_launch, k_to_internal<<<(n_pad + 255) / 256, 256, 0, p->stream>>>(n_pad, p->i2c.p, p->V.p, dst);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int linprob_store_solution_internal(xfemm_linprob *p)
{
    const int n_pad = p->lay.n_pad;
    ++g_launch, k_from_internal<<<(n_pad + 255) / 256, 256, 0, p->stream>>>(n_pad, p->i2c.p, p->x.p, p->V.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
} // namespace xb

extern "C" {

int xfemm_linprob_pcg(xfemm_linprob *p, double precision, int warm, int precond, int max_iter,
                      xfemm_solve_info *info)
{
    XB_ARG(p, "null handle");
    XB_ARG(precond == XFEMM_PRECOND_JACOBI || precond == XFEMM_PRECOND_MCSSOR || precond == XFEMM_PRECOND_AMG, "unknown preconditioner");
    cudaStream_t st = p->stream;
    cudaEvent_t e0, e1, e2;
    XB_CUDA(cudaEventCreate(&e0)); XB_CUDA(cudaEventCreate(&e1)); XB_CUDA(cudaEventCreate(&e2));
    XB_CUDA(cudaEventRecord(e0, st));
    int singular = -1;
    int rc = prepare_solver(p, precond, &singular);
    if (rc != XFEMM_OK) return rc;
    xfemm_solve_info local;
    memset(&local, 0, sizeof local);
    local.precond = precond;
    local.colors = precond == XFEMM_PRECOND_AMG ? 1 + (int)p->amg->coarse.size() : p->lay.ncolors; // AMG: levels
    if (singular >= 0)
    {
        // reference: "singular flag tripped" → PCGSolve returns false (spars.cpp:245-249)
        set_error("singular flag tripped at internal row %d of %d", singular, p->n);
        local.status = XFEMM_ERR_SINGULAR;
        if (info) *info = local;
        return XFEMM_ERR_SINGULAR;
    }
    const int n_pad = p->lay.n_pad;
    ++xb::g_launch, k_to_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->b.p, p->bint.p);
    if (warm) ++xb::g_launch, k_to_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->V.p, p->x.p);
    if (max_iter <= 0) max_iter = env_int("XFEMM_MAX_ITER", 200000);
    PcgArgs a = linprob_make_args(p, precision, warm ? 1 : 0, max_iter);
    int grid = 0;
    AmgArgs g;
    void *kargs[] = {&a, &g};
    XB_CUDA(cudaEventRecord(e1, st));
    if (precond == XFEMM_PRECOND_JACOBI)
    {
        // The reference's three-phase recurrence (u=Ap | x,r update + z.r | p update).  A
        // Chronopoulos–Gear two-barrier arrangement was tried and dropped: it saves one barrier per
        // iteration but moves 20 n more bytes; measured on B200 at 2 M DOF: 0.103 vs 0.093 ms/iteration.
        const int threads = env_int("XFEMM_PCG_THREADS", 1024);
        rc = coop_grid(p, k_pcg<XFEMM_PRECOND_JACOBI>, threads, &grid, STAGE_SMEM_BYTES);
        if (rc != XFEMM_OK) return rc;
        a.part = p->partials.p;
        if (ensure_tiles(p, grid)) { a.tiles = p->tiles.p; a.tile_ptr = p->tile_ptr.p; }
        ++xb::g_launch;
        XB_CUDA(cudaLaunchCooperativeKernel((void *)k_pcg<XFEMM_PRECOND_JACOBI>, dim3(grid), dim3(threads), kargs, STAGE_SMEM_BYTES, st));
    }
    else if (precond == XFEMM_PRECOND_AMG)
    {
        const int threads = env_int("XFEMM_PCG_THREADS", 1024);
        const size_t smem = STAGE_SMEM_BYTES + AMG_TASK_SMEM_BYTES; // matrix-stream ring, then the task buffer of the coarse levels
        rc = coop_grid(p, k_pcg<XFEMM_PRECOND_AMG>, threads, &grid, smem);
        if (rc != XFEMM_OK) return rc;
        amg_fill_args(p, *p->amg, g, grid);
        g.task_smem_offset = (int)STAGE_SMEM_BYTES;
        a.part = p->partials.p;
        if (ensure_tiles(p, grid)) { a.tiles = p->tiles.p; a.tile_ptr = p->tile_ptr.p; }
        ++xb::g_launch; XB_CUDA(cudaLaunchCooperativeKernel((void *)k_pcg<XFEMM_PRECOND_AMG>, dim3(grid), dim3(threads), kargs, smem, st));
    }
    else
    {
        rc = coop_grid(p, k_pcg<XFEMM_PRECOND_MCSSOR>, 512, &grid);
        if (rc != XFEMM_OK) return rc;
        a.part = p->partials.p;
        ++xb::g_launch; XB_CUDA(cudaLaunchCooperativeKernel((void *)k_pcg<XFEMM_PRECOND_MCSSOR>, dim3(grid), dim3(512), kargs, 0, st));
    }
    XB_CUDA(cudaEventRecord(e2, st));
    XB_CUDA(cudaMemcpyAsync(p->h_mail, p->d_mail.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    if (precond == XFEMM_PRECOND_AMG && g.stamps) amg_print_trace(*p->amg, st);
    if (precond == XFEMM_PRECOND_AMG && g.stamps && g.trace_detail)
    {
        // per block: SELL entries / wide rows of its chunk of the finest level (to set against the per-block times)
        const std::vector<int> &sp = p->lay.slice_ptr;
        fprintf(stderr, "[amg trace]   per block (k entries / wide rows):");
        for (int b = 0; b < grid; b++)
        {
            const int s0 = (int)((long long)p->lay.nslices * b / grid), s1 = (int)((long long)p->lay.nslices * (b + 1) / grid);
            int nw = 0;
            for (int r : p->lay.wide_row) if (r >= s0 * 32 && r < s1 * 32) nw++;
            fprintf(stderr, " %d/%d", (sp[s1] - sp[s0]) / 1000, nw);
        }
        fprintf(stderr, "\n");
    }
    local.iterations = (int)p->h_mail[0];
    local.er = p->h_mail[1];
    local.status = (int)p->h_mail[2];
    const bool trivial = (p->h_mail[3] == 0.0); // res_o == 0 → reference returns without touching V
    if (!trivial)
    {
        ++xb::g_launch, k_from_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->x.p, p->V.p);
        // true relative residual ||b - A x|| / ||b||
        int sb = std::max(1, std::min((p->lay.nslices + 7) / 8, p->num_sms * 8));
        (void)sb;
        launch_spmv(p, p->x.p, p->u.p);
        int nb = std::max(1, std::min((n_pad + 255) / 256, 512));
        if (p->max_grid < nb) { XB_CUDA(p->partials.alloc((size_t)4 * nb)); p->max_grid = nb; }
        ++xb::g_launch, k_resnorm_partial<<<nb, 256, 0, st>>>(n_pad, p->bint.p, p->u.p, p->partials.p);
        ++xb::g_launch, k_reduce2_final<<<1, 256, 0, st>>>(nb, p->partials.p, p->d_mail.p + 10);
        XB_CUDA(cudaMemcpyAsync(p->h_mail + 10, p->d_mail.p + 10, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        XB_CUDA(cudaStreamSynchronize(st));
        XB_CUDA(cudaGetLastError());
        local.relres = p->h_mail[11] > 0 ? sqrt(p->h_mail[10] / p->h_mail[11]) : 0.0;
    }
    float ms_setup = 0, ms_solve = 0;
    cudaEventElapsedTime(&ms_setup, e0, e1);
    cudaEventElapsedTime(&ms_solve, e1, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    local.setup_ms = ms_setup;
    local.solve_ms = ms_solve;
    // algorithmic bytes per iteration, SURVEY §8(d): Krylov part 12 nnz + 124 n, MC-SSOR adds 12 nnz + 40 n
    const double nn = (double)p->n, nz = (double)p->nnz;
    local.bytes_per_iter = 12.0 * nz + 124.0 * nn + (precond == XFEMM_PRECOND_MCSSOR ? 12.0 * nz + 40.0 * nn : 0.0);
    if (info) *info = local;
    if (local.status != XFEMM_OK)
    {
        set_error("PCG did not converge: %d iterations, er=%g", local.iterations, local.er);
        return local.status;
    }
    return XFEMM_OK;
}

int xfemm_linprob_multa(xfemm_linprob *p, const double *x, double *y)
{
    XB_ARG(p && x && y, "null pointer");
    int singular = -1;
    int precond = p->layout_precond >= 0 ? p->layout_precond : XFEMM_PRECOND_JACOBI;
    int rc = prepare_solver(p, precond, &singular);
    if (rc != XFEMM_OK) return rc;
    cudaStream_t st = p->stream;
    const int n_pad = p->lay.n_pad;
    // stage through V-sized scratch: Vold is free between Newton steps only, so use r/u
    XB_CUDA(cudaMemcpyAsync(p->Vold.p, x, (size_t)p->n * sizeof(double), cudaMemcpyHostToDevice, st));
    ++xb::g_launch, k_to_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->Vold.p, p->r.p);
    int sb = std::max(1, std::min((p->lay.nslices + 7) / 8, p->num_sms * 8));
    (void)sb;
    launch_spmv(p, p->r.p, p->u.p);
    ++xb::g_launch, k_from_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->u.p, p->Vold.p);
    XB_CUDA(cudaMemcpyAsync(y, p->Vold.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

int xfemm_linprob_multpc(xfemm_linprob *p, int precond, const double *x, double *y)
{
    XB_ARG(p && x && y, "null pointer");
    int singular = -1;
    int rc = prepare_solver(p, precond, &singular);
    if (rc != XFEMM_OK) return rc;
    cudaStream_t st = p->stream;
    const int n_pad = p->lay.n_pad;
    XB_CUDA(cudaMemcpyAsync(p->Vold.p, x, (size_t)p->n * sizeof(double), cudaMemcpyHostToDevice, st));
    ++xb::g_launch, k_to_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->Vold.p, p->r.p);
    PcgArgs a = linprob_make_args(p, 0, 0, 0);
    int grid = 0;
    AmgArgs g;
    void *kargs[] = {&a, &g};
    if (precond == XFEMM_PRECOND_JACOBI)
    {
        rc = coop_grid(p, k_precond_only<XFEMM_PRECOND_JACOBI>, 512, &grid);
        if (rc != XFEMM_OK) return rc;
        ++xb::g_launch; XB_CUDA(cudaLaunchCooperativeKernel((void *)k_precond_only<XFEMM_PRECOND_JACOBI>, dim3(grid), dim3(512), kargs, 0, st));
    }
    else if (precond == XFEMM_PRECOND_AMG)
    {
        rc = coop_grid(p, k_precond_only<XFEMM_PRECOND_AMG>, 512, &grid, AMG_TASK_SMEM_BYTES);
        if (rc != XFEMM_OK) return rc;
        amg_fill_args(p, *p->amg, g, grid);
        g.task_smem_offset = 0;
        a.part = p->partials.p;
        ++xb::g_launch; XB_CUDA(cudaLaunchCooperativeKernel((void *)k_precond_only<XFEMM_PRECOND_AMG>, dim3(grid), dim3(512), kargs, AMG_TASK_SMEM_BYTES, st));
    }
    else
    {
        rc = coop_grid(p, k_precond_only<XFEMM_PRECOND_MCSSOR>, 512, &grid);
        if (rc != XFEMM_OK) return rc;
        ++xb::g_launch; XB_CUDA(cudaLaunchCooperativeKernel((void *)k_precond_only<XFEMM_PRECOND_MCSSOR>, dim3(grid), dim3(512), kargs, 0, st));
    }
    ++xb::g_launch, k_from_internal<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, p->i2c.p, p->z.p, p->Vold.p);
    XB_CUDA(cudaMemcpyAsync(y, p->Vold.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

int xfemm_linprob_set_coords(xfemm_linprob *p, const double *xy)
{
    XB_ARG(p, "null handle");
    p->d_xy.release();
    if (!xy) { p->h_xy.clear(); return XFEMM_OK; }
    p->h_xy.resize(2 * (size_t)p->n);
    parallel_for(p->n, [&](int i0, int i1, int) {
        for (size_t k = 2 * (size_t)i0; k < 2 * (size_t)i1; k++) p->h_xy[k] = (float)xy[k];
    });
    if (p->layout_precond >= 0)
    {
        // a layout built before the hint arrived is rebuilt on the next solve
        p->layout_precond = -1;
        delete p->amg;
        p->amg = nullptr;
        p->tile_grid = -1;
    }
    return XFEMM_OK;
}

int xfemm_linprob_amg_levels(xfemm_linprob *p, int max_levels, int *n_rows, long long *nnz)
{
    if (!p || !p->amg) return 0;
    const int nl = (int)p->amg->level_n.size();
    for (int l = 0; l < nl && l < max_levels; l++)
    {
        if (n_rows) n_rows[l] = p->amg->level_n[l];
        if (nnz) nnz[l] = p->amg->level_nnz[l];
    }
    return nl;
}

int xfemm_linprob_bench_spmv(xfemm_linprob *p, int reps, double *ms_per_rep)
{
    XB_ARG(p && reps > 0 && ms_per_rep, "arguments");
    int singular = -1;
    int precond = p->layout_precond >= 0 ? p->layout_precond : XFEMM_PRECOND_JACOBI;
    int rc = prepare_solver(p, precond, &singular);
    if (rc != XFEMM_OK) return rc;
    cudaStream_t st = p->stream;
    cudaEvent_t e0, e1;
    XB_CUDA(cudaEventCreate(&e0)); XB_CUDA(cudaEventCreate(&e1));
    int sb = std::max(1, std::min((p->lay.nslices + 7) / 8, p->num_sms * 8));
    for (int w = 0; w < 3; w++)
        launch_spmv(p, p->bint.p, p->u.p);
    XB_CUDA(cudaEventRecord(e0, st));
    for (int w = 0; w < reps; w++)
        launch_spmv(p, p->bint.p, p->u.p);
    XB_CUDA(cudaEventRecord(e1, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms_per_rep = ms / reps;
    return XFEMM_OK;
}

int xfemm_linprob_save_solution(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    XB_CUDA(cudaMemcpyAsync(p->Vold.p, p->V.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    return XFEMM_OK;
}

int xfemm_linprob_restore_solution(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    XB_CUDA(cudaMemcpyAsync(p->V.p, p->Vold.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    return XFEMM_OK;
}

int xfemm_linprob_newton_norms(xfemm_linprob *p, double *dv2, double *v2)
{
    XB_ARG(p && dv2 && v2, "null pointer");
    cudaStream_t st = p->stream;
    int nb = std::max(1, std::min((p->n + 255) / 256, 512));
    if (p->max_grid < nb) { XB_CUDA(p->partials.alloc((size_t)4 * nb)); p->max_grid = nb; }
    ++xb::g_launch, k_newton_norms<<<nb, 256, 0, st>>>(p->n, p->V.p, p->Vold.p, p->partials.p);
    ++xb::g_launch, k_reduce2_final<<<1, 256, 0, st>>>(nb, p->partials.p, p->d_mail.p + 12);
    XB_CUDA(cudaMemcpyAsync(p->h_mail + 12, p->d_mail.p + 12, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    *dv2 = p->h_mail[12];
    *v2 = p->h_mail[13];
    return XFEMM_OK;
}

int xfemm_linprob_relax(xfemm_linprob *p, double relax)
{
    XB_ARG(p, "null handle");
    ++xb::g_launch, k_relax<<<(p->n + 255) / 256, 256, 0, p->stream>>>(p->n, p->V.p, p->Vold.p, relax);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

// ---- solution history for sweeps: the converged solutions of the last two problems solved on this
// handle, and an initial guess for the next one by linear extrapolation (equidistant parameter steps)
__global__ void k_hist_guess(int n, const double *__restrict__ h1, const double *__restrict__ h2,
                             const double *__restrict__ h3, double *V)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (h3) V[i] = 3.0 * (h1[i] - h2[i]) + h3[i]; // quadratic through the last three
    else if (h2) V[i] = 2.0 * h1[i] - h2[i];       // linear through the last two
    else V[i] = h1[i];
}
int xfemm_linprob_history_push(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    if (p->hist1.n != (size_t)p->n)
    {
        XB_CUDA(p->hist1.alloc(p->n)); XB_CUDA(p->hist2.alloc(p->n)); XB_CUDA(p->hist3.alloc(p->n));
        p->hist_count = 0;
    }
    std::swap(p->hist2.p, p->hist3.p); // hist3 <- hist2 <- hist1 <- current
    std::swap(p->hist1.p, p->hist2.p);
    XB_CUDA(cudaMemcpyAsync(p->hist1.p, p->V.p, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
    if (p->hist_count < 3) p->hist_count++;
    return XFEMM_OK;
}
int xfemm_linprob_history_guess(xfemm_linprob *p, int extrapolate)
{
    XB_ARG(p, "null handle");
    if (p->hist_count <= 0) return XFEMM_OK; // nothing known: V stays as it is
    const double *h2 = (extrapolate >= 1 && p->hist_count >= 2) ? p->hist2.p : nullptr;
    const double *h3 = (extrapolate >= 2 && p->hist_count >= 3) ? p->hist3.p : nullptr;
    ++xb::g_launch, k_hist_guess<<<(p->n + 255) / 256, 256, 0, p->stream>>>(p->n, p->hist1.p, h2, h3, p->V.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int xfemm_linprob_history_clear(xfemm_linprob *p)
{
    XB_ARG(p, "null handle");
    p->hist_count = 0;
    return XFEMM_OK;
}

int xfemm_linprob_set_fill_mask(xfemm_linprob *p, const unsigned char *fill)
{
    XB_ARG(p && fill, "null pointer");
    XB_CUDA(p->fill_mask.upload(fill, (size_t)p->nnz, p->stream));
    XB_CUDA(p->created.alloc((size_t)p->nnz));
    XB_CUDA(p->created.zero(p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}

int xfemm_linprob_get_exists_mask(xfemm_linprob *p, unsigned char *exists)
{
    XB_ARG(p && exists, "null pointer");
    if (!p->fill_mask.n) { memset(exists, 1, (size_t)p->nnz); return XFEMM_OK; }
    std::vector<unsigned char> f((size_t)p->nnz), c((size_t)p->nnz);
    XB_CUDA(cudaMemcpyAsync(f.data(), p->fill_mask.p, f.size(), cudaMemcpyDeviceToHost, p->stream));
    XB_CUDA(cudaMemcpyAsync(c.data(), p->created.p, c.size(), cudaMemcpyDeviceToHost, p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    for (size_t s = 0; s < f.size(); s++) exists[s] = (!f[s] || c[s]) ? 1 : 0;
    return XFEMM_OK;
}

void xfemm_transfer_bytes(long long *h2d, long long *d2h, int reset)
{
    if (h2d) *h2d = xb::g_h2d_bytes.load();
    if (d2h) *d2h = xb::g_d2h_bytes.load();
    if (reset) { xb::g_h2d_bytes = 0; xb::g_d2h_bytes = 0; }
}
long long xfemm_launch_count(void) { return xb::g_launch.load(); }

void xfemm_free(void *ptr) { free(ptr); }

int xfemm_host_register(void *ptr, size_t bytes)
{
    XB_ARG(ptr && bytes > 0, "null pointer");
    XB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
    return XFEMM_OK;
}
int xfemm_host_unregister(void *ptr)
{
    XB_ARG(ptr, "null pointer");
    XB_CUDA(cudaHostUnregister(ptr));
    return XFEMM_OK;
}

} // extern "C"
