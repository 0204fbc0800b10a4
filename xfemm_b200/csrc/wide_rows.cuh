// Wide rows of the solver layout (layout.h): rows kept out of the SELL-32 slices and stored
// row-wise in segments of <= 256 entries.  One warp reduces one segment, a second step adds the
// segments of a row in a fixed order (deterministic), so the off-diagonal part of a wide row
// lands in y[row] BEFORE the block's regular SELL pass, which then adds it for the rows flagged
// in the per-slice mask.  Shared by the real (double) and complex (double2) solvers.
#pragma once
#include "common.cuh"
#include "layout.h"
#include <algorithm>

namespace xb {

struct WideView
{
    int n_rows = 0;
    const int *row = nullptr, *rseg = nullptr, *sptr = nullptr, *col = nullptr;
    const unsigned *mask = nullptr; // per slice
};

__device__ __forceinline__ double w_zero(const double *) { return 0.0; }
__device__ __forceinline__ double2 w_zero(const double2 *) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double w_fma(double a, double x, double acc) { return fma(a, x, acc); }
__device__ __forceinline__ double2 w_fma(double2 a, double2 x, double2 acc)
{
    acc.x = fma(a.x, x.x, fma(-a.y, x.y, acc.x));
    acc.y = fma(a.x, x.y, fma(a.y, x.x, acc.y));
    return acc;
}
__device__ __forceinline__ double w_add(double a, double b) { return a + b; }
__device__ __forceinline__ double2 w_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double w_warp_sum(double v) { return warp_sum(v); }
__device__ __forceinline__ double2 w_warp_sum(double2 v) { return make_double2(warp_sum(v.x), warp_sum(v.y)); }

// wide rows whose internal row lies in [r0, r1): index range [k0, k1)
__device__ __forceinline__ void wide_range(const WideView &w, int r0, int r1, int &k0, int &k1)
{
    int lo = 0, hi = w.n_rows;
    while (lo < hi) { int m = (lo + hi) >> 1; if (w.row[m] < r0) lo = m + 1; else hi = m; }
    k0 = lo;
    hi = w.n_rows;
    while (lo < hi) { int m = (lo + hi) >> 1; if (w.row[m] < r1) lo = m + 1; else hi = m; }
    k1 = lo;
}

// y[row] = sum_j a_ij x_j (off-diagonals) for the wide rows [k0,k1), by all threads of the block.
// Ends with the row results written; the caller must __syncthreads() before other warps read them.
template <typename T>
__device__ __forceinline__ void wide_rows_block(const WideView &w, const T *__restrict__ val, T *seg_acc, int k0, int k1,
                                                const T *x, T *y)
{
    if (k0 >= k1) return; // uniform per block
    const int lane = threadIdx.x & 31, warp_b = threadIdx.x >> 5, nwarps_b = blockDim.x >> 5;
    const int sg0 = w.rseg[k0], sg1 = w.rseg[k1];
    for (int sg = sg0 + warp_b; sg < sg1; sg += nwarps_b)
    {
        const int q1 = w.sptr[sg + 1];
        T acc = w_zero((const T *)nullptr);
        // a segment has at most 256 entries: eight per lane, all loads in flight together (see wide_rows_warp)
        for (int qb = w.sptr[sg] + lane; qb < q1; qb += 256)
        {
            T a[8];
            int j[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                const int q = qb + 32 * u;
                const bool ok = q < q1;
                a[u] = ok ? val[q] : w_zero((const T *)nullptr);
                j[u] = ok ? w.col[q] : 0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) acc = w_fma(a[u], x[j[u]], acc);
        }
        acc = w_warp_sum(acc);
        if (lane == 0) seg_acc[sg] = acc;
    }
    __syncthreads();
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x)
    {
        T s = w_zero((const T *)nullptr);
        for (int sg = w.rseg[k]; sg < w.rseg[k + 1]; sg++) s = w_add(s, seg_acc[sg]);
        y[w.row[k]] = s;
    }
}

// The wide rows [k0,k1) dealt out to the warps of the block, last warp first: a warp forms the off-diagonal sum of a
// whole row (its segments are contiguous) and lane 0 hands it to epi(row, sum).  No block-wide step: the other warps
// of the block go on with the regular slices meanwhile (stage.cuh: `before`), which then SKIP the rows flagged in the
// per-slice mask — the wide warp is the only writer of a wide row.
// stride > 1: the rows k0, k0 + stride, ... (the caller deals the wide rows of a whole level out over the blocks —
// k0 = first row of the range + block index, stride = blocks — when the epilogue touches nothing but the row itself:
// the handful of blocks whose chunks hold the air-gap ring would otherwise be the stragglers of every pass).
template <class Epi>
__device__ __forceinline__ void wide_rows_warp(const WideView &w, const double *__restrict__ val, int k0, int k1, const double *x, Epi epi,
                                               int stride = 1)
{
    const int lane = threadIdx.x & 31, warp_b = threadIdx.x >> 5, nwarps_b = blockDim.x >> 5;
    for (int k = k0 + stride * (nwarps_b - 1 - warp_b); k < k1; k += stride * nwarps_b)
    {
        const int row = w.row[k];
        const int q0 = w.sptr[w.rseg[k]], q1 = w.sptr[w.rseg[k + 1]];
        double acc = 0.0;
        // eight entries per lane and step, all loads of a step in flight together (a plain loop with a run-time trip
        // count of six or seven goes through the compiler's remainder loop: one dependent round trip pair per entry)
        for (int qb = q0 + lane; qb < q1; qb += 256)
        {
            double a[8];
            int j[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
            {
                const int q = qb + 32 * u;
                const bool ok = q < q1;
                a[u] = ok ? val[q] : 0.0;
                j[u] = ok ? w.col[q] : row;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) acc = fma(a[u], x[j[u]], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) epi(row, acc);
    }
}

// stand-alone launch for the non-persistent SpMV kernels: block b takes an even share of the wide rows
template <typename T>
__global__ void k_wide_pre(WideView w, const T *__restrict__ val, T *seg_acc, const T *x, T *y)
{
    const int k0 = (int)((long long)w.n_rows * blockIdx.x / gridDim.x);
    const int k1 = (int)((long long)w.n_rows * (blockIdx.x + 1) / gridDim.x);
    wide_rows_block<T>(w, val, seg_acc, k0, k1, x, y);
}

template <typename T>
__global__ void k_wide_fill(int len, const int *__restrict__ src, const T *__restrict__ vals, T *__restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) out[i] = vals[src[i]];
}

// device copy of the wide-row part of a layout + its values
template <typename T>
struct WideDev
{
    DevBuf<int> row, rseg, sptr, col, src;
    DevBuf<unsigned> mask;
    DevBuf<T> val, acc;
    int n_rows = 0, n_seg = 0, len = 0;

    cudaError_t upload(const SellLayout &L, cudaStream_t st)
    {
        n_rows = (int)L.wide_row.size();
        n_seg = n_rows ? (int)L.wide_sptr.size() - 1 : 0;
        len = (int)L.wide_col.size();
        if (!n_rows) { for (auto *b : {&row, &rseg, &sptr, &col, &src}) b->release(); mask.release(); val.release(); acc.release(); return cudaSuccess; }
        cudaError_t e;
        if ((e = row.upload(L.wide_row.data(), L.wide_row.size(), st)) != cudaSuccess) return e;
        if ((e = rseg.upload(L.wide_rseg.data(), L.wide_rseg.size(), st)) != cudaSuccess) return e;
        if ((e = sptr.upload(L.wide_sptr.data(), L.wide_sptr.size(), st)) != cudaSuccess) return e;
        if ((e = col.upload(L.wide_col.data(), L.wide_col.size(), st)) != cudaSuccess) return e;
        if ((e = src.upload(L.wide_src.data(), L.wide_src.size(), st)) != cudaSuccess) return e;
        if ((e = mask.upload(L.slice_wide.data(), L.slice_wide.size(), st)) != cudaSuccess) return e;
        if ((e = val.alloc(len)) != cudaSuccess) return e;
        return acc.alloc(n_seg);
    }
    WideView view() const
    {
        WideView w;
        w.n_rows = n_rows; w.row = row.p; w.rseg = rseg.p; w.sptr = sptr.p; w.col = col.p; w.mask = mask.p;
        return w;
    }
    // gather the values of the wide rows out of the CSR value array
    void fill(const T *vals, cudaStream_t st)
    {
        if (!n_rows) return;
        ++g_launch, k_wide_fill<T><<<(len + 255) / 256, 256, 0, st>>>(len, src.p, vals, val.p);
    }
    // y[row] = off-diagonal part of every wide row (to be followed by a SpMV kernel that adds it)
    void pre(const T *x, T *y, cudaStream_t st)
    {
        if (!n_rows) return;
        int blocks = std::max(1, std::min(n_rows, 592));
        ++g_launch, k_wide_pre<T><<<blocks, 256, 0, st>>>(view(), val.p, acc.p, x, y);
    }
};

} // namespace xb
