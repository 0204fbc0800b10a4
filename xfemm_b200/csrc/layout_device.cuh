// Solver layout (internal row order + SELL-32, layout.h) of the uncoloured kind built ON THE DEVICE from the CSR
// pattern that already lives there.  Same result, array for array, as build_sell_layout() on the host (the
// XFEMM_LAYOUT_CHECK=1 self-test compares them); at 2 M rows the host version costs ~0.12 s on 32 threads plus the
// upload of 2 x 56 MB of slot tables, this one a few milliseconds.  What stays on the host: the handful of wide rows
// (build_wide_part) and small tables the host-side planners read (i2c, c2i, slice_ptr).
//
//   order   = rows along the Z-curve through the node coordinates (stable radix sort by 22-bit cell), or as forced
//   i2c     = order, windows of `sigma` rows stably sorted by descending row length          (radix sort, one key)
//   widths  = longest row of every slice of 32 rows, slice_ptr = exclusive scan x 32
//   slots   = one warp per slice, lane = row; the entries of a row in ascending internal column
#pragma once
#include "common.cuh"
#include "layout.h"
#include "wide_rows.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <vector>

namespace xb {

static __global__ void k_lay_cells(int n, const float *__restrict__ xy, float x0, float y0, float span, unsigned *cell, int *idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto spread = [](unsigned v) -> unsigned long long {
        unsigned long long x = v & 0xffffffffull;
        x = (x | (x << 16)) & 0x0000ffff0000ffffull;
        x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
        x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
        x = (x | (x << 2)) & 0x3333333333333333ull;
        x = (x | (x << 1)) & 0x5555555555555555ull;
        return x;
    };
    const float scale = (float)((1 << 11) - 1);
    // (same float expressions as morton_order(): subtraction, IEEE division, multiplication — nothing to contract)
    const unsigned qx = (unsigned)(__fmul_rn(__fdiv_rn(__fsub_rn(xy[2 * (size_t)i], x0), span), scale));
    const unsigned qy = (unsigned)(__fmul_rn(__fdiv_rn(__fsub_rn(xy[2 * (size_t)i + 1], y0), span), scale));
    cell[i] = (unsigned)(spread(qx) | (spread(qy) << 1));
    idx[i] = i;
}

static __global__ void k_lay_iota(int n, int *idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = i;
}

// key of position k of the base order: window number, then descending row length (wide rows count as empty)
static __global__ void k_lay_window_keys(int n, const int *__restrict__ order, const int *__restrict__ rowptr, int sigma,
                                         int wide_thr, unsigned *key)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int i = order[k];
    int len = rowptr[i + 1] - rowptr[i];
    if (wide_thr > 0 && len - 1 > wide_thr) len = 0;
    key[k] = ((unsigned)(k / sigma) << 12) | (unsigned)(4095 - min(len, 4095));
}

static __global__ void k_lay_numbering(int n, int n_pad, const int *__restrict__ order, int *i2c, int *c2i)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    const int i = r < n ? order[r] : -1;
    i2c[r] = i;
    if (i >= 0) c2i[i] = r;
}

// one warp per slice: width = longest row (off-diagonals, wide rows excluded); also the diagonal's CSR index
static __global__ void k_lay_widths(int nslices, const int *__restrict__ i2c, const int *__restrict__ rowptr,
                                    const int *__restrict__ col, int wide_thr, int *width, int *width32, int *diag_src)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = s * 32 + lane, i = i2c[r];
    int len = 0, dq = -1;
    if (i >= 0)
    {
        const int q0 = rowptr[i], q1 = rowptr[i + 1];
        for (int q = q0; q < q1; q++)
            if (col[q] == i) dq = q; else len++;
        if (wide_thr > 0 && q1 - q0 - 1 > wide_thr) len = 0;
    }
    diag_src[r] = dq;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) { width[s] = len; width32[s] = len * 32; }
    if (s == 0 && lane == 0) width32[nslices] = 0;
}

// one warp per slice, lane = row: padding slots point at the row itself, then every entry goes to the position its
// internal column has among the row's entries
static __global__ void k_lay_slots(int nslices, const int *__restrict__ i2c, const int *__restrict__ c2i,
                                   const int *__restrict__ rowptr, const int *__restrict__ col,
                                   const int *__restrict__ slice_ptr, int wide_thr, int *sell_col, int *sell_src)
{
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5;
    const int r = s * 32 + lane, i = i2c[r];
    for (int k = 0; k < w; k++) { sell_col[base + k * 32 + lane] = r; sell_src[base + k * 32 + lane] = -1; }
    if (i < 0) return;
    const int q0 = rowptr[i], q1 = rowptr[i + 1];
    if (wide_thr > 0 && q1 - q0 - 1 > wide_thr) return;
    constexpr int LOCAL = 32;
    int ic[LOCAL];
    const int nloc = min(q1 - q0, LOCAL);
    for (int e = 0; e < nloc; e++) { const int j = col[q0 + e]; ic[e] = j == i ? 0x7fffffff : c2i[j]; } // (the diagonal never counts)
    for (int q = q0; q < q1; q++)
    {
        const int j = col[q];
        if (j == i) continue;
        const int me = (q - q0 < LOCAL) ? ic[q - q0] : c2i[j];
        int pos = 0;
        for (int e = 0; e < nloc; e++) pos += ic[e] < me ? 1 : 0;
        for (int q2 = q0 + LOCAL; q2 < q1; q2++) pos += (col[q2] != i && c2i[col[q2]] < me) ? 1 : 0;
        sell_col[base + pos * 32 + lane] = me;
        sell_src[base + pos * 32 + lane] = q;
    }
}

template <class T>
static cudaError_t lay_fetch(std::vector<T> &h, const T *d, size_t count, cudaStream_t st)
{
    h.resize(count);
    g_d2h_bytes += (long long)(count * sizeof(T));
    return count ? copy_d2h(h.data(), d, count * sizeof(T), st) : cudaSuccess; // (big ones through the pinned ring)
}

// Fills p->lay (host part) and the device tables i2c, c2i, slice_ptr, slice_wl, slice_wu, sell_col, sell_src, diag_src,
// color_slice, wide of the uncoloured layout of a real or complex problem (same field names).  `sigma`, `wide_thr` as
// for build_sell_layout(); base order: `forced` (taken as is), else the Z-curve through `h_xy` (x,y per row; its device
// copy is left in *d_xy), else the caller's numbering.
template <class Problem>
int build_layout_device(Problem *p, int sigma, int wide_thr, const std::vector<int> *forced, const std::vector<float> *h_xy,
                        DevBuf<float> *d_xy)
{
    const bool have_forced = forced && (int)forced->size() == p->n;
    const bool use_coords = !have_forced && h_xy && d_xy && (int)h_xy->size() == 2 * p->n;
    const bool timing = getenv("XFEMM_TIMING") != nullptr && p->n > 100000;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    cudaStream_t st = p->stream;
    const int n = p->n;
    SellLayout &L = p->lay;
    if (n <= 0 || (long long)n + 31 > 0x7fffffffll) { set_error("device layout: bad row count %d", n); return XFEMM_ERR_ARG; }
    const int n_pad = (n + 31) & ~31, nslices = n_pad / 32;
    const int TB = 256;
    const unsigned gn = (unsigned)((n + TB - 1) / TB);

    DevBuf<int> ordA, ordB, w32;
    DevBuf<unsigned> keyA, keyB;
    DevBuf<unsigned char> tmp;
    size_t tmp_bytes = 0;
    {
        NoArena scratch; // temporaries stay out of a level's L2-pinned arena
        XB_CUDA(ordA.alloc(n)); XB_CUDA(ordB.alloc(n)); XB_CUDA(keyA.alloc(n)); XB_CUDA(keyB.alloc(n)); XB_CUDA(w32.alloc(nslices + 1));
        size_t b1 = 0, b2 = 0;
        XB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, keyA.p, keyB.p, ordA.p, ordB.p, n, 0, 32, st));
        XB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, w32.p, w32.p, nslices + 1, st));
        tmp_bytes = std::max(b1, b2);
        XB_CUDA(tmp.alloc(tmp_bytes + 256));
        if (use_coords && d_xy->n != 2 * (size_t)n)
            XB_CUDA(d_xy->upload(h_xy->data(), 2 * (size_t)n, st)); // (a coarse level has it already)
    }

    // ---- base order
    int *order = ordA.p, *spare = ordB.p;
    if (have_forced)
    {
        g_h2d_bytes += (long long)n * 4;
        XB_CUDA(cudaMemcpyAsync(order, forced->data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
        sigma = 1;
    }
    else if (use_coords)
    {
        // bounding box on the host threads (exact minima/maxima: the order of the comparisons does not matter)
        const float *xy = h_xy->data();
        std::vector<float> bb(4 * 64, 0.f);
        std::vector<unsigned char> used(64, 0);
        parallel_for(n, [&](int i0, int i1, int t) {
            float x0 = xy[2 * (size_t)i0], x1 = x0, y0 = xy[2 * (size_t)i0 + 1], y1 = y0;
            for (int i = i0; i < i1; i++)
            {
                x0 = std::min(x0, xy[2 * (size_t)i]); x1 = std::max(x1, xy[2 * (size_t)i]);
                y0 = std::min(y0, xy[2 * (size_t)i + 1]); y1 = std::max(y1, xy[2 * (size_t)i + 1]);
            }
            bb[4 * t] = x0; bb[4 * t + 1] = x1; bb[4 * t + 2] = y0; bb[4 * t + 3] = y1; used[t] = 1;
        });
        float x0 = 0, x1 = 0, y0 = 0, y1 = 0;
        bool first = true;
        for (int t = 0; t < 64; t++)
            if (used[t])
            {
                if (first) { x0 = bb[4 * t]; x1 = bb[4 * t + 1]; y0 = bb[4 * t + 2]; y1 = bb[4 * t + 3]; first = false; }
                else { x0 = std::min(x0, bb[4 * t]); x1 = std::max(x1, bb[4 * t + 1]); y0 = std::min(y0, bb[4 * t + 2]); y1 = std::max(y1, bb[4 * t + 3]); }
            }
        const float span = std::max(std::max(x1 - x0, y1 - y0), 1e-30f);
        ++g_launch, k_lay_cells<<<gn, TB, 0, st>>>(n, d_xy->p, x0, y0, span, keyA.p, ordB.p);
        XB_CUDA(cudaGetLastError());
        XB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keyA.p, keyB.p, ordB.p, ordA.p, n, 0, 22, st)); // stable
        ++g_launch;
    }
    else
    {
        ++g_launch, k_lay_iota<<<gn, TB, 0, st>>>(n, order);
        XB_CUDA(cudaGetLastError());
    }

    // ---- windows of sigma rows, longest rows first (stable)
    if (sigma > 1)
    {
        int bits = 12;
        for (unsigned nw = (unsigned)((n + sigma - 1) / sigma); nw > 0; nw >>= 1) bits++;
        if (bits > 32) { set_error("device layout: %d rows in windows of %d do not fit the sort key", n, sigma); return XFEMM_ERR_ARG; }
        ++g_launch, k_lay_window_keys<<<gn, TB, 0, st>>>(n, order, p->rowptr.p, sigma, wide_thr, keyA.p);
        XB_CUDA(cudaGetLastError());
        XB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keyA.p, keyB.p, order, spare, n, 0, bits, st));
        ++g_launch;
        std::swap(order, spare);
    }

    // ---- numbering, widths, slice offsets
    XB_CUDA(p->i2c.alloc(n_pad)); XB_CUDA(p->c2i.alloc(n));
    XB_CUDA(p->slice_ptr.alloc(nslices + 1)); XB_CUDA(p->slice_wl.alloc(nslices)); XB_CUDA(p->slice_wu.alloc(nslices));
    XB_CUDA(p->diag_src.alloc(n_pad));
    XB_CUDA(p->slice_wu.zero(st));
    ++g_launch, k_lay_numbering<<<(unsigned)((n_pad + TB - 1) / TB), TB, 0, st>>>(n, n_pad, order, p->i2c.p, p->c2i.p);
    XB_CUDA(cudaGetLastError());
    const unsigned gs = (unsigned)(((size_t)nslices * 32 + TB - 1) / TB);
    ++g_launch, k_lay_widths<<<gs, TB, 0, st>>>(nslices, p->i2c.p, p->rowptr.p, p->col.p, wide_thr, p->slice_wl.p, w32.p, p->diag_src.p);
    XB_CUDA(cudaGetLastError());
    // (a total beyond 2^31 slots would wrap: checked below against the widths)
    size_t scan_bytes = tmp_bytes;
    XB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, w32.p, p->slice_ptr.p, nslices + 1, st));
    ++g_launch;

    // ---- host copies of the small tables (planners: transfer maps, stream tiles, partitions)
    XB_CUDA(lay_fetch(L.slice_ptr, p->slice_ptr.p, (size_t)nslices + 1, st));
    XB_CUDA(lay_fetch(L.slice_wl, p->slice_wl.p, (size_t)nslices, st));
    XB_CUDA(lay_fetch(L.i2c, p->i2c.p, (size_t)n_pad, st));
    XB_CUDA(lay_fetch(L.c2i, p->c2i.p, (size_t)n, st));
    XB_CUDA(cudaStreamSynchronize(st));
    long long total = 0;
    for (int s = 0; s < nslices; s++) total += 32ll * L.slice_wl[s];
    if (total > 0x7fffffffll || total != (long long)L.slice_ptr[nslices])
    {
        set_error("device layout: %lld SELL slots do not fit 32-bit offsets", total);
        return XFEMM_ERR_ARG;
    }
    L.n_pad = n_pad; L.nslices = nslices; L.ncolors = 0; L.n_extra = 0; L.sell_len = total;
    L.slice_wu.assign(nslices, 0);
    L.color_slice = {0, nslices};
    XB_CUDA(p->color_slice.upload(L.color_slice.data(), 2, st));

    // ---- slots
    XB_CUDA(p->sell_col.alloc((size_t)total));
    XB_CUDA(p->sell_src.alloc((size_t)total));
    ++g_launch, k_lay_slots<<<gs, TB, 0, st>>>(nslices, p->i2c.p, p->c2i.p, p->rowptr.p, p->col.p, p->slice_ptr.p, wide_thr, p->sell_col.p, p->sell_src.p);
    XB_CUDA(cudaGetLastError());

    // ---- the wide rows stay a host job (a few hundred rows)
    std::vector<int> wide_rows;
    if (wide_thr > 0)
    {
        const int *rp = p->h_rowptr.data();
        std::vector<std::vector<int>> part(64);
        parallel_for(n, [&](int i0, int i1, int t) {
            for (int i = i0; i < i1; i++) if (rp[i + 1] - rp[i] - 1 > wide_thr) part[t].push_back(i);
        });
        for (auto &v : part) wide_rows.insert(wide_rows.end(), v.begin(), v.end());
    }
    build_wide_part(p->h_rowptr.data(), p->h_col.data(), wide_rows, L);
    XB_CUDA(p->wide.upload(L, st));
    XB_CUDA(cudaStreamSynchronize(st));
    if (timing) fprintf(stderr, "[xfemm timing]     device layout n=%d: %.4f s\n", n, now() - t0);
    return XFEMM_OK;
}

// XFEMM_LAYOUT_CHECK: the device tables against a host layout of the same problem
template <class Problem>
int check_layout_device(Problem *p, const SellLayout &H)
{
    cudaStream_t st = p->stream;
    const SellLayout &L = p->lay;
    auto fail = [&](const char *what) { set_error("device layout differs from the host layout: %s (n=%d)", what, p->n); return XFEMM_ERR_ARG; };
    if (L.n_pad != H.n_pad || L.nslices != H.nslices || L.sell_len != H.sell_len) return fail("sizes");
    if (L.i2c != H.i2c) return fail("i2c");
    if (L.c2i != H.c2i) return fail("c2i");
    if (L.slice_ptr != H.slice_ptr) return fail("slice_ptr");
    if (L.slice_wl != H.slice_wl) return fail("slice widths");
    if (L.wide_row != H.wide_row || L.wide_rseg != H.wide_rseg || L.wide_sptr != H.wide_sptr || L.slice_wide != H.slice_wide) return fail("wide rows");
    std::vector<int> a, b, c;
    XB_CUDA(lay_fetch(a, p->sell_col.p, (size_t)L.sell_len, st));
    XB_CUDA(lay_fetch(b, p->sell_src.p, (size_t)L.sell_len, st));
    XB_CUDA(lay_fetch(c, p->diag_src.p, (size_t)L.n_pad, st));
    XB_CUDA(cudaStreamSynchronize(st));
    if (a != H.sell_col) return fail("sell_col");
    if (b != H.sell_src) return fail("sell_src");
    if (c != H.diag_src) return fail("diag_src");
    return XFEMM_OK;
}

} // namespace xb
