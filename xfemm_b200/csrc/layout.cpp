// Host-side construction of the solver layout of a linear problem: internal row order
// (multicolour classes, then length-sorted windows) and the split SELL-32 structure.
// Pure integer work, done once per sparsity pattern and reused by every solve on it
// (Newton steps, rotor-angle sweeps).
#include "layout.h"
#include <algorithm>
#include <numeric>
#include <cstdlib>

namespace xb {

// Greedy distance-1 colouring in the caller's (Cuthill–McKee) row order.
int greedy_coloring(int n, const int *rowptr, const int *col, std::vector<int> &color)
{
    color.assign(n, -1);
    int ncol = 0;
    std::vector<int> mark(64, -1);
    for (int i = 0; i < n; i++)
    {
        for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
        {
            int j = col[s];
            if (j == i) continue;
            int c = color[j];
            if (c >= 0)
            {
                if (c >= (int)mark.size()) mark.resize(2 * c + 2, -1);
                mark[c] = i;
            }
        }
        int c = 0;
        while (c < (int)mark.size() && mark[c] == i) c++;
        if (c >= (int)mark.size()) mark.resize(2 * c + 2, -1);
        color[i] = c;
        if (c + 1 > ncol) ncol = c + 1;
    }
    return ncol;
}

void build_sell_layout(int n, const int *rowptr, const int *col, const std::vector<int> *color_in,
                       int ncolors, int sigma, SellLayout &L)
{
    // ---- internal order: by colour (if any), inside a colour keep the caller's order, then
    // sort windows of `sigma` rows by (lower count, upper count) so slices are homogeneous.
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::vector<int> color_start; // in rows (unpadded sequence positions)
    if (color_in && ncolors > 0)
    {
        const std::vector<int> &color = *color_in;
        std::vector<int> cnt(ncolors + 1, 0);
        for (int i = 0; i < n; i++) cnt[color[i] + 1]++;
        for (int c = 0; c < ncolors; c++) cnt[c + 1] += cnt[c];
        color_start = cnt;
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int i = 0; i < n; i++) order[pos[color[i]]++] = i;
    }
    else
    {
        ncolors = 0;
        color_start = {0, n};
    }
    const int ngroups = (int)color_start.size() - 1;

    // provisional internal index (without padding) to classify lower/upper entries
    // lower = neighbour placed earlier in the internal order.  With colours this is exactly
    // "neighbour has a lower colour" because a colour class has no internal couplings.
    std::vector<int> rank(n);
    for (int k = 0; k < n; k++) rank[order[k]] = k;
    std::vector<unsigned char> nlow(n), nup(n);
    for (int i = 0; i < n; i++)
    {
        int lo = 0, up = 0;
        for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
        {
            int j = col[s];
            if (j == i) continue;
            if (rank[j] < rank[i]) lo++; else up++;
        }
        nlow[i] = (unsigned char)std::min(lo, 255);
        nup[i] = (unsigned char)std::min(up, 255);
    }
    // NOTE: sorting inside a window changes ranks inside the window; for the colour layout
    // this never flips lower/upper (same colour ⇒ not coupled).  For the uncoloured layout
    // lower/upper is only a storage split (SpMV uses both halves), so a flip is harmless as
    // long as the classification used for storage is the final one → recompute after sorting.
    if (sigma > 1)
    {
        for (int g = 0; g < ngroups; g++)
        {
            for (int w0 = color_start[g]; w0 < color_start[g + 1]; w0 += sigma)
            {
                int w1 = std::min(w0 + sigma, color_start[g + 1]);
                std::stable_sort(order.begin() + w0, order.begin() + w1, [&](int a, int b) {
                    int ka = (rowptr[a + 1] - rowptr[a]), kb = (rowptr[b + 1] - rowptr[b]);
                    if (ncolors > 0)
                    {
                        ka = nlow[a] * 256 + nup[a];
                        kb = nlow[b] * 256 + nup[b];
                    }
                    return ka > kb;
                });
            }
        }
    }

    // ---- padded internal numbering: every colour starts on a slice boundary
    L.ncolors = ncolors;
    L.color_slice.assign(1, 0);
    L.i2c.clear();
    for (int g = 0; g < ngroups; g++)
    {
        for (int k = color_start[g]; k < color_start[g + 1]; k++) L.i2c.push_back(order[k]);
        while (L.i2c.size() % 32) L.i2c.push_back(-1);
        L.color_slice.push_back((int)(L.i2c.size() / 32));
    }
    L.n_pad = (int)L.i2c.size();
    L.nslices = L.n_pad / 32;
    L.c2i.assign(n, -1);
    for (int r = 0; r < L.n_pad; r++) if (L.i2c[r] >= 0) L.c2i[L.i2c[r]] = r;

    // ---- split SELL: per slice [lower block | upper block], diagonal kept apart
    L.slice_ptr.assign(L.nslices + 1, 0);
    L.slice_wl.assign(L.nslices, 0);
    std::vector<int> lo_cnt(L.n_pad, 0), up_cnt(L.n_pad, 0);
    for (int r = 0; r < L.n_pad; r++)
    {
        int i = L.i2c[r];
        if (i < 0) continue;
        for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
        {
            int j = col[s];
            if (j == i) continue;
            if (L.c2i[j] < r) lo_cnt[r]++; else up_cnt[r]++;
        }
    }
    long long total = 0;
    for (int s = 0; s < L.nslices; s++)
    {
        int wl = 0, wu = 0;
        for (int l = 0; l < 32; l++) { wl = std::max(wl, lo_cnt[s * 32 + l]); wu = std::max(wu, up_cnt[s * 32 + l]); }
        L.slice_wl[s] = wl;
        L.slice_ptr[s] = (int)total;
        total += (long long)(wl + wu) * 32;
    }
    L.slice_ptr[L.nslices] = (int)total;
    L.sell_len = total;
    L.sell_col.assign(total, 0);
    L.sell_src.assign(total, -1);
    L.diag_src.assign(L.n_pad, -1);
    std::vector<std::pair<int, int>> lo, up; // (internal col, csr index)
    for (int s = 0; s < L.nslices; s++)
    {
        const int base = L.slice_ptr[s], wl = L.slice_wl[s];
        const int w = (L.slice_ptr[s + 1] - base) / 32;
        for (int l = 0; l < 32; l++)
        {
            const int r = s * 32 + l;
            for (int k = 0; k < w; k++) L.sell_col[base + k * 32 + l] = r; // padding points at self
            int i = L.i2c[r];
            if (i < 0) continue;
            lo.clear(); up.clear();
            for (int q = rowptr[i]; q < rowptr[i + 1]; q++)
            {
                int j = col[q];
                if (j == i) { L.diag_src[r] = q; continue; }
                int jc = L.c2i[j];
                if (jc < r) lo.emplace_back(jc, q); else up.emplace_back(jc, q);
            }
            std::sort(lo.begin(), lo.end());
            std::sort(up.begin(), up.end());
            for (size_t k = 0; k < lo.size(); k++)
            {
                L.sell_col[base + (int)k * 32 + l] = lo[k].first;
                L.sell_src[base + (int)k * 32 + l] = lo[k].second;
            }
            for (size_t k = 0; k < up.size(); k++)
            {
                L.sell_col[base + (wl + (int)k) * 32 + l] = up[k].first;
                L.sell_src[base + (wl + (int)k) * 32 + l] = up[k].second;
            }
        }
    }
}

} // namespace xb
