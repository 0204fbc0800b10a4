// Host-side construction of the solver layout of a linear problem: internal row order
// (multicolour classes, then length-sorted windows) and the split SELL-32 structure.
// Pure integer work, done once per sparsity pattern and reused by every solve on it
// (Newton steps, rotor-angle sweeps).
#include "layout.h"
#include <algorithm>
#include <numeric>
#include <cstdlib>

namespace xb {

// Greedy distance-1 colouring in the caller's (Cuthill–McKee) row order with a cap on the
// number of colours.  A first-order triangulation needs 5-7 colours; the few rows that carry
// air-gap-element 10-cliques or periodic fill would push a plain greedy colouring to 12-13 and
// every extra colour costs two grid-wide barriers per preconditioner application.  When a row
// finds all `max_colors` colours taken it joins the colour with the fewest clashes; the clashing
// couplings are reported by the layout as "extra" entries: they stay in the matrix (SpMV) but
// are left out of the SSOR sweeps, i.e. the preconditioner is the multicolour SSOR of the matrix
// without those couplings — still symmetric positive definite for any positive diagonal.
int greedy_coloring(int n, const int *rowptr, const int *col, int max_colors, std::vector<int> &color)
{
    color.assign(n, -1);
    int ncol = 0;
    if (max_colors < 2) max_colors = 2;
    std::vector<int> cnt(max_colors, 0);
    for (int i = 0; i < n; i++)
    {
        std::fill(cnt.begin(), cnt.end(), 0);
        for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
        {
            int j = col[s];
            if (j == i) continue;
            int c = color[j];
            if (c >= 0) cnt[c]++;
        }
        int best = 0;
        for (int c = 0; c < max_colors; c++)
        {
            if (cnt[c] == 0) { best = c; break; }
            if (cnt[c] < cnt[best]) best = c;
        }
        color[i] = best;
        if (best + 1 > ncol) ncol = best + 1;
    }
    return ncol;
}

void build_sell_layout(int n, const int *rowptr, const int *col, const std::vector<int> *color_in,
                       int ncolors, int sigma, SellLayout &L, int wide_threshold, const int *base_order)
{
    // rows with more than `wide_threshold` off-diagonals leave the SELL structure (see layout.h)
    std::vector<unsigned char> wide(n, 0);
    bool any_wide = false;
    if (wide_threshold > 0)
        for (int i = 0; i < n; i++)
            if (rowptr[i + 1] - rowptr[i] - 1 > wide_threshold) { wide[i] = 1; any_wide = true; }

    // ---- internal order: by colour (if any), inside a colour keep the caller's order, then
    // sort windows of `sigma` rows by their entry counts so slices are homogeneous.
    // `base_order` (optional): a locality-preserving sequence of the rows (e.g. along a space-filling
    // curve through the node coordinates) to start from instead of the caller's numbering
    std::vector<int> order(n);
    if (base_order) std::copy(base_order, base_order + n, order.begin());
    else std::iota(order.begin(), order.end(), 0);
    std::vector<int> color_start; // in rows (unpadded sequence positions)
    const bool colored = color_in && ncolors > 0;
    if (colored)
    {
        const std::vector<int> &color = *color_in;
        std::vector<int> cnt(ncolors + 1, 0);
        for (int i = 0; i < n; i++) cnt[color[i] + 1]++;
        for (int c = 0; c < ncolors; c++) cnt[c + 1] += cnt[c];
        color_start = cnt;
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        std::vector<int> seq(order);
        for (int k = 0; k < n; k++) order[pos[color[seq[k]]]++] = seq[k];
    }
    else
    {
        ncolors = 0;
        color_start = {0, n};
    }
    const int ngroups = (int)color_start.size() - 1;

    // classification of an off-diagonal (i,j): 0 lower (j earlier in the internal order = lower
    // colour), 1 upper, 2 extra (same colour: excluded from the SSOR sweeps).  Sorting inside a
    // window only permutes rows of one colour, so for coloured layouts the class never depends
    // on the final position; the uncoloured layout keeps all off-diagonals in one block.
    auto klass = [&](int i, int j, const std::vector<int> &rank) -> int {
        if (colored)
        {
            if (wide[j]) return 2; // the partner row keeps no sweep coupling either → M stays symmetric
            int ci = (*color_in)[i], cj = (*color_in)[j];
            if (ci == cj) return 2;
            return cj < ci ? 0 : 1;
        }
        (void)rank;
        return 0; // uncoloured: one block per slice (a lower/upper split would only add padding)
    };
    std::vector<int> rank(n);
    for (int k = 0; k < n; k++) rank[order[k]] = k;
    if (sigma > 1)
    {
        std::vector<int> key(n);
        parallel_for(n, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++)
            {
                int c[3] = {0, 0, 0};
                if (!wide[i] && colored)
                    for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
                        if (col[s] != i) c[klass(i, col[s], rank)]++;
                key[i] = colored ? (std::min(c[0], 63) * 64 + std::min(c[1], 63)) * 64 + std::min(c[2], 63)
                                 : (wide[i] ? 0 : rowptr[i + 1] - rowptr[i]);
            }
        });
        std::vector<std::pair<int, int>> windows;
        for (int g = 0; g < ngroups; g++)
            for (int w0 = color_start[g]; w0 < color_start[g + 1]; w0 += sigma)
                windows.emplace_back(w0, std::min(w0 + sigma, color_start[g + 1]));
        parallel_for((int)windows.size(), [&](int a0, int a1, int) {
            for (int a = a0; a < a1; a++)
                std::stable_sort(order.begin() + windows[a].first, order.begin() + windows[a].second,
                                 [&](int x, int y) { return key[x] > key[y]; });
        }, 64);
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

    // ---- split SELL: per slice [lower | upper | extra] blocks, diagonal kept apart
    L.slice_ptr.assign(L.nslices + 1, 0);
    L.slice_wl.assign(L.nslices, 0);
    L.slice_wu.assign(L.nslices, 0);
    L.n_extra = 0;
    std::vector<int> width(3 * (size_t)L.nslices, 0);
    std::vector<long long> extra_part(64, 0);
    parallel_for(L.nslices, [&](int s0, int s1, int tid) {
        long long ex = 0;
        for (int s = s0; s < s1; s++)
            for (int r = s * 32; r < s * 32 + 32; r++)
            {
                int i = L.i2c[r];
                if (i < 0) continue;
                int c[3] = {0, 0, 0};
                if (!wide[i])
                    for (int q = rowptr[i]; q < rowptr[i + 1]; q++)
                        if (col[q] != i) c[klass(i, col[q], L.c2i)]++;
                ex += c[2];
                for (int q = 0; q < 3; q++) width[3 * (size_t)s + q] = std::max(width[3 * (size_t)s + q], c[q]);
            }
        extra_part[tid] += ex;
    });
    for (long long e : extra_part) L.n_extra += e;
    long long total = 0;
    for (int s = 0; s < L.nslices; s++)
    {
        L.slice_wl[s] = width[3 * (size_t)s];
        L.slice_wu[s] = width[3 * (size_t)s + 1];
        L.slice_ptr[s] = (int)total;
        total += (long long)(width[3 * (size_t)s] + width[3 * (size_t)s + 1] + width[3 * (size_t)s + 2]) * 32;
    }
    L.slice_ptr[L.nslices] = (int)total;
    L.sell_len = total;
    L.sell_col.assign(total, 0);
    L.sell_src.assign(total, -1);
    L.diag_src.assign(L.n_pad, -1);
    parallel_for(L.nslices, [&](int sb, int se, int) {
    std::vector<std::pair<int, int>> part[3]; // (internal col, csr index)
    for (int s = sb; s < se; s++)
    {
        const int base = L.slice_ptr[s];
        const int off[3] = {0, L.slice_wl[s], L.slice_wl[s] + L.slice_wu[s]};
        const int w = (L.slice_ptr[s + 1] - base) / 32;
        for (int l = 0; l < 32; l++)
        {
            const int r = s * 32 + l;
            for (int k = 0; k < w; k++) L.sell_col[base + k * 32 + l] = r; // padding points at self
            int i = L.i2c[r];
            if (i < 0) continue;
            for (auto &p : part) p.clear();
            for (int q = rowptr[i]; q < rowptr[i + 1]; q++)
            {
                int j = col[q];
                if (j == i) { L.diag_src[r] = q; continue; }
                if (wide[i]) continue;
                part[klass(i, j, L.c2i)].emplace_back(L.c2i[j], q);
            }
            for (int t = 0; t < 3; t++)
            {
                std::sort(part[t].begin(), part[t].end());
                for (size_t k = 0; k < part[t].size(); k++)
                {
                    L.sell_col[base + (off[t] + (int)k) * 32 + l] = part[t][k].first;
                    L.sell_src[base + (off[t] + (int)k) * 32 + l] = part[t][k].second;
                }
            }
        }
    }
    });

    std::vector<int> wide_rows;
    if (any_wide)
        for (int i = 0; i < n; i++) if (wide[i]) wide_rows.push_back(i);
    build_wide_part(rowptr, col, wide_rows, L);
}

// The wide rows (caller numbers, any order) of a layout whose i2c/c2i are known: row-wise storage in segments
// (layout.h), ascending internal row.
void build_wide_part(const int *rowptr, const int *col, const std::vector<int> &wide_rows, SellLayout &L)
{
    L.wide_row.clear(); L.wide_rseg.assign(1, 0); L.wide_sptr.assign(1, 0);
    L.wide_col.clear(); L.wide_src.clear(); L.slice_wide.clear();
    if (wide_rows.empty()) return;
    L.slice_wide.assign(L.nslices, 0u);
    for (int i : wide_rows) L.wide_row.push_back(L.c2i[i]);
    std::sort(L.wide_row.begin(), L.wide_row.end());
    std::vector<std::pair<int, int>> ent;
    for (int r : L.wide_row)
    {
        const int i = L.i2c[r];
        L.slice_wide[r / 32] |= 1u << (r % 32);
        ent.clear();
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++)
            if (col[q] != i) ent.emplace_back(L.c2i[col[q]], q);
        std::sort(ent.begin(), ent.end());
        for (size_t k = 0; k < ent.size(); k++)
        {
            if (k && k % SellLayout::WIDE_SEG == 0) L.wide_sptr.push_back((int)L.wide_col.size());
            L.wide_col.push_back(ent[k].first);
            L.wide_src.push_back(ent[k].second);
        }
        L.wide_sptr.push_back((int)L.wide_col.size());
        L.wide_rseg.push_back((int)L.wide_sptr.size() - 1);
    }
}

// Z-order (Morton) sequence of points: rows that are close in the plane end up close in memory, so
// the gathers of a slice hit few sectors and the transfer maps between multigrid levels become local.
void morton_order(int n, const float *xy, std::vector<int> &order)
{
    order.resize(n);
    if (n == 0) return;
    float x0 = xy[0], x1 = xy[0], y0 = xy[1], y1 = xy[1];
    for (int i = 1; i < n; i++)
    {
        x0 = std::min(x0, xy[2 * (size_t)i]); x1 = std::max(x1, xy[2 * (size_t)i]);
        y0 = std::min(y0, xy[2 * (size_t)i + 1]); y1 = std::max(y1, xy[2 * (size_t)i + 1]);
    }
    const float span = std::max(std::max(x1 - x0, y1 - y0), 1e-30f);
    auto spread = [](unsigned v) -> unsigned long long {
        unsigned long long x = v & 0xffffffffull;
        x = (x | (x << 16)) & 0x0000ffff0000ffffull;
        x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
        x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
        x = (x | (x << 2)) & 0x3333333333333333ull;
        x = (x | (x << 1)) & 0x5555555555555555ull;
        return x;
    };
    // 11 bits per axis -> 4 M cells; a stable counting sort by cell keeps the caller's (Cuthill-McKee,
    // already local) order inside a cell and costs three linear passes instead of a comparison sort
    const int BITS = 11;
    const float scale = (float)((1 << BITS) - 1);
    std::vector<unsigned> cell(n);
    parallel_for(n, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++)
        {
            const unsigned qx = (unsigned)((xy[2 * (size_t)i] - x0) / span * scale);
            const unsigned qy = (unsigned)((xy[2 * (size_t)i + 1] - y0) / span * scale);
            cell[i] = (unsigned)(spread(qx) | (spread(qy) << 1));
        }
    });
    std::vector<int> start((size_t)(1u << (2 * BITS)) + 1, 0);
    for (int i = 0; i < n; i++) start[cell[i] + 1]++;
    for (size_t c = 0; c + 1 < start.size(); c++) start[c + 1] += start[c];
    for (int i = 0; i < n; i++) order[start[cell[i]]++] = i;
}

} // namespace xb

#include <unordered_map>
namespace xb {

int build_pbc_plan(int n, const int *rowptr, const int *col, int npairs, const int *ni, const int *nj,
                   const int *type, std::vector<int> &level_ptr, std::vector<int> &item_ptr,
                   std::vector<int> &meta, std::vector<int> &items, int err[3])
{
    auto slot = [&](int row, int c) -> int {
        const int *b = col + rowptr[row], *e = col + rowptr[row + 1];
        const int *it = std::lower_bound(b, e, c);
        return (it != e && *it == c) ? (int)(it - col) : -1;
    };
    std::vector<int> iptr(npairs + 1, 0), m0((size_t)6 * npairs), it0, level(npairs, 0);
    std::unordered_map<long long, int> last; // slot / rhs id → last level that touched it
    std::vector<int> ks;
    int nlevels = 0;
    for (int k = 0; k < npairs; k++)
    {
        int i = ni[k], j = nj[k];
        if (i < 0 || i >= n || j < 0 || j >= n || i == j) { err[0] = i; err[1] = j; err[2] = -1; return -1; }
        if (j < i) std::swap(i, j);
        ks.clear();
        ks.insert(ks.end(), col + rowptr[i], col + rowptr[i + 1]);
        ks.insert(ks.end(), col + rowptr[j], col + rowptr[j + 1]);
        std::sort(ks.begin(), ks.end());
        ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
        int lv = 0;
        auto touch = [&](long long id) { auto f = last.find(id); if (f != last.end()) lv = std::max(lv, f->second + 1); };
        const size_t first = it0.size();
        for (int kk : ks)
        {
            if (kk == i || kk == j) continue;
            int s[4] = {slot(kk, i), slot(i, kk), slot(kk, j), slot(j, kk)};
            if (s[0] < 0 || s[1] < 0 || s[2] < 0 || s[3] < 0) { err[0] = i; err[1] = j; err[2] = kk; return -1; }
            for (int t = 0; t < 4; t++) { it0.push_back(s[t]); touch(s[t]); }
        }
        const int sii = slot(i, i), sjj = slot(j, j);
        touch(sii); touch(sjj); touch(-1 - (long long)i); touch(-1 - (long long)j);
        for (size_t q = first; q < it0.size(); q++) last[it0[q]] = lv;
        last[sii] = lv; last[sjj] = lv; last[-1 - (long long)i] = lv; last[-1 - (long long)j] = lv;
        level[k] = lv;
        nlevels = std::max(nlevels, lv + 1);
        iptr[k + 1] = (int)(it0.size() / 4);
        int *m = &m0[(size_t)6 * k];
        m[0] = i; m[1] = j; m[2] = type[k]; m[3] = sii; m[4] = sjj; m[5] = 0;
    }
    std::vector<int> order(npairs);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return level[a] < level[b]; });
    level_ptr.assign(nlevels + 1, 0);
    item_ptr.assign(npairs + 1, 0);
    meta.assign((size_t)6 * npairs, 0);
    items.clear();
    items.reserve(it0.size());
    for (int q = 0; q < npairs; q++)
    {
        int k = order[q];
        level_ptr[level[k] + 1]++;
        items.insert(items.end(), it0.begin() + (size_t)iptr[k] * 4, it0.begin() + (size_t)iptr[k + 1] * 4);
        item_ptr[q + 1] = (int)(items.size() / 4);
        for (int t = 0; t < 6; t++) meta[(size_t)6 * q + t] = m0[(size_t)6 * k + t];
    }
    for (int l = 0; l < nlevels; l++) level_ptr[l + 1] += level_ptr[l];
    if (items.empty()) items.assign(4, 0);
    return nlevels;
}

} // namespace xb
