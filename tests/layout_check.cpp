// CPU check of the host-side set-up code that only the GPU tests exercise otherwise: the SELL layout with
// wide rows, the Z-curve order, the aggregation and the Galerkin maps of the multigrid hierarchy.
// Prints "ok" or a diagnostic; tests/test_host_parity.py compiles and runs it.
#include "../xfemm_b200/csrc/layout.h"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

namespace xb {
int aggregate(int n, const int *rp, const int *col, int wide_thr, std::vector<int> &agg);
void coarse_pattern(int n, const int *rp, const int *col, const std::vector<int> &agg, int nc, std::vector<int> &crp,
                    std::vector<int> &ccol, std::vector<int> &gptr, std::vector<int> &gsrc);
}
#define CHECK(c, ...) do { if (!(c)) { printf("FAILED %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); return 1; } } while (0)

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 60; // N x N grid, 5-point stencil + a few rows coupled to everything
    const int n = N * N;
    std::vector<std::set<int>> adj(n);
    for (int j = 0; j < N; j++)
        for (int i = 0; i < N; i++)
        {
            int k = j * N + i;
            adj[k].insert(k);
            if (i + 1 < N) { adj[k].insert(k + 1); adj[k + 1].insert(k); }
            if (j + 1 < N) { adj[k].insert(k + N); adj[k + N].insert(k); }
        }
    for (int w : {7, n / 2, n - 3}) // wide rows (think: air-gap ring, circuit row)
        for (int k = 0; k < n; k += 3) { adj[w].insert(k); adj[k].insert(w); }
    std::vector<int> rp(n + 1, 0), col;
    for (int i = 0; i < n; i++) { col.insert(col.end(), adj[i].begin(), adj[i].end()); rp[i + 1] = (int)col.size(); }
    std::vector<float> xy(2 * (size_t)n);
    for (int k = 0; k < n; k++) { xy[2 * k] = (float)(k % N); xy[2 * k + 1] = (float)(k / N); }

    for (int variant = 0; variant < 3; variant++)
    {
        std::vector<int> zorder, color;
        const int *base = nullptr;
        if (variant >= 1) { xb::morton_order(n, xy.data(), zorder); base = zorder.data(); }
        if (variant >= 1)
        {
            std::vector<char> seen(n, 0);
            for (int v : zorder) { CHECK(v >= 0 && v < n && !seen[v], "Z-order is not a permutation"); seen[v] = 1; }
        }
        int nc = 0;
        if (variant == 2) nc = xb::greedy_coloring(n, rp.data(), col.data(), 7, color);
        xb::SellLayout L;
        xb::build_sell_layout(n, rp.data(), col.data(), variant == 2 ? &color : nullptr, nc, 256, L, 64, base);
        CHECK(L.n_pad % 32 == 0 && L.n_pad >= n && (int)L.i2c.size() == L.n_pad, "padding");
        for (int i = 0; i < n; i++) CHECK(L.c2i[i] >= 0 && L.i2c[L.c2i[i]] == i, "c2i/i2c are not inverse at %d", i);
        // every off-diagonal CSR slot appears exactly once (SELL or wide storage), the diagonal in diag_src
        std::vector<int> hit(col.size(), 0);
        for (long long q = 0; q < L.sell_len; q++)
            if (L.sell_src[q] >= 0)
            {
                hit[L.sell_src[q]]++;
                const int slice = (int)(std::upper_bound(L.slice_ptr.begin(), L.slice_ptr.end(), (int)q) - L.slice_ptr.begin()) - 1;
                const int row = slice * 32 + (int)((q - L.slice_ptr[slice]) % 32);
                const int i = L.i2c[row];
                CHECK(i >= 0 && L.sell_src[q] >= rp[i] && L.sell_src[q] < rp[i + 1], "SELL slot in the wrong row");
                CHECK(L.sell_col[q] == L.c2i[col[L.sell_src[q]]], "SELL column mismatch");
            }
        for (size_t q = 0; q < L.wide_src.size(); q++) { hit[L.wide_src[q]]++; CHECK(L.wide_col[q] == L.c2i[col[L.wide_src[q]]], "wide column"); }
        for (int r = 0; r < L.n_pad; r++) if (L.diag_src[r] >= 0) hit[L.diag_src[r]]++;
        for (size_t q = 0; q < col.size(); q++) CHECK(hit[q] == 1, "CSR slot %zu covered %d times (variant %d)", q, hit[q], variant);
        CHECK(L.wide_row.size() == 3, "expected 3 wide rows, got %zu", L.wide_row.size());
        for (size_t k = 0; k + 1 < L.wide_rseg.size(); k++)
            for (int sg = L.wide_rseg[k]; sg < L.wide_rseg[k + 1]; sg++)
                CHECK(L.wide_sptr[sg + 1] - L.wide_sptr[sg] <= xb::SellLayout::WIDE_SEG && L.wide_sptr[sg + 1] > L.wide_sptr[sg], "segment size");
        long long padded = L.sell_len, real = (long long)col.size() - n - (long long)L.wide_src.size();
        CHECK(padded >= real && padded <= real * 1.35 + 4096, "SELL padding too large: %lld slots for %lld entries", padded, real);
    }

    // ---- aggregation + Galerkin maps over three levels
    std::vector<int> r = rp, c = col;
    int nn = n;
    for (int lev = 0; lev < 3; lev++)
    {
        std::vector<int> agg, crp, ccol, gptr, gsrc;
        if (nn <= 64) break; // the hierarchy builder stops here as well (dense solve on the coarsest level)
        const int nc = xb::aggregate(nn, r.data(), c.data(), 64, agg);
        CHECK(nc > 0 && nc < nn, "no coarsening at level %d (%d rows)", lev, nn);
        std::vector<int> size(nc, 0);
        for (int i = 0; i < nn; i++) { CHECK(agg[i] >= -1 && agg[i] < nc, "aggregate id"); if (agg[i] >= 0) size[agg[i]]++; }
        for (int I = 0; I < nc; I++) CHECK(size[I] > 0, "empty aggregate %d", I);
        for (int i = 0; i < nn; i++) CHECK(agg[i] >= 0 || r[i + 1] - r[i] <= 1, "coupled row %d left out", i);
        xb::coarse_pattern(nn, r.data(), c.data(), agg, nc, crp, ccol, gptr, gsrc);
        // reference: brute-force P^T |A| P pattern and slot lists
        std::map<std::pair<int, int>, std::vector<int>> ref;
        for (int i = 0; i < nn; i++)
            for (int s = r[i]; s < r[i + 1]; s++)
                if (agg[i] >= 0 && agg[c[s]] >= 0) ref[{agg[i], agg[c[s]]}].push_back(s);
        CHECK((size_t)crp[nc] == ref.size(), "coarse nnz %d vs %zu", crp[nc], ref.size());
        for (int I = 0; I < nc; I++)
            for (int t = crp[I]; t < crp[I + 1]; t++)
            {
                if (t > crp[I]) CHECK(ccol[t] > ccol[t - 1], "coarse columns not sorted");
                auto f = ref.find({I, ccol[t]});
                CHECK(f != ref.end(), "coarse entry (%d,%d) should not exist", I, ccol[t]);
                std::vector<int> mine(gsrc.begin() + gptr[t], gsrc.begin() + gptr[t + 1]), want = f->second;
                std::sort(mine.begin(), mine.end()); std::sort(want.begin(), want.end());
                CHECK(mine == want, "Galerkin slot list of (%d,%d) differs", I, ccol[t]);
                CHECK(std::binary_search(ccol.begin() + crp[ccol[t]], ccol.begin() + crp[ccol[t] + 1], I), "coarse pattern not symmetric");
            }
        r = crp; c = ccol; nn = nc;
    }
    printf("ok\n");
    return 0;
}
