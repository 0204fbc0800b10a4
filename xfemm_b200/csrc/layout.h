// Solver layout (internal row order + split SELL-32) built on the host once per pattern.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>

namespace xb {

// Host-side parallel loop: f(begin, end, thread_index) over [0,n) in contiguous chunks.  The set-up
// code (layouts, multigrid hierarchy, pattern) is integer work over millions of rows; a handful
// of host threads takes it off the critical path.  XFEMM_HOST_THREADS overrides the thread count.
template <class F>
inline void parallel_for(int n, F f, int min_per_thread = 16384)
{
    int nt = (int)std::thread::hardware_concurrency();
    // one process per GPU: share the cores between the ranks of this node (torchrun exports LOCAL_WORLD_SIZE)
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) nt /= std::max(1, atoi(lw));
    if (const char *e = getenv("XFEMM_HOST_THREADS")) nt = atoi(e);
    nt = std::max(1, std::min(nt, 32));
    nt = std::min(nt, std::max(1, n / std::max(1, min_per_thread)));
    if (nt <= 1) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=]() { f((int)((long long)n * t / nt), (int)((long long)n * (t + 1) / nt), t); });
    for (auto &x : th) x.join();
}

struct SellLayout
{
    int n_pad = 0, nslices = 0, ncolors = 0;
    long long sell_len = 0;
    long long n_extra = 0;        // couplings excluded from the SSOR sweeps (same-colour clashes)
    std::vector<int> i2c;         // internal row -> caller row (-1 = padding row)
    std::vector<int> c2i;         // caller row -> internal row
    std::vector<int> color_slice; // first slice of each colour (ncolors+1), or {0,nslices}
    std::vector<int> slice_ptr;   // nslices+1, offsets into sell_* (multiples of 32)
    std::vector<int> slice_wl;    // width of the "lower" block of each slice
    std::vector<int> slice_wu;    // width of the "upper" block (then the "extra" block to the end)
    std::vector<int> sell_col;    // internal column per slot (padding: own row)
    std::vector<int> sell_src;    // CSR index per slot (-1 padding)
    std::vector<int> diag_src;    // CSR index of the diagonal per internal row (-1 padding row)
    // "wide" rows (more off-diagonals than the threshold: air-gap rings with the angle-independent
    // pattern, circuit rows of the harmonic formulations) are kept OUT of the SELL slices — one of
    // them would stretch a whole slice and a single warp would walk it — and stored row-wise in
    // segments of <= WIDE_SEG entries that the warps of the owning block share.
    static constexpr int WIDE_SEG = 256;
    std::vector<int> wide_row;    // internal row per wide row, ascending
    std::vector<int> wide_rseg;   // first segment of each wide row (n_wide+1)
    std::vector<int> wide_sptr;   // segment offsets into wide_col/wide_src (n_seg+1)
    std::vector<int> wide_col;    // internal column per entry
    std::vector<int> wide_src;    // CSR index per entry
    std::vector<unsigned> slice_wide; // per slice: bit l set = row 32*s+l is a wide row (empty if none)
};

int greedy_coloring(int n, const int *rowptr, const int *col, int max_colors, std::vector<int> &color);
void build_sell_layout(int n, const int *rowptr, const int *col, const std::vector<int> *color,
                       int ncolors, int sigma, SellLayout &L, int wide_threshold = 64, const int *base_order = nullptr);
void build_wide_part(const int *rowptr, const int *col, const std::vector<int> &wide_rows, SellLayout &L);
void morton_order(int n, const float *xy, std::vector<int> &order);

} // namespace xb

namespace xb {
// Plan for applying Periodicity/AntiPeriodicity pairs in list order on a fixed CSR pattern
// (reference: spars.cpp:366-474, cspars.cpp:677-758): pairs are grouped into dependency levels
// (pairs of a level touch disjoint entries), items are (slot(k,i), slot(i,k), slot(k,j), slot(j,k)),
// meta is (i, j, type, slot(i,i), slot(j,j), 0) per pair.  Returns the number of levels, -1 on a
// missing slot (err_i/err_j/err_k tell which).
int build_pbc_plan(int n, const int *rowptr, const int *col, int npairs, const int *ni, const int *nj,
                   const int *type, std::vector<int> &level_ptr, std::vector<int> &item_ptr,
                   std::vector<int> &meta, std::vector<int> &items, int err[3]);
} // namespace xb
