// Host side of the aggregation-multigrid preconditioner (see amg.cuh): hierarchy construction from
// the sparsity pattern (once per pattern) and the per-solve Galerkin value refresh on the device.
#include "amg.cuh"
#include "amg2.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

namespace xb {

int linprob_refresh_values(xfemm_linprob *p, int *singular_row); // linprob.cu
int linprob_refresh_values_on(xfemm_linprob *p, cudaStream_t st);
// amg_setup.cu
int amg_coarse_pattern_device(xfemm_linprob *cur, const int *d_agg, int nc, std::vector<int> &crp, std::vector<int> &ccol,
                              DevBuf<int> &gptr, DevBuf<int> &gsrc, cudaStream_t st);
int amg_centroids_device(xfemm_linprob *cur, xfemm_linprob *c, const int *d_agg, cudaStream_t st);
int amg_transfer_maps_device(xfemm_linprob *cur, xfemm_linprob *c, const int *d_agg, AmgHierarchy::Transfer &T, cudaStream_t st);

AmgHierarchy::~AmgHierarchy()
{
    for (auto *c : coarse) xfemm_linprob_destroy(c);
    xfer.clear();
    if (arena.base) dev_free(arena.base);
}

// scope guard: DevBuf allocations of this thread come from `a` while alive
struct ArenaScope
{
    DevArena *prev;
    explicit ArenaScope(DevArena *a) : prev(g_arena) { if (a && a->base) g_arena = a; }
    ~ArenaScope() { g_arena = prev; }
};

// Greedy distance-1 aggregation on the pattern.  Rows without off-diagonals stay out (-1): the
// smoother alone is exact for them.  Wide rows (air-gap rings, see layout.h) never seed an
// aggregate — their neighbourhood is the whole ring — they join the aggregate of a mesh neighbour.
int aggregate(int n, const int *rp, const int *col, int wide_thr, std::vector<int> &agg)
{
    agg.assign(n, -1);
    int na = 0;
    auto is_wide = [&](int i) { return rp[i + 1] - rp[i] - 1 > wide_thr; };
    // pass 1: a row whose whole neighbourhood is still free becomes the root of a new aggregate
    for (int i = 0; i < n; i++)
    {
        if (agg[i] >= 0 || rp[i + 1] - rp[i] <= 1 || is_wide(i)) continue;
        bool free_nb = true;
        for (int s = rp[i]; s < rp[i + 1] && free_nb; s++) free_nb = agg[col[s]] < 0;
        if (!free_nb) continue;
        for (int s = rp[i]; s < rp[i + 1]; s++) agg[col[s]] = na;
        agg[i] = na++;
    }
    // pass 2: leftovers join the neighbouring aggregate (of pass 1) they are connected to most — every row decides from
    // the pass-1 state alone, so the host threads share the rows
    std::vector<int> agg1(agg);
    parallel_for(n, [&](int i0, int i1, int) {
        std::vector<std::pair<int, int>> cnt;
        for (int i = i0; i < i1; i++)
        {
            if (agg1[i] >= 0 || rp[i + 1] - rp[i] <= 1) continue;
            const bool wi = is_wide(i);
            cnt.clear();
            for (int s = rp[i]; s < rp[i + 1]; s++)
            {
                const int j = col[s];
                if (j == i || agg1[j] < 0 || (wi && is_wide(j))) continue;
                bool found = false;
                for (auto &c : cnt) if (c.first == agg1[j]) { c.second++; found = true; break; }
                if (!found) cnt.emplace_back(agg1[j], 1);
            }
            int best = -1, bc = 0;
            for (auto &c : cnt) if (c.second > bc || (c.second == bc && c.first < best)) { best = c.first; bc = c.second; }
            if (best >= 0) agg[i] = best;
        }
    }, 65536);
    // pass 3: what is still free forms aggregates with its free neighbours
    for (int i = 0; i < n; i++)
    {
        if (agg[i] >= 0 || rp[i + 1] - rp[i] <= 1) continue;
        const bool wi = is_wide(i);
        agg[i] = na;
        if (!wi)
            for (int s = rp[i]; s < rp[i + 1]; s++)
                if (agg[col[s]] < 0 && rp[col[s] + 1] - rp[col[s]] > 1) agg[col[s]] = na;
        na++;
    }
    return na;
}

// The first coarsening only needs the pattern: a host thread starts on it the moment a problem is created, next to
// the uploads, the assembly set-up and the device layout; amg_build() collects the result.  Small problems and
// XFEMM_NO_EARLY_AGG skip it.
void linprob_start_aggregation(xfemm_linprob *p)
{
    if (p->n < 20000 || getenv("XFEMM_NO_EARLY_AGG") || p->agg_thread.joinable()) return;
    const char *ew = getenv("XFEMM_WIDE_ROW");
    p->agg_wide_thr = ew ? atoi(ew) : 64;
    p->agg_thread = std::thread([p]() { p->agg_early_nc = aggregate(p->n, p->h_rowptr.data(), p->h_col.data(), p->agg_wide_thr, p->agg_early); });
}

// coarse pattern (sorted CSR) + for every coarse slot the fine slots that sum into it
void coarse_pattern(int n, const int *rp, const int *col, const std::vector<int> &agg, int nc,
                           std::vector<int> &crp, std::vector<int> &ccol, std::vector<int> &gptr, std::vector<int> &gsrc)
{
    // members of each aggregate
    std::vector<int> mptr(nc + 1, 0), mem(n);
    for (int i = 0; i < n; i++) if (agg[i] >= 0) mptr[agg[i] + 1]++;
    for (int I = 0; I < nc; I++) mptr[I + 1] += mptr[I];
    {
        std::vector<int> pos(mptr.begin(), mptr.end() - 1);
        for (int i = 0; i < n; i++) if (agg[i] >= 0) mem[pos[agg[i]]++] = i;
    }
    // Every coarse row is built from its own members only, so the three passes below run in
    // parallel over coarse rows (per-thread marker arrays); results do not depend on the thread count.
    // pass 1: distinct coarse columns per coarse row
    crp.assign(nc + 1, 0);
    std::vector<std::vector<int>> tcols(64);
    std::vector<std::pair<int, int>> trange(64, {0, 0});
    parallel_for(nc, [&](int I0, int I1, int tid) {
        std::vector<int> mark(nc, -1), cols, &out = tcols[tid];
        trange[tid] = {I0, I1};
        for (int I = I0; I < I1; I++)
        {
            cols.clear();
            for (int k = mptr[I]; k < mptr[I + 1]; k++)
            {
                const int i = mem[k];
                for (int s = rp[i]; s < rp[i + 1]; s++)
                {
                    const int J = agg[col[s]];
                    if (J >= 0 && mark[J] != I) { mark[J] = I; cols.push_back(J); }
                }
            }
            std::sort(cols.begin(), cols.end());
            out.insert(out.end(), cols.begin(), cols.end());
            crp[I + 1] = (int)cols.size();
        }
    }, 4096);
    for (int I = 0; I < nc; I++) crp[I + 1] += crp[I];
    const int nnzc = crp[nc];
    ccol.resize(nnzc);
    for (int t = 0; t < 64; t++)
        if (!tcols[t].empty()) std::copy(tcols[t].begin(), tcols[t].end(), ccol.begin() + crp[trange[t].first]);
    // pass 2: how many fine slots feed each coarse slot
    gptr.assign(nnzc + 1, 0);
    parallel_for(nc, [&](int I0, int I1, int) {
        std::vector<int> pos(nc, -1);
        for (int I = I0; I < I1; I++)
        {
            for (int t = crp[I]; t < crp[I + 1]; t++) pos[ccol[t]] = t;
            for (int k = mptr[I]; k < mptr[I + 1]; k++)
            {
                const int i = mem[k];
                for (int s = rp[i]; s < rp[i + 1]; s++)
                {
                    const int J = agg[col[s]];
                    if (J >= 0) gptr[pos[J] + 1]++;
                }
            }
        }
    }, 4096);
    for (int t = 0; t < nnzc; t++) gptr[t + 1] += gptr[t];
    // pass 3: the fine slots of each coarse slot, in ascending fine-slot order within a member row
    gsrc.assign(gptr[nnzc], 0);
    parallel_for(nc, [&](int I0, int I1, int) {
        std::vector<int> pos(nc, -1), fill;
        for (int I = I0; I < I1; I++)
        {
            fill.assign(gptr.begin() + crp[I], gptr.begin() + crp[I + 1]);
            for (int t = crp[I]; t < crp[I + 1]; t++) pos[ccol[t]] = t - crp[I];
            for (int k = mptr[I]; k < mptr[I + 1]; k++)
            {
                const int i = mem[k];
                for (int s = rp[i]; s < rp[i + 1]; s++)
                {
                    const int J = agg[col[s]];
                    if (J >= 0) gsrc[fill[pos[J]]++] = s;
                }
            }
        }
    }, 4096);
}

// coarse row of the column of every SELL slot (team levels: prolongation fused into the post-smoothing pass)
__global__ void k_amg_cagg(long long len, const int *__restrict__ sell_col, const int *__restrict__ agg, int *__restrict__ cagg)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < len) cagg[i] = agg[sell_col[i]];
}

__global__ void k_amg_galerkin(int nnzc, const int *__restrict__ gptr, const int *__restrict__ gsrc,
                               const double *__restrict__ fine_vals, double *__restrict__ coarse_vals)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnzc) return;
    double s = 0.0;
    for (int k = gptr[t]; k < gptr[t + 1]; k++) s += fine_vals[gsrc[k]];
    coarse_vals[t] = s;
}

// The same with one WARP per coarse slot, for the small levels: the couplings of the air-gap ring funnel into a
// few coarse slots whose gather lists hold thousands of entries, and a single thread walking such a list (two
// dependent loads per entry) took 0.2 ms on the last level.  Lane-strided partial sums, fixed-order shuffle tree.
__global__ void k_amg_galerkin_warp(int nnzc, const int *__restrict__ gptr, const int *__restrict__ gsrc,
                                    const double *__restrict__ fine_vals, double *__restrict__ coarse_vals)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= nnzc) return;
    double s = 0.0;
    for (int k = gptr[t] + lane; k < gptr[t + 1]; k += 32) s += fine_vals[gsrc[k]];
    s = warp_sum(s);
    if (lane == 0) coarse_vals[t] = s;
}

// Dense inverse of the coarsest level by Gauss-Jordan without pivoting (the level is symmetric positive definite),
// in place in shared memory: at step k column k of the work matrix takes the k-th column of the inverse-in-progress —
// the operations of an elimination on [A | I], without storing the unit columns.  One block; at most AMG_DENSE_MAX
// rows (128 x 128 doubles = 128 KB of dynamic shared memory).  Rows/columns in the internal order of the level;
// padding rows are identity rows.
constexpr int AMG_DENSE_MAX = 128;
constexpr int AMG_TEAM_ROWS_MAX = 1 << 18; // no level with more rows is ever given to the team
__global__ void __launch_bounds__(1024)
k_amg_dense_inverse(int n, int n_pad, const int *__restrict__ rowptr, const int *__restrict__ col,
                    const double *__restrict__ vals, const int *__restrict__ c2i, double *inv)
{
    extern __shared__ double dense_w[]; // n_pad x n_pad, then n_pad for the pivot column
    double *W = dense_w, *colk = dense_w + (size_t)n_pad * n_pad;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int e = tid; e < n_pad * n_pad; e += nthr) W[e] = (e / n_pad == e % n_pad) ? 1.0 : 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += nthr) W[c2i[i] * n_pad + c2i[i]] = 0.0;
    __syncthreads();
    for (int i = tid; i < n; i += nthr) // one thread per row: its entries are distinct slots
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++) W[c2i[i] * n_pad + c2i[col[q]]] = vals[q];
    __syncthreads();
    for (int k = 0; k < n_pad; k++)
    {
        const double piv = W[k * n_pad + k];
        for (int r = tid; r < n_pad; r += nthr) colk[r] = W[r * n_pad + k];
        __syncthreads();
        // row k: divided by the pivot; its own column holds the unit column's 1 / piv
        for (int c = tid; c < n_pad; c += nthr) W[k * n_pad + c] = (c == k ? 1.0 : W[k * n_pad + c]) / piv;
        __syncthreads();
        for (int e = tid; e < n_pad * n_pad; e += nthr)
        {
            const int r = e / n_pad, c = e % n_pad;
            if (r != k) W[e] = fma(-colk[r], W[k * n_pad + c], c == k ? 0.0 : W[e]);
        }
        __syncthreads();
    }
    for (int e = tid; e < n_pad * n_pad; e += nthr) inv[e] = W[e];
}

int amg_build(xfemm_linprob *fine, std::unique_ptr<AmgHierarchy> &out, std::thread *fine_layout)
{
    std::unique_ptr<AmgHierarchy> h(new AmgHierarchy);
    const char *e = getenv("XFEMM_AMG_MIN_COARSE");
    const int min_coarse = e ? atoi(e) : 64;
    const char *ew = getenv("XFEMM_WIDE_ROW");
    const int wide_thr = ew ? atoi(ew) : 64;
    // levels with at most this many rows are cycled by a team of blocks with its own barrier (amg_team_cycle)
    const char *etr = getenv("XFEMM_AMG_TEAM_ROWS");
    h->team_rows = std::min(AMG_TEAM_ROWS_MAX, etr ? atoi(etr) : 65536);
    xfemm_linprob *cur = fine;
    h->level_n.push_back(fine->n);
    h->level_nnz.push_back((int)fine->nnz);
    const bool on_device = !getenv("XFEMM_HOST_HIERARCHY");
    const bool check = on_device && getenv("XFEMM_HIERARCHY_CHECK");
    auto differs = [&](const char *what, int level) { set_error("device hierarchy set-up differs from the host's: %s (level %d)", what, level); return XFEMM_ERR_ARG; };
    while ((int)h->coarse.size() + 1 < AMG_MAX_LEVELS && cur->n > min_coarse)
    {
        std::vector<int> agg, crp, ccol, gptr, gsrc;
        const bool timing = getenv("XFEMM_TIMING") != nullptr;
        auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        const double ta = now();
        // the aggregates: a sequential greedy pass on the host — started when the pattern arrived (linprob_start_aggregation)
        int nc;
        if (cur->agg_thread.joinable() && cur->agg_wide_thr == wide_thr)
        {
            cur->agg_thread.join();
            agg.swap(cur->agg_early);
            nc = cur->agg_early_nc;
        }
        else
        {
            if (cur->agg_thread.joinable()) cur->agg_thread.join();
            nc = aggregate(cur->n, cur->h_rowptr.data(), cur->h_col.data(), wide_thr, agg);
        }
        if (nc < 1 || nc > cur->n * 0.7) break; // coarsening stalled
        const double tb = now();
        cudaStream_t st = fine->stream;
        h->xfer.emplace_back(new AmgHierarchy::Transfer);
        AmgHierarchy::Transfer &T = *h->xfer.back();
        DevBuf<int> d_agg;
        if (on_device)
        {
            XB_CUDA(d_agg.upload(agg.data(), agg.size(), st));
            int rc = amg_coarse_pattern_device(cur, d_agg.p, nc, crp, ccol, T.gptr, T.gsrc, st);
            if (rc != XFEMM_OK) return rc;
        }
        if (!on_device || check)
        {
            std::vector<int> crp2, ccol2;
            coarse_pattern(cur->n, cur->h_rowptr.data(), cur->h_col.data(), agg, nc, check ? crp2 : crp, check ? ccol2 : ccol, gptr, gsrc);
            if (check)
            {
                if (crp2 != crp || ccol2 != ccol) return differs("coarse pattern", (int)h->coarse.size());
                std::vector<int> a(gptr.size()), b(gsrc.size());
                XB_CUDA(cudaMemcpy(a.data(), T.gptr.p, a.size() * sizeof(int), cudaMemcpyDeviceToHost));
                XB_CUDA(cudaMemcpy(b.data(), T.gsrc.p, b.size() * sizeof(int), cudaMemcpyDeviceToHost));
                if (T.gptr.n != gptr.size() || a != gptr || b != gsrc) return differs("Galerkin gather lists", (int)h->coarse.size());
            }
            else
            {
                XB_CUDA(T.gptr.upload(gptr.data(), gptr.size(), st));
                XB_CUDA(T.gsrc.upload(gsrc.data(), gsrc.size(), st));
            }
        }
        T.nnz_coarse = (long long)ccol.size();
        const double tc = now();
        // levels from the third on live in the L2-pinned arena (see AmgHierarchy::arena)
        // XFEMM_AMG_PIN_L1: the second level's solver layout (SELL stream, vectors) as well — not its CSR arrays
        const bool pin_l1 = h->coarse.empty() && getenv("XFEMM_AMG_PIN_L1") && !getenv("XFEMM_AMG_NO_PIN");
        const bool pinned = h->coarse.size() >= 1 && !getenv("XFEMM_AMG_NO_PIN");
        if ((pinned || pin_l1) && !h->arena.base)
        {
            const size_t cap = pin_l1 ? (size_t)(1.2 * (18.0 * ccol.size() + 100.0 * nc) + 0.25 * (46.0 * ccol.size() + 150.0 * nc)) + (2u << 20)
                                      : (size_t)(1.5 * (34.0 * ccol.size() + 150.0 * nc + 12.0 * cur->n)) + (2u << 20);
            if (dev_alloc((void **)&h->arena.base, cap) == cudaSuccess) { h->arena.cap = cap; h->arena.used = 0; }
            else { h->arena.base = nullptr; (void)cudaGetLastError(); }
        }
        // the layout of the finest level (internal order, device copy of the coordinates) is needed from here on
        if (fine_layout && fine_layout->joinable()) fine_layout->join();
        if (cur == fine && fine->layout_precond < 0) { set_error("the solver layout of the finest level is missing"); return XFEMM_ERR_ARG; }
        xfemm_linprob *c = nullptr;
        int rc;
        {
            ArenaScope scope(pinned ? &h->arena : nullptr);
            rc = xfemm_linprob_create(nc, crp.data(), ccol.data(), &c);
            if (rc != XFEMM_OK) return rc;
            h->coarse.push_back(c);
            // small levels (the team's): the task-split SpMV pass (amg_spmv_tasks) takes slices of any width up to 256
            // columns in its stride, so their long rows (air-gap ring couplings) stay in the SELL slices — no
            // wide-row pre-pass, and the prolongation can be fused into the post-smoothing pass
            if (nc <= h->team_rows && !getenv("XFEMM_AMG_TEAM_WIDE_ROWS")) c->wide_threshold = 224;
            if ((int)cur->h_xy.size() == 2 * cur->n)
            {
                // coarse "node" = centroid of its aggregate: the coarse level follows the same Z-curve
                const bool dev_xy = on_device && cur->d_xy.n == 2 * (size_t)cur->n;
                if (dev_xy)
                {
                    rc = amg_centroids_device(cur, c, d_agg.p, st);
                    if (rc != XFEMM_OK) return rc;
                }
                if (!dev_xy || check)
                {
                    std::vector<double> sx(nc, 0.), sy(nc, 0.);
                    std::vector<int> cnt(nc, 0);
                    for (int i = 0; i < cur->n; i++)
                        if (agg[i] >= 0) { sx[agg[i]] += cur->h_xy[2 * (size_t)i]; sy[agg[i]] += cur->h_xy[2 * (size_t)i + 1]; cnt[agg[i]]++; }
                    std::vector<float> hxy(2 * (size_t)nc);
                    for (int I = 0; I < nc; I++)
                    {
                        hxy[2 * (size_t)I] = (float)(sx[I] / std::max(1, cnt[I]));
                        hxy[2 * (size_t)I + 1] = (float)(sy[I] / std::max(1, cnt[I]));
                    }
                    if (dev_xy && hxy != c->h_xy) return differs("centroids", (int)h->coarse.size());
                    c->h_xy.swap(hxy);
                }
            }
            {
                ArenaScope hot(pin_l1 ? &h->arena : nullptr);
                rc = linprob_build_layout(c, XFEMM_PRECOND_JACOBI);
            }
            if (rc != XFEMM_OK) return rc;
        }
        // transfer maps in the internal (SELL) row orders of both levels
        const SellLayout &LF = cur->lay, &LC = c->lay;
        {
            ArenaScope scope(pinned ? &h->arena : nullptr);
            if (on_device)
            {
                rc = amg_transfer_maps_device(cur, c, d_agg.p, T, st);
                if (rc != XFEMM_OK) return rc;
            }
            if (!on_device || check)
            {
                std::vector<int> agg_int(LF.n_pad, -1), mptr(LC.n_pad + 1, 0), mem;
                for (int r = 0; r < LF.n_pad; r++)
                {
                    const int i = LF.i2c[r];
                    if (i >= 0 && agg[i] >= 0) { agg_int[r] = LC.c2i[agg[i]]; mptr[agg_int[r] + 1]++; }
                }
                for (int R = 0; R < LC.n_pad; R++) mptr[R + 1] += mptr[R];
                mem.assign(std::max(1, mptr[LC.n_pad]), 0);
                {
                    std::vector<int> pos(mptr.begin(), mptr.end() - 1);
                    for (int r = 0; r < LF.n_pad; r++) if (agg_int[r] >= 0) mem[pos[agg_int[r]]++] = r;
                }
                if (check)
                {
                    std::vector<int> a(agg_int.size()), b(mptr.size()), m2(mptr[LC.n_pad]);
                    XB_CUDA(cudaMemcpy(a.data(), T.agg.p, a.size() * sizeof(int), cudaMemcpyDeviceToHost));
                    XB_CUDA(cudaMemcpy(b.data(), T.mptr.p, b.size() * sizeof(int), cudaMemcpyDeviceToHost));
                    XB_CUDA(cudaMemcpy(m2.data(), T.mem.p, m2.size() * sizeof(int), cudaMemcpyDeviceToHost));
                    mem.resize(m2.size());
                    if (a != agg_int || b != mptr || m2 != mem) return differs("transfer maps", (int)h->coarse.size() - 1);
                }
                else
                {
                    XB_CUDA(T.agg.upload(agg_int.data(), agg_int.size(), st));
                    XB_CUDA(T.mptr.upload(mptr.data(), mptr.size(), st));
                    XB_CUDA(T.mem.upload(mem.data(), mem.size(), st));
                }
            }
            // a level the team may cycle (amg_fill_args decides): coarse row per SELL slot, unless it has wide rows
            if (cur != fine && cur->n <= h->team_rows && cur->wide.n_rows == 0 && LF.sell_len > 0)
            {
                XB_CUDA(T.cagg.alloc((size_t)LF.sell_len));
                ++g_launch, k_amg_cagg<<<(unsigned)((LF.sell_len + 255) / 256), 256, 0, st>>>(LF.sell_len, cur->sell_col.p, T.agg.p, T.cagg.p);
                XB_CUDA(cudaGetLastError());
            }
        }
        XB_CUDA(cudaStreamSynchronize(st));
        if (timing && cur->n > 100000)
            fprintf(stderr, "[xfemm timing]     level %d -> %d rows: aggregate %.3f s, coarse pattern %.3f s, level problem + layout + transfer maps %.3f s\n",
                    cur->n, nc, tb - ta, tc - tb, now() - tc);
        h->level_n.push_back(nc);
        h->level_nnz.push_back((int)ccol.size());
        cur = c;
    }
    XB_CUDA(h->team_ctr.alloc(64));
    XB_CUDA(h->team_ctr.zero(fine->stream));
    // exact solve on the coarsest level when it is small enough for a dense inverse
    if (!h->coarse.empty() && h->coarse.back()->lay.n_pad <= AMG_DENSE_MAX && !getenv("XFEMM_AMG_NO_DENSE"))
    {
        ArenaScope scope(h->arena.base ? &h->arena : nullptr);
        const size_t np = (size_t)h->coarse.back()->lay.n_pad;
        XB_CUDA(h->dense_inv.alloc(np * np));
    }
    // pin the arena in L2 for everything launched on the solver stream
    if (h->arena.base && h->arena.used > 0)
    {
        int dev = 0, max_persist = 0, max_window = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t want = std::min<size_t>(h->arena.used, (size_t)std::max(0, max_persist));
        if (want > 0 && max_window > 0)
        {
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = h->arena.base;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(h->arena.used, (size_t)max_window);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            if (cudaStreamSetAttribute(fine->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) (void)cudaGetLastError();
        }
        if (getenv("XFEMM_TIMING"))
            fprintf(stderr, "[xfemm timing] multigrid arena %.1f MB of %.1f MB pinned in L2 (device limit %.1f MB)\n",
                    h->arena.used / 1048576.0, h->arena.cap / 1048576.0, max_persist / 1048576.0);
    }
    out = std::move(h);
    return XFEMM_OK;
}

// level l+1 values = P^T A_l P (gather-sum), then the SELL / diagonal refill of that level
int amg_refresh(xfemm_linprob *fine, AmgHierarchy &h)
{
    // everything on the finest level's stream, nothing read back: one queue of small kernels, no host round trips
    cudaStream_t st = fine->stream;
    xfemm_linprob *cur = fine;
    for (size_t l = 0; l < h.coarse.size(); l++)
    {
        xfemm_linprob *c = h.coarse[l];
        AmgHierarchy::Transfer &T = *h.xfer[l];
        const int nnzc = (int)T.nnz_coarse;
        if (nnzc > 65536) ++g_launch, k_amg_galerkin<<<(nnzc + 255) / 256, 256, 0, st>>>(nnzc, T.gptr.p, T.gsrc.p, cur->vals.p, c->vals.p);
        else ++g_launch, k_amg_galerkin_warp<<<(unsigned)(((size_t)nnzc * 32 + 255) / 256), 256, 0, st>>>(nnzc, T.gptr.p, T.gsrc.p, cur->vals.p, c->vals.p);
        XB_CUDA(cudaGetLastError());
        int rc = linprob_refresh_values_on(c, st);
        if (rc != XFEMM_OK) return rc;
        cur = c;
    }
    if (h.dense_inv.p)
    {
        xfemm_linprob *c = h.coarse.back();
        const int np = c->lay.n_pad;
        const size_t smem = ((size_t)np * np + np) * sizeof(double);
        static bool attr_set[64] = {}; // per device: function attributes belong to the device's context
        const int dev = fine->device >= 0 && fine->device < 64 ? fine->device : 0;
        if (!attr_set[dev])
        {
            XB_CUDA(cudaFuncSetAttribute(k_amg_dense_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(((size_t)AMG_DENSE_MAX * AMG_DENSE_MAX + AMG_DENSE_MAX) * sizeof(double))));
            attr_set[dev] = true;
        }
        ++g_launch, k_amg_dense_inverse<<<1, 1024, smem, st>>>(c->n, np, c->rowptr.p, c->col.p, c->vals.p, c->c2i.p, h.dense_inv.p);
        XB_CUDA(cudaGetLastError());
    }
    return XFEMM_OK;
}

void amg_print_trace(AmgHierarchy &h, cudaStream_t st)
{
    if (!h.stamps.p) return;
    std::vector<unsigned long long> t(AMG_MAX_STAMPS + 16 + AMG_BLOCK_STAMPS, 0);
    cudaMemcpyAsync(t.data(), h.stamps.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[amg trace] levels:");
    for (size_t l = 0; l < h.level_n.size(); l++)
        fprintf(stderr, " %d/%d(w%d)", h.level_n[l], h.level_nnz[l], l == 0 ? -1 : h.coarse[l - 1]->wide.n_rows);
    fprintf(stderr, "\n[amg trace] phase us:");
    double total = 0;
    for (int k = 1; k < AMG_MAX_STAMPS - 8 && t[k] > t[k - 1]; k++)
    {
        fprintf(stderr, " %.1f", (t[k] - t[k - 1]) * 1e-3);
        total += (t[k] - t[k - 1]) * 1e-3;
    }
    fprintf(stderr, "  (cycle %.1f us)\n", total);
    const unsigned long long *dd = t.data() + AMG_MAX_STAMPS;
    if (dd[6] > dd[0])
        fprintf(stderr, "[amg trace] first team SpMV pass, block 0 (us): table %.2f, row-operand issue %.2f, tasks %.2f, sync %.2f, sums+epilogue %.2f, sync %.2f\n",
                (dd[1] - dd[0]) * 1e-3, (dd[2] - dd[1]) * 1e-3, (dd[3] - dd[2]) * 1e-3, (dd[4] - dd[3]) * 1e-3, (dd[5] - dd[4]) * 1e-3, (dd[6] - dd[5]) * 1e-3);
    for (int ph = 0; ph < 3; ph++)
    {
        const unsigned long long *bs = t.data() + AMG_MAX_STAMPS + 16 + 160 * ph, t0 = bs[159];
        if (!t0) continue;
        double mn = 1e30, mx = 0, sum = 0; int cnt = 0, arg = -1;
        for (int b = 0; b < 159; b++) if (bs[b] > t0) { double d = (bs[b] - t0) * 1e-3; mn = std::min(mn, d); if (d > mx) { mx = d; arg = b; } sum += d; cnt++; }
        if (cnt) fprintf(stderr, "[amg trace] %s: blocks finish after min %.1f / mean %.1f / max %.1f us (block %d last, %d blocks)\n",
                         ph == 0 ? "CG SpMV" : ph == 1 ? "level-0 residual SpMV" : "level-0 post-smoothing SpMV", mn, sum / cnt, mx, arg, cnt);
        if (cnt && ph == 0)
        {
            fprintf(stderr, "[amg trace]   per block (us):");
            for (int b = 0; b < 159; b++) if (bs[b] > t0) fprintf(stderr, " %.0f", (bs[b] - t0) * 1e-3);
            fprintf(stderr, "\n");
        }
    }
    const unsigned long long *c = t.data() + AMG_MAX_STAMPS - 8; // CG part of the last iteration
    if (c[4] > c[0])
        fprintf(stderr, "[amg trace] cg part us: spmv+dot %.1f, residual update %.1f, (v-cycle %.1f), direction+solution update %.1f\n",
                (c[1] - c[0]) * 1e-3, (c[2] - c[1]) * 1e-3, (c[3] - c[2]) * 1e-3, (c[4] - c[3]) * 1e-3);
}

static void fill_level(AmgLev &L, xfemm_linprob *p)
{
    L.n_pad = p->lay.n_pad; L.nslices = p->lay.nslices;
    L.slice_ptr = p->slice_ptr.p; L.sell_col = p->sell_col.p;
    L.sell_val = p->sell_val.p; L.diag = p->diag.p; L.dinv = p->dinv.p;
    L.wide = p->wide.view(); L.wide_val = p->wide.val.p; L.wide_acc = p->wide.acc.p;
    L.x = p->x.p; L.b = p->bint.p; L.t = p->u.p;
}

void amg_fill_args(xfemm_linprob *fine, AmgHierarchy &h, AmgArgs &g, int grid)
{
    g.nlev = 1 + (int)h.coarse.size();
    const char *e;
    g.omega = (e = getenv("XFEMM_AMG_OMEGA")) ? atof(e) : 0.8;
    g.over = (e = getenv("XFEMM_AMG_OVER")) ? atof(e) : 2.6;
    g.coarse_sweeps = ((e = getenv("XFEMM_AMG_COARSE_SWEEPS")) ? atoi(e) : 9) | 1;
    g.team_from = std::max(1, g.nlev - 1);
    for (int l = 1; l < g.nlev; l++)
        if (h.level_n[l] <= h.team_rows) { g.team_from = l; break; }
    g.team_blocks = std::max(1, std::min(grid, (e = getenv("XFEMM_AMG_TEAM_BLOCKS")) ? atoi(e) : 48));
    // the smallest levels (at most this many rows) are block 0's alone; never the first team level
    const int solo_rows = (e = getenv("XFEMM_AMG_SOLO_ROWS")) ? atoi(e) : 0; // measured neutral at 2 M nodes (0.294 vs 0.293 ms/iteration): off by default
    g.solo_from = 1 << 20;
    for (int l = g.team_from + 1; l < g.nlev; l++)
        if (h.level_n[l] <= solo_rows) { g.solo_from = l; break; }
    // restriction by scatter (AmgArgs::scatter_restrict): real one-component cycles with an exact coarsest solve, unless
    // the run has to be reproducible bit for bit
    g.scatter_restrict = (h.dense_inv.p && !fine->deterministic && !getenv("XFEMM_DETERMINISTIC") && !getenv("XFEMM_AMG_ORDERED_RESTRICT")) ? 1 : 0;
    g.team_ctr = h.team_ctr.p;
    g.grid_ctr = getenv("XFEMM_CG_GRID_SYNC") ? nullptr : h.team_ctr.p + 32;
    // both barriers count from zero in every launch; the launch follows on this stream
    if (cudaMemsetAsync(h.team_ctr.p, 0, 64 * sizeof(unsigned), fine->stream) != cudaSuccess) (void)cudaGetLastError();
    g.coarse_inv = h.dense_inv.p;
    g.stamps = nullptr;
    if (getenv("XFEMM_AMG_TRACE"))
    {
        if (!h.stamps.p && h.stamps.alloc(AMG_MAX_STAMPS + 16 + AMG_BLOCK_STAMPS) == cudaSuccess) h.stamps.zero(fine->stream);
        g.stamps = h.stamps.p;
        g.trace_detail = atoi(getenv("XFEMM_AMG_TRACE")) >= 2 ? atoi(getenv("XFEMM_AMG_TRACE")) - 1 : 0;
    }
    fill_level(g.lev[0], fine);
    g.lev[0].b = fine->r.p; // CG residual in, z out
    g.lev[0].x = fine->w.p;
    g.lev[0].t = fine->z.p;
    for (size_t l = 0; l < h.coarse.size(); l++)
    {
        fill_level(g.lev[l + 1], h.coarse[l]);
        g.lev[l].agg = h.xfer[l]->agg.p;
        g.lev[l].mptr = h.xfer[l]->mptr.p;
        g.lev[l].mem = h.xfer[l]->mem.p;
        g.lev[l].sell_cagg = ((int)l >= g.team_from && !getenv("XFEMM_AMG_NO_FUSED_PROLONGATION")) ? h.xfer[l]->cagg.p : nullptr;
    }
}

// vectors of the second right-hand side (amg2.cuh): the work vectors of each level's problem that the first
// component leaves unused.  `fine` = the real companion of a complex problem: none of its own vectors is in use.
void amg_fill_args2(xfemm_linprob *fine, AmgHierarchy &h, AmgArgs2 &g2)
{
    g2.lev[0].b = fine->bint.p;
    g2.lev[0].x = fine->x.p;
    g2.lev[0].t = fine->u.p;
    for (size_t l = 0; l < h.coarse.size(); l++)
    {
        xfemm_linprob *c = h.coarse[l];
        g2.lev[l + 1].x = c->p.p;
        g2.lev[l + 1].b = c->r.p;
        g2.lev[l + 1].t = c->z.p;
    }
}

} // namespace xb
