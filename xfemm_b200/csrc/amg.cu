// Host side of the aggregation-multigrid preconditioner (see amg.cuh): hierarchy construction from
// the sparsity pattern (once per pattern) and the per-solve Galerkin value refresh on the device.
#include "amg.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace xb {

int linprob_refresh_values(xfemm_linprob *p, int *singular_row); // linprob.cu

AmgHierarchy::~AmgHierarchy()
{
    for (auto *c : coarse) xfemm_linprob_destroy(c);
}

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
    // pass 2: leftovers join the neighbouring aggregate (of pass 1) they are connected to most
    std::vector<int> agg1(agg);
    std::vector<std::pair<int, int>> cnt;
    for (int i = 0; i < n; i++)
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

__global__ void k_amg_galerkin(int nnzc, const int *__restrict__ gptr, const int *__restrict__ gsrc,
                               const double *__restrict__ fine_vals, double *__restrict__ coarse_vals)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nnzc) return;
    double s = 0.0;
    for (int k = gptr[t]; k < gptr[t + 1]; k++) s += fine_vals[gsrc[k]];
    coarse_vals[t] = s;
}

int amg_build(xfemm_linprob *fine, std::unique_ptr<AmgHierarchy> &out)
{
    std::unique_ptr<AmgHierarchy> h(new AmgHierarchy);
    const char *e = getenv("XFEMM_AMG_MIN_COARSE");
    const int min_coarse = e ? atoi(e) : 64;
    const char *ew = getenv("XFEMM_WIDE_ROW");
    const int wide_thr = ew ? atoi(ew) : 64;
    xfemm_linprob *cur = fine;
    h->level_n.push_back(fine->n);
    h->level_nnz.push_back((int)fine->nnz);
    while ((int)h->coarse.size() + 1 < AMG_MAX_LEVELS && cur->n > min_coarse)
    {
        std::vector<int> agg, crp, ccol, gptr, gsrc;
        const int nc = aggregate(cur->n, cur->h_rowptr.data(), cur->h_col.data(), wide_thr, agg);
        if (nc < 1 || nc > cur->n * 0.7) break; // coarsening stalled
        coarse_pattern(cur->n, cur->h_rowptr.data(), cur->h_col.data(), agg, nc, crp, ccol, gptr, gsrc);
        xfemm_linprob *c = nullptr;
        int rc = xfemm_linprob_create(nc, crp.data(), ccol.data(), &c);
        if (rc != XFEMM_OK) return rc;
        h->coarse.push_back(c);
        rc = linprob_build_layout(c, XFEMM_PRECOND_JACOBI);
        if (rc != XFEMM_OK) return rc;
        h->xfer.emplace_back(new AmgHierarchy::Transfer);
        AmgHierarchy::Transfer &T = *h->xfer.back();
        T.nnz_coarse = (long long)ccol.size();
        cudaStream_t st = fine->stream;
        XB_CUDA(T.gptr.upload(gptr.data(), gptr.size(), st));
        XB_CUDA(T.gsrc.upload(gsrc.data(), gsrc.size(), st));
        // transfer maps in the internal (SELL) row orders of both levels
        const SellLayout &LF = cur->lay, &LC = c->lay;
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
        XB_CUDA(T.agg.upload(agg_int.data(), agg_int.size(), st));
        XB_CUDA(T.mptr.upload(mptr.data(), mptr.size(), st));
        XB_CUDA(T.mem.upload(mem.data(), mem.size(), st));
        XB_CUDA(cudaStreamSynchronize(st));
        h->level_n.push_back(nc);
        h->level_nnz.push_back((int)ccol.size());
        cur = c;
    }
    out = std::move(h);
    return XFEMM_OK;
}

// level l+1 values = P^T A_l P (gather-sum), then the SELL / diagonal refill of that level
int amg_refresh(xfemm_linprob *fine, AmgHierarchy &h)
{
    cudaStream_t st = fine->stream;
    xfemm_linprob *cur = fine;
    for (size_t l = 0; l < h.coarse.size(); l++)
    {
        xfemm_linprob *c = h.coarse[l];
        AmgHierarchy::Transfer &T = *h.xfer[l];
        const int nnzc = (int)T.nnz_coarse;
        ++g_launch, k_amg_galerkin<<<(nnzc + 255) / 256, 256, 0, st>>>(nnzc, T.gptr.p, T.gsrc.p, cur->vals.p, c->vals.p);
        XB_CUDA(cudaGetLastError());
        // the coarse problem owns its own stream: order its refill behind the Galerkin kernel
        XB_CUDA(cudaStreamSynchronize(st));
        int singular = -1;
        int rc = linprob_refresh_values(c, &singular);
        if (rc != XFEMM_OK) return rc;
        XB_CUDA(cudaStreamSynchronize(c->stream));
        cur = c;
    }
    return XFEMM_OK;
}

void amg_print_trace(AmgHierarchy &h, cudaStream_t st)
{
    if (!h.stamps.p) return;
    std::vector<unsigned long long> t(AMG_MAX_STAMPS, 0);
    cudaMemcpyAsync(t.data(), h.stamps.p, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[amg trace] levels:");
    for (size_t l = 0; l < h.level_n.size(); l++) fprintf(stderr, " %d/%d", h.level_n[l], h.level_nnz[l]);
    fprintf(stderr, "\n[amg trace] phase us:");
    double total = 0;
    for (int k = 1; k < AMG_MAX_STAMPS && t[k] > t[k - 1]; k++)
    {
        fprintf(stderr, " %.1f", (t[k] - t[k - 1]) * 1e-3);
        total += (t[k] - t[k - 1]) * 1e-3;
    }
    fprintf(stderr, "  (cycle %.1f us)\n", total);
}

static void fill_level(AmgLev &L, xfemm_linprob *p)
{
    L.n_pad = p->lay.n_pad; L.nslices = p->lay.nslices;
    L.slice_ptr = p->slice_ptr.p; L.sell_col = p->sell_col.p;
    L.sell_val = p->sell_val.p; L.diag = p->diag.p; L.dinv = p->dinv.p;
    L.wide = p->wide.view(); L.wide_val = p->wide.val.p; L.wide_acc = p->wide.acc.p;
    L.x = p->x.p; L.b = p->bint.p; L.t = p->u.p;
}

void amg_fill_args(xfemm_linprob *fine, AmgHierarchy &h, AmgArgs &g)
{
    g.nlev = 1 + (int)h.coarse.size();
    const char *e;
    g.omega = (e = getenv("XFEMM_AMG_OMEGA")) ? atof(e) : 0.8;
    g.over = (e = getenv("XFEMM_AMG_OVER")) ? atof(e) : 2.6;
    g.coarse_sweeps = ((e = getenv("XFEMM_AMG_COARSE_SWEEPS")) ? atoi(e) : 9) | 1;
    g.stamps = nullptr;
    if (getenv("XFEMM_AMG_TRACE"))
    {
        if (!h.stamps.p && h.stamps.alloc(AMG_MAX_STAMPS) == cudaSuccess) h.stamps.zero(fine->stream);
        g.stamps = h.stamps.p;
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
    }
}

} // namespace xb
