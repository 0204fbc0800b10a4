// The multigrid V-cycle of amg.cuh applied to TWO right-hand sides at once (complex solver: the real and the
// imaginary part of the COCG residual go through the real hierarchy of |Re A| + |Im A|, clinprob.cu).
// Same phases and barriers as amg_vcycle; every matrix entry read from the stream (shared-memory ring on the finest
// level, SELL slices below) feeds two accumulators, so the second component costs no matrix traffic and no extra
// grid barrier.  Kept apart from amg.cuh so that the real solver's kernel is not touched.
#pragma once
#include "amg.cuh"

namespace xb {

// (AmgVec2 / AmgArgs2: vectors of the second component per level, amg.cuh)
void amg_fill_args2(xfemm_linprob *fine, AmgHierarchy &h, AmgArgs2 &g2);

struct AmgPre2 { double d, xa, xb; };

// staged_slices (stage.cuh) with two gathered vectors: epi(slice, row, acc_a, acc_b, pre(row))
template <class Pre, class Epi>
__device__ __forceinline__ void staged_slices2(StageSmem &S, const int *__restrict__ slice_ptr, const double *sell_val,
                                               const int *sell_col, int s0, int s1, const double *xa, const double *xb_,
                                               int *s_next, Pre pre, Epi epi)
{
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0)
    {
        for (int st = 0; st < STAGE_COUNT; st++) { mbar_init(&S.full[st], 1); S.remaining[st] = 0; S.current[st] = -1; }
        fence_mbar_init();
        fence_proxy_async();
        *s_next = s0;
        for (int k = 0; k < S.ntiles && k < STAGE_COUNT; k++) stage_issue(S, k, sell_val, sell_col);
    }
    __syncthreads();
    int k = 0;
    while (true)
    {
        int s = 0;
        if (lane == 0) s = atomicAdd(s_next, 1);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= s1) break;
        const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5;
        const auto pv = pre(s * 32 + lane);
        while (s >= S.tile[k].y) k++;
        const int st = k % STAGE_COUNT;
        while (*(volatile int *)&S.current[st] != k) {}
        mbar_wait(&S.full[st], (uint32_t)(k / STAGE_COUNT) & 1u);
        const int off = base - S.tile[k].z + lane;
        const double *v = S.val + (size_t)st * STAGE_TILE_ENT + off;
        const int *c = S.col + (size_t)st * STAGE_TILE_ENT + off;
        double acca = 0.0, accb = 0.0;
#pragma unroll 8
        for (int q = 0; q < w; q++)
        {
            const double a = v[q * 32];
            const int j = c[q * 32];
            acca = fma(a, xa[j], acca);
            accb = fma(a, xb_[j], accb);
        }
        epi(s, s * 32 + lane, acca, accb, pv);
        __syncwarp();
        if (lane == 0)
        {
            const int left = atomicSub(&S.remaining[st], 1) - 1;
            if (left == 0 && k + STAGE_COUNT < S.ntiles)
            {
                fence_proxy_async();
                stage_issue(S, k + STAGE_COUNT, sell_val, sell_col);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int st = 0; st < STAGE_COUNT; st++) mbar_inval(&S.full[st]);
    __syncthreads();
}

// epi(row, (A xa)[row], (A xb)[row]) for every own row of the block's slices of level L
template <class Epi>
__device__ __forceinline__ void amg_spmv_phase2(const AmgLev &L, int s0, int s1, int wk0, int wk1, double *xa,
                                                double *xb_, double *tmpa, double *tmpb, int *s_next, Epi epi,
                                                StageSmem *S = nullptr)
{
    const int lane = threadIdx.x & 31;
    if (wk1 > wk0)
    {
        wide_rows_block<double>(L.wide, L.wide_val, L.wide_acc, wk0, wk1, xa, tmpa);
        __syncthreads(); // the segment scratch is shared by the two passes
        wide_rows_block<double>(L.wide, L.wide_val, L.wide_acc, wk0, wk1, xb_, tmpb);
    }
    if (S && S->ntiles > 0)
    {
        staged_slices2(*S, L.slice_ptr, L.sell_val, L.sell_col, s0, s1, xa, xb_, s_next,
                       [&](int row) { return AmgPre2{L.diag[row], xa[row], xb_[row]}; },
                       [&](int s, int row, double acca, double accb, AmgPre2 p) {
                           if (wk1 > wk0 && ((L.wide.mask[s] >> lane) & 1u)) { acca += tmpa[row]; accb += tmpb[row]; }
                           epi(row, fma(p.d, p.xa, acca), fma(p.d, p.xb, accb));
                       });
        return;
    }
    if (threadIdx.x == 0) *s_next = s0;
    __syncthreads();
    while (true)
    {
        int s = 0;
        if (lane == 0) s = atomicAdd(s_next, 1);
        s = __shfl_sync(0xffffffffu, s, 0);
        if (s >= s1) break;
        const int row = s * 32 + lane;
        const int base = L.slice_ptr[s], w = (L.slice_ptr[s + 1] - base) >> 5;
        const double *v = L.sell_val + base + lane;
        const int *c = L.sell_col + base + lane;
        double acca = 0.0, accb = 0.0;
#pragma unroll 8
        for (int k = 0; k < w; k++)
        {
            const double a = __ldcs(v + k * 32);
            const int j = __ldcs(c + k * 32);
            acca = fma(a, xa[j], acca);
            accb = fma(a, xb_[j], accb);
        }
        if (wk1 > wk0 && ((L.wide.mask[s] >> lane) & 1u)) { acca += tmpa[row]; accb += tmpb[row]; }
        const double d = L.diag[row];
        epi(row, fma(d, xa[row], acca), fma(d, xb_[row], accb));
    }
    __syncthreads();
}

// (t, t2) = M^-1 (b, b2) on level 0.  The caller has stored the right-hand sides in lev[0].b / lev2[0].b and the
// pre-smoothed iterates omega D^-1 b in lev[0].x / lev2[0].x and synchronised the grid.
__device__ __forceinline__ void amg_vcycle2(const AmgArgs &g, const AmgArgs &sg, const AmgArgs2 &g2, GridBar &grid, int *s_next,
                                            int *s_wk, StageSmem *S0)
{
    const int G = gridDim.x, B = blockIdx.x;
    const int last = g.nlev - 1;
    const int lc = g.team_from;
    const int tid_g = B * blockDim.x + threadIdx.x, nthr_g = G * blockDim.x;
    const double w = g.omega, over = g.over;
    if (last == 0)
    {
        // degenerate one-level hierarchy: M^-1 = damped Jacobi
        const AmgLev &L = g.lev[0];
        const int r0 = amg_chunk(L.nslices, B, G) * 32, r1 = amg_chunk(L.nslices, B + 1, G) * 32;
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) { L.t[i] = L.x[i]; g2.lev[0].t[i] = g2.lev[0].x[i]; }
        return;
    }
    for (int l = 0; l < lc; l++)
    {
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        const AmgVec2 &L2 = g2.lev[l];
        const int s0 = amg_chunk(L.nslices, B, G), s1 = amg_chunk(L.nslices, B + 1, G);
        const AmgV<2> lv = amg_vecs<2>(g, &g2, l), cv = amg_vecs<2>(g, &g2, l + 1);
        if (l == 0)
            amg_spmv_phase2(L, s0, s1, s_wk[0], s_wk[1], L.x, L2.x, L.t, L2.t, s_next,
                            [&](int row, double ax, double bx) { L.t[row] = L.b[row] - ax; L2.t[row] = L2.b[row] - bx; }, S0);
        else
            amg_spmv_tasks<2, false>(sg.lev[l], lv, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], AMG_EPI_RESIDUAL, w, over, xb_dyn_smem + g.task_smem_offset,
                                         amg_table(g, s_wk, l));
        grid.sync();
        amg_restrict_level<2>(sg.lev[l], sg.lev[l + 1], lv, cv, tid_g, nthr_g, w);
        grid.sync();
    }
    if (B < g.team_blocks) amg_team_cycle<2>(sg, &g2, s_wk);
    grid.sync();
    for (int l = lc - 1; l >= 0; l--)
    {
        const AmgLev &L = g.lev[l];
        const AmgVec2 &L2 = g2.lev[l];
        const int s0 = amg_chunk(L.nslices, B, G), s1 = amg_chunk(L.nslices, B + 1, G);
        const AmgV<2> lv = amg_vecs<2>(g, &g2, l), cv = amg_vecs<2>(g, &g2, l + 1);
        amg_prolong_rows<2>(sg.lev[l], lv, cv, s0 * 32, s1 * 32, threadIdx.x, blockDim.x, over);
        grid.sync();
        if (l == 0)
            amg_spmv_phase2(L, s0, s1, s_wk[0], s_wk[1], L.x, L2.x, L.t, L2.t, s_next,
                            [&](int row, double ax, double bx) {
                                const double wd = w * L.dinv[row];
                                L.t[row] = fma(wd, L.b[row] - ax, L.x[row]);
                                L2.t[row] = fma(wd, L2.b[row] - bx, L2.x[row]);
                            },
                            S0);
        else
        {
            amg_spmv_tasks<2, false>(sg.lev[l], lv, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], AMG_EPI_POST, w, over, xb_dyn_smem + g.task_smem_offset,
                                     amg_table(g, s_wk, l));
            grid.sync();
        }
    }
}

} // namespace xb
