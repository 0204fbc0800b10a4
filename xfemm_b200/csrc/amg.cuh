// Aggregation multigrid preconditioner for the real symmetric solver (XFEMM_PRECOND_AMG).
//
// The reference preconditions PCGSolve with serial SSOR (spars.cpp:197-236); on a GPU that
// recurrence does not parallelise, Jacobi costs ~3.5x its iterations and multicolour SSOR pays two
// grid barriers per colour.  A V-cycle over a hierarchy of plain (unsmoothed, piecewise-constant)
// aggregates needs 20-40x fewer CG iterations than Jacobi on these first-order triangle meshes
// and is made of the same building blocks as the solver itself: SELL-32 SpMV passes, row-wise
// vector passes and grid barriers, all inside the ONE persistent PCG kernel.
//
//   hierarchy (host, once per sparsity pattern): greedy distance-1 aggregation on the pattern,
//       coarse pattern = aggregate graph, per coarse CSR slot the list of fine slots that sum into
//       it (Galerkin product P^T A P with P = aggregate indicator → a gather-sum, deterministic).
//   values (device, every solve): level l+1 values from level l by that gather-sum, then the usual
//       SELL / diagonal refill of each level.
//   cycle: damped-Jacobi pre- and post-smoothing (same weight → symmetric), coarse correction
//       scaled by an over-correction factor (plain aggregation under-estimates the correction),
//       coarsest level (<= 64 rows or stalled coarsening) by a fixed number of Jacobi sweeps in one
//       block.  M^-1 is a fixed symmetric positive definite operator, so CG applies.
#pragma once
#include "linprob.cuh"
#include "stage.cuh"
#include <cooperative_groups.h>
#include <memory>
#include <thread>

namespace xb {

constexpr int AMG_MAX_LEVELS = 12;

// device view of one level, internal (SELL) row order
struct AmgLev
{
    int n_pad = 0, nslices = 0;
    const int *slice_ptr = nullptr, *sell_col = nullptr;
    const double *sell_val = nullptr, *diag = nullptr, *dinv = nullptr;
    WideView wide;
    const double *wide_val = nullptr;
    double *wide_acc = nullptr;
    double *x = nullptr, *b = nullptr, *t = nullptr; // iterate, right-hand side, result / scratch
    const int *agg = nullptr;                          // row -> row of the next level (-1: not represented)
    const int *mptr = nullptr, *mem = nullptr;         // row of the NEXT level -> member rows of this level
};

struct AmgArgs
{
    int nlev = 0;
    int coarse_sweeps = 9;  // odd: the result of the coarsest level lands in its t buffer
    double omega = 0.8, over = 2.6;
    int local_from = 1;     // first level handled by one block alone (small levels), <= nlev-1
    const double *coarse_inv = nullptr; // dense inverse of the coarsest level (n_pad x n_pad, internal order) or null
    unsigned long long *stamps = nullptr; // optional (XFEMM_AMG_TRACE): %globaltimer at every phase boundary of a cycle
    AmgLev lev[AMG_MAX_LEVELS];
};

constexpr int AMG_MAX_STAMPS = 8 * AMG_MAX_LEVELS;
__device__ __forceinline__ void amg_stamp(const AmgArgs &g, int &k, int limit = AMG_MAX_STAMPS - 8)
{
    if (g.stamps && blockIdx.x == 0 && threadIdx.x == 0 && k < limit)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g.stamps[k++] = t;
    }
}

// host side: one coarse level = a complete linear problem (pattern, values, SELL layout, vectors)
struct AmgHierarchy
{
    struct Transfer
    {
        DevBuf<int> gptr, gsrc; // coarse CSR slot -> contributing fine CSR slots
        DevBuf<int> agg;        // fine internal row -> coarse internal row
        DevBuf<int> mptr, mem;  // coarse internal row -> fine internal rows
        long long nnz_coarse = 0;
    };
    std::vector<xfemm_linprob *> coarse; // levels 1..
    std::vector<std::unique_ptr<Transfer>> xfer; // xfer[l]: level l -> l+1
    std::vector<int> level_n, level_nnz;
    DevBuf<unsigned long long> stamps; // XFEMM_AMG_TRACE
    // The levels from the third on are latency-bound (a few thousand rows, dependent loads); the
    // streams of the finest level would evict them from L2 every cycle.  Their buffers are carved
    // from this one range, which the solver stream pins in L2 (access policy window, persisting).
    DevArena arena;
    DevBuf<double> dense_work, dense_inv; // coarsest level: [A | I] scratch and A^-1 (n_pad x n_pad), if small enough
    ~AmgHierarchy();
};
void amg_print_trace(AmgHierarchy &h, cudaStream_t st); // phase timings of the last cycle to stderr

// fine_layout: optional host thread that is still building the solver layout of `fine`; the aggregation and the
// coarse pattern only need the sparsity pattern, the thread is joined before the transfer maps read the layout
int amg_build(xfemm_linprob *fine, std::unique_ptr<AmgHierarchy> &out, std::thread *fine_layout = nullptr);
int amg_refresh(xfemm_linprob *fine, AmgHierarchy &h);
void amg_fill_args(xfemm_linprob *fine, AmgHierarchy &h, AmgArgs &g);

// --------------------------------------------------------------------------------------------
// device: grid-wide V-cycle.  Every block owns a contiguous chunk of slices of every level.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double amg_sell_row(const AmgLev &L, int s, int lane, const double *x)
{
    const int base = L.slice_ptr[s], w = (L.slice_ptr[s + 1] - base) >> 5;
    double acc = 0.0;
    const double *v = L.sell_val + base + lane;
    const int *c = L.sell_col + base + lane;
#pragma unroll 8
    for (int k = 0; k < w; k++) acc = fma(__ldcs(v + k * 32), x[__ldcs(c + k * 32)], acc);
    return acc;
}

// y-phase over the block's slices of level L: calls epi(row, (A x)[row]) for every own row
template <class Epi>
__device__ __forceinline__ void amg_spmv_phase(const AmgLev &L, int s0, int s1, int wk0, int wk1, const double *x,
                                               double *wide_tmp, int *s_next, Epi epi, StageSmem *S = nullptr,
                                               const double *pf0 = nullptr, const double *pf1 = nullptr)
{
    const int lane = threadIdx.x & 31;
    wide_rows_block<double>(L.wide, L.wide_val, L.wide_acc, wk0, wk1, x, wide_tmp);
    if (S && S->ntiles > 0)
    {
        // finest level: the matrix stream comes through the shared-memory ring (stage.cuh)
        // pf0/pf1: row-indexed arrays the caller's epilogue reads (prefetched while the warp waits for its tile)
        staged_slices(*S, L.slice_ptr, L.sell_val, L.sell_col, s0, s1, x, s_next,
                      [&](int row) {
                          if (pf0) prefetch_l1(pf0 + row);
                          if (pf1) prefetch_l1(pf1 + row);
                          return make_double2(L.diag[row], x[row]);
                      },
                      [&](int s, int row, double acc, double2 dx) {
                          if (wk1 > wk0 && ((L.wide.mask[s] >> lane) & 1u)) acc += wide_tmp[row];
                          acc = fma(dx.x, dx.y, acc);
                          epi(row, acc);
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
        double acc = amg_sell_row(L, s, lane, x);
        if (wk1 > wk0 && ((L.wide.mask[s] >> lane) & 1u)) acc += wide_tmp[row];
        acc = fma(L.diag[row], x[row], acc);
        epi(row, acc);
    }
    __syncthreads(); // s_next is reused by the next phase
}

// Levels first..last handled by ONE block with __syncthreads only.  In: lev[first].b and .x
// (pre-smoothed).  Out: lev[first].t.
__device__ __forceinline__ void amg_block_cycle(const AmgArgs &g, int first)
{
    const int last = g.nlev - 1;
    const int lane = threadIdx.x & 31, warp_b = threadIdx.x >> 5, nwarps_b = blockDim.x >> 5;
    const double w = g.omega, over = g.over;
    auto spmv = [&](const AmgLev &L, const double *src, double *dst, auto epi) {
        wide_rows_block<double>(L.wide, L.wide_val, L.wide_acc, 0, L.wide.n_rows, src, dst);
        __syncthreads();
        for (int s = warp_b; s < L.nslices; s += nwarps_b)
        {
            const int row = s * 32 + lane;
            double acc = amg_sell_row(L, s, lane, src);
            if (L.wide.n_rows && ((L.wide.mask[s] >> lane) & 1u)) acc += dst[row];
            acc = fma(L.diag[row], src[row], acc);
            epi(row, acc);
        }
        __syncthreads();
    };
    for (int l = first; l < last; l++)
    {
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        spmv(L, L.x, L.t, [&](int row, double ax) { L.t[row] = L.b[row] - ax; });
        for (int I = threadIdx.x; I < C.n_pad; I += blockDim.x)
        {
            double s = 0.0;
            for (int k = L.mptr[I]; k < L.mptr[I + 1]; k++) s += L.t[L.mem[k]];
            C.b[I] = s;
            C.x[I] = w * C.dinv[I] * s;
        }
        __syncthreads();
    }
    {
        const AmgLev &L = g.lev[last];
        if (g.coarse_inv)
        {
            // coarsest level: exact solve with the dense inverse prepared by amg_refresh (t = A^-1 b)
            const int n = L.n_pad;
            for (int r = warp_b; r < n; r += nwarps_b)
            {
                double acc = 0.0;
                for (int c = lane; c < n; c += 32) acc = fma(g.coarse_inv[(size_t)r * n + c], L.b[c], acc);
                acc = warp_sum(acc);
                if (lane == 0) L.t[r] = acc;
            }
            __syncthreads();
        }
        else
        {
            // damped-Jacobi sweeps x -> t -> x ... (odd count: ends in t)
            double *src = L.x, *dst = L.t;
            for (int it = 0; it < g.coarse_sweeps; it++)
            {
                double *d = dst;
                const double *sv = src;
                spmv(L, sv, d, [&](int row, double ax) { d[row] = fma(w * L.dinv[row], L.b[row] - ax, sv[row]); });
                double *tmp = src; src = dst; dst = tmp;
            }
        }
    }
    for (int l = last - 1; l >= first; l--)
    {
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        for (int i = threadIdx.x; i < L.n_pad; i += blockDim.x)
        {
            const int I = L.agg[i];
            if (I >= 0) L.x[i] = fma(over, C.t[I], L.x[i]);
        }
        __syncthreads();
        spmv(L, L.x, L.t, [&](int row, double ax) { L.t[row] = fma(w * L.dinv[row], L.b[row] - ax, L.x[row]); });
    }
}

// z = M^-1 r.  Level 0: b = r (input), t = z (output), x = scratch.  Returns this thread's partial of r.z.
// x0_ready: the caller already stored the pre-smoothed x = omega D^-1 r of level 0 (fused into its
// residual update) and synchronised the grid.
__device__ __forceinline__ double amg_vcycle(const AmgArgs &g, cooperative_groups::grid_group &grid, int *s_next,
                                             const int *s_wk /* [2*AMG_MAX_LEVELS] in shared memory */,
                                             StageSmem *S0 /* ring of the finest level, or null */, bool x0_ready = false)
{
    const int G = gridDim.x, B = blockIdx.x;
    const int last = g.nlev - 1;
    const int lc = g.local_from; // levels lc..last belong to block 0 alone (1 <= lc <= last unless nlev == 1)
    const double w = g.omega, over = g.over;
    double dot = 0.0;
    int sk = 0;
    amg_stamp(g, sk);
    // ---- down: pre-smooth from a zero guess, residual, restrict
    if (!x0_ready || last == 0)
    {
        const AmgLev &L = g.lev[0];
        const int r0 = (int)((long long)L.nslices * B / G) * 32, r1 = (int)((long long)L.nslices * (B + 1) / G) * 32;
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) L.x[i] = w * L.dinv[i] * L.b[i];
        if (last == 0)
        {
            // degenerate one-level hierarchy: M^-1 = damped Jacobi
            for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) { L.t[i] = L.x[i]; dot = fma(L.b[i], L.x[i], dot); }
            return dot;
        }
        grid.sync();
        amg_stamp(g, sk);
    }
    for (int l = 0; l < lc; l++)
    {
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        const int s0 = (int)((long long)L.nslices * B / G), s1 = (int)((long long)L.nslices * (B + 1) / G);
        // t = b - A x   (t doubles as the scratch the wide rows leave their partial sums in)
        amg_spmv_phase(L, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], L.x, L.t, s_next,
                       [&](int row, double ax) { L.t[row] = L.b[row] - ax; }, l == 0 ? S0 : nullptr, L.b);
        grid.sync();
        amg_stamp(g, sk);
        // coarse right-hand side = sum over the aggregate ; coarse pre-smooth fused
        const int c0 = (int)((long long)C.nslices * B / G) * 32, c1 = (int)((long long)C.nslices * (B + 1) / G) * 32;
        for (int I = c0 + threadIdx.x; I < c1; I += blockDim.x)
        {
            double s = 0.0;
            for (int k = L.mptr[I]; k < L.mptr[I + 1]; k++) s += L.t[L.mem[k]];
            C.b[I] = s;
            C.x[I] = w * C.dinv[I] * s;
        }
        grid.sync();
        amg_stamp(g, sk);
    }
    // ---- the small levels: one block, no grid barriers; result in lev[lc].t
    if (B == 0) amg_block_cycle(g, lc);
    grid.sync();
    amg_stamp(g, sk);
    // ---- up: add the (over-)corrected coarse result (a row-wise gather: a scatter from the coarse rows
    // was measured 3x slower, its read-modify-writes serialise), then post-smooth into t
    for (int l = lc - 1; l >= 0; l--)
    {
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        const int s0 = (int)((long long)L.nslices * B / G), s1 = (int)((long long)L.nslices * (B + 1) / G);
        for (int i = s0 * 32 + threadIdx.x; i < s1 * 32; i += blockDim.x)
        {
            const int I = L.agg[i];
            if (I >= 0) L.x[i] = fma(over, C.t[I], L.x[i]);
        }
        grid.sync();
        amg_stamp(g, sk);
        if (l > 0)
        {
            amg_spmv_phase(L, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], L.x, L.t, s_next,
                           [&](int row, double ax) { L.t[row] = fma(w * L.dinv[row], L.b[row] - ax, L.x[row]); });
            grid.sync();
            amg_stamp(g, sk);
        }
        else
        {
            amg_spmv_phase(L, s0, s1, s_wk[0], s_wk[1], L.x, L.t, s_next, [&](int row, double ax) {
                const double br = L.b[row], z = fma(w * L.dinv[row], br - ax, L.x[row]);
                L.t[row] = z;
                dot = fma(br, z, dot);
            }, S0, L.b, L.dinv);
        }
    }
    amg_stamp(g, sk);
    return dot;
}

} // namespace xb
