// Aggregation multigrid preconditioner for the real symmetric solver (XFEMM_PRECOND_AMG).
//
// The reference preconditions PCGSolve with serial SSOR (spars.cpp:197-236); on a GPU that
// recurrence does not parallelise, Jacobi costs ~3.5x its iterations and multicolour SSOR pays two
// grid barriers per colour.  A V-cycle over a hierarchy of plain (unsmoothed, piecewise-constant)
// aggregates needs 20-40x fewer CG iterations than Jacobi on these first-order triangle meshes
// and is made of the same building blocks as the solver itself: SELL-32 SpMV passes, row-wise
// vector passes and barriers, all inside the ONE persistent PCG kernel.
//
//   hierarchy (host, once per sparsity pattern): greedy distance-1 aggregation on the pattern,
//       coarse pattern = aggregate graph, per coarse CSR slot the list of fine slots that sum into
//       it (Galerkin product P^T A P with P = aggregate indicator → a gather-sum, deterministic).
//   values (device, every solve): level l+1 values from level l by that gather-sum, then the usual
//       SELL / diagonal refill of each level.
//   cycle: damped-Jacobi pre- and post-smoothing (same weight → symmetric), coarse correction
//       scaled by an over-correction factor (plain aggregation under-estimates the correction),
//       coarsest level (<= 128 rows) by its dense inverse, else by Jacobi sweeps.  M^-1 is a fixed
//       symmetric positive definite operator, so CG applies.
//
// Who runs which level.  The big levels (0 .. team_from-1) are cycled by the whole grid: every block
// owns a contiguous chunk of the slices of every level, phases are separated by grid barriers.  The
// small levels (team_from .. last; a few ten thousand rows and less — they hold ~1 % of the data but
// every phase on them is a chain of two or three dependent L2 accesses plus a barrier) belong to a
// TEAM of the first `team_blocks` blocks, which synchronise among themselves with a counter barrier
// (team_sync: ~1/5 of the arrivals of a grid barrier) while the other blocks wait at the one grid
// barrier behind the team's sub-cycle.  On team levels the prolongation is fused into the
// post-smoothing pass (x + over*P*t_c is formed on the fly from a coarse-column copy of the SELL
// columns), which removes one pass and one barrier per level.  All small-level loops load in
// predicated batches of eight so that a slice costs two round trips, not one per entry.
#pragma once
#include "linprob.cuh"
#include "stage.cuh"
#include <cooperative_groups.h>
#include <memory>
#include <thread>
#include <type_traits>

namespace xb {

constexpr int AMG_MAX_LEVELS = 12;
// per-block scratch of the kernels that run the cycle (shared memory): wide-row ranges of the block's chunk
// of every level (2 ints per level) + the team barrier's epoch
constexpr int AMG_SWK_EPOCH = 2 * AMG_MAX_LEVELS;     // [+0] team barrier epoch, [+1] grid barrier epoch
constexpr int AMG_SWK_TAB = 2 * AMG_MAX_LEVELS + 2;   // per level: offset of the block's resident task table (ints), -1: none
constexpr int AMG_SWK_SIZE = 3 * AMG_MAX_LEVELS + 2;
// dynamic shared memory of the kernels that run the cycle: [matrix-stream ring (stage.cuh)] [task buffer] [task tables]
constexpr int AMG_TASK_MAXS = 480;                 // slices per group of amg_spmv_tasks
constexpr int AMG_TASK_PART_BYTES = 32 * 1024;     // partial sums: MAXT tasks x 32 rows x NC doubles
constexpr int AMG_TASK_HDR_INTS = 4 + 2 * (AMG_TASK_MAXS + 1) + 128;
constexpr int AMG_TASK_HDR_BYTES = ((AMG_TASK_HDR_INTS * 4 + 127) / 128) * 128;
constexpr int AMG_TABLE_INTS = 2048;               // resident task tables of the block's chunks of the coarse levels
constexpr int AMG_TASK_SMEM_BYTES = AMG_TASK_HDR_BYTES + AMG_TASK_PART_BYTES + AMG_TABLE_INTS * 4;
extern __shared__ __align__(128) unsigned char xb_dyn_smem[]; // the kernels' dynamic shared memory

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
    const int *sell_cagg = nullptr;                    // team levels without wide rows: agg[sell_col[slot]] (fused prolongation)
};

struct AmgArgs
{
    int nlev = 0;
    int coarse_sweeps = 9;  // odd: the result of the coarsest level lands in its t buffer
    double omega = 0.8, over = 2.6;
    int team_from = 1;      // first level handled by the team of blocks (small levels), 1 <= team_from <= nlev-1
    int team_blocks = 1;    // blocks 0..team_blocks-1 form the team
    // Restriction by scatter (one-component cycles with a dense coarsest solve): the residual pass of level l adds
    // w * dinv_c[I] * r_i straight into the coarse iterate x_c[I] (red.add.f64 in L2) instead of storing r and running
    // a restriction pass behind a barrier; the coarse right-hand side is recovered row-locally as x_c * diag_c / w.
    // One phase and one barrier less per level; the order of the additions is not fixed, so deterministic runs
    // (XFEMM_DETERMINISTIC, option deterministic_assembly) keep the ordered restriction pass.
    int scatter_restrict = 0;
    int solo_from = 1 << 20; // levels >= solo_from (a few hundred rows) are cycled by block 0 ALONE with __syncthreads: on them
                            // a barrier between blocks costs more than the whole level
    unsigned *team_ctr = nullptr; // arrival counter of the team barrier (device memory, zeroed before the launch)
    unsigned *grid_ctr = nullptr; // arrival counter of the grid barrier (GridBar), or null: cooperative-groups barrier
    const double *coarse_inv = nullptr; // dense inverse of the coarsest level (n_pad x n_pad, internal order) or null
    int task_smem_offset = 0; // where the task buffer of amg_spmv_tasks starts in the kernel's dynamic shared memory
    int trace_detail = 0;   // XFEMM_AMG_TRACE=2: a second stamp in front of every team barrier
    unsigned long long *stamps = nullptr; // optional (XFEMM_AMG_TRACE): %globaltimer at every phase boundary of a cycle
    AmgLev lev[AMG_MAX_LEVELS];
};

// vectors of a second right-hand side per level (complex solver: real and imaginary part of the residual go
// through ONE cycle, amg2.cuh); the first component uses AmgLev::x/b/t
struct AmgVec2 { double *x = nullptr, *b = nullptr, *t = nullptr; };
struct AmgArgs2 { AmgVec2 lev[AMG_MAX_LEVELS]; };

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

// XFEMM_AMG_TRACE=2: every block records when it finished phase `ph` (0 CG SpMV, 1 level-0 residual, 2 level-0
// post-smoothing) of the last iteration, and when the phase started (block 0 only, slot 160*ph+159)
constexpr int AMG_BLOCK_STAMPS = 3 * 160;
__device__ __forceinline__ void amg_block_stamp(const AmgArgs &g, int ph, bool start = false)
{
    if (g.stamps && g.trace_detail && threadIdx.x == 0 && blockIdx.x < 159)
    {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (!start) g.stamps[AMG_MAX_STAMPS + 16 + 160 * ph + blockIdx.x] = t;
        else if (blockIdx.x == 0) g.stamps[AMG_MAX_STAMPS + 16 + 160 * ph + 159] = t;
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
        DevBuf<int> cagg;       // fine SELL slot -> coarse internal row of its column (team levels without wide rows)
        long long nnz_coarse = 0;
    };
    std::vector<xfemm_linprob *> coarse; // levels 1..
    std::vector<std::unique_ptr<Transfer>> xfer; // xfer[l]: level l -> l+1
    std::vector<int> level_n, level_nnz;
    int team_rows = 65536;             // levels with at most this many rows belong to the team (XFEMM_AMG_TEAM_ROWS)
    DevBuf<unsigned long long> stamps; // XFEMM_AMG_TRACE
    DevBuf<unsigned> team_ctr;         // barrier counters: [0] team, [32] grid (one 128 B line each)
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
// grid = blocks of the cooperative launch the arguments are for (the team cannot be larger)
void amg_fill_args(xfemm_linprob *fine, AmgHierarchy &h, AmgArgs &g, int grid = 1 << 30);

// --------------------------------------------------------------------------------------------
// device: per-kernel set-up, the team barrier
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int amg_chunk(int nslices, int b, int nb) { return (int)((long long)nslices * b / nb); }

// Warp-collective: task table of the group of slices that starts at `cur` (see amg_spmv_tasks): as many slices as fit
// `maxs` slices and `maxt` tasks of `batch` columns.  first_task / slice_base get one entry per slice plus the end,
// task_info one (slice of the group << 16 | batch number) per task.  Returns the slices covered; *ntasks the tasks.
__device__ __forceinline__ int amg_task_table(const int *slice_ptr, int cur, int s1, int maxs, int maxt, int batch, int *first_task,
                                              int *slice_base, int *task_info, int *ntasks)
{
    const int lane = threadIdx.x & 31;
    int nt_total = 0, ns = 0;
    while (cur + ns < s1 && ns < maxs)
    {
        const int s = cur + ns + lane;
        int nt = 0, a = 0, b = 0;
        if (s < s1) { a = slice_ptr[s]; b = slice_ptr[s + 1]; nt = (((b - a) >> 5) + batch - 1) / batch; }
        int inc = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        const bool fits = s < s1 && nt_total + inc <= maxt && ns + lane < maxs;
        const int cnt = __popc(__ballot_sync(0xffffffffu, fits)); // a prefix of the lanes: inc is non-decreasing
        if (fits)
        {
            const int t0 = nt_total + inc - nt;
            first_task[ns + lane] = t0;
            slice_base[ns + lane] = a;
            for (int q = 0; q < nt; q++) task_info[t0 + q] = ((ns + lane) << 16) | q;
            if (lane == cnt - 1) slice_base[ns + lane + 1] = b;
        }
        const int add = __shfl_sync(0xffffffffu, inc, cnt > 0 ? cnt - 1 : 0);
        if (cnt > 0) nt_total += add;
        ns += cnt;
        if (cnt < 32) break;
    }
    if (lane == 0) first_task[ns] = nt_total;
    *ntasks = nt_total;
    return ns;
}

// scatter_restrict: the coarse iterates collect additions and have to start from zero — every block clears its share of
// all of them (a few hundred thousand doubles).  Called at kernel start and when a cycle no longer needs them.
__device__ __forceinline__ void amg_zero_coarse_iterates(const AmgArgs &g)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int l = 1; l < g.nlev; l++)
    {
        double *x = g.lev[l].x;
        const int n = g.lev[l].n_pad;
        for (int i = tid; i < n; i += nthr) x[i] = 0.0;
    }
}

// blocks that share level l of the team's sub-cycle: the team, or block 0 alone on the smallest levels
__device__ __forceinline__ int amg_team_width(const AmgArgs &g, int l) { return l >= g.solo_from ? 1 : g.team_blocks; }

// Once per kernel, whole block, before the first cycle.  `sg` = the block's copy of the arguments in shared memory
// (raw storage of sizeof(AmgArgs)): the coarse-level code is out of line and would otherwise read the kernel
// parameters through generic loads — a round trip to L2 each after every barrier; shared memory is a few cycles.
//   - wide-row ranges of this block's chunk of every level (grid partition for the big levels, team partition for
//     the team levels), the barrier epochs
//   - resident task tables (amg_spmv_tasks) of the block's chunk of every coarse level that fits one group
template <int NC>
__device__ __forceinline__ const AmgArgs &amg_kernel_init(const AmgArgs &g, unsigned char *sg_raw, int *s_wk)
{
    const int G = gridDim.x, B = blockIdx.x;
    {
        const unsigned *src = (const unsigned *)&g;
        unsigned *dst = (unsigned *)sg_raw;
        for (int i = threadIdx.x; i < (int)(sizeof(AmgArgs) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x < g.nlev)
    {
        const int l = threadIdx.x;
        const AmgLev &L = g.lev[l];
        int k0 = 0, k1 = 0;
        if (L.wide.n_rows)
        {
            if (l < g.team_from) wide_range(L.wide, amg_chunk(L.nslices, B, G) * 32, amg_chunk(L.nslices, B + 1, G) * 32, k0, k1);
            else if (B < amg_team_width(g, l))
                wide_range(L.wide, amg_chunk(L.nslices, B, amg_team_width(g, l)) * 32, amg_chunk(L.nslices, B + 1, amg_team_width(g, l)) * 32, k0, k1);
        }
        s_wk[2 * l] = k0; s_wk[2 * l + 1] = k1;
        s_wk[AMG_SWK_TAB + l] = -1;
    }
    if (threadIdx.x == 0) { s_wk[AMG_SWK_EPOCH] = 0; s_wk[AMG_SWK_EPOCH + 1] = 0; }
    if (NC == 1 && g.scatter_restrict) amg_zero_coarse_iterates(g); // (the callers' first barrier precedes the first cycle)
    __syncthreads();
    if (threadIdx.x < 32)
    {
        constexpr int BATCH = (NC == 1) ? 8 : 4;
        constexpr int MAXT = AMG_TASK_PART_BYTES / (32 * 8 * NC);
        int *tab = (int *)(xb_dyn_smem + g.task_smem_offset + AMG_TASK_HDR_BYTES + AMG_TASK_PART_BYTES);
        int used = 0;
        for (int l = 1; l < g.nlev; l++)
        {
            const AmgLev &L = g.lev[l];
            int s0 = 0, s1 = 0;
            if (l < g.team_from) { s0 = amg_chunk(L.nslices, B, G); s1 = amg_chunk(L.nslices, B + 1, G); }
            else if (B < amg_team_width(g, l)) { s0 = amg_chunk(L.nslices, B, amg_team_width(g, l)); s1 = amg_chunk(L.nslices, B + 1, amg_team_width(g, l)); }
            const int nsl = s1 - s0, need = 2 + 2 * (nsl + 1) + MAXT;
            if (nsl <= 0 || nsl > AMG_TASK_MAXS || used + need > AMG_TABLE_INTS) continue;
            int *t = tab + used, ntasks = 0;
            const int ns = amg_task_table(L.slice_ptr, s0, s1, nsl, MAXT, BATCH, t + 2, t + 2 + (nsl + 1), t + 2 + 2 * (nsl + 1), &ntasks);
            if (ns == nsl) // the whole chunk is one group: keep the table
            {
                if (threadIdx.x == 0) { t[0] = ns; t[1] = ntasks; s_wk[AMG_SWK_TAB + l] = used; }
                used += need;
            }
        }
    }
    __syncthreads();
    return *reinterpret_cast<const AmgArgs *>(sg_raw);
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Counter barrier among `nblocks` blocks that are all resident (cooperative launch).  Monotone arrival counter:
// barrier number k is complete when the counter has reached k * nblocks; *s_epoch (shared memory) holds k * nblocks.
// Release/acquire at GPU scope on the counter itself: a red.release + ld.acquire pair costs a fraction of the
// sequentially-consistent fences (MEMBAR.SC) that __threadfence() / cooperative-groups' grid.sync() compile to.
__device__ __forceinline__ void counter_barrier(unsigned *ctr, int *s_epoch, int nblocks)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned target = (unsigned)(*s_epoch += nblocks);
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        while ((int)(ld_acquire_gpu(ctr) - target) < 0) {}
    }
    __syncthreads();
}

// Barrier among the team's blocks
__device__ __forceinline__ void team_sync(const AmgArgs &g, int *s_wk) { counter_barrier(g.team_ctr, s_wk + AMG_SWK_EPOCH, g.team_blocks); }

// Grid barrier with the interface of cooperative_groups::grid_group::sync(): the counter barrier when the kernel
// was given a counter, cooperative groups otherwise
struct GridBar
{
    cooperative_groups::grid_group grid;
    unsigned *ctr;
    int *s_epoch;
    __device__ __forceinline__ void sync()
    {
        if (ctr) counter_barrier(ctr, s_epoch, gridDim.x);
        else grid.sync();
    }
};

// --------------------------------------------------------------------------------------------
// device: building blocks shared by the one- and the two-component cycle
// --------------------------------------------------------------------------------------------
template <int NC>
struct AmgV
{
    double *x[NC], *b[NC], *t[NC];
};
template <int NC>
__device__ __forceinline__ AmgV<NC> amg_vecs(const AmgArgs &g, const AmgArgs2 *g2, int l)
{
    AmgV<NC> v;
    v.x[0] = g.lev[l].x; v.b[0] = g.lev[l].b; v.t[0] = g.lev[l].t;
    if (NC > 1) { v.x[NC - 1] = g2->lev[l].x; v.b[NC - 1] = g2->lev[l].b; v.t[NC - 1] = g2->lev[l].t; }
    return v;
}

// coarse right-hand side of row I of level C = sum of the fine residuals of its members (ascending), coarse
// pre-smoothing from a zero guess fused.  Member loads in predicated batches of eight.
template <int NC>
__device__ __forceinline__ void amg_restrict_row(const AmgLev &L, const AmgLev &C, const AmgV<NC> &lv, const AmgV<NC> &cv, int I,
                                                 double w)
{
    constexpr int MB = (NC == 1) ? 16 : 8; // members per step: indices first, then all their residuals — two round trips
    const int k0 = L.mptr[I], k1 = L.mptr[I + 1];
    const double wd = w * C.dinv[I];
    double s[NC];
#pragma unroll
    for (int m = 0; m < NC; m++) s[m] = 0.0;
    for (int k = k0; k < k1; k += MB)
    {
        int mi[MB];
#pragma unroll
        for (int q = 0; q < MB; q++) mi[q] = (k + q < k1) ? L.mem[k + q] : -1;
        double tv[NC][MB];
#pragma unroll
        for (int q = 0; q < MB; q++)
#pragma unroll
            for (int m = 0; m < NC; m++) tv[m][q] = mi[q] >= 0 ? lv.t[m][mi[q]] : 0.0;
#pragma unroll
        for (int q = 0; q < MB; q++)
#pragma unroll
            for (int m = 0; m < NC; m++) s[m] += tv[m][q];
    }
#pragma unroll
    for (int m = 0; m < NC; m++) { cv.b[m][I] = s[m]; cv.x[m][I] = wd * s[m]; }
}

// the restriction pass of a whole level by `nthr` threads (own register allocation: out of line)
template <int NC>
__device__ __noinline__ void amg_restrict_level(const AmgLev &L, const AmgLev &C, AmgV<NC> lv, AmgV<NC> cv, int tid, int nthr, double w)
{
    for (int I = tid; I < C.n_pad; I += nthr) amg_restrict_row<NC>(L, C, lv, cv, I, w);
}

// x[i] += over * ct[agg[i]] for the rows [i0,i1) by `nthr` threads (this thread = `tid`), four rows per thread and
// step with the loads ahead of the stores
template <int NC>
__device__ __noinline__ void amg_prolong_rows(const AmgLev &L, AmgV<NC> lv, AmgV<NC> cv, int i0, int i1, int tid, int nthr, double over)
{
    for (int i = i0 + tid; i < i1; i += 4 * nthr)
    {
        int I[4];
        double xv[NC][4];
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const int r = i + q * nthr;
            I[q] = r < i1 ? L.agg[r] : -1;
#pragma unroll
            for (int m = 0; m < NC; m++) xv[m][q] = r < i1 ? lv.x[m][r] : 0.0;
        }
        double cvv[NC][4];
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int m = 0; m < NC; m++) cvv[m][q] = I[q] >= 0 ? cv.t[m][I[q]] : 0.0;
#pragma unroll
        for (int q = 0; q < 4; q++)
        {
            const int r = i + q * nthr;
            if (I[q] >= 0)
#pragma unroll
                for (int m = 0; m < NC; m++) lv.x[m][r] = fma(over, cvv[m][q], xv[m][q]);
        }
    }
}

// SpMV pass of every level below the finest (whose matrix is not staged through the shared-memory ring), for a
// block's slices [s0,s1) — its chunk of the grid partition on the big levels, of the team partition on the small ones.
//
// These levels are latency-bound, and what sets the latency is the WIDEST slice: a warp that walks a slice alone pays
// two dependent round trips (entries, then gathers) per batch of columns, and the couplings of the air-gap ring make
// a few slices of every coarse level 5-20 batches wide.  So a slice is cut into TASKS of BATCH columns; the block's
// warps take tasks, not slices, leave the per-row partial sums of a task in shared memory, and after one
// __syncthreads the warp that owns a slice adds its tasks in ascending order (deterministic) and applies the
// epilogue.  Whatever the width: one round trip for the entries, one for the gathers.
//   AMG_EPI_RESIDUAL  t = b - A x'            AMG_EPI_POST  t = x' + w D^-1 (b - A x')
//   PROL: x' = x + over * P * (t of the next level) formed on the fly (prolongation fused); else x' = x
// Reads v.x, v.b (cv.t), writes v.t.  Out of line on purpose: own register allocation, one copy of the code.
//   AMG_EPI_SCATTER   b = x * diag / w (stored), then x_c[agg] += w * dinv_c[agg] * (b - A x)   — scatter_restrict
enum { AMG_EPI_RESIDUAL = 0, AMG_EPI_POST = 1, AMG_EPI_SCATTER = 2 };

template <int NC, bool PROL>
__device__ __noinline__ void amg_spmv_tasks(const AmgLev &L, AmgV<NC> v, AmgV<NC> cv, int s0, int s1, int wk0, int wk1, int kind,
                                            double w, double over, unsigned char *tsm, const int *table = nullptr,
                                            unsigned long long *dbg = nullptr, const double *cdinv = nullptr)
{
    auto mark = [&](int k) {
        if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[k] = t; }
    };
    mark(0);
    constexpr int BATCH = (NC == 1) ? 8 : 4;
    constexpr int MAXT = AMG_TASK_PART_BYTES / (32 * 8 * NC); // 128 / 64
    int *hdr = (int *)tsm;                       // [0] slices of the group, [1] tasks of the group
    const int *first_task = hdr + 4;             // [MAXS+1] first task of each slice of the group
    const int *slice_base = first_task + AMG_TASK_MAXS + 1; // [MAXS+1] SELL offset of each slice of the group (+ the end)
    const int *task_info = slice_base + AMG_TASK_MAXS + 1;  // [128] (slice of the group << 16) | batch
    double *part = (double *)(tsm + AMG_TASK_HDR_BYTES);    // [task][component][lane]
    const int lane = threadIdx.x & 31, warp_b = threadIdx.x >> 5, nwarps_b = blockDim.x >> 5;
    if (table)
    {
        // resident table of the whole chunk (built once per kernel by amg_kernel_init)
        first_task = table + 2;
        slice_base = first_task + (s1 - s0 + 1);
        task_info = slice_base + (s1 - s0 + 1);
    }
    int cur = s0;
    while (cur < s1) // uniform over the block
    {
        int ns, ntasks;
        if (table) { ns = table[0]; ntasks = table[1]; }
        else
        {
            // ---- warp 0: task table of the next group of slices
            if (warp_b == 0)
            {
                int nt = 0;
                const int n = amg_task_table(L.slice_ptr, cur, s1, AMG_TASK_MAXS, MAXT, BATCH, hdr + 4, hdr + 4 + AMG_TASK_MAXS + 1,
                                             hdr + 4 + 2 * (AMG_TASK_MAXS + 1), &nt);
                if (lane == 0) { hdr[0] = n; hdr[1] = nt; }
            }
            __syncthreads();
            ns = hdr[0]; ntasks = hdr[1];
        }
        mark(1);
        if (ns == 0) break; // a single slice wider than MAXT * BATCH columns: the host never builds one (amg_build checks)
        // row operands of this warp's first slice: in flight during the task phase
        double d0 = 0.0, wd0 = 0.0, xr0[NC], br0[NC];
        if (warp_b < ns)
        {
            const int row = (cur + warp_b) * 32 + lane;
            d0 = L.diag[row];
            if (kind == AMG_EPI_POST) wd0 = w * L.dinv[row];
            int I = -1;
            if (PROL) I = L.agg[row];
#pragma unroll
            for (int m = 0; m < NC; m++)
            {
                xr0[m] = v.x[m][row];
                br0[m] = (kind == AMG_EPI_SCATTER) ? xr0[m] * d0 / w : v.b[m][row];
                if (PROL) { if (I >= 0) xr0[m] = fma(over, cv.t[m][I], xr0[m]); }
            }
        }
        mark(2);
        // rows kept out of the SELL slices (layout.h), dealt out to the warps from the last one down, alongside the
        // tasks: their off-diagonal sums go to v.t, where the owner of the slice picks them up after the barrier
        if (wk1 > wk0 && cur == s0)
        {
#pragma unroll
            for (int m = 0; m < NC; m++)
            {
                double *tm = v.t[m];
                wide_rows_warp(L.wide, L.wide_val, wk0, wk1, v.x[m], [&](int row, double off) { tm[row] = off; });
            }
        }
        // ---- tasks: BATCH columns of one slice, lane = row
        for (int t = warp_b; t < ntasks; t += nwarps_b)
        {
            const int info = task_info[t], sl = info >> 16, k0 = (info & 0xffff) * BATCH;
            const int base = slice_base[sl], wdt = (slice_base[sl + 1] - base) >> 5;
            const int row = (cur + sl) * 32 + lane;
            const double *vv = L.sell_val + base + lane;
            const int *cc = L.sell_col + base + lane;
            const int *ca = PROL ? L.sell_cagg + base + lane : nullptr;
            double a[BATCH];
            int j[BATCH], J[BATCH];
#pragma unroll
            for (int q = 0; q < BATCH; q++)
            {
                const bool ok = k0 + q < wdt;
                a[q] = ok ? vv[(k0 + q) * 32] : 0.0;
                j[q] = ok ? cc[(k0 + q) * 32] : row;
                if (PROL) J[q] = ok ? ca[(k0 + q) * 32] : -1;
            }
            double acc[NC];
#pragma unroll
            for (int m = 0; m < NC; m++) acc[m] = 0.0;
#pragma unroll
            for (int q = 0; q < BATCH; q++)
            {
#pragma unroll
                for (int m = 0; m < NC; m++)
                {
                    double xv = v.x[m][j[q]];
                    if (PROL) { if (J[q] >= 0) xv = fma(over, cv.t[m][J[q]], xv); }
                    acc[m] = fma(a[q], xv, acc[m]);
                }
            }
#pragma unroll
            for (int m = 0; m < NC; m++) part[(t * NC + m) * 32 + lane] = acc[m];
        }
        mark(3);
        __syncthreads();
        mark(4);
        // ---- the owner of a slice adds its tasks in ascending order and applies the epilogue
        for (int sl = warp_b; sl < ns; sl += nwarps_b)
        {
            const int s = cur + sl, row = s * 32 + lane;
            double d = d0, wd = wd0, xr[NC], br[NC];
#pragma unroll
            for (int m = 0; m < NC; m++) { xr[m] = xr0[m]; br[m] = br0[m]; }
            if (sl != warp_b)
            {
                d = L.diag[row];
                if (kind == AMG_EPI_POST) wd = w * L.dinv[row];
                int I = -1;
                if (PROL) I = L.agg[row];
#pragma unroll
                for (int m = 0; m < NC; m++)
                {
                    xr[m] = v.x[m][row];
                    br[m] = (kind == AMG_EPI_SCATTER) ? xr[m] * d / w : v.b[m][row];
                    if (PROL) { if (I >= 0) xr[m] = fma(over, cv.t[m][I], xr[m]); }
                }
            }
            double acc[NC];
#pragma unroll
            for (int m = 0; m < NC; m++) acc[m] = 0.0;
            for (int t = first_task[sl]; t < first_task[sl + 1]; t++)
#pragma unroll
                for (int m = 0; m < NC; m++) acc[m] += part[(t * NC + m) * 32 + lane];
            if (wk1 > wk0 && ((L.wide.mask[s] >> lane) & 1u))
#pragma unroll
                for (int m = 0; m < NC; m++) acc[m] += v.t[m][row];
#pragma unroll
            for (int m = 0; m < NC; m++)
            {
                const double res = br[m] - fma(d, xr[m], acc[m]);
                if (kind == AMG_EPI_SCATTER)
                {
                    v.b[m][row] = br[m]; // the post-smoothing pass of this level reads it
                    const int I = L.agg[row];
                    if (I >= 0) atomicAdd(cv.x[m] + I, w * cdinv[I] * res);
                }
                else v.t[m][row] = (kind == AMG_EPI_POST) ? fma(wd, res, xr[m]) : res;
            }
        }
        cur += ns;
        mark(5);
        __syncthreads(); // the table and the partial sums are reused by the next group
        mark(6);
    }
}

// the block's resident task table of level l, or null
__device__ __forceinline__ const int *amg_table(const AmgArgs &g, const int *s_wk, int l)
{
    const int off = s_wk[AMG_SWK_TAB + l];
    return off >= 0 ? (const int *)(xb_dyn_smem + g.task_smem_offset + AMG_TASK_HDR_BYTES + AMG_TASK_PART_BYTES) + off : nullptr;
}

// --------------------------------------------------------------------------------------------
// device: the team's sub-cycle over the levels team_from..last.  In: b and the pre-smoothed x of level
// team_from.  Out: t of level team_from.  Called by the blocks 0..team_blocks-1 only; the caller follows it
// with a grid barrier.
// --------------------------------------------------------------------------------------------
template <int NC>
__device__ __noinline__ void amg_team_cycle(const AmgArgs &g, const AmgArgs2 *g2, int *s_wk, int *stamp_k = nullptr)
{
    const int first = g.team_from, last = g.nlev - 1;
    const int B = blockIdx.x;
    const int lane = threadIdx.x & 31, warp_b = threadIdx.x >> 5, nwarps_b = blockDim.x >> 5;
    const double w = g.omega, over = g.over;
    unsigned char *tsm = xb_dyn_smem + g.task_smem_offset;
    int sk = stamp_k ? *stamp_k : AMG_MAX_STAMPS; // XFEMM_AMG_TRACE: a stamp behind every team barrier
    // barrier behind a phase on level l: among the blocks that share the level
    auto sync = [&](int l) {
        if (l >= g.solo_from) { __syncthreads(); return; }
        if (g.stamps && g.trace_detail) { __syncthreads(); amg_stamp(g, sk); } // block 0's own part of the phase ends here
        team_sync(g, s_wk);
        amg_stamp(g, sk);
    };
    // SpMV pass over this block's slices of team level l
    auto spmv = [&](int l, bool prol, const AmgV<NC> &v, const AmgV<NC> &cv, int kind) {
        const AmgLev &L = g.lev[l];
        const int T = amg_team_width(g, l);
        const int s0 = amg_chunk(L.nslices, B, T), s1 = amg_chunk(L.nslices, B + 1, T);
        unsigned long long *dbg = (g.stamps && g.trace_detail && B == 0 && l == first && kind == AMG_EPI_RESIDUAL) ? g.stamps + AMG_MAX_STAMPS : nullptr;
        const int *table = s_wk[AMG_SWK_TAB + l] >= 0 ? (const int *)(tsm + AMG_TASK_HDR_BYTES + AMG_TASK_PART_BYTES) + s_wk[AMG_SWK_TAB + l] : nullptr;
        const double *cdinv = l + 1 < g.nlev ? g.lev[l + 1].dinv : nullptr;
        if (prol) amg_spmv_tasks<NC, true>(L, v, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], kind, w, over, tsm, table);
        else amg_spmv_tasks<NC, false>(L, v, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], kind, w, over, tsm, table, dbg, cdinv);
    };
    const bool scatter = NC == 1 && g.scatter_restrict != 0;
    // does this block take part in level l?
    auto in = [&](int l) { return B < amg_team_width(g, l); };

    // ---- down
    for (int l = first; l < last; l++)
    {
        if (!in(l)) break; // the rest of the way down and back up to here is block 0's
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        const AmgV<NC> lv = amg_vecs<NC>(g, g2, l), cv = amg_vecs<NC>(g, g2, l + 1);
        const int T = amg_team_width(g, l);
        if (scatter)
        {
            // residual and restriction in one pass: the coarse iterate collects the rows' contributions
            spmv(l, false, lv, cv, AMG_EPI_SCATTER);
            sync(l);
            continue;
        }
        spmv(l, false, lv, cv, AMG_EPI_RESIDUAL);
        sync(l);
        amg_restrict_level<NC>(L, C, lv, cv, B * blockDim.x + threadIdx.x, T * blockDim.x, w);
        sync(l); // (the last team level hands over to block 0 through a team barrier: everybody passes it)
    }
    // ---- coarsest level
    if (in(last))
    {
        const AmgLev &L = g.lev[last];
        const AmgV<NC> lv = amg_vecs<NC>(g, g2, last);
        const int T = amg_team_width(g, last);
        if (g.coarse_inv)
        {
            // exact solve with the dense inverse prepared by amg_refresh: t = A^-1 b, one warp per row
            const int n = L.n_pad;
            for (int r = B * nwarps_b + warp_b; r < n; r += T * nwarps_b)
            {
                double acc[NC];
#pragma unroll
                for (int m = 0; m < NC; m++) acc[m] = 0.0;
                for (int c = lane; c < n; c += 32)
                {
                    const double a = g.coarse_inv[(size_t)r * n + c];
                    // (scatter_restrict: the right-hand side is implied by the collected iterate)
                    const double sc = (scatter && last > 0) ? L.diag[c] / w : 0.0;
#pragma unroll
                    for (int m = 0; m < NC; m++) acc[m] = fma(a, (scatter && last > 0) ? lv.x[m][c] * sc : lv.b[m][c], acc[m]);
                }
#pragma unroll
                for (int m = 0; m < NC; m++)
                {
                    const double vsum = warp_sum(acc[m]);
                    if (lane == 0) lv.t[m][r] = vsum;
                }
            }
        }
        else
        {
            // damped-Jacobi sweeps x -> t -> x ... (odd count: ends in t)
            for (int it = 0; it < g.coarse_sweeps; it++)
            {
                AmgV<NC> sv = lv;
                if (it & 1)
#pragma unroll
                    for (int m = 0; m < NC; m++) { sv.x[m] = lv.t[m]; sv.t[m] = lv.x[m]; }
                spmv(last, false, sv, sv, AMG_EPI_POST);
                if (it + 1 < g.coarse_sweeps) sync(last);
            }
        }
        if (last > first) sync(last);
    }
    // ---- up: t = x' + w D^-1 (b - A x'),  x' = x + over * P * t_coarse
    for (int l = last - 1; l >= first; l--)
    {
        if (!in(l))
            continue;
        if (l + 1 >= g.solo_from && l < g.solo_from)
        {
            // block 0 comes back from its solo levels: the team waits for it here (block 0's stores to the vectors of
            // level l + 1 are ordered by the barrier's release)
            team_sync(g, s_wk);
            amg_stamp(g, sk);
        }
        const AmgLev &L = g.lev[l];
        const AmgV<NC> lv = amg_vecs<NC>(g, g2, l), cv = amg_vecs<NC>(g, g2, l + 1);
        const int T = amg_team_width(g, l);
        if (L.sell_cagg) spmv(l, true, lv, cv, AMG_EPI_POST);
        else
        {
            amg_prolong_rows<NC>(L, lv, cv, 0, L.n_pad, B * blockDim.x + threadIdx.x, T * blockDim.x, over);
            sync(l);
            spmv(l, false, lv, cv, AMG_EPI_POST);
        }
        if (l > first) sync(l);
    }
    if (stamp_k) *stamp_k = sk;
}

// --------------------------------------------------------------------------------------------
// device: grid-wide part of the one-component cycle
// --------------------------------------------------------------------------------------------
// y-phase over the block's slices of the FINEST level: calls epi(row, (A x)[row]) for every own row; the matrix
// stream comes through the shared-memory ring (stage.cuh) when the block has tiles
template <class Epi>
__device__ __forceinline__ void amg_spmv_phase(const AmgLev &L, int s0, int s1, int wk0, int wk1, double *x,
                                               double *wide_tmp, int *s_next, Epi epi, StageSmem *S = nullptr,
                                               const double *pf0 = nullptr, const double *pf1 = nullptr, int wr0 = -1, int wr1 = 0,
                                               int wstride = 1)
{
    // the wide rows this block takes (before its warps join the stream): wr0, wr0 + wstride, ... < wr1.  Default: the
    // wide rows of the whole level dealt out over the blocks; the partitioned solver passes the block's own ones
    // (stride 1), because a block pushes the rows of its chunk to the peers without a barrier in between.
    if (wr0 < 0) { wr0 = (int)blockIdx.x; wr1 = L.wide.n_rows; wstride = (int)gridDim.x; }
    const int lane = threadIdx.x & 31;
    if (S && S->ntiles > 0)
    {
        // pf0/pf1: row-indexed arrays the caller's epilogue reads (prefetched while the warp waits for its tile).
        // The wide rows of the chunk go to a few warps BEFORE they join the stream (no block-wide pre-pass in front of
        // it: a single wide row used to hold its whole block up by ~9 us); the stream skips the flagged rows.
        staged_slices(*S, L.slice_ptr, L.sell_val, L.sell_col, s0, s1, x, s_next,
                      [&](int row) {
                          if (pf0) prefetch_l1(pf0 + row);
                          if (pf1) prefetch_l1(pf1 + row);
                          return make_double2(L.diag[row], x[row]);
                      },
                      [&](int s, int row, double acc, double2 dx) {
                          if (L.wide.n_rows && ((L.wide.mask[s] >> lane) & 1u)) return;
                          epi(row, fma(dx.x, dx.y, acc));
                      },
                      [&]() {
                          // wide rows [wr0, wr1) of the level (all of them, or a rank's), dealt out over the blocks
                          if (wr1 > wr0)
                              wide_rows_warp(L.wide, L.wide_val, wr0, wr1, x,
                                             [&](int row, double off) { epi(row, fma(L.diag[row], x[row], off)); }, wstride);
                      });
        return;
    }
    wide_rows_block<double>(L.wide, L.wide_val, L.wide_acc, wk0, wk1, x, wide_tmp);
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
        double acc = 0.0;
#pragma unroll 8
        for (int k = 0; k < w; k++) acc = fma(__ldcs(v + k * 32), x[__ldcs(c + k * 32)], acc);
        if (wk1 > wk0 && ((L.wide.mask[s] >> lane) & 1u)) acc += wide_tmp[row];
        epi(row, fma(L.diag[row], x[row], acc));
    }
    __syncthreads(); // s_next is reused by the next phase
}

// z = M^-1 r.  Level 0: b = r (input), t = z (output), x = scratch.  Returns this thread's partial of r.z.
// x0_ready: the caller already stored the pre-smoothed x = omega D^-1 r of level 0 (fused into its
// residual update) and synchronised the grid.
__device__ __forceinline__ double amg_vcycle(const AmgArgs &g, const AmgArgs &sg /* copy in shared memory, for the out-of-line passes */,
                                             GridBar &grid, int *s_next,
                                             int *s_wk /* [AMG_SWK_SIZE] in shared memory, see amg_kernel_init */,
                                             StageSmem *S0 /* ring of the finest level, or null */, bool x0_ready = false)
{
    const int G = gridDim.x, B = blockIdx.x;
    const int last = g.nlev - 1;
    const int lc = g.team_from; // levels lc..last belong to the team (1 <= lc <= last unless nlev == 1)
    const int tid_g = B * blockDim.x + threadIdx.x, nthr_g = G * blockDim.x;
    const double w = g.omega, over = g.over;
    double dot = 0.0;
    int sk = 0;
    amg_stamp(g, sk);
    // ---- down: pre-smooth from a zero guess, residual, restrict
    if (!x0_ready || last == 0)
    {
        const AmgLev &L = g.lev[0];
        const int r0 = amg_chunk(L.nslices, B, G) * 32, r1 = amg_chunk(L.nslices, B + 1, G) * 32;
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
        const int s0 = amg_chunk(L.nslices, B, G), s1 = amg_chunk(L.nslices, B + 1, G);
        const AmgV<1> lv = amg_vecs<1>(g, nullptr, l), cv = amg_vecs<1>(g, nullptr, l + 1);
        // t = b - A x   (t doubles as the scratch the wide rows leave their partial sums in)
        if (l == 0)
        {
            amg_block_stamp(g, 1, true);
            if (g.scatter_restrict)
            {
                const AmgLev &C = g.lev[1];
                amg_spmv_phase(L, s0, s1, s_wk[0], s_wk[1], L.x, L.t, s_next, [&](int row, double ax) {
                    const int I = L.agg[row];
                    if (I >= 0) atomicAdd(C.x + I, w * C.dinv[I] * (L.b[row] - ax));
                }, S0, L.b);
            }
            else
                amg_spmv_phase(L, s0, s1, s_wk[0], s_wk[1], L.x, L.t, s_next, [&](int row, double ax) { L.t[row] = L.b[row] - ax; }, S0, L.b);
            amg_block_stamp(g, 1);
        }
        else
            amg_spmv_tasks<1, false>(sg.lev[l], lv, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1],
                                     g.scatter_restrict ? AMG_EPI_SCATTER : AMG_EPI_RESIDUAL, w, over,
                                     xb_dyn_smem + g.task_smem_offset, amg_table(g, s_wk, l), nullptr, sg.lev[l + 1].dinv);
        grid.sync();
        amg_stamp(g, sk);
        if (g.scatter_restrict) continue; // the residual pass has restricted already
        // coarse right-hand side = sum over the aggregate ; coarse pre-smooth fused.  Rows are dealt out
        // thread by thread over the whole grid (coalesced pointer loads, no tail of idle blocks).
        amg_restrict_level<1>(sg.lev[l], sg.lev[l + 1], lv, cv, tid_g, nthr_g, w);
        grid.sync();
        amg_stamp(g, sk);
    }
    // ---- the small levels: the team, counter barriers; result in lev[lc].t
    if (B < g.team_blocks) amg_team_cycle<1>(sg, nullptr, s_wk, &sk);
    grid.sync();
    amg_stamp(g, sk);
    // ---- up: add the (over-)corrected coarse result (a row-wise gather: a scatter from the coarse rows
    // was measured 3x slower, its read-modify-writes serialise), then post-smooth into t
    for (int l = lc - 1; l >= 0; l--)
    {
        const AmgLev &L = g.lev[l];
        const int s0 = amg_chunk(L.nslices, B, G), s1 = amg_chunk(L.nslices, B + 1, G);
        const AmgV<1> lv = amg_vecs<1>(g, nullptr, l), cv = amg_vecs<1>(g, nullptr, l + 1);
        // own rows only: the post-smoothing pass below gathers x across blocks, hence the barrier
        amg_prolong_rows<1>(sg.lev[l], lv, cv, s0 * 32, s1 * 32, threadIdx.x, blockDim.x, over);
        // the coarse iterates have served this cycle (their levels are done): ready to collect the next one's
        if (l == 0 && g.scatter_restrict) amg_zero_coarse_iterates(g);
        grid.sync();
        amg_stamp(g, sk);
        if (l > 0)
        {
            amg_spmv_tasks<1, false>(sg.lev[l], lv, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], AMG_EPI_POST, w, over,
                                     xb_dyn_smem + g.task_smem_offset, amg_table(g, s_wk, l));
            grid.sync();
            amg_stamp(g, sk);
        }
        else
        {
            amg_block_stamp(g, 2, true);
            amg_spmv_phase(L, s0, s1, s_wk[0], s_wk[1], L.x, L.t, s_next, [&](int row, double ax) {
                const double br = L.b[row], z = fma(w * L.dinv[row], br - ax, L.x[row]);
                L.t[row] = z;
                dot = fma(br, z, dot);
            }, S0, L.b, L.dinv);
            amg_block_stamp(g, 2);
        }
    }
    amg_stamp(g, sk);
    return dot;
}

} // namespace xb
