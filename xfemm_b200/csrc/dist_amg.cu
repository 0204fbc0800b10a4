// xfemm_dpcg_* — ONE mesh solved by several GPUs of a node: row-partitioned multigrid-preconditioned CG
// (BASELINE config 5b).  Same recurrence and stopping rule as the single-GPU solver
// (reference: CBigLinProb::PCGSolve, cfemm/libfemm/spars.cpp:238-316), same preconditioner — the SAME
// multigrid hierarchy, so the iteration count does not depend on the number of GPUs.
//
// Layout.  Every rank (one process per GPU) holds the whole problem — pattern, values, hierarchy; assembly and the
// Galerkin refresh are replicated, they are a fraction of a percent of a solve — but WORKS on a contiguous range of
// the slices of every big level (the finest and the ones below it that the whole grid cycles, amg.cuh): rank r of W
// owns the slices [n*r/W, n*(r+1)/W) of the level's internal (Z-curve) order, its blocks split that range.  The small
// levels (the team's, a few ten thousand rows and less) are cycled by every rank redundantly: replicating them costs
// nothing and keeps the hierarchy global.
//
// Data path.  No collective library in the loop and no host: ONE persistent cooperative kernel per GPU.  The vectors
// that cross ranks live in a symmetric arena that every rank maps from every other (CUDA IPC, NVLink/NVSwitch peer
// access).  After the phase that produces a vector, a block stores the entries of ITS rows that other ranks read —
// halo columns of their rows, members of their aggregates, their prolongation sources; lists built once per pattern —
// straight into the peers' copies of that vector (same index: the arenas are symmetric), the grid barrier of the
// phase orders the stores, block 0 raises a sequence number in every peer's flag array (st.release.sys) and every
// block waits for the peers' flags (ld.acquire.sys) before the phase that consumes the vector.  The two dot products
// of an iteration travel the same way (one slot per rank, summed in rank order by everybody: bit-identical scalars on
// all ranks, so the data-dependent control flow stays uniform across the node).  Per iteration: 8-12 halo pushes of a
// few thousand doubles, 2 scalar exchanges, ~10 us of exposed NVLink latency in total.
//
// Safety.  Every wait has a deadline (%globaltimer); a rank that misses one raises an abort flag, stops waiting, and
// all ranks leave the iteration with XFEMM_ERR_DIVERGED at the next uniform check — a lost rank costs seconds, not a
// hung GPU.
#include "linprob.cuh"
#include "amg.cuh"
#include "pcg_args.cuh"
#include <cooperative_groups.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

namespace cg = cooperative_groups;

namespace xb {

constexpr int DIST_MAX_WORLD = 8;
constexpr int DIST_MAX_PART = 4;                 // partitioned levels at most
constexpr int DIST_LISTS = 3 * DIST_MAX_PART;    // per level: H (halo of x / p), R (residuals for the restriction), P (prolongation sources)
enum { DL_H = 0, DL_R = 1, DL_P = 2 };

struct DistArgs
{
    int rank = 0, world = 1, lp = 1;             // levels 0..lp-1 are partitioned, lp.. are cycled by every rank
    char *peer[DIST_MAX_WORLD] = {};             // comm arena of every rank (own: peer[rank])
    size_t off_flags = 0, off_slots = 0;         // u64 arrival counters[4][8] ; double slots[4][world]
    size_t off_x[DIST_MAX_PART] = {}, off_t[DIST_MAX_PART] = {}; // lev[l].x / lev[l].t
    size_t off_p = 0, off_xs = 0;                // CG direction, solution
    size_t off_bc = 0, off_xc = 0;               // b and x of level lp
    const int *send_idx = nullptr;               // all lists, rows ascending within a list
    const int *send_ptr = nullptr;               // [(DIST_LISTS * world) + 1]
    int *abort_flag = nullptr;                   // own memory: a wait missed its deadline
    unsigned long long timeout_ns = 0;
    unsigned long long seq0 = 0;                 // exchange steps / scalar exchanges of the earlier solves on this handle: the
    int ar0 = 0;                                 // arrival counters count on monotonically and are never reset
    int peer_grid[DIST_MAX_WORLD] = {};          // blocks of every rank's kernel: rank q has finished step s when its counter
                                                 // here has reached s * peer_grid[q]
    int slices0[AMG_MAX_LEVELS] = {};            // first slice of this rank per level (partitioned levels + level lp)
    int slices1[AMG_MAX_LEVELS] = {};
};

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_sys_add(unsigned long long *p, unsigned long long v)
{
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Arrival counters: four per source rank, step s uses counter s & 3.  A block that runs ahead announces step s + 1 on
// another counter than the stragglers of step s, so a count can only be completed by the blocks it is meant for
// (steps four apart cannot overlap: every step is a rendezvous of all ranks).
__device__ __forceinline__ int ctr_index(unsigned long long s) { return (int)(s & 3ull); }
__device__ __forceinline__ unsigned long long ctr_uses(unsigned long long s) { return (s >> 2) + ((s & 3ull) ? 1ull : 0ull); }

// per-block state of the exchange machinery (shared memory)
struct DistSmem
{
    int rng[DIST_LISTS][DIST_MAX_WORLD][2]; // this block's part of every send list: [k0,k1) into send_idx
};

// One per thread.  seq / ar count the exchange steps and the scalar exchanges done so far: every thread of every
// block of every rank runs through the same sequence, so the counters need no storage of their own.
struct DistCtx
{
    const DistArgs &d;
    DistSmem &sm;
    GridBar &grid;
    mutable unsigned long long seq;
    mutable int ar;

    __device__ __forceinline__ double *vec(int q, size_t off) const { return (double *)(d.peer[q] + off); }
    // counter of step s that rank `src` raises in the arena of rank q
    __device__ __forceinline__ unsigned long long *flags(int q, unsigned long long s, int src) const
    {
        return (unsigned long long *)(d.peer[q] + d.off_flags) + ctr_index(s) * DIST_MAX_WORLD + src;
    }

    // rows of this block that other ranks read: store them into the peers' copies of the vector at `off`
    __device__ __forceinline__ void push(int list, size_t off) const
    {
        if (d.world == 1) return;
        __syncthreads(); // the block's own stores to the vector are visible to all its threads
        const double *src = vec(d.rank, off);
        for (int q = 0; q < d.world; q++)
        {
            if (q == d.rank) continue;
            double *dst = vec(q, off);
            const int k1 = sm.rng[list][q][1];
            for (int k = sm.rng[list][q][0] + (int)threadIdx.x; k < k1; k += blockDim.x)
            {
                const int i = d.send_idx[k];
                dst[i] = src[i];
            }
        }
    }
    // a contiguous range of rows to every peer
    __device__ __forceinline__ void push_range(size_t off, int i0, int i1) const
    {
        if (d.world == 1) return;
        __syncthreads();
        const double *src = vec(d.rank, off);
        for (int q = 0; q < d.world; q++)
        {
            if (q == d.rank) continue;
            double *dst = vec(q, off);
            for (int i = i0 + (int)threadIdx.x; i < i1; i += blockDim.x) dst[i] = src[i];
        }
    }
    // every block, every rank, same order: wait until all peers have finished exchange step `s`
    __device__ __forceinline__ void wait_peers(unsigned long long s) const
    {
        if ((int)threadIdx.x < d.world && (int)threadIdx.x != d.rank)
        {
            const unsigned long long *f = flags(d.rank, s, threadIdx.x);
            const unsigned long long target = ctr_uses(s) * (unsigned long long)d.peer_grid[threadIdx.x];
            const unsigned long long t0 = gtime();
            while (ld_acquire_sys(f) < target)
            {
                if (*(volatile int *)d.abort_flag) break;
                if (gtime() - t0 > d.timeout_ns) { atomicExch(d.abort_flag, 1); break; }
            }
        }
        __syncthreads();
    }
    // End of a phase whose pushes the next phase of the peers consumes.  Every block announces itself to every
    // peer as soon as ITS rows are out (a release at system scope on the peer's arrival counter: the block's peer
    // stores are ordered before it) — the announcements travel while the local grid barrier, which the phase needs
    // anyway, completes — and then waits until every block of every peer has announced the same step.
    __device__ __forceinline__ void exchange() const
    {
        if (d.world == 1) { grid.sync(); return; }
        const unsigned long long s = ++seq;
        __syncthreads();
        if ((int)threadIdx.x < d.world && (int)threadIdx.x != d.rank) red_release_sys_add(flags(threadIdx.x, s, d.rank), 1ull);
        grid.sync();
        wait_peers(s);
    }
    // sum of `local` (identical in every block of this rank, final: a grid barrier lies behind it) over the ranks,
    // added in rank order on every rank: bit-identical everywhere
    __device__ __forceinline__ double allreduce(double local) const
    {
        if (d.world == 1) return local;
        const int slot = (ar & 3) * d.world;
        const unsigned long long s = ++seq;
        ar++;
        if (blockIdx.x == 0 && (int)threadIdx.x < d.world && (int)threadIdx.x != d.rank)
        {
            ((double *)(d.peer[threadIdx.x] + d.off_slots))[slot + d.rank] = local;
            red_release_sys_add(flags(threadIdx.x, s, d.rank), (unsigned long long)gridDim.x); // one block speaks for all
        }
        wait_peers(s);
        const double *mine = (const double *)(d.peer[d.rank] + d.off_slots) + slot;
        double v = 0.0;
        for (int q = 0; q < d.world; q++) v += (q == d.rank) ? local : ld_relaxed_sys(mine + q);
        return v;
    }
};

// this block's chunk [s0,s1) of this rank's slices of level l
__device__ __forceinline__ void dist_chunk(const DistArgs &d, int l, int &s0, int &s1)
{
    const int a = d.slices0[l], n = d.slices1[l] - a;
    s0 = a + (int)((long long)n * blockIdx.x / gridDim.x);
    s1 = a + (int)((long long)n * (blockIdx.x + 1) / gridDim.x);
}

// ---------------------------------------------------------------------------------------------------
// the V-cycle with the big levels partitioned over the ranks.  z = M^-1 r: level 0 b = r, t = z, x = scratch
// (pre-smoothed iterate stored and pushed by the caller).  Returns this thread's partial of r.z (own rows).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dist_vcycle(const AmgArgs &g, const AmgArgs &sg, const DistCtx &X, int *s_next, int *s_wk, StageSmem *S0,
                                              int wr0, int wr1 /* wide rows of this block's chunk of the finest level */)
{
    const DistArgs &d = X.d;
    const int lp = d.lp;
    const double w = g.omega, over = g.over;
    unsigned char *tsm = xb_dyn_smem + g.task_smem_offset;
    double dot = 0.0;
    // ---- down.  In: x of level 0 (pre-smoothed, halo exchanged)
    for (int l = 0; l < lp; l++)
    {
        const AmgLev &L = g.lev[l];
        const AmgLev &C = g.lev[l + 1];
        int s0, s1, c0, c1;
        dist_chunk(d, l, s0, s1);
        dist_chunk(d, l + 1, c0, c1);
        const AmgV<1> lv = amg_vecs<1>(g, nullptr, l), cv = amg_vecs<1>(g, nullptr, l + 1);
        if (l == 0)
            amg_spmv_phase(L, s0, s1, s_wk[0], s_wk[1], L.x, L.t, s_next, [&](int row, double ax) { L.t[row] = L.b[row] - ax; }, S0, L.b,
                           nullptr, wr0, wr1, 1);
        else
            amg_spmv_tasks<1, false>(sg.lev[l], lv, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], AMG_EPI_RESIDUAL, w, over, tsm, amg_table(g, s_wk, l));
        X.push(3 * l + DL_R, d.off_t[l]);
        X.exchange();
        // coarse right-hand side of the coarse rows of THIS block's chunk of level l+1 (members may be rows of peers)
        for (int I = c0 * 32 + (int)threadIdx.x; I < c1 * 32; I += blockDim.x) amg_restrict_row<1>(sg.lev[l], sg.lev[l + 1], lv, cv, I, w);
        if (l + 1 < lp) X.push(3 * (l + 1) + DL_H, d.off_x[l + 1]);
        else
        {
            // the first level every rank cycles on its own: everybody gets everybody's part of b and x
            X.push_range(d.off_bc, c0 * 32, c1 * 32);
            X.push_range(d.off_xc, c0 * 32, c1 * 32);
        }
        (void)C;
        X.exchange();
    }
    // ---- the small levels, on every rank alike
    if ((int)blockIdx.x < g.team_blocks) amg_team_cycle<1>(sg, nullptr, s_wk, nullptr);
    X.grid.sync();
    // ---- up
    for (int l = lp - 1; l >= 0; l--)
    {
        const AmgLev &L = g.lev[l];
        int s0, s1;
        dist_chunk(d, l, s0, s1);
        const AmgV<1> lv = amg_vecs<1>(g, nullptr, l), cv = amg_vecs<1>(g, nullptr, l + 1);
        amg_prolong_rows<1>(sg.lev[l], lv, cv, s0 * 32, s1 * 32, threadIdx.x, blockDim.x, over);
        X.push(3 * l + DL_H, d.off_x[l]);
        X.exchange();
        if (l > 0)
        {
            amg_spmv_tasks<1, false>(sg.lev[l], lv, cv, s0, s1, s_wk[2 * l], s_wk[2 * l + 1], AMG_EPI_POST, w, over, tsm, amg_table(g, s_wk, l));
            X.push(3 * l + DL_P, d.off_t[l]);
            X.exchange();
        }
        else
        {
            amg_spmv_phase(L, s0, s1, s_wk[0], s_wk[1], L.x, L.t, s_next, [&](int row, double ax) {
                const double br = L.b[row], z = fma(w * L.dinv[row], br - ax, L.x[row]);
                L.t[row] = z;
                dot = fma(br, z, dot);
            }, S0, L.b, L.dinv, wr0, wr1, 1);
        }
    }
    return dot;
}

// ---------------------------------------------------------------------------------------------------
// the kernel: k_pcg<AMG> (linprob.cu) with every phase restricted to this rank's rows and the exchanges in between
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
k_dpcg(const __grid_constant__ PcgArgs a, const __grid_constant__ AmgArgs g, const __grid_constant__ DistArgs d)
{
    __shared__ double sm[33];
    __shared__ int s_next;
    __shared__ int s_wk[AMG_SWK_SIZE];
    __shared__ __align__(16) unsigned char sg_raw[sizeof(AmgArgs)];
    __shared__ DistSmem ds;
    GridBar grid{cg::this_grid(), g.grid_ctr, s_wk + AMG_SWK_EPOCH + 1};
    const int lane = threadIdx.x & 31;
    const int G = gridDim.x, B = blockIdx.x;
    (void)lane;
    int s0, s1;
    dist_chunk(d, 0, s0, s1);
    const int r0 = s0 * 32, r1 = s1 * 32;
    double *partA = a.part, *partB = a.part + G;
    int wk0 = 0, wk1 = 0; // wide rows of this block's chunk of the finest level
    if (a.wide.n_rows) wide_range(a.wide, r0, r1, wk0, wk1);
    const int wr0 = wk0, wr1 = wk1;
    StageSmem S = stage_carve(xb_dyn_smem);
    if (a.tiles)
    {
        const int t0 = a.tile_ptr[B], t1 = a.tile_ptr[B + 1];
        S.ntiles = t1 - t0;
        S.tile = a.tiles + t0;
    }
    // ---- per-kernel set-up: shared copy of the arguments, wide-row ranges and task tables of this block's chunks
    {
        const unsigned *src = (const unsigned *)&g;
        unsigned *dst = (unsigned *)sg_raw;
        for (int i = threadIdx.x; i < (int)(sizeof(AmgArgs) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if ((int)threadIdx.x < g.nlev)
    {
        const int l = threadIdx.x;
        const AmgLev &L = g.lev[l];
        int k0 = 0, k1 = 0;
        if (L.wide.n_rows)
        {
            if (l < d.lp) { int c0, c1; dist_chunk(d, l, c0, c1); wide_range(L.wide, c0 * 32, c1 * 32, k0, k1); }
            else if (B < amg_team_width(g, l))
                wide_range(L.wide, amg_chunk(L.nslices, B, amg_team_width(g, l)) * 32, amg_chunk(L.nslices, B + 1, amg_team_width(g, l)) * 32, k0, k1);
        }
        s_wk[2 * l] = k0; s_wk[2 * l + 1] = k1;
        s_wk[AMG_SWK_TAB + l] = -1;
    }
    if (threadIdx.x == 0) { s_wk[AMG_SWK_EPOCH] = 0; s_wk[AMG_SWK_EPOCH + 1] = 0; }
    // this block's part of every send list: the entries whose row lies in the block's chunk of the list's level
    for (int e = threadIdx.x; e < DIST_LISTS * DIST_MAX_WORLD; e += blockDim.x)
    {
        const int list = e / DIST_MAX_WORLD, q = e % DIST_MAX_WORLD, l = list / 3;
        int lo = 0, hi = 0;
        if (q < d.world && q != d.rank && l < d.lp)
        {
            const int k0 = d.send_ptr[list * d.world + q], k1 = d.send_ptr[list * d.world + q + 1];
            int c0, c1;
            dist_chunk(d, l, c0, c1);
            const int ra = c0 * 32, rb = c1 * 32;
            int x = k0, y = k1;
            while (x < y) { const int m = (x + y) >> 1; if (d.send_idx[m] < ra) x = m + 1; else y = m; }
            lo = x; y = k1;
            while (x < y) { const int m = (x + y) >> 1; if (d.send_idx[m] < rb) x = m + 1; else y = m; }
            hi = x;
        }
        ds.rng[list][q][0] = lo; ds.rng[list][q][1] = hi;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
        constexpr int MAXT = AMG_TASK_PART_BYTES / (32 * 8);
        int *tab = (int *)(xb_dyn_smem + g.task_smem_offset + AMG_TASK_HDR_BYTES + AMG_TASK_PART_BYTES);
        int used = 0;
        for (int l = 1; l < g.nlev; l++)
        {
            const AmgLev &L = g.lev[l];
            int c0 = 0, c1 = 0;
            if (l < d.lp) dist_chunk(d, l, c0, c1);
            else if (B < amg_team_width(g, l)) { c0 = amg_chunk(L.nslices, B, amg_team_width(g, l)); c1 = amg_chunk(L.nslices, B + 1, amg_team_width(g, l)); }
            const int nsl = c1 - c0, need = 2 + 2 * (nsl + 1) + MAXT;
            if (nsl <= 0 || nsl > AMG_TASK_MAXS || used + need > AMG_TABLE_INTS) continue;
            int *t = tab + used, ntasks = 0;
            const int ns = amg_task_table(L.slice_ptr, c0, c1, nsl, MAXT, 8, t + 2, t + 2 + (nsl + 1), t + 2 + 2 * (nsl + 1), &ntasks);
            if (ns == nsl)
            {
                if (threadIdx.x == 0) { t[0] = ns; t[1] = ntasks; s_wk[AMG_SWK_TAB + l] = used; }
                used += need;
            }
        }
    }
    __syncthreads();
    const AmgArgs &sg = *reinterpret_cast<const AmgArgs *>(sg_raw);
    const DistCtx X{d, ds, grid, d.seq0, d.ar0};
    const AmgLev &L0 = g.lev[0];

    // z = M^-1 r for the r in a.r (own rows): pre-smoothed iterate, its halo, the cycle; returns the global r.z
    auto precond_dot = [&](double *part_out) -> double {
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) L0.x[i] = g.omega * a.dinv[i] * a.r[i];
        X.push(DL_H, d.off_x[0]);
        X.exchange();
        double part = dist_vcycle(g, sg, X, &s_next, s_wk, &S, wr0, wr1);
        part = block_sum(part, sm);
        if (threadIdx.x == 0) part_out[B] = part;
        grid.sync();
        return X.allreduce(reduce_partials(part_out, G, sm));
    };

    // ---- res_o = b . M^-1 b   (spars.cpp:257-259)
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
    {
        a.r[i] = a.b[i];
        if (!a.warm) a.x[i] = 0.0;
    }
    __syncthreads();
    const double res_o = precond_dot(partA);
    if (res_o == 0.0)
    {
        if (B == 0 && threadIdx.x == 0)
        {
            a.mail[0] = 0; a.mail[1] = 0; a.mail[2] = XFEMM_OK; a.mail[3] = 0; a.mail[4] = 0;
            a.mail[5] = (double)X.seq; a.mail[6] = (double)*(volatile int *)d.abort_flag; a.mail[7] = (double)X.ar;
        }
        return;
    }
    // ---- r = b - A x
    if (a.warm)
    {
        X.push(DL_H, d.off_xs);
        X.exchange();
        wide_rows_block<double>(a.wide, a.wide_val, a.wide_acc, wk0, wk1, a.x, a.u);
        __syncthreads();
        const int nwarps_b = blockDim.x >> 5, warp_b = threadIdx.x >> 5;
        for (int s = s0 + warp_b; s < s1; s += nwarps_b)
        {
            const int base = a.slice_ptr[s], wd = (a.slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
            double acc = 0.0;
            for (int k = 0; k < wd; k++) acc = fma(a.sell_val[base + lane + k * 32], a.x[a.sell_col[base + lane + k * 32]], acc);
            if (wk1 > wk0 && ((a.wide.mask[s] >> lane) & 1u)) acc += a.u[row];
            a.r[row] = a.b[row] - fma(a.diag[row], a.x[row], acc);
        }
        __syncthreads();
    }
    // ---- z = M^-1 r ; p = z ; res = z.r
    double res = precond_dot(partB);
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) a.p[i] = a.z[i];
    X.push(DL_H, d.off_p);
    X.exchange();

    int it = 0, status = XFEMM_OK;
    double er = 0.0;
    while (true)
    {
        // u = A p (own rows) ; p.Ap
        double part = 0.0;
        if (S.ntiles > 0)
        {
            staged_slices(S, a.slice_ptr, a.sell_val, a.sell_col, s0, s1, a.p, &s_next,
                          [&](int row) { return make_double2(a.diag[row], a.p[row]); },
                          [&](int s, int row, double acc, double2 dp) {
                              if (a.wide.n_rows && ((a.wide.mask[s] >> lane) & 1u)) return; // a wide row: see `before`
                              acc = fma(dp.x, dp.y, acc);
                              a.u[row] = acc;
                              part = fma(dp.y, acc, part);
                          },
                          [&]() {
                              // the wide rows of this block's own chunk (see amg_spmv_phase)
                              if (wk1 > wk0)
                                  wide_rows_warp(a.wide, a.wide_val, wk0, wk1, a.p, [&](int row, double off) {
                                      const double pr = a.p[row], acc = fma(a.diag[row], pr, off);
                                      a.u[row] = acc;
                                      part = fma(pr, acc, part);
                                  });
                          });
        }
        else
        {
            wide_rows_block<double>(a.wide, a.wide_val, a.wide_acc, wk0, wk1, a.p, a.u);
            __syncthreads();
            const int nwarps_b = blockDim.x >> 5, warp_b = threadIdx.x >> 5;
            for (int s = s0 + warp_b; s < s1; s += nwarps_b)
            {
                const int base = a.slice_ptr[s], wd = (a.slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
                double acc = 0.0;
                for (int k = 0; k < wd; k++) acc = fma(a.sell_val[base + lane + k * 32], a.p[a.sell_col[base + lane + k * 32]], acc);
                if (wk1 > wk0 && ((a.wide.mask[s] >> lane) & 1u)) acc += a.u[row];
                const double pr = a.p[row];
                acc = fma(a.diag[row], pr, acc);
                a.u[row] = acc;
                part = fma(pr, acc, part);
            }
        }
        part = block_sum(part, sm);
        if (threadIdx.x == 0) partA[B] = part;
        grid.sync();
        const double pAp = X.allreduce(reduce_partials(partA, G, sm));
        const int aborted = *(volatile int *)d.abort_flag; // read behind a barrier and an exchange: the same in every block
        if (!(pAp > 0.0) || aborted) { status = XFEMM_ERR_DIVERGED; break; }
        const double del = res / pAp;
        // r -= del u ; pre-smoothed iterate of the finest level
        for (int i = r0 + threadIdx.x; i < r1; i += 4 * blockDim.x)
        {
            double uv[4], rv[4], dv[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int k = i + q * blockDim.x;
                const bool ok = k < r1;
                uv[q] = ok ? a.u[k] : 0.0; rv[q] = ok ? a.r[k] : 0.0; dv[q] = ok ? a.dinv[k] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int k = i + q * blockDim.x;
                if (k < r1)
                {
                    const double ri = fma(-del, uv[q], rv[q]);
                    a.r[k] = ri;
                    L0.x[k] = g.omega * dv[q] * ri;
                }
            }
        }
        X.push(DL_H, d.off_x[0]);
        X.exchange();
        part = dist_vcycle(g, sg, X, &s_next, s_wk, &S, wr0, wr1);
        part = block_sum(part, sm);
        if (threadIdx.x == 0) partB[B] = part;
        grid.sync();
        const double res_new = X.allreduce(reduce_partials(partB, G, sm));
        const double rho = res_new / res;
        res = res_new;
        for (int i = r0 + threadIdx.x; i < r1; i += 4 * blockDim.x)
        {
            double pv[4], xv[4], zv[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int k = i + q * blockDim.x;
                const bool ok = k < r1;
                pv[q] = ok ? a.p[k] : 0.0; xv[q] = ok ? a.x[k] : 0.0; zv[q] = ok ? a.z[k] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                const int k = i + q * blockDim.x;
                if (k < r1) { a.x[k] = fma(del, pv[q], xv[q]); a.p[k] = fma(rho, pv[q], zv[q]); }
            }
        }
        it++;
        er = sqrt(res / res_o);
        if (!(er == er)) { status = XFEMM_ERR_DIVERGED; break; }
        if (er <= a.precision) break;
        if (it >= a.max_iter) { status = XFEMM_ERR_DIVERGED; break; }
        X.push(DL_H, d.off_p);
        X.exchange();
    }
    if (B == 0 && threadIdx.x == 0)
    {
        a.mail[0] = it; a.mail[1] = er; a.mail[2] = status; a.mail[3] = res_o; a.mail[4] = res;
        a.mail[5] = (double)X.seq; a.mail[6] = (double)*(volatile int *)d.abort_flag; a.mail[7] = (double)X.ar;
    }
}

// own rows of the solution to every peer (after the solve: the peers' kernels no longer read x; own rows of theirs
// are not touched), then the flag: one more exchange step, signalled by the last block to finish
__global__ void k_dist_share_rows(int world, int rank, size_t off, int i0, int i1, DistArgs d, unsigned long long s, unsigned *done)
{
    const double *src = (const double *)(d.peer[rank] + off);
    for (int q = 0; q < world; q++)
    {
        if (q == rank) continue;
        double *dst = (double *)(d.peer[q] + off);
        for (int i = i0 + blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += gridDim.x * blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0)
    {
        fence_sys();
        last = atomicAdd(done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && (int)threadIdx.x < world && (int)threadIdx.x != rank)
    {
        fence_sys();
        red_release_sys_add((unsigned long long *)(d.peer[threadIdx.x] + d.off_flags) + ctr_index(s) * DIST_MAX_WORLD + rank,
                            (unsigned long long)d.peer_grid[rank]);
        if (threadIdx.x == 0 || (rank == 0 && threadIdx.x == 1)) *done = 0;
    }
}
// wait for every peer's rows (flag >= s), then x (arena) -> the problem's internal solution vector
__global__ void k_dist_collect(int world, int rank, DistArgs d, unsigned long long s, int n, const double *__restrict__ src, double *__restrict__ dst)
{
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank)
    {
        const unsigned long long *f = (const unsigned long long *)(d.peer[rank] + d.off_flags) + ctr_index(s) * DIST_MAX_WORLD + threadIdx.x;
        const unsigned long long target = ctr_uses(s) * (unsigned long long)d.peer_grid[threadIdx.x];
        const unsigned long long t0 = gtime();
        while (ld_acquire_sys(f) < target)
            if (gtime() - t0 > d.timeout_ns) { atomicExch(d.abort_flag, 1); break; }
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = __ldcg(src + i);
}

} // namespace xb

using namespace xb;

// =====================================================================================================
// host side
// =====================================================================================================
struct xfemm_dpcg
{
    xfemm_linprob *p = nullptr; // not owned: the resident global problem
    int rank = 0, world = 1, lp = 1;
    char *arena = nullptr;
    size_t arena_bytes = 0;
    char *peer[DIST_MAX_WORLD] = {};
    bool connected = false;
    DistArgs d;
    DevBuf<int> send_idx, send_ptr, abort_flag;
    DevBuf<unsigned> done_ctr;
    std::vector<int> h_send_ptr;
    std::vector<long long> list_sizes; // entries per list (all peers), for the statistics
    int n_pad0 = 0;
    unsigned long long seq = 0; // exchange steps so far on this handle (same on every rank)
    int ar = 0;
    bool broken = false;        // a solve missed a deadline: the sequence numbers of the ranks may differ now
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double last_gather_ms = 0;
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// owner of slice s of a level with n slices split as [n*r/W, n*(r+1)/W)
inline int slice_owner(int s, int n, int W)
{
    int r = std::min(W - 1, (int)(((long long)s * W) / std::max(1, n))); // first guess, then fix up
    while (r > 0 && s < (int)((long long)n * r / W)) r--;
    while (r + 1 < W && s >= (int)((long long)n * (r + 1) / W)) r++;
    return r;
}

} // namespace

extern "C" {

// Plan + arena of rank `rank` of `world` for the resident problem `p` (its multigrid hierarchy must have been
// prepared: a solve or xfemm_linprob_prebuild with XFEMM_PRECOND_AMG).  Replaces the Jacobi-only row-block path of
// round 1 (xfemm_dist_*) for the multigrid preconditioner; reference recurrence: cfemm/libfemm/spars.cpp:238-316.
int xfemm_dpcg_create(xfemm_linprob *p, int rank, int world, xfemm_dpcg **out)
{
    XB_ARG(p && out, "null pointer");
    XB_ARG(world >= 1 && world <= DIST_MAX_WORLD && rank >= 0 && rank < world, "rank / world out of range (at most 8 ranks)");
    int rc = linprob_prepare(p, XFEMM_PRECOND_AMG);
    if (rc != XFEMM_OK) return rc;
    AmgHierarchy &h = *p->amg;
    AmgArgs g;
    amg_fill_args(p, h, g, 1 << 20);
    g.scatter_restrict = 0; // (this kernel runs the ordered restriction pass)
    const int nlev = g.nlev;
    XB_ARG(nlev >= 2, "the hierarchy has a single level: nothing to partition");
    int lp = g.team_from;
    XB_ARG(lp <= DIST_MAX_PART, "too many big levels for the partitioned cycle");
    std::unique_ptr<xfemm_dpcg> D(new xfemm_dpcg);
    D->p = p; D->rank = rank; D->world = world; D->lp = lp;
    std::vector<xfemm_linprob *> lev(nlev);
    lev[0] = p;
    for (int l = 1; l < nlev; l++) lev[l] = h.coarse[l - 1];
    D->n_pad0 = p->lay.n_pad;
    // ---- arena layout (identical on every rank: same pattern, same deterministic hierarchy)
    size_t off = 0;
    DistArgs &d = D->d;
    d.rank = rank; d.world = world; d.lp = lp;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    d.off_flags = take(4 * DIST_MAX_WORLD * sizeof(unsigned long long));
    d.off_slots = take(4 * DIST_MAX_WORLD * sizeof(double));
    for (int l = 0; l < lp; l++)
    {
        d.off_x[l] = take((size_t)lev[l]->lay.n_pad * 8);
        d.off_t[l] = take((size_t)lev[l]->lay.n_pad * 8);
    }
    d.off_p = take((size_t)p->lay.n_pad * 8);
    d.off_xs = take((size_t)p->lay.n_pad * 8);
    d.off_bc = take((size_t)lev[lp]->lay.n_pad * 8);
    d.off_xc = take((size_t)lev[lp]->lay.n_pad * 8);
    D->arena_bytes = off;
    XB_CUDA(cudaMalloc((void **)&D->arena, D->arena_bytes)); // its own allocation: exported whole through CUDA IPC
    XB_CUDA(cudaMemset(D->arena, 0, D->arena_bytes));
    D->peer[rank] = D->arena;
    for (int l = 0; l <= lp; l++)
    {
        d.slices0[l] = (int)((long long)lev[l]->lay.nslices * rank / world);
        d.slices1[l] = (int)((long long)lev[l]->lay.nslices * (rank + 1) / world);
    }
    // blocks of every rank's kernel (the GPUs of a node are alike): what xfemm_dpcg_solve launches
    for (int q = 0; q < world; q++)
    {
        const int own = (int)((long long)p->lay.nslices * (q + 1) / world) - (int)((long long)p->lay.nslices * q / world);
        d.peer_grid[q] = std::max(1, std::min(p->num_sms, (own + 31) / 32));
    }
    // ---- send lists
    // internal-order transfer maps of the partitioned levels (device copies exist; bring them to the host once)
    std::vector<std::vector<int>> agg(lp);
    for (int l = 0; l < lp; l++)
    {
        agg[l].resize(lev[l]->lay.n_pad);
        XB_CUDA(cudaMemcpy(agg[l].data(), h.xfer[l]->agg.p, agg[l].size() * sizeof(int), cudaMemcpyDeviceToHost));
    }
    std::vector<std::vector<std::vector<int>>> lists(DIST_LISTS, std::vector<std::vector<int>>(world));
    for (int l = 0; l < lp; l++)
    {
        xfemm_linprob *P = lev[l];
        const SellLayout &Lh = P->lay;
        const int ns = Lh.nslices, nsc = lev[l + 1]->lay.nslices;
        const int my0 = d.slices0[l] * 32, my1 = d.slices1[l] * 32;
        // H: my rows that are columns of rows of rank q.  Rows of peers are scanned in parallel; marks per (peer, row).
        {
            std::vector<std::vector<unsigned char>> mark(world);
            for (int q = 0; q < world; q++) if (q != rank) mark[q].assign((size_t)(my1 - my0), 0);
            parallel_for(Lh.n_pad, [&](int a0, int a1, int) {
                for (int ri = a0; ri < a1; ri++)
                {
                    if (ri >= my0 && ri < my1) continue;
                    const int i = Lh.i2c[ri];
                    if (i < 0) continue;
                    const int q = slice_owner(ri >> 5, ns, world);
                    for (int s = P->h_rowptr[i]; s < P->h_rowptr[i + 1]; s++)
                    {
                        const int rj = Lh.c2i[P->h_col[s]];
                        if (rj >= my0 && rj < my1) mark[q][(size_t)(rj - my0)] = 1; // benign race: all writers store 1
                    }
                }
            });
            for (int q = 0; q < world; q++)
                if (q != rank)
                    for (int r = my0; r < my1; r++) if (mark[q][(size_t)(r - my0)]) lists[3 * l + DL_H][q].push_back(r);
        }
        // R: my rows whose aggregate (a row of level l+1) is restricted by rank q
        for (int r = my0; r < my1; r++)
        {
            const int I = agg[l][r];
            if (I < 0) continue;
            const int q = slice_owner(I >> 5, nsc, world);
            if (q != rank) lists[3 * l + DL_R][q].push_back(r);
        }
        // P (lists of level l+1, used when level l+1 is partitioned too): my rows of level l+1 that rows of rank q on
        // level l aggregate into
        if (l + 1 < lp)
        {
            const int c0 = d.slices0[l + 1] * 32, c1 = d.slices1[l + 1] * 32;
            std::vector<std::vector<unsigned char>> mark(world);
            for (int q = 0; q < world; q++) if (q != rank) mark[q].assign((size_t)(c1 - c0), 0);
            for (int ri = 0; ri < Lh.n_pad; ri++)
            {
                if (ri >= my0 && ri < my1) continue;
                const int I = agg[l][ri];
                if (I < c0 || I >= c1) continue;
                mark[slice_owner(ri >> 5, ns, world)][(size_t)(I - c0)] = 1;
            }
            for (int q = 0; q < world; q++)
                if (q != rank)
                    for (int I = c0; I < c1; I++) if (mark[q][(size_t)(I - c0)]) lists[3 * (l + 1) + DL_P][q].push_back(I);
        }
    }
    std::vector<int> idx;
    D->h_send_ptr.assign((size_t)DIST_LISTS * world + 1, 0);
    D->list_sizes.assign(DIST_LISTS, 0);
    for (int li = 0; li < DIST_LISTS; li++)
        for (int q = 0; q < world; q++)
        {
            idx.insert(idx.end(), lists[li][q].begin(), lists[li][q].end());
            D->h_send_ptr[(size_t)li * world + q + 1] = (int)idx.size();
            D->list_sizes[li] += (long long)lists[li][q].size();
        }
    if (idx.empty()) idx.push_back(0);
    XB_CUDA(D->send_idx.upload(idx.data(), idx.size(), p->stream));
    XB_CUDA(D->send_ptr.upload(D->h_send_ptr.data(), D->h_send_ptr.size(), p->stream));
    XB_CUDA(D->abort_flag.alloc(1));
    XB_CUDA(D->abort_flag.zero(p->stream));
    XB_CUDA(D->done_ctr.alloc(1));
    XB_CUDA(D->done_ctr.zero(p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    d.send_idx = D->send_idx.p; d.send_ptr = D->send_ptr.p; d.abort_flag = D->abort_flag.p;
    const char *et = getenv("XFEMM_DIST_TIMEOUT_MS");
    d.timeout_ns = (unsigned long long)(et ? atoi(et) : 20000) * 1000000ull;
    XB_CUDA(cudaEventCreate(&D->e0));
    XB_CUDA(cudaEventCreate(&D->e1));
    if (world == 1) D->connected = true;
    *out = D.release();
    return XFEMM_OK;
}

int xfemm_dpcg_destroy(xfemm_dpcg *D)
{
    if (!D) return XFEMM_OK;
    cudaDeviceSynchronize();
    for (int q = 0; q < D->world; q++)
        if (q != D->rank && D->peer[q]) cudaIpcCloseMemHandle(D->peer[q]);
    if (D->arena) cudaFree(D->arena);
    if (D->e0) cudaEventDestroy(D->e0);
    if (D->e1) cudaEventDestroy(D->e1);
    delete D;
    return XFEMM_OK;
}

// 64-byte CUDA IPC handle of this rank's arena: hand it to the other ranks (any channel: torch.distributed, MPI,
// files) and pass all of them, in rank order, to xfemm_dpcg_connect
int xfemm_dpcg_ipc_handle(xfemm_dpcg *D, void *handle64)
{
    XB_ARG(D && handle64, "null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t hnd;
    XB_CUDA(cudaIpcGetMemHandle(&hnd, D->arena));
    memcpy(handle64, &hnd, 64);
    return XFEMM_OK;
}

int xfemm_dpcg_connect(xfemm_dpcg *D, const void *handles)
{
    XB_ARG(D && handles, "null pointer");
    for (int q = 0; q < D->world; q++)
    {
        if (q == D->rank) continue;
        cudaIpcMemHandle_t hnd;
        memcpy(&hnd, (const char *)handles + 64 * q, 64);
        void *ptr = nullptr;
        XB_CUDA(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
        D->peer[q] = (char *)ptr;
    }
    D->connected = true;
    return XFEMM_OK;
}

// Rendezvous through files for hosts without a messaging layer (the C++ front end: FSolver with the options
// dist_rank / dist_world): every rank writes its handle to <dir>/<job>.<rank> and reads the others'.
int xfemm_dpcg_connect_files(xfemm_dpcg *D, const char *dir, const char *job, int timeout_s)
{
    XB_ARG(D && dir && job, "null pointer");
    if (D->world == 1) return XFEMM_OK;
    char mine[64];
    int rc = xfemm_dpcg_ipc_handle(D, mine);
    if (rc != XFEMM_OK) return rc;
    auto path = [&](int q) { return std::string(dir) + "/" + job + "." + std::to_string(q); };
    {
        const std::string tmp = path(D->rank) + ".tmp";
        std::ofstream f(tmp, std::ios::binary);
        f.write(mine, 64);
        f.close();
        if (rename(tmp.c_str(), path(D->rank).c_str()) != 0) { set_error("cannot write %s", path(D->rank).c_str()); return XFEMM_ERR_IO; }
    }
    std::vector<char> all((size_t)64 * D->world, 0);
    memcpy(all.data() + 64 * D->rank, mine, 64);
    const auto t0 = std::chrono::steady_clock::now();
    for (int q = 0; q < D->world; q++)
    {
        if (q == D->rank) continue;
        while (true)
        {
            std::ifstream f(path(q), std::ios::binary);
            if (f && f.read(all.data() + 64 * q, 64)) break;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s)
            {
                set_error("rank %d never published its arena handle (%s)", q, path(q).c_str());
                return XFEMM_ERR_IO;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
    }
    return xfemm_dpcg_connect(D, all.data());
}

int xfemm_dpcg_stats(xfemm_dpcg *D, int *lp, long long *list_entries /* [12] */, long long *arena_bytes)
{
    XB_ARG(D, "null handle");
    if (lp) *lp = D->lp;
    if (list_entries) for (int i = 0; i < DIST_LISTS; i++) list_entries[i] = D->list_sizes[i];
    if (arena_bytes) *arena_bytes = (long long)D->arena_bytes;
    return XFEMM_OK;
}

// The solve.  Every rank calls it with the same arguments after a host-side barrier (the kernels wait for each
// other on the device).  b = the resident right-hand side of the problem, the solution lands in the problem's V on
// EVERY rank (own rows computed, the others received).  info->solve_ms: the persistent kernel alone (CUDA events).
int xfemm_dpcg_solve(xfemm_dpcg *D, double precision, int warm, int max_iter, xfemm_solve_info *info)
{
    XB_ARG(D, "null handle");
    XB_ARG(D->connected, "xfemm_dpcg_connect has not been called");
    xfemm_linprob *p = D->p;
    cudaStream_t st = p->stream;
    int rc = linprob_prepare(p, XFEMM_PRECOND_AMG); // value refresh of every level (replicated)
    if (rc != XFEMM_OK) return rc;
    rc = linprob_load_rhs_internal(p);
    if (rc != XFEMM_OK) return rc;
    const int n_pad = p->lay.n_pad;
    if (max_iter <= 0) max_iter = 200000;
    PcgArgs a = linprob_make_args(p, precision, warm ? 1 : 0, max_iter);
    AmgArgs g;
    DistArgs d = D->d;
    for (int q = 0; q < D->world; q++) d.peer[q] = D->peer[q];
    const size_t smem = STAGE_SMEM_BYTES + AMG_TASK_SMEM_BYTES;
    XB_CUDA(cudaFuncSetAttribute(k_dpcg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    XB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dpcg, 1024, smem));
    if (per_sm < 1) { set_error("k_dpcg does not fit on an SM"); return XFEMM_ERR_CUDA; }
    const int grid = d.peer_grid[D->rank];
    if (grid > p->max_grid) { XB_CUDA(p->partials.alloc((size_t)4 * grid)); p->max_grid = grid; }
    amg_fill_args(p, *p->amg, g, grid);
    g.scatter_restrict = 0; // (this kernel runs the ordered restriction pass)
    g.task_smem_offset = (int)STAGE_SMEM_BYTES;
    g.stamps = nullptr;
    if (g.team_from != D->lp) { set_error("the hierarchy changed since xfemm_dpcg_create"); return XFEMM_ERR_ARG; }
    // the vectors that cross ranks live in the arena
    char *A = D->arena;
    a.p = (double *)(A + d.off_p);
    a.x = (double *)(A + d.off_xs);
    a.part = p->partials.p;
    for (int l = 0; l < D->lp; l++) { g.lev[l].x = (double *)(A + d.off_x[l]); g.lev[l].t = (double *)(A + d.off_t[l]); }
    a.z = g.lev[0].t;
    a.w = g.lev[0].x;
    g.lev[D->lp].b = (double *)(A + d.off_bc);
    g.lev[D->lp].x = (double *)(A + d.off_xc);
    // stream tiles of this rank's slices of the finest level, cut for its blocks
    if (linprob_ensure_tiles_range(p, grid, d.slices0[0], d.slices1[0])) { a.tiles = p->tiles.p; a.tile_ptr = p->tile_ptr.p; }
    if (warm)
    {
        // the problem's V (caller order) → arena x (internal order)
        rc = linprob_load_guess_internal(p, a.x);
        if (rc != XFEMM_OK) return rc;
    }
    XB_ARG(!D->broken, "an earlier solve on this handle lost a rank: create a new handle");
    XB_CUDA(cudaMemsetAsync(D->abort_flag.p, 0, sizeof(int), st));
    d.seq0 = D->seq;
    d.ar0 = D->ar;
    XB_CUDA(cudaStreamSynchronize(st));
    void *kargs[] = {&a, &g, &d};
    XB_CUDA(cudaEventRecord(D->e0, st));
    ++g_launch;
    XB_CUDA(cudaLaunchCooperativeKernel((void *)k_dpcg, dim3(grid), dim3(1024), kargs, smem, st));
    XB_CUDA(cudaEventRecord(D->e1, st));
    XB_CUDA(cudaMemcpyAsync(p->h_mail, p->d_mail.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    xfemm_solve_info local;
    memset(&local, 0, sizeof local);
    local.precond = XFEMM_PRECOND_AMG;
    local.colors = g.nlev;
    local.iterations = (int)p->h_mail[0];
    local.er = p->h_mail[1];
    local.status = (int)p->h_mail[2];
    float ms = 0;
    cudaEventElapsedTime(&ms, D->e0, D->e1);
    local.solve_ms = ms;
    const double nn = (double)p->n, nz = (double)p->nnz;
    local.bytes_per_iter = 12.0 * nz + 124.0 * nn;
    if (info) *info = local;
    D->seq = (unsigned long long)p->h_mail[5];
    D->ar = (int)p->h_mail[7];
    if (p->h_mail[6] != 0.0)
    {
        D->broken = true;
        set_error("a rank missed an exchange deadline (peer lost or not launched)");
        return XFEMM_ERR_DIVERGED;
    }
    if (local.status != XFEMM_OK) { set_error("partitioned PCG did not converge: %d iterations, er=%g", local.iterations, local.er); return local.status; }
    (void)n_pad;
    return XFEMM_OK;
}

// After xfemm_dpcg_solve: every rank sends its rows of the solution to the others (peer stores + one more flag);
// xfemm_dpcg_finish waits for everybody's rows on the device and stores the complete solution in the problem's V.
// No host-side barrier is needed between the three calls.
int xfemm_dpcg_share_solution(xfemm_dpcg *D)
{
    XB_ARG(D, "null handle");
    XB_ARG(!D->broken, "an earlier solve on this handle lost a rank: create a new handle");
    xfemm_linprob *p = D->p;
    DistArgs d = D->d;
    for (int q = 0; q < D->world; q++) d.peer[q] = D->peer[q];
    D->seq += 1;
    if (D->world > 1)
    {
        ++g_launch;
        k_dist_share_rows<<<p->num_sms * 2, 512, 0, p->stream>>>(D->world, D->rank, d.off_xs, d.slices0[0] * 32, d.slices1[0] * 32, d,
                                                                  D->seq, D->done_ctr.p);
        XB_CUDA(cudaGetLastError());
    }
    return XFEMM_OK;
}
int xfemm_dpcg_finish(xfemm_dpcg *D)
{
    XB_ARG(D, "null handle");
    xfemm_linprob *p = D->p;
    DistArgs d = D->d;
    for (int q = 0; q < D->world; q++) d.peer[q] = D->peer[q];
    ++g_launch;
    k_dist_collect<<<p->num_sms * 2, 512, 0, p->stream>>>(D->world, D->rank, d, D->seq, p->lay.n_pad,
                                                           (const double *)(D->arena + D->d.off_xs), p->x.p);
    XB_CUDA(cudaGetLastError());
    int rc = linprob_store_solution_internal(p);
    if (rc != XFEMM_OK) return rc;
    int aborted = 0;
    XB_CUDA(cudaMemcpyAsync(&aborted, D->abort_flag.p, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    if (aborted) { D->broken = true; set_error("a rank never delivered its rows of the solution"); return XFEMM_ERR_DIVERGED; }
    return XFEMM_OK;
}

} // extern "C"
