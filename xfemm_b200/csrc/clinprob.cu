// xfemm_clinprob_* — GPU backend of CBigComplexLinProb (reference: cfemm/libfemm/cspars.cpp:119-1081):
// complex-symmetric systems of the time-harmonic formulations.  Same data layout as the real
// problem (CSR "boundary" representation written by assembly / boundary kernels, split SELL-32
// "solver" representation, see linprob.cuh) with double2 values, and the complex-symmetric
// preconditioned BiCG (= COCG, cspars.cpp:822-894: unconjugated inner products, stop on
// ||r||_2/||b||_2 <= Precision) as one persistent cooperative kernel, seeded like the reference by
// three CG steps on the normal equations of the preconditioned system (PCGSQStart, cspars.cpp:764-819).
#include "clinprob.cuh"
#include "layout_device.cuh"
#include "amg.cuh"
#include "amg2.cuh"
#include <complex>
#include <cooperative_groups.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace cg = cooperative_groups;
namespace xb {

__device__ __forceinline__ cd block_sum_c(cd v, double *smem)
{
    double a = block_sum(v.x, smem);
    double b = block_sum(v.y, smem);
    return make_double2(a, b);
}
__device__ __forceinline__ cd reduce_partials_c(const double *part, int count, double *smem)
{
    double a = reduce_partials(part, count, smem);
    double b = reduce_partials(part + count, count, smem);
    return make_double2(a, b);
}

__global__ void kc_sell_fill(long long len, const int *__restrict__ src, const cd *__restrict__ vals, cd *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < len; i += (long long)gridDim.x * blockDim.x)
    {
        int s = src[i];
        out[i] = s >= 0 ? vals[s] : make_double2(0., 0.);
    }
}
__global__ void kc_diag_fill(int n_pad, const int *__restrict__ diag_src, const cd *__restrict__ vals, cd *__restrict__ diag,
                             cd *__restrict__ dinv, int *singular_row)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int s = diag_src[r];
    cd d = s >= 0 ? vals[s] : make_double2(0., 0.);
    diag[r] = d;
    const bool zero = (d.x == 0.0 && d.y == 0.0);
    dinv[r] = zero ? make_double2(0., 0.) : cdiv(make_double2(1., 0.), d);
    if (s >= 0 && zero) atomicMin(singular_row, r);
}
__global__ void kc_to_internal(int n_pad, const int *__restrict__ i2c, const cd *__restrict__ in, cd *__restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int i = i2c[r];
    out[r] = i >= 0 ? in[i] : make_double2(0., 0.);
}
__global__ void kc_from_internal(int n_pad, const int *__restrict__ i2c, const cd *__restrict__ in, cd *__restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad) return;
    int i = i2c[r];
    if (i >= 0) out[i] = in[r];
}
__global__ void kc_conj(int n, cd *v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i].y = -v[i].y;
}
__global__ void kc_mul(int n, const cd *__restrict__ a, const cd *__restrict__ b, cd *out) // out = a .* b
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cmul(a[i], b[i]);
}
// y = y + alpha x  (alpha complex) ; optional second pair
__global__ void kc_axpy(int n, cd alpha, const cd *__restrict__ x, cd *y)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = cfma(alpha, x[i], y[i]);
}
__global__ void kc_xpay(int n, cd alpha, const cd *__restrict__ x, cd *y) // y = x + alpha y
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = cfma(alpha, y[i], x[i]);
}
// conj(x).y → out2[0..1] (single block finisher inside)
__global__ void kc_conjdot_partial(int n, const cd *__restrict__ x, const cd *__restrict__ y, double *part)
{
    __shared__ double sm[33];
    cd a = make_double2(0., 0.);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a = cfma(cconj(x[i]), y[i], a);
    a = block_sum_c(a, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = a.x; part[gridDim.x + blockIdx.x] = a.y; }
}
__global__ void kc_reduce2_final(int nblk, const double *part, double *out2)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) { a += part[i]; c += part[nblk + i]; }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { out2[0] = a; out2[1] = c; }
}

// values of the real companion problem of XFEMM_PRECOND_AMG: sr*Re(a_ij) + si*Im(a_ij) between mesh nodes, sr/si = the
// signs that make either part positive (semi-)definite (the time-harmonic formulations assemble
// -K - j*omega*sigma*M, stiffness K and mass M, so both are -1).  With P = K + omega*sigma*M the eigenvalues of
// P^-1 (K + j*omega*sigma*M) are (k + j m)/(k + m): on the segment between 1 and j whatever the mesh size and the
// frequency.  The appended circuit unknowns (rows >= n_nodes) are identity rows of the companion (the solver
// treats them with a Jacobi step); a diagonal that does not come out positive falls back to its modulus.
__global__ void kc_diag_sign(int n_nodes, const int *__restrict__ rowptr, const int *__restrict__ col,
                             const cd *__restrict__ vals, int *counts /* diagonals with Re>0, Re<0, Im>0, Im<0 */)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
        if (col[s] == i)
        {
            if (vals[s].x > 0) atomicAdd(counts, 1); else if (vals[s].x < 0) atomicAdd(counts + 1, 1);
            if (vals[s].y > 0) atomicAdd(counts + 2, 1); else if (vals[s].y < 0) atomicAdd(counts + 3, 1);
        }
}
__global__ void kc_fill_real(int n, int n_nodes, const int *__restrict__ rowptr, const int *__restrict__ col,
                             const int *__restrict__ src, const cd *__restrict__ vals, double *__restrict__ out,
                             double sr, double si)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
    {
        const int j = col[s];
        double v;
        if (i >= n_nodes) v = 1.0; // identity row (the companion pattern holds nothing but its diagonal)
        else
        {
            const cd a = vals[src[s]];
            v = sr * a.x + si * a.y;
            if (i == j && !(v > 0.0)) { v = sqrt(a.x * a.x + a.y * a.y); if (v == 0.0) v = 1.0; }
        }
        out[s] = v;
    }
}

__device__ __forceinline__ cd csell_row(const cd *__restrict__ val, const int *__restrict__ col, int base, int k0, int k1,
                                        int lane, const cd *x)
{
    cd acc = make_double2(0., 0.);
    const cd *v = val + base + lane;
    const int *c = col + base + lane;
#pragma unroll 4
    for (int k = k0; k < k1; k++)
    {
        cd a = __ldcs(v + k * 32);
        int j = __ldcs(c + k * 32);
        acc = cfma(a, x[j], acc);
    }
    return acc;
}

__global__ void __launch_bounds__(256)
kc_spmv(int nslices, const int *__restrict__ slice_ptr, const int *__restrict__ sell_col, const cd *__restrict__ sell_val,
        const cd *__restrict__ diag, const cd *x, cd *y, const unsigned *__restrict__ wide_mask)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < nslices; s += nwarps)
    {
        const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
        cd acc = csell_row(sell_val, sell_col, base, 0, w, lane, x);
        if (wide_mask && ((wide_mask[s] >> lane) & 1u)) acc = w_add(acc, y[row]); // left there by k_wide_pre
        y[row] = cfma(diag[row], x[row], acc);
    }
}

struct CocgArgs
{
    int n_pad, nslices, ncolors;
    const int *slice_ptr, *slice_wl, *slice_wu, *sell_col, *color_slice;
    const cd *sell_val, *diag, *dinv, *b;
    cd *x, *r, *p, *u, *z;
    double *part; // 6*gridDim
    double precision, omega;
    int max_iter, warm;
    double *mail;
    WideView wide;
    const cd *wide_val;
    cd *wide_acc;
    // XFEMM_PRECOND_AMG: rows of the appended circuit unknowns (caller index >= n_nodes) keep the Jacobi step
    const int *i2c;
    int n_nodes;
    double prec_sign; // sign that makes the real part of the matrix positive definite
    const int4 *tiles;   // stream tiles of the real companion's finest level (stage.cuh), or null
    const int *tile_ptr;
};

__device__ __forceinline__ int cnext_slice(int *ctr, int lane)
{
    int v = 0;
    if (lane == 0) v = atomicAdd(ctr, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}

// z = M^-1 r ; returns this thread's partial of the unconjugated z.r
template <int PRECOND>
__device__ __forceinline__ cd cprecond_apply(const CocgArgs &a, const AmgArgs &g, const AmgArgs &sg, const AmgArgs2 &g2, GridBar &grid, int r0,
                                             int r1, int warp_g, int nwarps_g, int lane, int *s_next, int *s_wk, StageSmem *S)
{
    cd dot = make_double2(0., 0.);
    if (PRECOND == XFEMM_PRECOND_AMG)
    {
        // Multigrid V-cycle of the REAL matrix P = |Re A| + |Im A| between mesh nodes (stiffness + omega*sigma*mass,
        // symmetric positive definite; kc_fill_real), applied to the real and the imaginary part of r — both in ONE
        // cycle (amg2.cuh).  A real symmetric M keeps the preconditioned operator complex symmetric, which is all
        // COCG needs.
        const AmgLev &L0 = g.lev[0];
        const AmgVec2 &V0 = g2.lev[0];
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            // circuit unknowns enter the cycle with a zero residual: their identity rows are outside the hierarchy's
            // couplings, and their own z is replaced by the Jacobi step below
            const bool circ = a.i2c[i] >= a.n_nodes;
            const cd ri = circ ? make_double2(0., 0.) : a.r[i];
            const double wd = g.omega * L0.dinv[i];
            L0.b[i] = ri.x; L0.x[i] = wd * ri.x;
            V0.b[i] = ri.y; V0.x[i] = wd * ri.y;
        }
        grid.sync();
        amg_vcycle2(g, sg, g2, grid, s_next, s_wk, S);
        // every block post-smoothed its own rows of level 0: no barrier needed before reading them back
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            cd zi = make_double2(a.prec_sign * L0.t[i], a.prec_sign * V0.t[i]);
            if (a.i2c[i] >= a.n_nodes) zi = cmul(a.r[i], a.dinv[i]); // circuit unknowns: Jacobi
            a.z[i] = zi;
            dot = cfma(zi, a.r[i], dot);
        }
        return dot;
    }
    if (PRECOND == XFEMM_PRECOND_JACOBI)
    {
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            cd ri = a.r[i], zi = cmul(ri, a.dinv[i]);
            a.z[i] = zi;
            dot = cfma(zi, ri, dot);
        }
        return dot;
    }
    // multicolour SSOR (cspars.cpp:439-480 in colour order)
    const double w = a.omega, c0 = w * (2.0 - w);
    const int m = a.ncolors;
    {
        const int rc0 = a.color_slice[1] * 32;
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            cd v = cscale(a.r[i], c0);
            if (i < rc0) v = cmul(v, a.dinv[i]);
            a.z[i] = v;
        }
    }
    grid.sync();
    for (int c = 1; c < m; c++)
    {
        const int sa = a.color_slice[c], sb = a.color_slice[c + 1];
        for (int s = sa + warp_g; s < sb; s += nwarps_g)
        {
            const int base = a.slice_ptr[s], wl = a.slice_wl[s], row = s * 32 + lane;
            cd acc = csell_row(a.sell_val, a.sell_col, base, 0, wl, lane, a.z);
            cd v = cmul(csub(a.z[row], cscale(acc, w)), a.dinv[row]);
            a.z[row] = v;
            if (c == m - 1) dot = cfma(a.r[row], v, dot);
        }
        grid.sync();
    }
    for (int c = m - 2; c >= 0; c--)
    {
        const int sa = a.color_slice[c], sb = a.color_slice[c + 1];
        for (int s = sa + warp_g; s < sb; s += nwarps_g)
        {
            const int base = a.slice_ptr[s], wl = a.slice_wl[s], row = s * 32 + lane;
            cd acc = csell_row(a.sell_val, a.sell_col, base, wl, wl + a.slice_wu[s], lane, a.z);
            cd v = csub(a.z[row], cmul(cscale(acc, w), a.dinv[row]));
            a.z[row] = v;
            dot = cfma(a.r[row], v, dot);
        }
        if (c > 0) grid.sync();
    }
    if (m == 1)
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) dot = cfma(a.r[i], a.z[i], dot);
    return dot;
}

template <int PRECOND>
__global__ void __launch_bounds__(512, 1) k_cocg(const __grid_constant__ CocgArgs a, const __grid_constant__ AmgArgs g, const __grid_constant__ AmgArgs2 g2)
{
    __shared__ double sm[33];
    __shared__ int s_next;
    __shared__ int s_wk[AMG_SWK_SIZE];
    __shared__ __align__(16) unsigned char sg_raw[PRECOND == XFEMM_PRECOND_AMG ? sizeof(AmgArgs) : 16];
    const AmgArgs *sgp = &g; // multigrid: the block's shared-memory copy of g for the out-of-line passes (amg_kernel_init)
    GridBar grid{cg::this_grid(), PRECOND == XFEMM_PRECOND_AMG ? g.grid_ctr : nullptr, s_wk + AMG_SWK_EPOCH + 1};
    extern __shared__ __align__(128) unsigned char dyn_smem[];
    StageSmem S = stage_carve(dyn_smem); // ring for the matrix stream of the real companion's finest level
    const int lane = threadIdx.x & 31;
    const int nwarps_b = blockDim.x >> 5, warp_b = threadIdx.x >> 5;
    const int warp_g = blockIdx.x * nwarps_b + warp_b, nwarps_g = gridDim.x * nwarps_b;
    const int G = gridDim.x;
    const int s0 = (int)((long long)a.nslices * blockIdx.x / G);
    const int s1 = (int)((long long)a.nslices * (blockIdx.x + 1) / G);
    const int r0 = s0 * 32, r1 = s1 * 32;
    double *partA = a.part, *partB = a.part + 2 * G, *partC = a.part + 4 * G;
    int wk0 = 0, wk1 = 0; // wide rows (circuit rows, air-gap rings) of this block's chunk
    if (a.wide.n_rows) wide_range(a.wide, r0, r1, wk0, wk1);
    if (PRECOND == XFEMM_PRECOND_AMG)
    {
        if (a.tiles)
        {
            S.ntiles = a.tile_ptr[blockIdx.x + 1] - a.tile_ptr[blockIdx.x];
            S.tile = a.tiles + a.tile_ptr[blockIdx.x];
        }
        sgp = &amg_kernel_init<2>(g, sg_raw, s_wk);
    }

    // ||b||^2 ; r = b ; x = 0 unless warm
    double pn = 0.0;
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
    {
        cd bi = a.b[i];
        pn = fma(bi.x, bi.x, fma(bi.y, bi.y, pn));
        a.r[i] = bi;
        if (!a.warm) a.x[i] = make_double2(0., 0.);
    }
    pn = block_sum(pn, sm);
    if (threadIdx.x == 0) partC[blockIdx.x] = pn;
    grid.sync();
    const double normb2 = reduce_partials(partC, G, sm);
    if (normb2 == 0.0)
    {
        // b = 0: the solution is x = 0 (the reference would divide by zero here)
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) a.x[i] = make_double2(0., 0.);
        if (blockIdx.x == 0 && threadIdx.x == 0) { a.mail[0] = 0; a.mail[1] = 0; a.mail[2] = XFEMM_OK; a.mail[3] = 0; }
        return;
    }
    if (a.warm)
    {
        wide_rows_block<cd>(a.wide, a.wide_val, a.wide_acc, wk0, wk1, a.x, a.u);
        if (threadIdx.x == 0) s_next = s0;
        __syncthreads();
        for (int s = cnext_slice(&s_next, lane); s < s1; s = cnext_slice(&s_next, lane))
        {
            const int base = a.slice_ptr[s], w = (a.slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
            cd acc = csell_row(a.sell_val, a.sell_col, base, 0, w, lane, a.x);
            if (wk1 > wk0 && ((a.wide.mask[s] >> lane) & 1u)) acc = w_add(acc, a.u[row]);
            a.r[row] = csub(a.b[row], cfma(a.diag[row], a.x[row], acc));
        }
        grid.sync();
    }
    cd part = cprecond_apply<PRECOND>(a, g, *sgp, g2, grid, r0, r1, warp_g, nwarps_g, lane, &s_next, s_wk, &S);
    part = block_sum_c(part, sm);
    if (threadIdx.x == 0) { partB[blockIdx.x] = part.x; partB[G + blockIdx.x] = part.y; }
    grid.sync();
    cd res = reduce_partials_c(partB, G, sm);
    for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) a.p[i] = a.z[i];
    grid.sync();

    int it = 0, status = XFEMM_OK;
    double er = 0.0;
    while (true)
    {
        // u = A p ; p.u (unconjugated)
        part = make_double2(0., 0.);
        wide_rows_block<cd>(a.wide, a.wide_val, a.wide_acc, wk0, wk1, a.p, a.u);
        if (threadIdx.x == 0) s_next = s0;
        __syncthreads();
        for (int s = cnext_slice(&s_next, lane); s < s1; s = cnext_slice(&s_next, lane))
        {
            const int base = a.slice_ptr[s], w = (a.slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
            cd acc = csell_row(a.sell_val, a.sell_col, base, 0, w, lane, a.p);
            if (wk1 > wk0 && ((a.wide.mask[s] >> lane) & 1u)) acc = w_add(acc, a.u[row]);
            const cd pr = a.p[row];
            acc = cfma(a.diag[row], pr, acc);
            a.u[row] = acc;
            part = cfma(pr, acc, part);
        }
        part = block_sum_c(part, sm);
        if (threadIdx.x == 0) { partA[blockIdx.x] = part.x; partA[G + blockIdx.x] = part.y; }
        grid.sync();
        const cd pAp = reduce_partials_c(partA, G, sm);
        if (pAp.x == 0.0 && pAp.y == 0.0) { status = XFEMM_ERR_DIVERGED; break; }
        const cd del = cdiv(res, pAp);
        // x += del p ; r -= del u ; ||r||^2
        pn = 0.0;
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x)
        {
            a.x[i] = cfma(del, a.p[i], a.x[i]);
            cd ri = csub(a.r[i], cmul(del, a.u[i]));
            a.r[i] = ri;
            pn = fma(ri.x, ri.x, fma(ri.y, ri.y, pn));
        }
        pn = block_sum(pn, sm);
        if (threadIdx.x == 0) partC[blockIdx.x] = pn;
        if (PRECOND != XFEMM_PRECOND_JACOBI) grid.sync();
        part = cprecond_apply<PRECOND>(a, g, *sgp, g2, grid, r0, r1, warp_g, nwarps_g, lane, &s_next, s_wk, &S);
        part = block_sum_c(part, sm);
        if (threadIdx.x == 0) { partB[blockIdx.x] = part.x; partB[G + blockIdx.x] = part.y; }
        grid.sync();
        const cd res_new = reduce_partials_c(partB, G, sm);
        const double rn2 = reduce_partials(partC, G, sm);
        if (res.x == 0.0 && res.y == 0.0) { status = XFEMM_ERR_DIVERGED; break; }
        const cd rho = cdiv(res_new, res);
        res = res_new;
        for (int i = r0 + threadIdx.x; i < r1; i += blockDim.x) a.p[i] = cfma(rho, a.p[i], a.z[i]);
        it++;
        er = sqrt(rn2 / normb2);
        if (!(er == er)) { status = XFEMM_ERR_DIVERGED; break; }
        if (er <= a.precision) break;
        if (it >= a.max_iter) { status = XFEMM_ERR_DIVERGED; break; }
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.mail[0] = it; a.mail[1] = er; a.mail[2] = status; a.mail[3] = normb2; }
}

// ---- boundary conditions -------------------------------------------------------------------
__global__ void kc_fill_int(int n, int *p, int v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void kc_dir_scatter(int count, const int *__restrict__ nodes, int *dir_idx)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dir_idx[nodes[k]] = k;
}
// SetValue (cspars.cpp:482-537): entry {k,i} goes when i is fixed and k is in the scanned band (cband, ccomplex.cuh);
// free row k: b[k] -= A[k,i] x_first(i).
__global__ void kc_dirichlet(int n, int n_nodes, const int *__restrict__ rowptr, const int *__restrict__ col, cd *vals, cd *b,
                             const int *__restrict__ dir_idx, const cd *__restrict__ xf, const cd *__restrict__ xl, int bdw)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int dk = dir_idx[k];
    cd bk = b[k], dkk = make_double2(0., 0.);
    for (int s = rowptr[k]; s < rowptr[k + 1]; s++)
    {
        const int c = col[s];
        if (c == k) { dkk = vals[s]; continue; }
        const int dc = dir_idx[c];
        const bool by_c = dc >= 0 && cband(k, c, bdw, n_nodes);
        const bool by_k = dk >= 0 && cband(c, k, bdw, n_nodes);
        if (by_c || by_k)
        {
            if (by_c && dk < 0) bk = csub(bk, cmul(vals[s], xf[dc]));
            vals[s] = make_double2(0., 0.);
        }
    }
    if (dk >= 0) bk = cmul(dkk, xl[dk]);
    b[k] = bk;
}
__global__ void __launch_bounds__(1024)
kc_periodic(int nlevels, const int *__restrict__ level_ptr, const int *__restrict__ item_ptr, const int *__restrict__ meta,
            const int *__restrict__ items, cd *vals, cd *b, const unsigned char *__restrict__ fill_mask, unsigned char *created)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    volatile cd *vv = vals;
    volatile cd *vb = b;
    for (int lv = 0; lv < nlevels; lv++)
    {
        for (int pr = level_ptr[lv] + warp; pr < level_ptr[lv + 1]; pr += nw)
        {
            const int *m = meta + 6 * pr;
            const double sgn = m[2] ? -1.0 : 1.0;
            for (int it = item_ptr[pr] + lane; it < item_ptr[pr + 1]; it += 32)
            {
                const int *q = items + 4 * it;
                cd v1 = make_double2(vv[q[0]].x, vv[q[0]].y), v2 = make_double2(vv[q[2]].x, vv[q[2]].y);
                if (v1.x != 0.0 || v1.y != 0.0 || v2.x != 0.0 || v2.y != 0.0)
                {
                    cd c = make_double2((v1.x + sgn * v2.x) / 2.0, (v1.y + sgn * v2.y) / 2.0);
                    vv[q[0]].x = c.x; vv[q[0]].y = c.y; vv[q[1]].x = c.x; vv[q[1]].y = c.y;
                    vv[q[2]].x = sgn * c.x; vv[q[2]].y = sgn * c.y; vv[q[3]].x = sgn * c.x; vv[q[3]].y = sgn * c.y;
                    if (fill_mask)
                        for (int t = 0; t < 4; t++)
                            if (fill_mask[q[t]]) created[q[t]] = 1;
                }
            }
            if (lane == 0)
            {
                cd c = make_double2((vv[m[3]].x + vv[m[4]].x) / 2.0, (vv[m[3]].y + vv[m[4]].y) / 2.0);
                vv[m[3]].x = c.x; vv[m[3]].y = c.y; vv[m[4]].x = c.x; vv[m[4]].y = c.y;
                cd cb = make_double2(0.5 * (vb[m[0]].x + sgn * vb[m[1]].x), 0.5 * (vb[m[0]].y + sgn * vb[m[1]].y));
                vb[m[0]].x = cb.x; vb[m[0]].y = cb.y; vb[m[1]].x = sgn * cb.x; vb[m[1]].y = sgn * cb.y;
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}
__global__ void kc_newton_norms(int n, const cd *__restrict__ V, const cd *__restrict__ Vold, double *part)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        cd v = V[i], d = csub(v, Vold[i]);
        a = fma(d.x, d.x, fma(d.y, d.y, a));
        c = fma(v.x, v.x, fma(v.y, v.y, c));
    }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = a; part[gridDim.x + blockIdx.x] = c; }
}
__global__ void kc_relax(int n, cd *V, const cd *__restrict__ Vold, double relax)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) V[i] = make_double2(relax * V[i].x + (1.0 - relax) * Vold[i].x, relax * V[i].y + (1.0 - relax) * Vold[i].y);
}
__global__ void kc_resnorm_partial(int n, const cd *__restrict__ b, const cd *__restrict__ u, double *part)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        cd bb = b[i], d = csub(bb, u[i]);
        a = fma(d.x, d.x, fma(d.y, d.y, a));
        c = fma(bb.x, bb.x, fma(bb.y, bb.y, c));
    }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = a; part[gridDim.x + blockIdx.x] = c; }
}

static inline int cenv_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

static int cbuild_layout(xfemm_clinprob *p, int precond)
{
    if (p->layout_precond == precond) return XFEMM_OK;
    SellLayout &L = p->lay;
    const int sigma = cenv_int("XFEMM_SELL_SIGMA", 1024);
    const int wide_thr = cenv_int("XFEMM_WIDE_ROW", 64);
    cudaStream_t st = p->stream;
    // the uncoloured layout comes from the device builder shared with the real problem (layout_device.cuh)
    const bool on_device = precond != XFEMM_PRECOND_MCSSOR && !cenv_int("XFEMM_HOST_LAYOUT", 0);
    const bool check = on_device && cenv_int("XFEMM_LAYOUT_CHECK", 0);
    if (on_device)
    {
        int rc = build_layout_device(p, sigma, wide_thr, nullptr, nullptr, nullptr);
        if (rc != XFEMM_OK) return rc;
    }
    if (!on_device || check)
    {
        SellLayout Hc;
        SellLayout &H = check ? Hc : L;
        if (precond == XFEMM_PRECOND_MCSSOR)
        {
            std::vector<int> color;
            int nc = greedy_coloring(p->n, p->h_rowptr.data(), p->h_col.data(), cenv_int("XFEMM_MAX_COLORS", 7), color);
            build_sell_layout(p->n, p->h_rowptr.data(), p->h_col.data(), &color, nc, sigma, H, wide_thr);
        }
        else
            build_sell_layout(p->n, p->h_rowptr.data(), p->h_col.data(), nullptr, 0, sigma, H, wide_thr);
        if (check)
        {
            int rc = check_layout_device(p, H);
            if (rc != XFEMM_OK) return rc;
        }
        else
        {
            XB_CUDA(p->i2c.upload(L.i2c.data(), L.i2c.size(), st));
            XB_CUDA(p->c2i.upload(L.c2i.data(), L.c2i.size(), st));
            XB_CUDA(p->slice_ptr.upload(L.slice_ptr.data(), L.slice_ptr.size(), st));
            XB_CUDA(p->slice_wl.upload(L.slice_wl.data(), L.slice_wl.size(), st));
            XB_CUDA(p->slice_wu.upload(L.slice_wu.data(), L.slice_wu.size(), st));
            XB_CUDA(p->sell_col.upload(L.sell_col.data(), L.sell_col.size(), st));
            XB_CUDA(p->sell_src.upload(L.sell_src.data(), L.sell_src.size(), st));
            XB_CUDA(p->diag_src.upload(L.diag_src.data(), L.diag_src.size(), st));
            XB_CUDA(p->color_slice.upload(L.color_slice.data(), L.color_slice.size(), st));
            XB_CUDA(p->wide.upload(L, st));
        }
    }
    XB_CUDA(p->sell_val.alloc(L.sell_len));
    for (auto *v : {&p->diag, &p->dinv, &p->x, &p->r, &p->p, &p->u, &p->z, &p->w, &p->bint})
    {
        XB_CUDA(v->alloc(L.n_pad));
        XB_CUDA(v->zero(st));
    }
    XB_CUDA(cudaStreamSynchronize(st));
    std::vector<int>().swap(L.sell_col);
    std::vector<int>().swap(L.sell_src);
    std::vector<int>().swap(L.diag_src);
    std::vector<int>().swap(L.wide_col);
    std::vector<int>().swap(L.wide_src);
    p->layout_precond = precond;
    return XFEMM_OK;
}

static int crefresh_values(xfemm_clinprob *p, int *singular_row)
{
    const SellLayout &L = p->lay;
    cudaStream_t st = p->stream;
    int *d_flag = (int *)(p->d_mail.p + 8);
    int big = 0x7fffffff;
    XB_CUDA(cudaMemcpyAsync(d_flag, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    int blocks = (int)std::max<long long>(1, std::min<long long>((L.sell_len + 255) / 256, (long long)p->num_sms * 16));
    ++g_launch, kc_sell_fill<<<blocks, 256, 0, st>>>(L.sell_len, p->sell_src.p, p->vals.p, p->sell_val.p);
    p->wide.fill(p->vals.p, st);
    ++g_launch, kc_diag_fill<<<(L.n_pad + 255) / 256, 256, 0, st>>>(L.n_pad, p->diag_src.p, p->vals.p, p->diag.p, p->dinv.p, d_flag);
    int flag = 0;
    XB_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    *singular_row = (flag == big) ? -1 : flag;
    return XFEMM_OK;
}

static CocgArgs cmake_args(xfemm_clinprob *p, double precision, int warm, int max_iter)
{
    CocgArgs a;
    a.n_pad = p->lay.n_pad; a.nslices = p->lay.nslices; a.ncolors = p->lay.ncolors;
    a.slice_ptr = p->slice_ptr.p; a.slice_wl = p->slice_wl.p; a.slice_wu = p->slice_wu.p; a.sell_col = p->sell_col.p;
    a.color_slice = p->color_slice.p;
    a.sell_val = p->sell_val.p; a.diag = p->diag.p; a.dinv = p->dinv.p; a.b = p->bint.p;
    a.x = p->x.p; a.r = p->r.p; a.p = p->p.p; a.u = p->u.p; a.z = p->z.p;
    a.part = p->partials.p;
    a.precision = precision;
    const char *om = getenv("XFEMM_SSOR_OMEGA");
    a.omega = om ? atof(om) : 1.2;
    a.max_iter = max_iter; a.warm = warm;
    a.mail = p->d_mail.p;
    a.wide = p->wide.view(); a.wide_val = p->wide.val.p; a.wide_acc = p->wide.acc.p;
    a.i2c = p->i2c.p; a.n_nodes = p->n_nodes; a.tiles = nullptr; a.tile_ptr = nullptr; a.prec_sign = 1.0;
    return a;
}

// Hermitian inner product conj(x).y on the device (two-stage, deterministic)
static int cconjdot(xfemm_clinprob *p, const cd *x, const cd *y, std::complex<double> *out)
{
    cudaStream_t st = p->stream;
    const int n = p->lay.n_pad;
    int nb = std::max(1, std::min((n + 255) / 256, 512));
    if (p->max_grid < nb) { XB_CUDA(p->partials.alloc((size_t)8 * nb)); p->max_grid = nb; }
    ++g_launch, kc_conjdot_partial<<<nb, 256, 0, st>>>(n, x, y, p->partials.p);
    ++g_launch, kc_reduce2_final<<<1, 256, 0, st>>>(nb, p->partials.p, p->d_mail.p + 12);
    XB_CUDA(cudaMemcpyAsync(p->h_mail + 12, p->d_mail.p + 12, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    *out = std::complex<double>(p->h_mail[12], p->h_mail[13]);
    return XFEMM_OK;
}

} // namespace xb

using namespace xb;

extern "C" {

int xfemm_clinprob_create(int n, int n_nodes, const int *rowptr, const int *colidx, xfemm_clinprob **out)
{
    XB_ARG(n > 0 && n_nodes >= 0 && n_nodes <= n && rowptr && colidx && out, "n, rowptr, colidx, out");
    if (xfemm_device_count() <= 0) { set_error("no CUDA device available: the xfemm_b200 backend has no CPU fallback"); return XFEMM_ERR_CUDA; }
    const long long nnz = rowptr[n];
    for (int i = 0; i < n; i++)
    {
        bool diag = false;
        for (int s = rowptr[i]; s < rowptr[i + 1]; s++)
        {
            int c = colidx[s];
            if (c < 0 || c >= n || (s > rowptr[i] && colidx[s - 1] >= c)) { set_error("row %d: columns must be sorted, unique and in range", i); return XFEMM_ERR_ARG; }
            if (c == i) diag = true;
        }
        if (!diag) { set_error("row %d has no diagonal entry", i); return XFEMM_ERR_ARG; }
    }
    xfemm_clinprob *p = new xfemm_clinprob;
    p->n = n; p->n_nodes = n_nodes; p->nnz = nnz;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    {
        delete p;
        set_error("cannot query the device");
        return XFEMM_ERR_CUDA;
    }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) { delete p; set_error("stream create failed"); return XFEMM_ERR_CUDA; }
    p->device = dev;
    p->h_rowptr.assign(rowptr, rowptr + n + 1);
    p->h_col.assign(colidx, colidx + nnz);
    auto body = [&]() -> int {
        XB_CUDA(p->rowptr.upload(rowptr, n + 1, p->stream));
        XB_CUDA(p->col.upload(colidx, nnz, p->stream));
        XB_CUDA(p->vals.alloc(nnz)); XB_CUDA(p->vals.zero(p->stream));
        for (auto *v : {&p->b, &p->V, &p->Vold}) { XB_CUDA(v->alloc(n)); XB_CUDA(v->zero(p->stream)); }
        XB_CUDA(p->d_mail.alloc(16)); XB_CUDA(p->d_mail.zero(p->stream));
        XB_CUDA(p->dir_idx.alloc(n));
        ++g_launch, kc_fill_int<<<(n + 255) / 256, 256, 0, p->stream>>>(n, p->dir_idx.p, -1);
        XB_CUDA(cudaStreamSynchronize(p->stream));
        return XFEMM_OK;
    };
    int rc = body();
    if (rc != XFEMM_OK) { xfemm_clinprob_destroy(p); return rc; }
    *out = p;
    return XFEMM_OK;
}

int xfemm_clinprob_destroy(xfemm_clinprob *p)
{
    if (!p) return XFEMM_OK;
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->re_prec) { xfemm_linprob_destroy(p->re_prec); p->re_prec = nullptr; }
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
    return XFEMM_OK;
}
int xfemm_clinprob_n(const xfemm_clinprob *p) { return p ? p->n : 0; }
// levels of the multigrid hierarchy of the real companion problem (XFEMM_PRECOND_AMG), as xfemm_linprob_amg_levels
int xfemm_clinprob_amg_levels(xfemm_clinprob *p, int max_levels, int *n_rows, long long *nnz)
{
    return (p && p->re_prec) ? xfemm_linprob_amg_levels(p->re_prec, max_levels, n_rows, nnz) : 0;
}
long long xfemm_clinprob_nnz(const xfemm_clinprob *p) { return p ? p->nnz : 0; }
int xfemm_clinprob_wipe(xfemm_clinprob *p)
{
    XB_ARG(p, "null handle");
    XB_CUDA(p->vals.zero(p->stream));
    XB_CUDA(p->b.zero(p->stream));
    if (p->newton)
        for (auto *v : {&p->vals_h, &p->vals_s, &p->vals_a}) XB_CUDA(v->zero(p->stream));
    return XFEMM_OK;
}
#define XC_IN(buf, src, cnt)                                                                              \
    XB_ARG(p && (src), "null pointer");                                                                   \
    xb::g_h2d_bytes += (long long)(cnt) * 16;                                                             \
    XB_CUDA(cudaMemcpyAsync((buf).p, (src), (size_t)(cnt) * 16, cudaMemcpyHostToDevice, p->stream));      \
    XB_CUDA(cudaStreamSynchronize(p->stream));                                                            \
    return XFEMM_OK
#define XC_OUT(buf, dst, cnt)                                                                             \
    XB_ARG(p && (dst), "null pointer");                                                                   \
    xb::g_d2h_bytes += (long long)(cnt) * 16;                                                             \
    XB_CUDA(cudaMemcpyAsync((dst), (buf).p, (size_t)(cnt) * 16, cudaMemcpyDeviceToHost, p->stream));      \
    XB_CUDA(cudaStreamSynchronize(p->stream));                                                            \
    return XFEMM_OK
int xfemm_clinprob_set_values(xfemm_clinprob *p, const double *v) { XC_IN(p->vals, v, p->nnz); }
int xfemm_clinprob_get_values(xfemm_clinprob *p, double *v) { XC_OUT(p->vals, v, p->nnz); }
int xfemm_clinprob_set_rhs(xfemm_clinprob *p, const double *b) { XC_IN(p->b, b, p->n); }
int xfemm_clinprob_get_rhs(xfemm_clinprob *p, double *b) { XC_OUT(p->b, b, p->n); }
int xfemm_clinprob_set_solution(xfemm_clinprob *p, const double *V) { XC_IN(p->V, V, p->n); }
int xfemm_clinprob_get_solution(xfemm_clinprob *p, double *V) { XC_OUT(p->V, V, p->n); }

int xfemm_clinprob_dirichlet(xfemm_clinprob *p, int count, const int *nodes, const double *x_first, const double *x_last, int bdw)
{
    XB_ARG(p, "null handle");
    if (count <= 0) return XFEMM_OK;
    XB_ARG(nodes && x_first && x_last, "null arrays");
    for (int k = 0; k < count; k++) XB_ARG(nodes[k] >= 0 && nodes[k] < p->n, "Dirichlet node out of range");
    cudaStream_t st = p->stream;
    XB_CUDA(p->dir_nodes.upload(nodes, count, st));
    XB_CUDA(p->dir_first.upload((const cd *)x_first, count, st));
    XB_CUDA(p->dir_last.upload((const cd *)x_last, count, st));
    ++g_launch, kc_dir_scatter<<<(count + 255) / 256, 256, 0, st>>>(count, p->dir_nodes.p, p->dir_idx.p);
    ++g_launch, kc_dirichlet<<<(p->n + 255) / 256, 256, 0, st>>>(p->n, p->n_nodes, p->rowptr.p, p->col.p, p->vals.p, p->b.p,
                                                                  p->dir_idx.p, p->dir_first.p, p->dir_last.p, bdw);
    if (p->newton) { int rc = cnewton_dirichlet(p, bdw); if (rc != XFEMM_OK) return rc; }
    ++g_launch, kc_fill_int<<<(p->n + 255) / 256, 256, 0, st>>>(p->n, p->dir_idx.p, -1);
    XB_CUDA(cudaGetLastError());
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

int xfemm_clinprob_periodic(xfemm_clinprob *p, int npairs, const int *ni, const int *nj, const int *type)
{
    XB_ARG(p, "null handle");
    if (npairs <= 0) return XFEMM_OK;
    XB_ARG(ni && nj && type, "null arrays");
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](int v) { h ^= (uint32_t)v; h *= 1099511628211ull; };
    for (int k = 0; k < npairs; k++) { mix(ni[k]); mix(nj[k]); mix(type[k]); }
    if (h != p->pbc_hash || npairs != p->pbc_pairs)
    {
        std::vector<int> level_ptr, item_ptr, meta, items;
        int err[3] = {0, 0, 0};
        int nl = build_pbc_plan(p->n, p->h_rowptr.data(), p->h_col.data(), npairs, ni, nj, type, level_ptr, item_ptr, meta, items, err);
        if (nl < 0)
        {
            set_error("periodic pair (%d,%d): pattern lacks the fill for row %d (or the pair is out of range)", err[0], err[1], err[2]);
            return XFEMM_ERR_ARG;
        }
        cudaStream_t st = p->stream;
        XB_CUDA(p->pbc_level_ptr.upload(level_ptr.data(), level_ptr.size(), st));
        XB_CUDA(p->pbc_item_ptr.upload(item_ptr.data(), item_ptr.size(), st));
        XB_CUDA(p->pbc_meta.upload(meta.data(), meta.size(), st));
        XB_CUDA(p->pbc_items.upload(items.data(), items.size(), st));
        XB_CUDA(cudaStreamSynchronize(st));
        p->pbc_hash = h; p->pbc_pairs = npairs; p->pbc_levels = nl;
    }
    ++g_launch, kc_periodic<<<1, 1024, 0, p->stream>>>(p->pbc_levels, p->pbc_level_ptr.p, p->pbc_item_ptr.p, p->pbc_meta.p,
                                                       p->pbc_items.p, p->vals.p, p->b.p,
                                                       p->fill_mask.n ? p->fill_mask.p : nullptr, p->created.p);
    if (p->newton) { int rc = cnewton_periodic(p, npairs, ni, nj); if (rc != XFEMM_OK) return rc; }
    XB_CUDA(cudaGetLastError());
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}

int xfemm_clinprob_set_fill_mask(xfemm_clinprob *p, const unsigned char *fill)
{
    XB_ARG(p && fill, "null pointer");
    XB_CUDA(p->fill_mask.upload(fill, (size_t)p->nnz, p->stream));
    XB_CUDA(p->created.alloc((size_t)p->nnz));
    XB_CUDA(p->created.zero(p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}
int xfemm_clinprob_get_exists_mask(xfemm_clinprob *p, unsigned char *exists)
{
    XB_ARG(p && exists, "null pointer");
    if (!p->fill_mask.n) { memset(exists, 1, (size_t)p->nnz); return XFEMM_OK; }
    std::vector<unsigned char> f((size_t)p->nnz), c((size_t)p->nnz);
    XB_CUDA(cudaMemcpyAsync(f.data(), p->fill_mask.p, f.size(), cudaMemcpyDeviceToHost, p->stream));
    XB_CUDA(cudaMemcpyAsync(c.data(), p->created.p, c.size(), cudaMemcpyDeviceToHost, p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    for (size_t s = 0; s < f.size(); s++) exists[s] = (!f[s] || c[s]) ? 1 : 0;
    return XFEMM_OK;
}

// Real companion problem of XFEMM_PRECOND_AMG (see kc_fill_real): the pattern between mesh nodes plus identity rows
// for the circuit unknowns, in the internal row order of this problem (the layout must exist).  Created once.
static int censure_companion(xfemm_clinprob *p)
{
    if (p->re_prec) return XFEMM_OK;
    const int n = p->n, nn = p->n_nodes;
    std::vector<int> rp(n + 1, 0);
    // row lengths, offsets, then the filtered columns and their source slots in parallel
    parallel_for(n, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++)
        {
            int len = 0;
            for (int s = p->h_rowptr[i]; s < p->h_rowptr[i + 1]; s++)
            {
                const int j = p->h_col[s];
                if (i < nn ? j < nn : j == i) len++;
            }
            rp[i + 1] = len;
        }
    });
    for (int i = 0; i < n; i++) rp[i + 1] += rp[i];
    std::vector<int> rc_((size_t)rp[n]), src((size_t)rp[n]);
    parallel_for(n, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++)
        {
            int w = rp[i];
            for (int s = p->h_rowptr[i]; s < p->h_rowptr[i + 1]; s++)
            {
                const int j = p->h_col[s];
                if (i < nn ? j < nn : j == i) { rc_[w] = j; src[w] = s; w++; }
            }
        }
    });
    int rc = xfemm_linprob_create(n, rp.data(), rc_.data(), &p->re_prec);
    if (rc != XFEMM_OK) return rc;
    for (int r : p->lay.i2c) if (r >= 0) p->re_prec->forced_order.push_back(r);
    XB_CUDA(p->re_src.upload(src.data(), src.size(), p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}

// Everything of the solver that depends on the sparsity pattern alone — the SELL layout and, for
// XFEMM_PRECOND_AMG, the real companion problem with its layout and multigrid hierarchy — built ahead of the first
// solve.  FSolver calls it from a host thread while the first assembly runs (as xfemm_linprob_prebuild).
int xfemm_clinprob_prebuild(xfemm_clinprob *p, int precond)
{
    XB_ARG(p, "null handle");
    XB_ARG(precond == XFEMM_PRECOND_JACOBI || precond == XFEMM_PRECOND_MCSSOR || precond == XFEMM_PRECOND_AMG, "unknown preconditioner");
    XB_CUDA(cudaSetDevice(p->device)); // a fresh host thread starts on device 0
    int rc = cbuild_layout(p, precond == XFEMM_PRECOND_AMG ? XFEMM_PRECOND_JACOBI : precond);
    if (rc != XFEMM_OK || precond != XFEMM_PRECOND_AMG) return rc;
    rc = censure_companion(p);
    if (rc != XFEMM_OK) return rc;
    return xfemm_linprob_prebuild(p->re_prec, XFEMM_PRECOND_AMG);
}

// PBCGSolveMod (cspars.cpp:1062-1081): cold start → 3 CG steps on the normal equations of the
// Jacobi-preconditioned system to leave the pathological starting points of COCG, then COCG.
int xfemm_clinprob_cocg(xfemm_clinprob *p, double precision, int warm, int precond, int max_iter, xfemm_solve_info *info)
{
    XB_ARG(p, "null handle");
    XB_ARG(precond == XFEMM_PRECOND_JACOBI || precond == XFEMM_PRECOND_MCSSOR || precond == XFEMM_PRECOND_AMG, "unknown preconditioner");
    cudaStream_t st = p->stream;
    cudaEvent_t e0, e1, e2;
    XB_CUDA(cudaEventCreate(&e0)); XB_CUDA(cudaEventCreate(&e1)); XB_CUDA(cudaEventCreate(&e2));
    XB_CUDA(cudaEventRecord(e0, st));
    int rc = cbuild_layout(p, precond == XFEMM_PRECOND_AMG ? XFEMM_PRECOND_JACOBI : precond); // AMG: the uncoloured layout
    if (rc != XFEMM_OK) return rc;
    // values unchanged since the previous call (the inner solves of KludgeSolve): keep everything prepared then
    const bool reuse = p->keep_prepared && p->prepared_precond == precond && (precond != XFEMM_PRECOND_AMG || p->re_prec);
    AmgArgs g;
    AmgArgs2 g2;
    double prec_sign = 1.0;
    if (precond == XFEMM_PRECOND_AMG)
    {
        rc = censure_companion(p);
        if (rc != XFEMM_OK) return rc;
        xfemm_linprob *R = p->re_prec;
        if (reuse) prec_sign = p->prepared_sign;
        else
        {
            int *d_cnt = (int *)(p->d_mail.p + 12);
            int h_cnt[4] = {0, 0, 0, 0};
            XB_CUDA(cudaMemsetAsync(d_cnt, 0, 4 * sizeof(int), st));
            ++g_launch, kc_diag_sign<<<(p->n_nodes + 255) / 256, 256, 0, st>>>(p->n_nodes, p->rowptr.p, p->col.p, p->vals.p, d_cnt);
            XB_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
            XB_CUDA(cudaStreamSynchronize(st));
            prec_sign = h_cnt[1] > h_cnt[0] ? -1.0 : 1.0;
            double im_sign = h_cnt[3] > h_cnt[2] ? -1.0 : 1.0;
            if (const char *e = getenv("XFEMM_COMPLEX_AMG_REAL_ONLY")) if (atoi(e)) im_sign = 0.0;
            ++g_launch, kc_fill_real<<<(p->n + 255) / 256, 256, 0, st>>>(p->n, p->n_nodes, R->rowptr.p, R->col.p, p->re_src.p, p->vals.p,
                                                                          R->vals.p, prec_sign, im_sign);
            XB_CUDA(cudaStreamSynchronize(st)); // the companion works on its own stream
            rc = linprob_prepare(R, XFEMM_PRECOND_AMG);
            if (rc != XFEMM_OK) return rc;
            XB_CUDA(cudaStreamSynchronize(R->stream));
            if (R->lay.n_pad != p->lay.n_pad || R->lay.i2c != p->lay.i2c)
            {
                set_error("internal row orders of the complex problem and its real companion differ");
                return XFEMM_ERR_ARG;
            }
        }
        amg_fill_args(R, *R->amg, g);
        g.scatter_restrict = 0; // (this kernel runs the ordered restriction pass)
        amg_fill_args2(R, *R->amg, g2);
        XB_CUDA(cudaStreamSynchronize(R->stream)); // the barrier counters are zeroed on the companion's stream
    }
    int singular = -1;
    if (!reuse)
    {
        p->prepared_precond = -1;
        rc = crefresh_values(p, &singular);
        if (rc != XFEMM_OK) return rc;
        if (singular < 0) { p->prepared_precond = precond; p->prepared_sign = prec_sign; }
    }
    xfemm_solve_info local;
    memset(&local, 0, sizeof local);
    local.precond = precond;
    local.colors = p->lay.ncolors;
    if (singular >= 0)
    {
        set_error("singular flag tripped.");
        local.status = XFEMM_ERR_SINGULAR;
        if (info) *info = local;
        return XFEMM_ERR_SINGULAR;
    }
    const int n_pad = p->lay.n_pad, nb = (n_pad + 255) / 256;
    const int sb = std::max(1, std::min((p->lay.nslices + 7) / 8, p->num_sms * 8));
    ++g_launch, kc_to_internal<<<nb, 256, 0, st>>>(n_pad, p->i2c.p, p->b.p, p->bint.p);
    XB_CUDA(cudaEventRecord(e1, st));
    if (warm)
        ++g_launch, kc_to_internal<<<nb, 256, 0, st>>>(n_pad, p->i2c.p, p->V.p, p->x.p);
    else if (cenv_int("XFEMM_COCG_NOSTART", 0) == 0)
    {
        // ---- PCGSQStart with M = D: operator T y = conj(A D^-1 conj(D^-1 A y)), rhs g = conj(A D^-1 conj(D^-1 b))
        auto spmv = [&](const cd *in, cd *out) {
            p->wide.pre(in, out, st);
            ++g_launch, kc_spmv<<<sb, 256, 0, st>>>(p->lay.nslices, p->slice_ptr.p, p->sell_col.p, p->sell_val.p, p->diag.p, in, out,
                                                    p->wide.n_rows ? p->wide.mask.p : nullptr);
        };
        auto applyT = [&](const cd *in, cd *out, cd *tmp) {
            spmv(in, tmp);                                                         // tmp = A in
            ++g_launch, kc_mul<<<nb, 256, 0, st>>>(n_pad, tmp, p->dinv.p, out);      // out = D^-1 tmp
            ++g_launch, kc_conj<<<nb, 256, 0, st>>>(n_pad, out);
            ++g_launch, kc_mul<<<nb, 256, 0, st>>>(n_pad, out, p->dinv.p, tmp);      // tmp = D^-1 conj(..)
            spmv(tmp, out);
            ++g_launch, kc_conj<<<nb, 256, 0, st>>>(n_pad, out);
        };
        cd *X = p->x.p, *R = p->r.p, *P = p->p.p, *U = p->u.p, *Z = p->z.p, *W = p->w.p;
        // g → R
        ++g_launch, kc_mul<<<nb, 256, 0, st>>>(n_pad, p->bint.p, p->dinv.p, Z);
        ++g_launch, kc_conj<<<nb, 256, 0, st>>>(n_pad, Z);
        ++g_launch, kc_mul<<<nb, 256, 0, st>>>(n_pad, Z, p->dinv.p, W);
        spmv(W, R);
        ++g_launch, kc_conj<<<nb, 256, 0, st>>>(n_pad, R);
        XB_CUDA(cudaMemsetAsync(X, 0, (size_t)n_pad * sizeof(cd), st));
        XB_CUDA(cudaMemcpyAsync(P, R, (size_t)n_pad * sizeof(cd), cudaMemcpyDeviceToDevice, st));
        std::complex<double> res, pAp, res_new;
        rc = cconjdot(p, R, R, &res);
        if (rc != XFEMM_OK) return rc;
        for (int k = 0; k < 3 && std::abs(res) > 0; k++)
        {
            applyT(P, U, W);
            rc = cconjdot(p, P, U, &pAp);
            if (rc != XFEMM_OK) return rc;
            if (std::abs(pAp) == 0) break;
            std::complex<double> del = res / pAp;
            ++g_launch, kc_axpy<<<nb, 256, 0, st>>>(n_pad, make_double2(del.real(), del.imag()), P, X);
            ++g_launch, kc_axpy<<<nb, 256, 0, st>>>(n_pad, make_double2(-del.real(), -del.imag()), U, R);
            rc = cconjdot(p, R, R, &res_new);
            if (rc != XFEMM_OK) return rc;
            std::complex<double> rho = res_new / res;
            res = res_new;
            ++g_launch, kc_xpay<<<nb, 256, 0, st>>>(n_pad, make_double2(rho.real(), rho.imag()), R, P);
        }
        warm = 1; // COCG continues from the seeded x (the reference calls PBCGSolve(2))
    }
    if (max_iter <= 0) max_iter = cenv_int("XFEMM_MAX_ITER", 200000);
    CocgArgs a = cmake_args(p, precision, warm ? 1 : 0, max_iter);
    a.prec_sign = prec_sign;
    void *kargs[] = {&a, &g, &g2};
    int per_sm = 0, grid = 0;
    void *kern = precond == XFEMM_PRECOND_JACOBI ? (void *)k_cocg<XFEMM_PRECOND_JACOBI>
                 : precond == XFEMM_PRECOND_AMG  ? (void *)k_cocg<XFEMM_PRECOND_AMG>
                                                 : (void *)k_cocg<XFEMM_PRECOND_MCSSOR>;
    const size_t dyn_smem = precond == XFEMM_PRECOND_AMG ? STAGE_SMEM_BYTES + AMG_TASK_SMEM_BYTES : 0;
    g.task_smem_offset = (int)STAGE_SMEM_BYTES;
    if (dyn_smem) XB_CUDA(cudaFuncSetAttribute((void (*)(CocgArgs, AmgArgs, AmgArgs2))kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    XB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (void (*)(CocgArgs, AmgArgs, AmgArgs2))kern, 512, dyn_smem));
    if (per_sm < 1) { set_error("cooperative kernel does not fit on an SM"); return XFEMM_ERR_CUDA; }
    per_sm = std::min(per_sm, cenv_int("XFEMM_PCG_BLOCKS_PER_SM", 2));
    grid = std::max(1, std::min(p->num_sms * per_sm, (p->lay.nslices + 15) / 16));
    if (p->max_grid < grid) { XB_CUDA(p->partials.alloc((size_t)8 * grid)); p->max_grid = grid; }
    a.part = p->partials.p;
    if (precond == XFEMM_PRECOND_AMG) g.team_blocks = std::min(g.team_blocks, grid); // the team lives inside this launch
    if (precond == XFEMM_PRECOND_AMG && linprob_ensure_tiles(p->re_prec, grid))
    {
        XB_CUDA(cudaStreamSynchronize(p->re_prec->stream));
        a.tiles = p->re_prec->tiles.p; a.tile_ptr = p->re_prec->tile_ptr.p;
    }
    ++g_launch;
    XB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(512), kargs, dyn_smem, st));
    XB_CUDA(cudaEventRecord(e2, st));
    XB_CUDA(cudaMemcpyAsync(p->h_mail, p->d_mail.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    local.iterations = (int)p->h_mail[0];
    local.er = p->h_mail[1];
    local.status = (int)p->h_mail[2];
    ++g_launch, kc_from_internal<<<nb, 256, 0, st>>>(n_pad, p->i2c.p, p->x.p, p->V.p);
    // true relative residual
    p->wide.pre(p->x.p, p->u.p, st);
    ++g_launch, kc_spmv<<<sb, 256, 0, st>>>(p->lay.nslices, p->slice_ptr.p, p->sell_col.p, p->sell_val.p, p->diag.p, p->x.p, p->u.p,
                                            p->wide.n_rows ? p->wide.mask.p : nullptr);
    int nb2 = std::max(1, std::min(nb, 512));
    if (p->max_grid < nb2) { XB_CUDA(p->partials.alloc((size_t)8 * nb2)); p->max_grid = nb2; }
    ++g_launch, kc_resnorm_partial<<<nb2, 256, 0, st>>>(n_pad, p->bint.p, p->u.p, p->partials.p);
    ++g_launch, kc_reduce2_final<<<1, 256, 0, st>>>(nb2, p->partials.p, p->d_mail.p + 10);
    XB_CUDA(cudaMemcpyAsync(p->h_mail + 10, p->d_mail.p + 10, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    local.relres = p->h_mail[11] > 0 ? sqrt(p->h_mail[10] / p->h_mail[11]) : 0.0;
    float ms_setup = 0, ms_solve = 0;
    cudaEventElapsedTime(&ms_setup, e0, e1);
    cudaEventElapsedTime(&ms_solve, e1, e2);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    local.setup_ms = ms_setup;
    local.solve_ms = ms_solve;
    const double nn = (double)p->n, nz = (double)p->nnz; // complex doubles every value/vector byte
    local.bytes_per_iter = 20.0 * nz + 244.0 * nn + (precond == XFEMM_PRECOND_MCSSOR ? 20.0 * nz + 80.0 * nn : 0.0);
    if (info) *info = local;
    if (local.status != XFEMM_OK)
    {
        set_error("COCG did not converge: %d iterations, er=%g", local.iterations, local.er);
        return local.status;
    }
    return XFEMM_OK;
}

int xfemm_clinprob_save_solution(xfemm_clinprob *p)
{
    XB_ARG(p, "null handle");
    XB_CUDA(cudaMemcpyAsync(p->Vold.p, p->V.p, (size_t)p->n * sizeof(cd), cudaMemcpyDeviceToDevice, p->stream));
    return XFEMM_OK;
}
int xfemm_clinprob_newton_norms(xfemm_clinprob *p, int count, double *dv2, double *v2)
{
    XB_ARG(p && dv2 && v2 && count >= 0 && count <= p->n, "arguments");
    cudaStream_t st = p->stream;
    int nb = std::max(1, std::min((count + 255) / 256, 512));
    if (p->max_grid < nb) { XB_CUDA(p->partials.alloc((size_t)8 * nb)); p->max_grid = nb; }
    ++g_launch, kc_newton_norms<<<nb, 256, 0, st>>>(count, p->V.p, p->Vold.p, p->partials.p);
    ++g_launch, kc_reduce2_final<<<1, 256, 0, st>>>(nb, p->partials.p, p->d_mail.p + 12);
    XB_CUDA(cudaMemcpyAsync(p->h_mail + 12, p->d_mail.p + 12, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    *dv2 = p->h_mail[12];
    *v2 = p->h_mail[13];
    return XFEMM_OK;
}
int xfemm_clinprob_relax(xfemm_clinprob *p, double relax)
{
    XB_ARG(p, "null handle");
    ++g_launch, kc_relax<<<(p->n + 255) / 256, 256, 0, p->stream>>>(p->n, p->V.p, p->Vold.p, relax);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

} // extern "C"
