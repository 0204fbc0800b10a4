// xfemm_dist_* — row-block partitioned PCG for one mesh spread over several GPUs
// (BASELINE config 5b).  Every rank assembles and boundary-conditions the global
// system on its own GPU (assembly is <1 % of a solve) and keeps a contiguous block of rows
// [r0,r1) of the Cuthill–McKee ordered matrix: owned rows + the halo columns they touch.  One
// PCG iteration = halo exchange of the search direction (packed by pack_dev, exchanged by the
// caller over NVLink with torch.distributed/NCCL, scattered by unpack_dev), a local SELL-32 SpMV
// fused with the p.Ap partial, and two fused vector kernels; the two scalars of an iteration are
// all-reduced by the caller and stay on the device (no host synchronisation inside the loop).
// Preconditioner: Jacobi (block-local by construction).
#include "linprob.cuh"
#include <algorithm>
#include <vector>

struct xfemm_dist
{
    xfemm_linprob *glob = nullptr; // not owned
    xfemm_linprob *loc = nullptr;  // owned: square (n_own + n_halo), halo rows are identity rows
    int r0 = 0, r1 = 0, n_own = 0, n_halo = 0;
    std::vector<int> halo_cols;    // global column ids, sorted
    xb::DevBuf<int> slot_map;      // local CSR slot -> global CSR slot (-1: identity diagonal of a halo row)
    xb::DevBuf<int> send_int, halo_int; // internal row indices
    int n_send = 0;
    xb::DevBuf<unsigned char> owned; // per internal row: 1 = owned row
    xb::DevBuf<double> partials;
    int nblk = 0;
    int precond = XFEMM_PRECOND_JACOBI; // or XFEMM_PRECOND_AMG: V-cycle of the rank's own diagonal block
};

namespace xb {

// res / pAp and res_new / res with the guards of the single-GPU kernel (k_pcg stops on pAp <= 0 and on NaN): an exactly
// zero residual (diagonal or fully fixed system) or a non-positive curvature gives a zero step, which freezes the
// iterate; the host sees the non-finite or stalled residual and reports XFEMM_ERR_DIVERGED / convergence.
__device__ __forceinline__ double guarded_ratio(double num, double den) { return (den > 0.0 && num == num) ? num / den : 0.0; }

__global__ void kd_copy_block(long long nnz_loc, const int *__restrict__ slot_map, const double *__restrict__ gvals, double *lvals)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nnz_loc) return;
    int s = slot_map[i];
    lvals[i] = s >= 0 ? gvals[s] : 1.0;
}
__global__ void kd_copy_rhs(int n_loc, int n_own, int r0, const double *__restrict__ gb, double *lb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_loc) lb[i] = i < n_own ? gb[r0 + i] : 0.0;
}
__global__ void kd_pack(int n, const int *__restrict__ idx, const double *__restrict__ v, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v[idx[i]];
}
__global__ void kd_unpack(int n, const int *__restrict__ idx, const double *__restrict__ in, double *v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[idx[i]] = in[i];
}
// block partial sums → out[slot] via a second tiny kernel (deterministic)
__global__ void kd_reduce(int nblk, const double *__restrict__ part, double *out, int nout)
{
    __shared__ double sm[33];
    for (int q = 0; q < nout; q++)
    {
        double a = 0;
        for (int i = threadIdx.x; i < nblk; i += blockDim.x) a += part[(size_t)q * nblk + i];
        a = block_sum(a, sm);
        if (threadIdx.x == 0) out[q] = a;
    }
}
// x = 0 ; r = b ; z = dinv r ; p = z ; partials: [b.dinv b] (= r.z as x = 0)
__global__ void kd_init(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ b,
                        const double *__restrict__ dinv, double *x, double *r, double *p, double *part)
{
    __shared__ double sm[33];
    double a = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
    {
        const double bi = owned[i] ? b[i] : 0.0, zi = bi * dinv[i];
        x[i] = 0.0; r[i] = bi; p[i] = zi;
        a = fma(bi, zi, a);
    }
    a = block_sum(a, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}
// u = A p over the local rows, partial of p.u over owned rows
__global__ void __launch_bounds__(256)
kd_spmv_dot(int nslices, const int *__restrict__ slice_ptr, const int *__restrict__ sell_col,
            const double *__restrict__ sell_val, const double *__restrict__ diag,
            const unsigned char *__restrict__ owned, const double *p, double *u, double *part,
            const unsigned *__restrict__ wide_mask)
{
    __shared__ double sm[33];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    double a = 0;
    for (int s = warp; s < nslices; s += nwarps)
    {
        const int base = slice_ptr[s], w = (slice_ptr[s + 1] - base) >> 5, row = s * 32 + lane;
        double acc = 0.0;
        const double *v = sell_val + base + lane;
        const int *c = sell_col + base + lane;
#pragma unroll 8
        for (int k = 0; k < w; k++) acc = fma(__ldcs(v + k * 32), p[__ldcs(c + k * 32)], acc);
        if (wide_mask && ((wide_mask[s] >> lane) & 1u)) acc += u[row]; // left there by k_wide_pre
        const double pr = p[row];
        acc = fma(diag[row], pr, acc);
        u[row] = acc;
        if (owned[row]) a = fma(pr, acc, a);
    }
    a = block_sum(a, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}
// alpha = g[0]/g[1] (res / pAp, both already all-reduced) ; x += alpha p ; r -= alpha u ; partial r.dinv r
__global__ void kd_update(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ g,
                          const double *__restrict__ dinv, const double *__restrict__ p, const double *__restrict__ u,
                          double *x, double *r, double *part)
{
    __shared__ double sm[33];
    const double alpha = guarded_ratio(g[0], g[1]);
    double a = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
    {
        if (!owned[i]) continue;
        x[i] = fma(alpha, p[i], x[i]);
        const double ri = fma(-alpha, u[i], r[i]);
        r[i] = ri;
        a = fma(ri * dinv[i], ri, a);
    }
    a = block_sum(a, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}
// beta = g[2]/g[0] (res_new / res) ; p = dinv r + beta p
__global__ void kd_direction(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ g,
                             const double *__restrict__ dinv, const double *__restrict__ r, double *p)
{
    const double beta = guarded_ratio(g[2], g[0]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        if (owned[i]) p[i] = fma(beta, p[i], r[i] * dinv[i]);
}
// multigrid mode: x += alpha p ; r -= alpha u (the preconditioner kernel follows)
__global__ void kd_update_xr(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ g,
                             const double *__restrict__ p, const double *__restrict__ u, double *x, double *r)
{
    const double alpha = guarded_ratio(g[0], g[1]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
    {
        if (!owned[i]) continue;
        x[i] = fma(alpha, p[i], x[i]);
        r[i] = fma(-alpha, u[i], r[i]);
    }
}
// partial of r.z over the owned rows
__global__ void kd_dot_rz(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ r,
                          const double *__restrict__ z, double *part)
{
    __shared__ double sm[33];
    double a = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        if (owned[i]) a = fma(r[i], z[i], a);
    a = block_sum(a, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}
// p = z + beta p (beta = g[2]/g[0]); first = 1: p = z
__global__ void kd_direction_z(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ g,
                               const double *__restrict__ z, double *p, int first)
{
    const double beta = first ? 0.0 : guarded_ratio(g[2], g[0]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
        if (owned[i]) p[i] = first ? z[i] : fma(beta, p[i], z[i]);
}
// x = 0 ; r = b (owned rows), 0 elsewhere
__global__ void kd_init_xr(int n_pad, const unsigned char *__restrict__ owned, const double *__restrict__ b, double *x,
                           double *r, double *p)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
    {
        x[i] = 0.0; p[i] = 0.0;
        r[i] = owned[i] ? b[i] : 0.0;
    }
}
__global__ void kd_mark_owned(int n_pad, int n_own, const int *__restrict__ i2c, unsigned char *owned)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_pad) owned[r] = (i2c[r] >= 0 && i2c[r] < n_own) ? 1 : 0;
}

} // namespace xb

using namespace xb;

extern "C" {

int xfemm_dist_destroy(xfemm_dist *d)
{
    if (!d) return XFEMM_OK;
    if (d->loc) xfemm_linprob_destroy(d->loc);
    delete d;
    return XFEMM_OK;
}

int xfemm_dist_create(xfemm_linprob *g, int r0, int r1, xfemm_dist **out)
{
    XB_ARG(g && out && r0 >= 0 && r1 > r0 && r1 <= g->n, "row block");
    xfemm_dist *d = new xfemm_dist;
    d->glob = g; d->r0 = r0; d->r1 = r1; d->n_own = r1 - r0;
    const std::vector<int> &rp = g->h_rowptr, &col = g->h_col;
    // halo columns: every column outside [r0,r1) touched by an owned row
    for (int i = r0; i < r1; i++)
        for (int s = rp[i]; s < rp[i + 1]; s++)
            if (col[s] < r0 || col[s] >= r1) d->halo_cols.push_back(col[s]);
    std::sort(d->halo_cols.begin(), d->halo_cols.end());
    d->halo_cols.erase(std::unique(d->halo_cols.begin(), d->halo_cols.end()), d->halo_cols.end());
    d->n_halo = (int)d->halo_cols.size();
    const int n_loc = d->n_own + d->n_halo;
    auto local_col = [&](int c) -> int {
        if (c >= r0 && c < r1) return c - r0;
        return d->n_own + (int)(std::lower_bound(d->halo_cols.begin(), d->halo_cols.end(), c) - d->halo_cols.begin());
    };
    std::vector<int> lrp(n_loc + 1, 0), lcol, smap;
    lcol.reserve((size_t)(rp[r1] - rp[r0]) + d->n_halo);
    std::vector<std::pair<int, int>> row;
    for (int i = 0; i < d->n_own; i++)
    {
        row.clear();
        for (int s = rp[r0 + i]; s < rp[r0 + i + 1]; s++) row.emplace_back(local_col(col[s]), s);
        std::sort(row.begin(), row.end());
        for (auto &e : row) { lcol.push_back(e.first); smap.push_back(e.second); }
        lrp[i + 1] = (int)lcol.size();
    }
    for (int h = 0; h < d->n_halo; h++)
    {
        lcol.push_back(d->n_own + h);
        smap.push_back(-1);
        lrp[d->n_own + h + 1] = (int)lcol.size();
    }
    int rc = xfemm_linprob_create(n_loc, lrp.data(), lcol.data(), &d->loc);
    if (rc != XFEMM_OK) { delete d; return rc; }
    cudaStream_t st = d->loc->stream;
    auto body = [&]() -> int {
        XB_CUDA(d->slot_map.upload(smap.data(), smap.size(), st));
        int rc2 = linprob_build_layout(d->loc, XFEMM_PRECOND_JACOBI);
        if (rc2 != XFEMM_OK) return rc2;
        const int n_pad = d->loc->lay.n_pad;
        XB_CUDA(d->owned.alloc(n_pad));
        ++g_launch, kd_mark_owned<<<(n_pad + 255) / 256, 256, 0, st>>>(n_pad, d->n_own, d->loc->i2c.p, d->owned.p);
        std::vector<int> hidx(std::max(1, d->n_halo));
        for (int h = 0; h < d->n_halo; h++) hidx[h] = d->loc->lay.c2i[d->n_own + h];
        XB_CUDA(d->halo_int.upload(hidx.data(), hidx.size(), st));
        d->nblk = std::max(1, std::min((n_pad + 255) / 256, d->loc->num_sms * 8));
        XB_CUDA(d->partials.alloc((size_t)d->nblk));
        XB_CUDA(cudaStreamSynchronize(st));
        XB_CUDA(cudaGetLastError());
        return XFEMM_OK;
    };
    rc = body();
    if (rc != XFEMM_OK) { xfemm_dist_destroy(d); return rc; }
    *out = d;
    return XFEMM_OK;
}

// preconditioner of the partitioned solve: XFEMM_PRECOND_JACOBI or XFEMM_PRECOND_AMG (multigrid V-cycle of
// each rank's own diagonal block — additive Schwarz without overlap; call before xfemm_dist_refresh)
int xfemm_dist_set_precond(xfemm_dist *d, int precond)
{
    XB_ARG(d && (precond == XFEMM_PRECOND_JACOBI || precond == XFEMM_PRECOND_AMG), "preconditioner");
    d->precond = precond;
    return XFEMM_OK;
}

int xfemm_dist_sizes(xfemm_dist *d, int *n_own, int *n_halo)
{
    XB_ARG(d, "null handle");
    if (n_own) *n_own = d->n_own;
    if (n_halo) *n_halo = d->n_halo;
    return XFEMM_OK;
}
int xfemm_dist_halo_columns(xfemm_dist *d, int *global_cols)
{
    XB_ARG(d && global_cols, "null pointer");
    std::copy(d->halo_cols.begin(), d->halo_cols.end(), global_cols);
    return XFEMM_OK;
}
// rows of THIS rank (global ids, in the order the peers expect them) whose p values are sent each iteration
int xfemm_dist_set_send_rows(xfemm_dist *d, int n_send, const int *global_rows)
{
    XB_ARG(d && n_send >= 0 && (n_send == 0 || global_rows), "send list");
    std::vector<int> idx(std::max(1, n_send));
    for (int k = 0; k < n_send; k++)
    {
        XB_ARG(global_rows[k] >= d->r0 && global_rows[k] < d->r1, "send row not owned by this rank");
        idx[k] = d->loc->lay.c2i[global_rows[k] - d->r0];
    }
    XB_CUDA(d->send_int.upload(idx.data(), idx.size(), d->loc->stream));
    XB_CUDA(cudaStreamSynchronize(d->loc->stream));
    d->n_send = n_send;
    return XFEMM_OK;
}

// pulls the owned rows (values + RHS) out of the assembled, boundary-conditioned global problem
int xfemm_dist_refresh(xfemm_dist *d)
{
    XB_ARG(d, "null handle");
    xfemm_linprob *L = d->loc, *G = d->glob;
    cudaStream_t st = L->stream;
    XB_CUDA(cudaStreamSynchronize(G->stream));
    ++g_launch, kd_copy_block<<<(int)((L->nnz + 255) / 256), 256, 0, st>>>(L->nnz, d->slot_map.p, G->vals.p, L->vals.p);
    ++g_launch, kd_copy_rhs<<<(L->n + 255) / 256, 256, 0, st>>>(L->n, d->n_own, d->r0, G->b.p, L->b.p);
    // SELL refill + diagonal + RHS in internal order (same helpers as the single-GPU solve)
    int rc = linprob_prepare(L, d->precond);
    if (rc != XFEMM_OK) return rc;
    return linprob_load_rhs_internal(L);
}

int xfemm_dist_init(xfemm_dist *d, double *scal_dev)
{
    XB_ARG(d && scal_dev, "null pointer");
    xfemm_linprob *L = d->loc;
    const int n_pad = L->lay.n_pad;
    if (d->precond == XFEMM_PRECOND_AMG)
    {
        // x = 0, r = b ; z = M^-1 r (V-cycle of the rank's diagonal block) ; p = z ; scal = local r.z
        ++g_launch, kd_init_xr<<<d->nblk, 256, 0, L->stream>>>(n_pad, d->owned.p, L->bint.p, L->x.p, L->r.p, L->p.p);
        int rc = linprob_apply_amg_resident(L);
        if (rc != XFEMM_OK) return rc;
        ++g_launch, kd_direction_z<<<d->nblk, 256, 0, L->stream>>>(n_pad, d->owned.p, scal_dev, L->z.p, L->p.p, 1);
        ++g_launch, kd_dot_rz<<<d->nblk, 256, 0, L->stream>>>(n_pad, d->owned.p, L->r.p, L->z.p, d->partials.p);
        ++g_launch, kd_reduce<<<1, 256, 0, L->stream>>>(d->nblk, d->partials.p, scal_dev, 1);
        XB_CUDA(cudaGetLastError());
        return XFEMM_OK;
    }
    ++g_launch, kd_init<<<d->nblk, 256, 0, L->stream>>>(n_pad, d->owned.p, L->bint.p, L->dinv.p, L->x.p, L->r.p, L->p.p, d->partials.p);
    ++g_launch, kd_reduce<<<1, 256, 0, L->stream>>>(d->nblk, d->partials.p, scal_dev, 1);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int xfemm_dist_pack(xfemm_dist *d, double *sendbuf_dev)
{
    XB_ARG(d, "null handle");
    if (d->n_send == 0) return XFEMM_OK;
    XB_ARG(sendbuf_dev, "null buffer");
    ++g_launch, kd_pack<<<(d->n_send + 255) / 256, 256, 0, d->loc->stream>>>(d->n_send, d->send_int.p, d->loc->p.p, sendbuf_dev);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int xfemm_dist_unpack(xfemm_dist *d, const double *recvbuf_dev)
{
    XB_ARG(d, "null handle");
    if (d->n_halo == 0) return XFEMM_OK;
    XB_ARG(recvbuf_dev, "null buffer");
    ++g_launch, kd_unpack<<<(d->n_halo + 255) / 256, 256, 0, d->loc->stream>>>(d->n_halo, d->halo_int.p, recvbuf_dev, d->loc->p.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int xfemm_dist_spmv_dot(xfemm_dist *d, double *scal_dev)
{
    XB_ARG(d && scal_dev, "null pointer");
    xfemm_linprob *L = d->loc;
    int sb = std::max(1, std::min((L->lay.nslices + 7) / 8, d->nblk));
    L->wide.pre(L->p.p, L->u.p, L->stream);
    ++g_launch, kd_spmv_dot<<<sb, 256, 0, L->stream>>>(L->lay.nslices, L->slice_ptr.p, L->sell_col.p, L->sell_val.p, L->diag.p,
                                                       d->owned.p, L->p.p, L->u.p, d->partials.p,
                                                       L->wide.n_rows ? L->wide.mask.p : nullptr);
    ++g_launch, kd_reduce<<<1, 256, 0, L->stream>>>(sb, d->partials.p, scal_dev, 1);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int xfemm_dist_update(xfemm_dist *d, const double *glob_dev, double *scal_dev)
{
    XB_ARG(d && glob_dev && scal_dev, "null pointer");
    xfemm_linprob *L = d->loc;
    if (d->precond == XFEMM_PRECOND_AMG)
    {
        ++g_launch, kd_update_xr<<<d->nblk, 256, 0, L->stream>>>(L->lay.n_pad, d->owned.p, glob_dev, L->p.p, L->u.p, L->x.p, L->r.p);
        int rc = linprob_apply_amg_resident(L);
        if (rc != XFEMM_OK) return rc;
        ++g_launch, kd_dot_rz<<<d->nblk, 256, 0, L->stream>>>(L->lay.n_pad, d->owned.p, L->r.p, L->z.p, d->partials.p);
        ++g_launch, kd_reduce<<<1, 256, 0, L->stream>>>(d->nblk, d->partials.p, scal_dev, 1);
        XB_CUDA(cudaGetLastError());
        return XFEMM_OK;
    }
    ++g_launch, kd_update<<<d->nblk, 256, 0, L->stream>>>(L->lay.n_pad, d->owned.p, glob_dev, L->dinv.p, L->p.p, L->u.p, L->x.p, L->r.p, d->partials.p);
    ++g_launch, kd_reduce<<<1, 256, 0, L->stream>>>(d->nblk, d->partials.p, scal_dev, 1);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
int xfemm_dist_direction(xfemm_dist *d, const double *glob_dev)
{
    XB_ARG(d && glob_dev, "null pointer");
    xfemm_linprob *L = d->loc;
    if (d->precond == XFEMM_PRECOND_AMG)
    {
        ++g_launch, kd_direction_z<<<d->nblk, 256, 0, L->stream>>>(L->lay.n_pad, d->owned.p, glob_dev, L->z.p, L->p.p, 0);
        XB_CUDA(cudaGetLastError());
        return XFEMM_OK;
    }
    ++g_launch, kd_direction<<<d->nblk, 256, 0, L->stream>>>(L->lay.n_pad, d->owned.p, glob_dev, L->dinv.p, L->r.p, L->p.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}
// owned part of the solution (n_own doubles, caller order) to the host
int xfemm_dist_get_solution(xfemm_dist *d, double *x_own)
{
    XB_ARG(d && x_own, "null pointer");
    xfemm_linprob *L = d->loc;
    int rc = linprob_store_solution_internal(L);
    if (rc != XFEMM_OK) return rc;
    xb::g_d2h_bytes += (long long)d->n_own * 8;
    XB_CUDA(cudaMemcpyAsync(x_own, L->V.p, (size_t)d->n_own * sizeof(double), cudaMemcpyDeviceToHost, L->stream));
    XB_CUDA(cudaStreamSynchronize(L->stream));
    return XFEMM_OK;
}
// the CUDA stream the kernels above are enqueued on (as an integer handle), so the caller can make
// its collectives wait on / be waited on by them
unsigned long long xfemm_dist_stream(xfemm_dist *d) { return d ? (unsigned long long)(uintptr_t)d->loc->stream : 0ull; }

} // extern "C"
