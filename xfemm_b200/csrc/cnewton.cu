// Newton-AC extension of the complex linear problem ([ACSolver]=1).
//
// With a B-H curve the time-harmonic Newton step is only REAL-linear in the unknowns: the reference
// (cfemm/libfemm/cspars.h:53-58) keeps three side matrices next to the complex-symmetric M —
// Mh (Hermitian), Ms (complex symmetric, applied to conj(x)) and Ma (anti-Hermitian) — so that the full
// operator is  x -> M x + Mh x + Ms conj(x) + Ma x  (MultA(...,-1), cspars.cpp:303-312), and solves with
// KludgeSolve (cspars.cpp:983-1056): at most ten COCG solves on M alone with the side terms moved to the
// right-hand side, each followed by a line search on the full residual.
//
// Here the side matrices live on the FULL CSR pattern of the problem (both triangles explicit), so the
// reference's flip rules (Put/Get conjugate on p>q for k=1, negate-conjugate for k=3; cspars.cpp:152-158,
// 292-296) are implicit in the stored values.  Everything below is per-Newton-pass work (<= ~25 vector
// passes), not the inner COCG loop, hence plain CSR kernels.
#include "clinprob.cuh"
#include <complex>
#include <algorithm>
#include <vector>

namespace xb {

// SetValue on the side matrices (cspars.cpp:511-535): column/row i of every side matrix is zeroed INCLUDING the
// diagonal; a free row k gets b[k] -= Mh[k,i] x + (Ms[k,i] + Ma[k,i]) conj(x)  (the reference conjugates x for
// k=3 as well, cspars.cpp:528-532, kept).
__global__ void kn_dirichlet_side(int n, int n_nodes, const int *__restrict__ rowptr, const int *__restrict__ col, cd *vh, cd *vs,
                                  cd *va, cd *b, const int *__restrict__ dir_idx, const cd *__restrict__ xf, int bdw)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int dk = dir_idx[k];
    const cd zero = make_double2(0., 0.);
    cd bk = b[k];
    for (int s = rowptr[k]; s < rowptr[k + 1]; s++)
    {
        const int c = col[s];
        if (c == k)
        {
            if (dk >= 0) { vh[s] = zero; vs[s] = zero; va[s] = zero; }
            continue;
        }
        const int dc = dir_idx[c];
        const bool by_c = dc >= 0 && cband(k, c, bdw, n_nodes);
        const bool by_k = dk >= 0 && cband(c, k, bdw, n_nodes);
        if (by_c || by_k)
        {
            if (by_c && dk < 0)
            {
                const cd x = xf[dc], xc = cconj(x);
                bk = csub(bk, cadd(cmul(vh[s], x), cmul(cadd(vs[s], va[s]), xc)));
            }
            vh[s] = zero; vs[s] = zero; va[s] = zero;
        }
    }
    if (dk < 0) b[k] = bk;
}

// Periodicity / AntiPeriodicity on a side matrix (cspars.cpp:648-668, 732-752).  On full storage the column
// operation  G[k,i] = c, G[k,j] = +-c  with c = (G[k,i] +- G[k,j])/2  and its mirror on the rows are the same
// linear map for all three symmetry classes; mirror = conj (h=1), identity (h=2), -conj (h=3).
// Diagonal block: c = (G[i,i] +- G[i,j] +- G[j,i] + G[j,j])/4 -> (i,i), (j,j) and +-c -> (i,j), (j,i).
__global__ void __launch_bounds__(1024)
kn_periodic_side(int nlevels, const int *__restrict__ level_ptr, const int *__restrict__ item_ptr, const int *__restrict__ meta,
                 const int *__restrict__ items, const int *__restrict__ ij, cd *vh, cd *vs, cd *va)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int lv = 0; lv < nlevels; lv++)
    {
        for (int pr = level_ptr[lv] + warp; pr < level_ptr[lv + 1]; pr += nw)
        {
            const int *m = meta + 6 * pr;
            const double sgn = m[2] ? -1.0 : 1.0;
            for (int h = 1; h <= 3; h++)
            {
                volatile cd *g = h == 1 ? vh : (h == 2 ? vs : va);
                for (int it = item_ptr[pr] + lane; it < item_ptr[pr + 1]; it += 32)
                {
                    const int *q = items + 4 * it;
                    const cd v1 = make_double2(g[q[0]].x, g[q[0]].y), v2 = make_double2(g[q[2]].x, g[q[2]].y);
                    if (v1.x != 0.0 || v1.y != 0.0 || v2.x != 0.0 || v2.y != 0.0)
                    {
                        const cd c = make_double2((v1.x + sgn * v2.x) / 2.0, (v1.y + sgn * v2.y) / 2.0);
                        // mirror entry (i,k) of a value c at (k,i)
                        const cd cm = h == 1 ? make_double2(c.x, -c.y) : (h == 2 ? c : make_double2(-c.x, c.y));
                        g[q[0]].x = c.x; g[q[0]].y = c.y; g[q[1]].x = cm.x; g[q[1]].y = cm.y;
                        g[q[2]].x = sgn * c.x; g[q[2]].y = sgn * c.y; g[q[3]].x = sgn * cm.x; g[q[3]].y = sgn * cm.y;
                    }
                }
                __syncwarp();
                if (lane == 0)
                {
                    const int sii = m[3], sjj = m[4], sij = ij[2 * pr], sji = ij[2 * pr + 1];
                    cd c = make_double2(g[sii].x + g[sjj].x, g[sii].y + g[sjj].y);
                    if (sij >= 0) { c.x += sgn * g[sij].x; c.y += sgn * g[sij].y; }
                    if (sji >= 0) { c.x += sgn * g[sji].x; c.y += sgn * g[sji].y; }
                    c.x /= 4.0; c.y /= 4.0;
                    g[sii].x = c.x; g[sii].y = c.y; g[sjj].x = c.x; g[sjj].y = c.y;
                    // Put(+-c,i,j,h) also defines (j,i): conj / same / -conj of +-c; c is real (h=1), arbitrary (h=2),
                    // imaginary (h=3) up to rounding, so the mirror rule is applied literally
                    const cd cij = make_double2(sgn * c.x, sgn * c.y);
                    const cd cji = h == 1 ? make_double2(cij.x, -cij.y) : (h == 2 ? cij : make_double2(-cij.x, cij.y));
                    if (sij >= 0) { g[sij].x = cij.x; g[sij].y = cij.y; }
                    if (sji >= 0) { g[sji].x = cji.x; g[sji].y = cji.y; }
                }
                __syncwarp();
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// y = [M x] + Mh x + Ms conj(x) + Ma x   (one warp per row; `vm` null: side terms only — the right-hand side
// correction of KludgeSolve, P + conj(U) + R of cspars.cpp:1026-1029)
__global__ void __launch_bounds__(256)
kn_mult(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const cd *__restrict__ vm, const cd *__restrict__ vh,
        const cd *__restrict__ vs, const cd *__restrict__ va, const cd *__restrict__ x, cd *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = w0; r < n; r += nw)
    {
        cd acc = make_double2(0., 0.);
        for (int s = rowptr[r] + lane; s < rowptr[r + 1]; s += 32)
        {
            const cd xv = x[col[s]];
            cd lin = cadd(vh[s], va[s]);
            if (vm) lin = cadd(lin, vm[s]);
            acc = cfma(lin, xv, acc);
            acc = cfma(vs[s], cconj(xv), acc);
        }
        acc.x = warp_sum(acc.x);
        acc.y = warp_sum(acc.y);
        if (lane == 0) y[r] = acc;
    }
}

// y = G x or conj(G) x for one value array on the full pattern (MultA / MultConjA with a selector, cspars.cpp:290-390)
__global__ void __launch_bounds__(256)
kn_mult_one(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const cd *__restrict__ g, int conj_g,
            const cd *__restrict__ x, cd *__restrict__ y)
{
    const int lane = threadIdx.x & 31;
    const long long w0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = w0; r < n; r += nw)
    {
        cd acc = make_double2(0., 0.);
        for (int s = rowptr[r] + lane; s < rowptr[r + 1]; s += 32)
        {
            cd a = g[s];
            if (conj_g) a.y = -a.y;
            acc = cfma(a, x[col[s]], acc);
        }
        acc.x = warp_sum(acc.x);
        acc.y = warp_sum(acc.y);
        if (lane == 0) y[r] = acc;
    }
}

__global__ void kn_sub(int n, const cd *__restrict__ a, const cd *__restrict__ b, cd *out) // out = a - b
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = csub(a[i], b[i]);
}
// V = v + c P ; r -= c U ; v = V   (cspars.cpp:1038-1043)
__global__ void kn_step(int n, double c, const cd *__restrict__ P, const cd *__restrict__ U, cd *V, cd *v, cd *r)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const cd nv = make_double2(fma(c, P[i].x, v[i].x), fma(c, P[i].y, v[i].y));
    V[i] = nv; v[i] = nv;
    r[i] = make_double2(fma(-c, U[i].x, r[i].x), fma(-c, U[i].y, r[i].y));
}
// part[0..] = Re(conj(x).y), part[grid..] = |y|^2
__global__ void kn_dots(int n, const cd *__restrict__ x, const cd *__restrict__ y, double *part)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const cd xv = x[i], yv = y[i];
        a = fma(xv.x, yv.x, fma(xv.y, yv.y, a));
        c = fma(yv.x, yv.x, fma(yv.y, yv.y, c));
    }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { part[blockIdx.x] = a; part[gridDim.x + blockIdx.x] = c; }
}
__global__ void kn_dots_final(int nblk, const double *part, double *out2)
{
    __shared__ double sm[33];
    double a = 0, c = 0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) { a += part[i]; c += part[nblk + i]; }
    a = block_sum(a, sm);
    c = block_sum(c, sm);
    if (threadIdx.x == 0) { out2[0] = a; out2[1] = c; }
}

// Re(conj(x).y) and |y|^2, deterministic two-stage reduction
static int ndots(xfemm_clinprob *p, const cd *x, const cd *y, double *re_xy, double *yy)
{
    cudaStream_t st = p->stream;
    const int n = p->n;
    const int nb = std::max(1, std::min((n + 255) / 256, 512));
    if (p->max_grid < nb) { XB_CUDA(p->partials.alloc((size_t)8 * nb)); p->max_grid = nb; }
    ++g_launch, kn_dots<<<nb, 256, 0, st>>>(n, x, y, p->partials.p);
    ++g_launch, kn_dots_final<<<1, 256, 0, st>>>(nb, p->partials.p, p->d_mail.p + 14);
    XB_CUDA(cudaMemcpyAsync(p->h_mail + 14, p->d_mail.p + 14, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    g_d2h_bytes += 16;
    if (re_xy) *re_xy = p->h_mail[14];
    if (yy) *yy = p->h_mail[15];
    return XFEMM_OK;
}

static int nmult(xfemm_clinprob *p, bool with_m, const cd *x, cd *y)
{
    const int blocks = std::max(1, std::min((p->n + 7) / 8, p->num_sms * 16));
    ++g_launch, kn_mult<<<blocks, 256, 0, p->stream>>>(p->n, p->rowptr.p, p->col.p, with_m ? p->vals.p : nullptr, p->vals_h.p,
                                                        p->vals_s.p, p->vals_a.p, x, y);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

int cnewton_dirichlet(xfemm_clinprob *p, int bdw)
{
    ++g_launch, kn_dirichlet_side<<<(p->n + 255) / 256, 256, 0, p->stream>>>(p->n, p->n_nodes, p->rowptr.p, p->col.p, p->vals_h.p,
                                                                              p->vals_s.p, p->vals_a.p, p->b.p, p->dir_idx.p,
                                                                              p->dir_first.p, bdw);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

int cnewton_periodic(xfemm_clinprob *p, int npairs, const int *ni, const int *nj)
{
    if (p->pbc_ij_hash != p->pbc_hash || (int)p->pbc_ij.n != 2 * npairs)
    {
        // the device plan is sorted by dependency level: read the pairs back from its meta table
        std::vector<int> meta((size_t)6 * npairs), ij((size_t)2 * npairs, -1);
        XB_CUDA(cudaMemcpyAsync(meta.data(), p->pbc_meta.p, meta.size() * sizeof(int), cudaMemcpyDeviceToHost, p->stream));
        XB_CUDA(cudaStreamSynchronize(p->stream));
        auto find = [&](int r, int c) {
            const int *b = p->h_col.data() + p->h_rowptr[r], *e = p->h_col.data() + p->h_rowptr[r + 1];
            const int *it = std::lower_bound(b, e, c);
            return (it != e && *it == c) ? (int)(it - p->h_col.data()) : -1;
        };
        for (int k = 0; k < npairs; k++)
        {
            const int i = meta[6 * k], j = meta[6 * k + 1];
            ij[2 * k] = find(i, j);
            ij[2 * k + 1] = find(j, i);
            if (ij[2 * k] < 0 || ij[2 * k + 1] < 0)
            {
                set_error("Newton AC solver: the pattern lacks the entry coupling the periodic pair (%d,%d)", i, j);
                return XFEMM_ERR_ARG;
            }
        }
        XB_CUDA(p->pbc_ij.upload(ij.data(), ij.size(), p->stream));
        XB_CUDA(cudaStreamSynchronize(p->stream));
        p->pbc_ij_hash = p->pbc_hash;
    }
    (void)ni; (void)nj;
    ++g_launch, kn_periodic_side<<<1, 1024, 0, p->stream>>>(p->pbc_levels, p->pbc_level_ptr.p, p->pbc_item_ptr.p, p->pbc_meta.p,
                                                            p->pbc_items.p, p->pbc_ij.p, p->vals_h.p, p->vals_s.p, p->vals_a.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

} // namespace xb

using namespace xb;

extern "C" {

int xfemm_clinprob_enable_newton(xfemm_clinprob *p)
{
    XB_ARG(p, "null handle");
    if (p->newton) return XFEMM_OK;
    for (auto *v : {&p->vals_h, &p->vals_s, &p->vals_a}) { XB_CUDA(v->alloc((size_t)p->nnz)); XB_CUDA(v->zero(p->stream)); }
    for (auto *v : {&p->nk_b, &p->nk_v, &p->nk_r, &p->nk_p, &p->nk_u}) { XB_CUDA(v->alloc((size_t)p->n)); XB_CUDA(v->zero(p->stream)); }
    XB_CUDA(cudaStreamSynchronize(p->stream));
    p->newton = true;
    return XFEMM_OK;
}

static DevBuf<cd> *side_of(xfemm_clinprob *p, int k) { return k == 1 ? &p->vals_h : (k == 2 ? &p->vals_s : &p->vals_a); }

int xfemm_clinprob_set_side_values(xfemm_clinprob *p, int k, const double *v)
{
    XB_ARG(p && v && k >= 1 && k <= 3, "handle, k in 1..3, values");
    XB_ARG(p->newton, "call xfemm_clinprob_enable_newton first");
    g_h2d_bytes += p->nnz * 16;
    XB_CUDA(cudaMemcpyAsync(side_of(p, k)->p, v, (size_t)p->nnz * 16, cudaMemcpyHostToDevice, p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}
int xfemm_clinprob_get_side_values(xfemm_clinprob *p, int k, double *v)
{
    XB_ARG(p && v && k >= 1 && k <= 3, "handle, k in 1..3, values");
    XB_ARG(p->newton, "call xfemm_clinprob_enable_newton first");
    g_d2h_bytes += p->nnz * 16;
    XB_CUDA(cudaMemcpyAsync(v, side_of(p, k)->p, (size_t)p->nnz * 16, cudaMemcpyDeviceToHost, p->stream));
    XB_CUDA(cudaStreamSynchronize(p->stream));
    return XFEMM_OK;
}

// y = M x + Mh x + Ms conj(x) + Ma x   (MultA(X,Y,-1), cspars.cpp:303-312), host vectors
int xfemm_clinprob_mult_newton(xfemm_clinprob *p, const double *x, double *y)
{
    XB_ARG(p && x && y, "null pointer");
    XB_ARG(p->newton, "call xfemm_clinprob_enable_newton first");
    cudaStream_t st = p->stream;
    g_h2d_bytes += (long long)p->n * 16; g_d2h_bytes += (long long)p->n * 16;
    XB_CUDA(cudaMemcpyAsync(p->nk_p.p, x, (size_t)p->n * 16, cudaMemcpyHostToDevice, st));
    int rc = nmult(p, true, p->nk_p.p, p->nk_u.p);
    if (rc != XFEMM_OK) return rc;
    XB_CUDA(cudaMemcpyAsync(y, p->nk_u.p, (size_t)p->n * 16, cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

// MultA(X,Y,k) / MultConjA(X,Y,k) of one stored matrix (cspars.cpp:290-390): k=0 M, 1 Mh, 2 Ms, 3 Ma; conj != 0
// multiplies by the conjugated matrix.  Host vectors of n complex.  (The reference's MultConjA adds one term
// twice for k=1, cspars.cpp:376-382 — an `else` bound to the wrong `if`; no caller uses that combination and it is
// not reproduced.)
int xfemm_clinprob_mult(xfemm_clinprob *p, int k, int conj, const double *x, double *y)
{
    XB_ARG(p && x && y && k >= 0 && k <= 3, "handle, k in 0..3, vectors");
    XB_ARG(k == 0 || p->newton, "side matrices need xfemm_clinprob_enable_newton");
    cudaStream_t st = p->stream;
    DevBuf<cd> xin, yout; // k=0 works on problems without the Newton scratch vectors
    XB_CUDA(xin.alloc((size_t)p->n)); XB_CUDA(yout.alloc((size_t)p->n));
    g_h2d_bytes += (long long)p->n * 16; g_d2h_bytes += (long long)p->n * 16;
    XB_CUDA(cudaMemcpyAsync(xin.p, x, (size_t)p->n * 16, cudaMemcpyHostToDevice, st));
    const int blocks = std::max(1, std::min((p->n + 7) / 8, p->num_sms * 16));
    ++g_launch, kn_mult_one<<<blocks, 256, 0, st>>>(p->n, p->rowptr.p, p->col.p, k == 0 ? p->vals.p : side_of(p, k)->p, conj ? 1 : 0,
                                                    xin.p, yout.p);
    XB_CUDA(cudaGetLastError());
    XB_CUDA(cudaMemcpyAsync(y, yout.p, (size_t)p->n * 16, cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

// KludgeSolve (cspars.cpp:983-1056).  b, V of the handle in, V out; b is left holding the last corrected
// right-hand side, as in the reference.  info->iterations = COCG iterations summed over the inner solves.
int xfemm_clinprob_kludge_solve(xfemm_clinprob *p, double precision, int warm, int precond, int max_iter, xfemm_solve_info *info)
{
    XB_ARG(p, "null handle");
    XB_ARG(p->newton, "call xfemm_clinprob_enable_newton first");
    cudaStream_t st = p->stream;
    const int n = p->n, nb = (n + 255) / 256;
    const size_t bytes = (size_t)n * sizeof(cd);
    xfemm_solve_info total;
    memset(&total, 0, sizeof total);
    total.precond = precond;
    cudaEvent_t e0, e1;
    XB_CUDA(cudaEventCreate(&e0)); XB_CUDA(cudaEventCreate(&e1));
    XB_CUDA(cudaEventRecord(e0, st));
    auto finish = [&](int rc) {
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        total.solve_ms = ms;
        total.status = rc;
        if (info) *info = total;
        return rc;
    };
    if (!warm) XB_CUDA(p->V.zero(st));
    double normb2 = 0, dummy = 0, rr = 0;
    int rc = ndots(p, p->b.p, p->b.p, &dummy, &normb2);
    if (rc != XFEMM_OK) return finish(rc);
    XB_CUDA(cudaMemcpyAsync(p->nk_b.p, p->b.p, bytes, cudaMemcpyDeviceToDevice, st)); // borig
    XB_CUDA(cudaMemcpyAsync(p->nk_v.p, p->V.p, bytes, cudaMemcpyDeviceToDevice, st)); // v
    // r = b - (M + side) V
    if ((rc = nmult(p, true, p->V.p, p->nk_u.p)) != XFEMM_OK) return finish(rc);
    ++g_launch, kn_sub<<<nb, 256, 0, st>>>(n, p->b.p, p->nk_u.p, p->nk_r.p);
    if ((rc = ndots(p, p->nk_r.p, p->nk_r.p, &dummy, &rr)) != XFEMM_OK) return finish(rc);
    double er = normb2 > 0 ? sqrt(rr / normb2) : 0.0;
    total.relres = er;
    if (er < precision || normb2 == 0.0) return finish(XFEMM_OK);
    // the matrix does not change between the inner solves: the first one prepares the solver representation and the
    // multigrid hierarchy, the others reuse them
    struct KeepPrepared { xfemm_clinprob *p; ~KeepPrepared() { p->keep_prepared = false; } } keep{p};
    p->keep_prepared = false;
    for (int k = 0; k < 10; k++)
    {
        // b = borig - (Mh V + Ms conj(V) + Ma V)
        if ((rc = nmult(p, false, p->V.p, p->nk_u.p)) != XFEMM_OK) return finish(rc);
        ++g_launch, kn_sub<<<nb, 256, 0, st>>>(n, p->nk_b.p, p->nk_u.p, p->b.p);
        xfemm_solve_info inner;
        rc = xfemm_clinprob_cocg(p, precision, 1, precond, (precond == XFEMM_PRECOND_AMG && max_iter <= 0) ? 3000 : max_iter, &inner);
        if (rc == XFEMM_ERR_DIVERGED && precond == XFEMM_PRECOND_AMG)
        {
            XB_CUDA(cudaMemcpyAsync(p->V.p, p->nk_v.p, bytes, cudaMemcpyDeviceToDevice, st));
            rc = xfemm_clinprob_cocg(p, precision, 1, XFEMM_PRECOND_JACOBI, max_iter, &inner);
        }
        total.iterations += inner.iterations;
        total.colors = inner.colors;
        if (rc != XFEMM_OK) return finish(rc);
        p->keep_prepared = true;
        // step length along the new direction that reduces the full residual the most
        ++g_launch, kn_sub<<<nb, 256, 0, st>>>(n, p->V.p, p->nk_v.p, p->nk_p.p);
        if ((rc = nmult(p, true, p->nk_p.p, p->nk_u.p)) != XFEMM_OK) return finish(rc);
        double ru = 0, uu = 0;
        if ((rc = ndots(p, p->nk_r.p, p->nk_u.p, &ru, &uu)) != XFEMM_OK) return finish(rc);
        const double c = uu != 0.0 ? ru / uu : 0.0;
        ++g_launch, kn_step<<<nb, 256, 0, st>>>(n, c, p->nk_p.p, p->nk_u.p, p->V.p, p->nk_v.p, p->nk_r.p);
        if ((rc = ndots(p, p->nk_r.p, p->nk_r.p, &dummy, &rr)) != XFEMM_OK) return finish(rc);
        er = sqrt(rr / normb2);
        total.relres = er;
        total.er = er;
        if (er < precision * 10.) break;
    }
    XB_CUDA(cudaGetLastError());
    return finish(XFEMM_OK);
}

} // extern "C"
