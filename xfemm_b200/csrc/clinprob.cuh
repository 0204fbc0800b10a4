// Device-side complex-symmetric linear problem (backend of CBigComplexLinProb; cspars.h:37-95).
#pragma once
#include "common.cuh"
#include "ccomplex.cuh"
#include "layout.h"
#include "wide_rows.cuh"
#include <vector>
#include <cstdint>

using xb::cd;
struct xfemm_linprob;

struct xfemm_clinprob
{
    int n = 0, n_nodes = 0, device = 0;
    long long nnz = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::vector<int> h_rowptr, h_col;
    xb::DevBuf<int> rowptr, col;
    xb::DevBuf<cd> vals, b, V, Vold;
    xb::SellLayout lay;
    int layout_precond = -1;
    xb::DevBuf<int> i2c, c2i, slice_ptr, slice_wl, slice_wu, sell_col, sell_src, diag_src, color_slice;
    xb::DevBuf<cd> sell_val, diag, dinv, x, r, p, u, z, w, bint;
    xb::DevBuf<double> partials, d_mail;
    xb::WideDev<cd> wide;
    // XFEMM_PRECOND_AMG: real companion problem carrying the multigrid hierarchy (owned).  Its matrix is
    // |Re A| + |Im A| between mesh nodes (stiffness + omega*sigma*mass, both positive semi-definite: the spectrum of
    // the preconditioned operator then lies on the segment between 1 and j); its pattern leaves the circuit
    // unknowns out, re_src[slot of the companion] = slot of this problem.
    xfemm_linprob *re_prec = nullptr;
    xb::DevBuf<int> re_src;
    int max_grid = 0;
    double h_mail[16];
    xb::DevBuf<int> dir_idx, dir_nodes;
    xb::DevBuf<cd> dir_first, dir_last;
    uint64_t pbc_hash = 0;
    int pbc_pairs = 0, pbc_levels = 0;
    xb::DevBuf<int> pbc_level_ptr, pbc_item_ptr, pbc_meta, pbc_items;
    xb::DevBuf<unsigned char> fill_mask, created;
    // Newton AC solver ([ACSolver]=1): the side matrices Mh (Hermitian), Ms (complex symmetric) and Ma
    // (anti-Hermitian) of cspars.h:53-55, on the same full CSR pattern as `vals`; empty until
    // xfemm_clinprob_enable_newton.  Wipe, Dirichlet and periodic conditions then act on them too (cnewton.cu).
    // KludgeSolve calls the COCG solver up to ten times on unchanged matrix values: while keep_prepared is set the
    // solver representation and the multigrid hierarchy prepared for `prepared_precond` are reused as they are
    bool keep_prepared = false;
    int prepared_precond = -1;
    double prepared_sign = 1.0;
    bool newton = false;
    xb::DevBuf<cd> vals_h, vals_s, vals_a;
    xb::DevBuf<int> pbc_ij;          // per periodic pair: slot(i,j), slot(j,i)
    uint64_t pbc_ij_hash = 0;
    xb::DevBuf<cd> nk_b, nk_v, nk_r, nk_p, nk_u; // scratch vectors of xfemm_clinprob_kludge_solve
};

namespace xb {
int cnewton_dirichlet(xfemm_clinprob *p, int bdw);                              // after kc_dirichlet, dir_idx still set
int cnewton_periodic(xfemm_clinprob *p, int npairs, const int *ni, const int *nj); // after kc_periodic, plan current
} // namespace xb

