// Kernel arguments of the persistent PCG kernels (linprob.cu: one GPU; dist_amg.cu: one mesh over several GPUs).
#pragma once
#include "linprob.cuh"
#include "wide_rows.cuh"

namespace xb {

struct PcgArgs
{
    int n_pad, nslices, ncolors;
    const int *slice_ptr, *slice_wl, *slice_wu, *sell_col, *color_slice;
    const double *sell_val, *diag, *dinv, *b;
    double *x, *r, *p, *u, *z, *w;
    double *part; // 2*gridDim
    double precision, omega;
    int max_iter, warm;
    double *mail; // [0]=iters [1]=er [2]=status [3]=res_o [4]=res
    WideView wide; // rows outside the SELL slices
    const double *wide_val;
    double *wide_acc;
    const int4 *tiles;   // slice-aligned tiles of the SELL stream per block (stage.cuh), or null
    const int *tile_ptr; // gridDim+1
};

PcgArgs linprob_make_args(xfemm_linprob *p, double precision, int warm, int max_iter);

} // namespace xb
