// Solver layout (internal row order + SELL-32, layout.h) of the uncoloured kind built ON THE DEVICE from the CSR
// pattern that already lives there.  Same result, array for array, as build_sell_layout() on the host (the
// XFEMM_LAYOUT_CHECK=1 self-test compares them); at 2 M rows the host version costs ~0.12 s on 32 threads plus the
// upload of 2 x 56 MB of slot tables, this one a few milliseconds.  What stays on the host: the handful of wide rows
// (build_wide_part) and small tables the host-side planners read (i2c, c2i, slice_ptr).
//
//   order   = rows along the Z-curve through the node coordinates (stable radix sort by 22-bit cell), or as forced
//   i2c     = order, windows of `sigma` rows stably sorted by descending row length          (radix sort, one key)
//   widths  = longest row of every slice of 32 rows, slice_ptr = exclusive scan x 32
//   slots   = one warp per slice, lane = row; the entries of a row in ascending internal column
#include "linprob.cuh"
#include "layout_device.cuh"

namespace xb {

int linprob_build_layout_device(xfemm_linprob *p, int sigma, int wide_thr, bool use_coords)
{
    return build_layout_device(p, sigma, wide_thr, &p->forced_order, use_coords ? &p->h_xy : nullptr, &p->d_xy);
}

int linprob_check_layout_device(xfemm_linprob *p, const SellLayout &H) { return check_layout_device(p, H); }

} // namespace xb
