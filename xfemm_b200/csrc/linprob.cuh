// Device-side real symmetric linear problem (backend of CBigLinProb; reference: cfemm/libfemm/spars.h:38-83).
#pragma once
#include "common.cuh"
#include "layout.h"
#include "wide_rows.cuh"
#include <vector>
#include <cstdint>
#include <thread>

// Data layout in HBM (all FP64 values, int32 indices):
//   "boundary" representation, numbered as the caller numbers unknowns (Cuthill–McKee order):
//       rowptr[n+1], col[nnz] (full CSR, both triangles, sorted), vals[nnz], b[n], V[n]
//     → written by the assembly kernels (atomics into CSR slots) and the boundary-condition
//       kernels; this is what xfemm_linprob_get_values() exports and the parity tests compare.
//   "solver" representation, internal row order (colour-major, then length-sorted windows):
//       split SELL-32: slices of 32 rows stored column-major, per slice [lower block | upper
//       block] (lower = neighbour earlier in the internal order = lower colour), diagonal kept
//       apart.  A warp owns a slice, lane = row, so every matrix load is a fully coalesced
//       256 B (values) / 128 B (columns) line.  sell_src[slot] = CSR index feeding the slot
//       (-1 = padding) so a refill after re-assembly is one gather pass.
//       Work vectors x,r,p,u,z,diag,dinv,bint live in the same internal order.
namespace xb { struct AmgHierarchy; }
struct xfemm_linprob
{
    int n = 0;
    long long nnz = 0;
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;

    // boundary representation
    std::vector<int> h_rowptr, h_col;
    std::vector<float> h_xy;       // optional node coordinates (x,y per row): locality hint for the internal order
    xb::DevBuf<float> d_xy;        // device copy of h_xy (left by the device layout builder for the multigrid set-up)
    // aggregates of the first coarsening, computed by a host thread from the moment the pattern is known
    // (xb::linprob_start_aggregation; joined by amg_build or the destructor)
    std::thread agg_thread;
    std::vector<int> agg_early;
    int agg_early_nc = 0, agg_wide_thr = -1;
    ~xfemm_linprob() { if (agg_thread.joinable()) agg_thread.join(); }
    int deterministic = 0;         // bit-reproducible solves asked for (xfemm_mag_set_deterministic): no scatter restriction
    int wide_threshold = -1;       // rows with more off-diagonals are kept out of the SELL slices (-1: XFEMM_WIDE_ROW, 64)
    std::vector<int> forced_order; // optional: the internal row order to use as is (a companion of another problem)
    xb::DevBuf<int> rowptr, col;
    xb::DevBuf<double> vals, b, V, Vold;
    xb::DevBuf<double> snap_vals, snap_b; // xfemm_linprob_snapshot: boundary-conditioned values + RHS to restore from

    // solver representation
    xb::SellLayout lay;            // host copy (small vectors kept; big ones cleared after upload)
    int layout_precond = -1;       // preconditioner the current layout was built for
    xb::DevBuf<int> i2c, c2i, slice_ptr, slice_wl, slice_wu, sell_col, sell_src, diag_src;
    xb::DevBuf<double> sell_val, diag, dinv;
    xb::DevBuf<double> x, r, p, u, z, w, bint;
    xb::DevBuf<double> partials;   // 2 * max_grid
    xb::DevBuf<int> color_slice;
    xb::WideDev<double> wide;      // rows kept out of the SELL slices (layout.h)
    xb::AmgHierarchy *amg = nullptr; // coarse levels of the multigrid preconditioner (amg.cuh), owned
    xb::DevBuf<int4> tiles;        // per-block tiles of the SELL stream for the shared-memory ring (stage.cuh)
    xb::DevBuf<int> tile_ptr;
    int tile_grid = -1;            // grid size the tiles were cut for
    int tile_first = -1, tile_end = -1; // and the range of slices
    int precond_grid = 0;          // cached grid of the stand-alone preconditioner kernel
    xb::DevBuf<double> hist1, hist2, hist3; // solutions of the last three problems solved on this handle (sweeps)
    int hist_count = 0;
    int max_grid = 0;

    // Dirichlet scratch
    xb::DevBuf<int> dir_idx, dir_nodes;
    xb::DevBuf<double> dir_first, dir_last;

    // periodic plan (cached per pair list)
    uint64_t pbc_hash = 0;
    int pbc_pairs = 0, pbc_levels = 0, pbc_max_level_pairs = 0;
    xb::DevBuf<int> pbc_level_ptr, pbc_item_ptr, pbc_meta, pbc_items;
    std::vector<int> h_pbc_level_ptr; // host copy: pairs per dependency level
    xb::DevBuf<int> pbc_wait;         // per pair (level-sorted): pairs that must be finished before it starts
    // long dependency chains: chains of pairs cut into segments that are worked through in shared memory (k_periodic_chains)
    int pbc_chains = 0;
    xb::DevBuf<int> pbc_chain_seg, pbc_seg_slot, pbc_seg_gid, pbc_seg_pair, pbc_seg_item;
    xb::DevBuf<int4> pbc_pair_tab;
    xb::DevBuf<ushort4> pbc_sitems;
    // the pairs' units grouped by dependency level (k_periodic_units): the usual path
    int pbc_unit_levels = 0;
    std::vector<int> h_pbc_unit_ptr;
    xb::DevBuf<int4> pbc_unit_slots;
    xb::DevBuf<int> pbc_unit_flags;
    xb::DevBuf<unsigned char> fill_mask, created; // optional, see xfemm_linprob_set_fill_mask

    // result mailbox
    xb::DevBuf<double> d_mail;
    double h_mail[16];
};

namespace xb {
int linprob_build_layout(xfemm_linprob *p, int precond);
void linprob_start_aggregation(xfemm_linprob *p); // amg.cu
int linprob_prepare(xfemm_linprob *p, int precond);
int linprob_load_rhs_internal(xfemm_linprob *p);
int linprob_store_solution_internal(xfemm_linprob *p);
int linprob_apply_amg_resident(xfemm_linprob *p);
bool linprob_ensure_tiles(xfemm_linprob *p, int grid);
bool linprob_ensure_tiles_range(xfemm_linprob *p, int grid, int first, int end);
int linprob_load_guess_internal(xfemm_linprob *p, double *dst);
} // namespace xb
