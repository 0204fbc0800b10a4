/* xfemm_b200 — C ABI of the B200-native fsolver hot path.
 *
 * Plain C types only (pointers, sizes, int status: 0 = ok, non-zero = error; text via
 * xfemm_last_error()).  No exceptions cross this boundary, no torch types appear in it.
 * One handle per problem, so batches (rotor-angle / frequency sweeps) and multi-GPU runs are
 * simply several handles; a handle is bound to the CUDA device that was current at creation.
 * All array arguments are HOST buffers unless the name ends in _dev.
 *
 * What each group replaces in the reference (paths relative to the xfemm tree):
 *   xfemm_linprob_*   cfemm/libfemm/spars.h:38-83    class CBigLinProb  (Create/Put/Get/Wipe/
 *                     SetValue/Periodicity/AntiPeriodicity/MultA/MultPC/Dot/PCGSolve,
 *                     spars.cpp:80-474)
 *   xfemm_clinprob_*  cfemm/libfemm/cspars.h:37-95   class CBigComplexLinProb (cspars.cpp:119-1081)
 *   xfemm_mag_*       cfemm/fsolver/static2d.cpp:53-1033  FSolver::Static2D
 *                     cfemm/fsolver/staticaxi.cpp:45-794  FSolver::StaticAxisymmetric
 *                     cfemm/fsolver/harmonic2d.cpp:38-890 FSolver::Harmonic2D
 *                     cfemm/fsolver/harmonicaxi.cpp:34-821 FSolver::HarmonicAxisymmetric
 *   xfemm_pattern_*   the sparsity structure the reference grows implicitly through
 *                     CBigLinProb::Put (spars.cpp:105-142) during assembly + Periodicity
 *   xfemm_renumber_*  cfemm/libfemm/cuthill.cpp:40-391 (Cuthill–McKee + element comb sort)
 *   xfemm_fsolver_*   cfemm/fsolver/main.cpp:39-84 and mfemm/mexfsolver.cpp:34-135 (file-level
 *                     entry: LoadProblemFile + runSolver, exit codes 0/1/2)
 */
#ifndef XFEMM_B200_H
#define XFEMM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XFEMM_OK 0
#define XFEMM_ERR_CUDA 1      /* CUDA runtime error or no usable device            */
#define XFEMM_ERR_ARG 2       /* invalid argument                                  */
#define XFEMM_ERR_SINGULAR 3  /* zero diagonal (reference: PCGSolve returns false) */
#define XFEMM_ERR_DIVERGED 4  /* NaN / breakdown / iteration cap                   */
#define XFEMM_ERR_IO 5

/* preconditioners for the Krylov solvers */
#define XFEMM_PRECOND_JACOBI 0
#define XFEMM_PRECOND_MCSSOR 1 /* multicolour SSOR (GPU-parallel stand-in for the serial SSOR) */
#define XFEMM_PRECOND_AMG 2    /* aggregation-multigrid V-cycle (real problems only) */

typedef struct xfemm_linprob xfemm_linprob;   /* real symmetric system on the device    */
typedef struct xfemm_clinprob xfemm_clinprob; /* complex symmetric system on the device */
typedef struct xfemm_mag xfemm_mag;           /* magnetics mesh + materials on the device */
typedef struct xfemm_fsolver xfemm_fsolver;   /* host-side FSolver object (file level)   */

/* ------------------------------------------------------------------ runtime */
int xfemm_cuda_init(int device);        /* cudaSetDevice + context warm-up; XFEMM_ERR_CUDA if no GPU */
int xfemm_device_count(void);
const char *xfemm_last_error(void);     /* thread-local message of the last failing call */
const char *xfemm_version(void);
/* accounting for benchmarks: bytes copied host→device / device→host by this library since the
 * last reset, and the number of kernels it has launched */
void xfemm_transfer_bytes(long long *h2d, long long *d2h, int reset);
long long xfemm_launch_count(void);

/* ------------------------------------------------------------------ solve statistics */
typedef struct xfemm_solve_info
{
    int iterations;        /* Krylov iterations performed                                  */
    int precond;           /* preconditioner actually used                                 */
    int colors;            /* colours of the multicolour ordering (0 for Jacobi)           */
    int status;            /* XFEMM_OK / XFEMM_ERR_SINGULAR / XFEMM_ERR_DIVERGED            */
    double er;             /* reference stopping measure sqrt((z.r)/(z0.b)) at exit         */
    double relres;         /* true ||b-Ax||2 / ||b||2 at exit                               */
    double solve_ms;       /* CUDA-event time of the Krylov loop (excl. format conversion) */
    double setup_ms;       /* CUDA-event time of per-solve setup (SELL fill, diag, …)      */
    double bytes_per_iter; /* algorithmic bytes one iteration moves (SURVEY §8d formula)   */
} xfemm_solve_info;

/* ------------------------------------------------------------------ real linear problem
 * Pattern: full (both triangles) CSR with sorted int32 columns, every row holds its diagonal.
 * Row/column numbering is the caller's (the Cuthill–McKee order of the reference). */
int xfemm_linprob_create(int n, const int *rowptr, const int *colidx, xfemm_linprob **out);
int xfemm_linprob_destroy(xfemm_linprob *p);
int xfemm_linprob_n(const xfemm_linprob *p);
int xfemm_linprob_get_pattern(xfemm_linprob *p, int *rowptr, int *colidx); /* host copy of the CSR pattern; either may be NULL */
long long xfemm_linprob_nnz(const xfemm_linprob *p);
int xfemm_linprob_wipe(xfemm_linprob *p);                               /* spars.cpp:348 Wipe  */
/* device-side copy of the values and the right-hand side (after boundary conditions) and its restore: the
 * position-independent part of a sweep is assembled once (FSolver, option age_union_pattern) */
int xfemm_linprob_snapshot(xfemm_linprob *p);
int xfemm_linprob_restore(xfemm_linprob *p);
int xfemm_linprob_set_values(xfemm_linprob *p, const double *vals_csr); /* H2D, CSR order      */
int xfemm_linprob_get_values(xfemm_linprob *p, double *vals_csr);       /* D2H                 */
int xfemm_linprob_set_rhs(xfemm_linprob *p, const double *b);
int xfemm_linprob_get_rhs(xfemm_linprob *p, double *b);
int xfemm_linprob_set_solution(xfemm_linprob *p, const double *V);
int xfemm_linprob_get_solution(xfemm_linprob *p, double *V);
/* spars.cpp:318 SetValue, applied for `count` nodes at once; x_first = value of the first
 * SetValue call on that node (used for the neighbours' RHS), x_last = value of the last call
 * (used for the node's own RHS); bdw = CBigLinProb::bdw (0 = whole matrix). */
int xfemm_linprob_dirichlet(xfemm_linprob *p, int count, const int *nodes, const double *x_first,
                            const double *x_last, int bdw);
/* spars.cpp:421/366 Periodicity (type 0) / AntiPeriodicity (type 1), applied in list order. */
int xfemm_linprob_periodic(xfemm_linprob *p, int npairs, const int *ni, const int *nj, const int *type);
/* spars.cpp:238 PCGSolve.  warm != 0 keeps the current solution as initial guess.
 * max_iter <= 0 → default cap (the reference has none). */
int xfemm_linprob_pcg(xfemm_linprob *p, double precision, int warm, int precond, int max_iter,
                      xfemm_solve_info *info);
int xfemm_linprob_multa(xfemm_linprob *p, const double *x, double *y);  /* spars.cpp:167 MultA */
int xfemm_linprob_multpc(xfemm_linprob *p, int precond, const double *x, double *y); /* :197   */
/* timing hooks for bench.py: run `reps` SpMVs / PCG iterations on resident data, device-timed */
int xfemm_linprob_bench_spmv(xfemm_linprob *p, int reps, double *ms_per_rep);
/* Solution history for parameter sweeps on one handle (rotor positions): _push records the current solution,
 * _guess overwrites V with the last recorded solution (extrapolate = 0), the linear (1) or quadratic (2)
 * extrapolation of the last two / three recorded solutions (equidistant steps assumed) as the initial guess of a
 * warm-started xfemm_linprob_pcg.  The converged result does not depend on the guess. */
int xfemm_linprob_history_push(xfemm_linprob *p);
int xfemm_linprob_history_guess(xfemm_linprob *p, int extrapolate);
int xfemm_linprob_history_clear(xfemm_linprob *p);
/* value-independent part of the set-up of a later xfemm_linprob_pcg with `precond` (solver layout,
 * multigrid hierarchy): depends on the sparsity pattern only and may run on a host thread while the
 * assembly / boundary-condition calls work on the same problem */
int xfemm_linprob_prebuild(xfemm_linprob *p, int precond);
/* optional locality hint: coordinates (x,y per row) of the unknowns.  The solver's internal row order then
 * follows a Z-curve through the plane instead of the caller's numbering (gathers of a slice hit fewer
 * sectors, multigrid transfers become local).  Results are unaffected.  xy = NULL removes the hint. */
int xfemm_linprob_set_coords(xfemm_linprob *p, const double *xy);
/* levels of the multigrid hierarchy built by the last XFEMM_PRECOND_AMG solve: rows and full-matrix
 * non-zeros per level (level 0 = the problem itself); returns the number of levels (0: none built) */
int xfemm_linprob_amg_levels(xfemm_linprob *p, int max_levels, int *n_rows, long long *nnz);

/* ------------------------------------------------------------------ sparsity pattern
 * Builds the structure CBigLinProb ends up with after assembly + boundary conditions:
 * mesh graph ∪ air-gap-element 10x10 blocks ∪ periodic fill (symbolic replay of
 * spars.cpp:421-474 in list order).  Output arrays are malloc'ed; free with xfemm_free. */
int xfemm_pattern_build(int n_rows /* nodes + appended circuit unknowns */,
                        int n_elems, const int *tri, int tri_stride /* ints per element: 3 or 4 */,
                        int n_age_cells, const int *age_nodes /* n_age_cells*10 */,
                        int n_extra, const int *extra_i, const int *extra_j /* extra symmetric couplings */,
                        int n_pbc, const int *pbc_i, const int *pbc_j,
                        int **rowptr_out, int **colidx_out,
                        unsigned char **fill_out /* optional: 1 = slot exists only through the periodic replay */);
/* the same result with the mesh-graph pass on the GPU (radix sort of (row,column) keys); needs a CUDA device —
 * this is what FSolver calls; xfemm_pattern_build is the host-only version */
int xfemm_pattern_build_device(int n_rows /* nodes + appended circuit unknowns */,
                        int n_elems, const int *tri, int tri_stride /* ints per element: 3 or 4 */,
                        int n_age_cells, const int *age_nodes /* n_age_cells*10 */,
                        int n_extra, const int *extra_i, const int *extra_j /* extra symmetric couplings */,
                        int n_pbc, const int *pbc_i, const int *pbc_j,
                        int **rowptr_out, int **colidx_out,
                        unsigned char **fill_out /* optional: 1 = slot exists only through the periodic replay */);
/* The reference creates periodic fill only where a value is non-zero at that moment
 * (spars.cpp:442: if v1!=0 || v2!=0), which the structural replay cannot know.  Hand the fill
 * flags to the problem and, after the boundary conditions ran, ask which slots the reference's
 * linked lists would contain: exists[s] = !fill[s] || written by the periodic kernel. */
int xfemm_linprob_set_fill_mask(xfemm_linprob *p, const unsigned char *fill);
int xfemm_linprob_get_exists_mask(xfemm_linprob *p, unsigned char *exists);
void xfemm_free(void *ptr);
/* Page-lock / release a caller-owned host array (cudaHostRegister): copies between it and the device then run by
 * DMA at full PCIe speed instead of being staged.  Worth it for arrays that are transferred many times (the result
 * arrays of a sweep); the host mirror's CBigLinProb does it for V and b from the second fetch on. */
int xfemm_host_register(void *ptr, size_t bytes);
int xfemm_host_unregister(void *ptr);

/* ------------------------------------------------------------------ renumbering (host, integer, bit-exact)
 * edges: n_edges*2 node pairs in .edge file order.  newnum_out[i] = new index of node i.
 * cuthill.cpp:94-391 */
int xfemm_renumber_cuthill(int n_nodes, int n_edges, const int *edges, int *newnum_out, int *bandwidth_out);
/* cuthill.cpp:40-85: the reference's (incomplete) comb sort on p0+p1+p2; order_out[k] = index of
 * the input element that ends up at position k. */
int xfemm_renumber_sort_elements(int n_elems, const int *tri, int *order_out);

/* ------------------------------------------------------------------ magnetics assembly
 * Per-label record: everything an element needs besides its geometry is a function of its
 * block label (reference: labellist[lbl] → blockproplist[BlockType], circproplist[InCircuit]). */
typedef struct xfemm_mag_label
{
    double mu1, mu2;      /* linear (Iter==0) relative permeabilities after lamination rules   */
    double Jsrc;          /* J_block + circuit term  [MA/m^2]   (static2d.cpp:483-504)          */
    double H_c;           /* coercivity [A/m]                                                 */
    double magdir;        /* magnetisation direction [deg] (per-element override: magdir_elem) */
    double lamfill;       /* lamination fill factor                                           */
    int lamtype;          /* lamination type                                                  */
    int bh_offset;        /* first point of this label's B-H table in the flat arrays, or -1  */
    int bh_points;        /* number of B-H points (0 = linear)                                */
    int is_external;      /* axisymmetric external region flag                                */
    /* harmonic-only fields */
    double mu1_im, mu2_im;  /* imaginary parts of the complex permeabilities                  */
    double sigma;           /* conductivity seen by the eddy term [MS/m]                      */
    double Jsrc_im;
    double dVs;             /* axisymmetric: circuit voltage-gradient source = dVs / R_element
                               (-100 dV sigma, staticaxi.cpp:338-339); 0 otherwise              */
    int circuit;           /* index of the circuit unknown row (harmonic Case 0), or -1       */
    int pad_;
} xfemm_mag_label;

typedef struct xfemm_mag_desc
{
    int axisymmetric;            /* 0 planar (static2d.cpp), 1 axisymmetric (staticaxi.cpp)   */
    int n_nodes, n_elems;
    const double *xy;            /* n_nodes*2, centimetres, solver (Cuthill–McKee) order        */
    const int *tri;              /* n_elems*4: p0,p1,p2,label                                  */
    const double *magdir_elem;   /* optional n_elems (NULL = use the label's)                  */
    int n_labels;
    const xfemm_mag_label *labels;
    int n_bh;                    /* flat B-H tables: B, H, slope (real parts for statics)       */
    const double *bh_B, *bh_H, *bh_slope;
    const double *bh_H_im, *bh_slope_im; /* optional (harmonic)                                */
    /* mixed (BdryFormat 2) edges: element index, local side, c0, c1   (static2d.cpp:460-480)  */
    int n_mixed;
    const int *mixed_elem_side;  /* n_mixed*2 */
    const double *mixed_c0c1;    /* n_mixed*2 */
    /* air-gap elements, flattened to ring cells (static2d.cpp:277-348)                          */
    int n_age_cells;
    const int *age_nodes;        /* n_age_cells*10                                             */
    const double *age_weights;   /* n_age_cells*10                                             */
    const int *age_matrix_id;    /* n_age_cells: which 10x10 matrix                            */
    int n_age_matrices;
    const double *age_matrices;  /* n_age_matrices*100, row-major, upper triangle used          */
    /* point currents  (static2d.cpp:819-825)                                                   */
    int n_point_src;
    const int *point_src_node;
    const double *point_src_val; /* 0.01*J                                                     */
    /* axisymmetric extras */
    double ext_zo, ext_ro, ext_ri; /* external region (staticaxi.cpp:611-617), in cm           */
    double r_eps;                  /* 1e-6*units: nodes with |r| below are on the axis          */
} xfemm_mag_desc;

/* Raw material record of a block label for the block integrals below (CMMaterialProp members) */
typedef struct xfemm_integral_label
{
    double mu_x, mu_y, lamfill, H_c, magdir;
    int lamtype, bh_offset, bh_points, pad_;
} xfemm_integral_label;

/* 10x10 matrix of one air-gap-element ring cell for mesh shifts ci, co and aspect ratio
 * K = dr/(R dtheta) (static2d.cpp:200-274); MG100 row-major, symmetric */
int xfemm_age_cell_matrix(double ci, double co, double K, double *MG100);
/* creates the device mesh and binds it to a linear problem whose pattern contains the mesh graph */
int xfemm_mag_create(const xfemm_mag_desc *desc, xfemm_linprob *lin, xfemm_mag **out);
int xfemm_mag_destroy(xfemm_mag *m);
/* replaces the per-label table (Newton-independent inputs such as circuit currents) */
int xfemm_mag_set_labels(xfemm_mag *m, int n_labels, const xfemm_mag_label *labels);
/* replaces the air-gap-element tables (rotor-angle sweeps reuse mesh + pattern) */
int xfemm_mag_set_age(xfemm_mag *m, int n_age_cells, const int *age_nodes, const double *age_weights,
                      const int *age_matrix_id, int n_age_matrices, const double *age_matrices);
/* One assembly pass into the bound linear problem: Wipe (iter>0), AGE blocks, element loop,
 * point currents.  iter==0 uses the linear permeabilities; iter>0 evaluates B from the current
 * solution, looks up nu(B), dnu/dB^2 and adds the Newton term (static2d.cpp:683-805).
 * elapsed_ms (optional) receives the CUDA-event time. */
int xfemm_mag_assemble_static(xfemm_mag *m, int iter, double *elapsed_ms);
/* the same in two parts: 1 = everything that does not depend on the rotor position (wipes first), 2 = the air-gap
 * element blocks added to the current values; 3 = both (= xfemm_mag_assemble_static) */
int xfemm_mag_assemble_static_parts(xfemm_mag *m, int iter, int parts, double *elapsed_ms);
/* element permeabilities of the last assembly pass (n_elems*2: mu1, mu2) */
int xfemm_mag_get_mu(xfemm_mag *m, double *mu12);
/* Newton bookkeeping on the device (static2d.cpp:953-989): saves V_old, returns
 * ||V-V_old||^2 and ||V||^2, blends V = relax*V + (1-relax)*V_old. */
int xfemm_linprob_save_solution(xfemm_linprob *p);
int xfemm_linprob_restore_solution(xfemm_linprob *p); /* V = V_old: the saved iterate as the starting guess of a retry */
int xfemm_linprob_newton_norms(xfemm_linprob *p, double *dv2, double *v2);
int xfemm_linprob_relax(xfemm_linprob *p, double relax);
/* planar: b = V*c ; axisymmetric: b = V*c*2*pi*r*0.01 (static2d.cpp:1018-1021, staticaxi.cpp:779-783) */
/* 1: deterministic assembly — the element kernels store their contributions in a scratch array and a gather pass adds
 * them per CSR slot / row in a fixed order (bit-reproducible run to run); 0 (default): FP64 atomics straight into the
 * slots (the sum order, hence the last bits, vary).  Reference scatter: cfemm/fsolver/static2d.cpp:807-815. */
int xfemm_mag_set_deterministic(xfemm_mag *m, int on);
int xfemm_mag_solution_to_A(xfemm_mag *m, double *A_host);
/* Per block label, from the solution resident on the device (planar magnetostatics): energy_area[l] = sum over the
 * label's elements of area [m^2] x stored energy density [J/m^3] (FPProc::BlockIntegral(2), cfemm/fpproc/fpproc.cpp:3762-3830,
 * CMMaterialProp::DoEnergy/GetEnergy, cfemm/libfemm/CMaterialProp.cpp:536-650) and a_area[l] = sum of area x mean nodal A
 * [Wb m] (the A.J integrand of BlockIntegral(0) / GetFluxLinkage, fpproc.cpp:3678-3686, 5157-5196). */
int xfemm_mag_label_integrals(xfemm_mag *m, int n_labels, const xfemm_integral_label *labels, double *energy_area, double *a_area);
/* DC torque through an air-gap element from the RESIDENT solution (no nodal values leave the device).
 * Replaces the post-processor's mo_gapintegral(name,0): cfemm/fpproc/fpproc.cpp:1391-1615 (flux density
 * per ring cell and its harmonics) and 5418-5446 (torque).  cell0/ncells: range in the flattened ring
 * tables; lengths in metres; br_bt (optional, 2*ncells): radial and tangential flux density per cell [T]. */
int xfemm_mag_gap_torque(xfemm_mag *m, int cell0, int ncells, double inner_shift, double outer_shift, double arc_step_rad,
                         double r_mid_m, double dr_m, int bdry_format, int m_harm, double depth_m, double *torque,
                         double *br_bt);

/* ------------------------------------------------------------------ complex-symmetric linear problem
 * Values, right-hand side and solution are interleaved (re,im) doubles — layout of the reference's
 * CComplex {double re, im} (cfemm/libfemm/liblua/femmcomplex.h:31).  n_nodes = number of mesh
 * unknowns; rows n_nodes..n-1 are circuit unknowns (always inside the SetValue scan range,
 * cspars.cpp:500-503). */
int xfemm_clinprob_create(int n, int n_nodes, const int *rowptr, const int *colidx, xfemm_clinprob **out);
int xfemm_clinprob_destroy(xfemm_clinprob *p);
int xfemm_clinprob_n(const xfemm_clinprob *p);
int xfemm_clinprob_amg_levels(xfemm_clinprob *p, int max_levels, int *n_rows, long long *nnz); /* of the real companion */
long long xfemm_clinprob_nnz(const xfemm_clinprob *p);
int xfemm_clinprob_wipe(xfemm_clinprob *p);                                /* cspars.cpp:539 Wipe */
int xfemm_clinprob_set_values(xfemm_clinprob *p, const double *vals_csr);
int xfemm_clinprob_get_values(xfemm_clinprob *p, double *vals_csr);
int xfemm_clinprob_set_rhs(xfemm_clinprob *p, const double *b);
int xfemm_clinprob_get_rhs(xfemm_clinprob *p, double *b);
int xfemm_clinprob_set_solution(xfemm_clinprob *p, const double *V);
int xfemm_clinprob_get_solution(xfemm_clinprob *p, double *V);
int xfemm_clinprob_dirichlet(xfemm_clinprob *p, int count, const int *nodes, const double *x_first,
                             const double *x_last, int bdw);                /* cspars.cpp:482 SetValue */
int xfemm_clinprob_periodic(xfemm_clinprob *p, int npairs, const int *ni, const int *nj, const int *type); /* :677 */
int xfemm_clinprob_set_fill_mask(xfemm_clinprob *p, const unsigned char *fill);
int xfemm_clinprob_get_exists_mask(xfemm_clinprob *p, unsigned char *exists);
/* cspars.cpp:1062 PBCGSolveMod for the complex-symmetric case: cold start = 3 CG steps on the
 * normal equations (PCGSQStart, :764) then COCG (PBCGSolve, :822); stop on ||r||/||b|| <= precision */
/* builds what depends on the pattern alone (layout; for XFEMM_PRECOND_AMG also the real companion problem and its
 * multigrid hierarchy) ahead of the first solve; safe to call from another host thread while assembly runs */
int xfemm_clinprob_prebuild(xfemm_clinprob *p, int precond);
int xfemm_clinprob_cocg(xfemm_clinprob *p, double precision, int warm, int precond, int max_iter,
                        xfemm_solve_info *info);
int xfemm_clinprob_save_solution(xfemm_clinprob *p);
int xfemm_clinprob_newton_norms(xfemm_clinprob *p, int count, double *dv2, double *v2); /* first `count` rows */
int xfemm_clinprob_relax(xfemm_clinprob *p, double relax);


/* ------------------------------------------------------------------ Newton AC solver ([ACSolver]=1)
 * Replaces the k=1,2,3 selector of CBigComplexLinProb::Put/Get/MultA/MultConjA (cspars.cpp:146-390), the side-matrix
 * branches of SetValue/Wipe/Periodicity/AntiPeriodicity (cspars.cpp:511-535, 556-590, 648-668, 732-752) and
 * KludgeSolve (cspars.cpp:983-1056).  The side matrices Mh (k=1, Hermitian), Ms (k=2, complex symmetric, applied to
 * conj(x)) and Ma (k=3, anti-Hermitian) are stored on the full CSR pattern of the problem, nnz complex values each. */
int xfemm_clinprob_enable_newton(xfemm_clinprob *p);
int xfemm_clinprob_set_side_values(xfemm_clinprob *p, int k, const double *v);
int xfemm_clinprob_get_side_values(xfemm_clinprob *p, int k, double *v);
/* y = G x (conj_g = 0) or conj(G) x for one stored matrix, k = 0 (M), 1, 2, 3: MultA / MultConjA with a selector */
int xfemm_clinprob_mult(xfemm_clinprob *p, int k, int conj_g, const double *x, double *y);
/* y = M x + Mh x + Ms conj(x) + Ma x  (MultA(X,Y,-1)); host vectors of n complex */
int xfemm_clinprob_mult_newton(xfemm_clinprob *p, const double *x, double *y);
/* KludgeSolve: <= 10 COCG solves on M with the side terms on the right-hand side, each followed by a line search on
 * the full residual; stops when ||r||/||b|| < 10*precision.  info->iterations sums the inner COCG iterations. */
int xfemm_clinprob_kludge_solve(xfemm_clinprob *p, double precision, int warm, int precond, int max_iter, xfemm_solve_info *info);

/* ------------------------------------------------------------------ time-harmonic assembly */
typedef struct xfemm_cmag_label
{
    double mu1[2], mu2[2]; /* complex relative permeabilities of the first pass (lamination / proximity models applied) */
    double J[2];           /* J_block + a-priori circuit current density [MA/m^2] (harmonic2d.cpp:518-530)             */
    double dVs[2];         /* axisymmetric a-priori voltage gradient term, divided by R on the device                 */
    double sigma_eddy;     /* conductivity of the eddy-current mass term (0: wound or laminated block)               */
    double sigma_circ;     /* conductivity coupling the element to its Case-2 circuit unknown                        */
    int circuit;           /* Case 2: index k of the circuit unknown (row n_nodes+k), else -1                        */
    int bh_offset, bh_points, lamtype, is_external, pad_;
} xfemm_cmag_label;

typedef struct xfemm_cmag_desc
{
    int axisymmetric;
    int n_nodes, n_elems;
    double omega;                 /* 2 pi f */
    const double *xy;             /* n_nodes*2, cm, solver order */
    const int *tri;               /* n_elems*4: p0,p1,p2,label */
    int n_labels;
    const xfemm_cmag_label *labels;
    int n_bh;
    const double *bh_B, *bh_H, *bh_H_im, *bh_slope, *bh_slope_im;
    /* boundary sides of type 2 (mixed) and 1 (small skin depth): element, local side and 7 doubles:
     * type, c0.re, c0.im, c1.re, c1.im, Kssd.re, Kssd.im (Kssd = (1+j)/(-ds*Mu*100), harmonic2d.cpp:502-514) */
    int n_sides;
    const int *side_elem_side;
    const double *side_coeffs;
    int n_age_cells;
    const int *age_nodes;
    const double *age_weights;
    const int *age_matrix_id;
    int n_age_matrices;
    const double *age_matrices;
    /* additive right-hand-side terms: point currents (-0.01 J) and total-current constraints of Case-2 circuits */
    int n_rhs_add;
    const int *rhs_add_row;
    const double *rhs_add_val;    /* complex */
    /* circuit rows whose diagonal is set to A(0,0) after the Dirichlet conditions (Case < 2) */
    int n_fix_diag;
    const int *fix_diag_rows;
    double ext_zo, ext_ro, ext_ri;
} xfemm_cmag_desc;

typedef struct xfemm_cmag xfemm_cmag;
int xfemm_cmag_create(const xfemm_cmag_desc *desc, xfemm_clinprob *lin, xfemm_cmag **out);
int xfemm_cmag_destroy(xfemm_cmag *m);
int xfemm_cmag_set_labels(xfemm_cmag *m, int n_labels, const xfemm_cmag_label *labels);
/* one assembly pass (Wipe, AGE blocks with the harmonic sign, elements, boundary sides, RHS terms);
 * iter>0 re-evaluates the permeability of B-H elements by successive approximation, or — after
 * xfemm_cmag_set_newton(m,1), which needs xfemm_clinprob_enable_newton on the problem — builds the Newton
 * matrices Mn (into M) and Mnh/Mns/Mna (into the side matrices), harmonic2d.cpp:611-636, harmonicaxi.cpp:520-548 */
int xfemm_cmag_assemble(xfemm_cmag *m, int iter, double *elapsed_ms);
int xfemm_cmag_set_newton(xfemm_cmag *m, int on);
int xfemm_cmag_fix_circuit_diagonals(xfemm_cmag *m);
int xfemm_cmag_get_mu(xfemm_cmag *m, double *mu4);               /* n_elems*4: mu1.re, mu1.im, mu2.re, mu2.im */
int xfemm_cmag_solution_to_A(xfemm_cmag *m, double *A_host);      /* n complex: A (nodes) and circuit results  */

/* ------------------------------------------------------------------ row-block partitioned PCG (one mesh, several GPUs)
 * Every rank holds the assembled, boundary-conditioned global problem `g` on its GPU and owns rows
 * [r0,r1) of it.  The caller (one process per GPU) moves the halo of the search direction and
 * all-reduces the scalars with torch.distributed/NCCL over NVLink; buffers named *_dev are DEVICE
 * pointers.  glob_dev = 3 doubles on the device: [res, pAp, res_new], already all-reduced. */
typedef struct xfemm_dist xfemm_dist;
int xfemm_dist_create(xfemm_linprob *g, int r0, int r1, xfemm_dist **out);
int xfemm_dist_destroy(xfemm_dist *d);
/* XFEMM_PRECOND_JACOBI (default) or XFEMM_PRECOND_AMG: multigrid V-cycle of each rank's own diagonal block
 * (additive Schwarz without overlap, no communication inside the cycle); call before xfemm_dist_refresh */
int xfemm_dist_set_precond(xfemm_dist *d, int precond);
int xfemm_dist_sizes(xfemm_dist *d, int *n_own, int *n_halo);
int xfemm_dist_halo_columns(xfemm_dist *d, int *global_cols);              /* sorted global ids this rank needs */
int xfemm_dist_set_send_rows(xfemm_dist *d, int n_send, const int *global_rows);
int xfemm_dist_refresh(xfemm_dist *d);                                     /* pull owned rows of values + RHS from g */
int xfemm_dist_init(xfemm_dist *d, double *scal_dev);                      /* x=0, r=b, p=D^-1 r ; scal = local r.D^-1 r */
int xfemm_dist_pack(xfemm_dist *d, double *sendbuf_dev);
int xfemm_dist_unpack(xfemm_dist *d, const double *recvbuf_dev);
int xfemm_dist_spmv_dot(xfemm_dist *d, double *scal_dev);                  /* u = A p ; scal = local p.u */
int xfemm_dist_update(xfemm_dist *d, const double *glob_dev, double *scal_dev); /* x,r update ; scal = local r.D^-1 r */
int xfemm_dist_direction(xfemm_dist *d, const double *glob_dev);           /* p = D^-1 r + (res_new/res) p */
int xfemm_dist_get_solution(xfemm_dist *d, double *x_own);
unsigned long long xfemm_dist_stream(xfemm_dist *d);                       /* cudaStream_t the kernels run on */

/* ------------------------------------------------------------------ one mesh over several GPUs, multigrid-preconditioned
 * (csrc/dist_amg.cu).  The recurrence and the stopping rule are CBigLinProb::PCGSolve's (cfemm/libfemm/spars.cpp:238-316),
 * the preconditioner is the SAME multigrid hierarchy as on one GPU (iteration counts do not depend on the rank
 * count).  One process per GPU; every rank holds the assembled, boundary-conditioned problem `p` and works on a
 * contiguous share of the rows of the big levels; halo entries and the two scalars of an iteration travel as peer
 * stores + flags over NVLink inside ONE persistent kernel per GPU (no collective call, no host in the loop).
 *   create  : plan (send lists) + symmetric communication arena of this rank
 *   ipc_handle / connect : exchange of the 64-byte CUDA IPC handles of the arenas through any channel the host
 *             has (torch.distributed, MPI ...), all `world` handles in rank order;  connect_files: rendezvous
 *             through files <dir>/<job>.<rank> for hosts without one (the C++ front end)
 *   solve   : every rank, same arguments; the kernels wait for each other on the device (deadline
 *             XFEMM_DIST_TIMEOUT_MS, default 20 s: a lost rank gives XFEMM_ERR_DIVERGED, not a hung GPU).
 *             info->solve_ms = the persistent kernel (CUDA events)
 *   share_solution, [host barrier], finish : own rows of x to every rank, then the complete solution in p's V */
typedef struct xfemm_dpcg xfemm_dpcg;
int xfemm_dpcg_create(xfemm_linprob *p, int rank, int world, xfemm_dpcg **out);
int xfemm_dpcg_destroy(xfemm_dpcg *d);
int xfemm_dpcg_ipc_handle(xfemm_dpcg *d, void *handle64);
int xfemm_dpcg_connect(xfemm_dpcg *d, const void *handles /* world x 64 bytes */);
int xfemm_dpcg_connect_files(xfemm_dpcg *d, const char *dir, const char *job, int timeout_s);
int xfemm_dpcg_stats(xfemm_dpcg *d, int *partitioned_levels, long long *list_entries /* [12] */, long long *arena_bytes);
int xfemm_dpcg_solve(xfemm_dpcg *d, double precision, int warm, int max_iter, xfemm_solve_info *info);
int xfemm_dpcg_share_solution(xfemm_dpcg *d);
int xfemm_dpcg_finish(xfemm_dpcg *d);

/* ------------------------------------------------------------------ file-level entry
 * The call a user of fsolver / mexfsolver makes: basename without ".fem";
 * returns 0 ok, 1 problem loading .fem, 2 solve failed (main.cpp:72-83, mexfsolver.cpp:122-134). */
int xfemm_fsolver_run(const char *basename, int verbose);
/* finer-grained variant used by the tests and the benchmark */
int xfemm_fsolver_open(const char *basename, int verbose, xfemm_fsolver **out); /* LoadProblemFile */
int xfemm_fsolver_load_mesh(xfemm_fsolver *s, int delete_files);                /* LoadMesh        */
int xfemm_fsolver_renumber(xfemm_fsolver *s, int delete_files);                 /* Cuthill         */
int xfemm_fsolver_solve(xfemm_fsolver *s);     /* Static2D / StaticAxisymmetric / Harmonic*       */
/* re-runs assembly + boundary conditions + solve (the whole Newton loop) on the device objects
 * left resident by xfemm_fsolver_solve: no host set-up, no mesh upload */
int xfemm_fsolver_resolve(xfemm_fsolver *s);
/* Rotor-angle sweeps without re-meshing: regenerates the ring tables of air-gap element
 * `age_index` for new inner (rotor) / outer (stator) angles in degrees, as fmesher would write
 * them into the .pbc file (cfemm/fmesher/writepoly.cpp:1853-1966).  Follow with
 * xfemm_fsolver_solve, or with xfemm_fsolver_resolve if the problem was solved with option
 * "age_union_pattern" = 1 (sparsity pattern independent of the angle).
 * Other options of xfemm_fsolver_set_option: "fetch_solution" = 0 (leave V and A on the device after a
 * solve), "assemble_only" = 1 (stop after assembly + boundary conditions), "locality_order" = 1/0/-1
 * (Z-curve row order inside the solver; -1 = only when the problem is going to be re-solved),
 * "warm_resolve" = 0 (xfemm_fsolver_resolve starts from zero instead of the solution left on the device),
 * "cache_static_part" = 0 (re-assemble and re-condition everything at every position; by default a linear problem
 * solved with "age_union_pattern" keeps its position-independent, boundary-conditioned part on the device and a
 * position only adds its air-gap blocks — used when no air-gap node is a fixed or periodic node). */
int xfemm_fsolver_set_rotor_angle(xfemm_fsolver *s, int age_index, double inner_angle_deg, double outer_angle_deg);
/* torque [N m] through air-gap element `age_index` from the resident magnetostatic solution
 * (= mo_gapintegral(name,0) of the reference's post-processor on the written .ans) */
int xfemm_fsolver_gap_torque(xfemm_fsolver *s, int age_index, double *torque);
int xfemm_fsolver_write(xfemm_fsolver *s);
/* Derived quantities of a planar magnetostatic solution from the device, = what the reference obtains by writing the
 * .ans file and asking its post-processor: stored energy [J] of all blocks (mo_groupselectblock(); mo_blockintegral(2))
 * and flux linkage [Wb] of circuit `circuit` of the .fem file (third value of mo_getcircuitproperties; circuits that
 * carry a current: FPProc::GetFluxLinkage's first branch, fpproc.cpp:5166-5196). */
/* rows / non-zeros of the levels of the multigrid hierarchy of the resident problem (real, or the real companion of a
 * time-harmonic problem); returns the number of levels, 0 if there is none */
int xfemm_fsolver_amg_levels(xfemm_fsolver *s, int max_levels, int *n_rows, long long *nnz);
int xfemm_fsolver_stored_energy(xfemm_fsolver *s, double *joules);
int xfemm_fsolver_flux_linkage(xfemm_fsolver *s, int circuit, double *webers);     /* WriteStatic2D / WriteHarmonic2D                  */
int xfemm_fsolver_close(xfemm_fsolver *s);
int xfemm_fsolver_counts(xfemm_fsolver *s, int *n_nodes, int *n_elems, int *n_pbc, int *n_age,
                         int *bandwidth, int *n_unknowns, int *is_complex);
xfemm_linprob *xfemm_fsolver_linprob(xfemm_fsolver *s);                  /* resident real system (owned by s), or NULL */
long long xfemm_fsolver_nnz(xfemm_fsolver *s);                           /* full-CSR nonzeros of the resident system */
int xfemm_fsolver_get_nodes(xfemm_fsolver *s, double *xy, int *bc);      /* solver order, cm      */
int xfemm_fsolver_get_elems(xfemm_fsolver *s, int *elems8);              /* p0 p1 p2 lbl blk e0-2 */
int xfemm_fsolver_get_pbc(xfemm_fsolver *s, int *pbc3);
int xfemm_fsolver_get_permutation(xfemm_fsolver *s, int *newnum);
int xfemm_fsolver_get_pattern(xfemm_fsolver *s, long long *nnz, int *rowptr, int *colidx); /* colidx NULL → only nnz */
/* value [deg] of a block label's magnetisation-direction expression (`MagDirFctn`) at the point (x, y) in problem
 * length units: what the reference obtains from its Lua interpreter per element (static2d.cpp:513-583).  Host only. */
int xfemm_magdir_eval(const char *expr, double x, double y, double *out);
int xfemm_fsolver_get_system(xfemm_fsolver *s, double *vals_csr, double *rhs); /* last assembled+BC'd system */
/* Newton AC ([ACSolver]=1, keep_system=1): side matrix k=1 (Mh), 2 (Ms), 3 (Ma) of that system on the same slots */
int xfemm_fsolver_get_side_matrix(xfemm_fsolver *s, int k, double *vals_csr);
int xfemm_fsolver_get_solution(xfemm_fsolver *s, double *V, double *A);  /* complex: 2 doubles/entry */
int xfemm_fsolver_get_mu(xfemm_fsolver *s, double *mu12);
/* stats: newton steps, per-step Krylov iterations and device ms (arrays sized >= 64) */
int xfemm_fsolver_get_stats(xfemm_fsolver *s, int *newton_steps, int *iters, double *solve_ms,
                            double *assemble_ms, double *relres);
int xfemm_fsolver_set_option(xfemm_fsolver *s, const char *key, double value); /* "precond", "max_iter", "keep_system" */

#ifdef __cplusplus
}
#endif
#endif
