// xfemm_fsolver_* — file-level C entry points around the FSolver class: what cfemm/fsolver/main.cpp
// and mfemm/mexfsolver.cpp do with the reference class, exposed with plain C types so that
// ctypes / MEX / Lua bindings (and this repo's tests and bench) can drive the same path.
#include "fsolver.h"
#include "magdir_expr.h"
#include "../../include/xfemm_b200.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace xb
{
void set_error(const char *fmt, ...);
void age_cell_matrix(double ci, double co, double K, double *MG);
}

namespace xb { void tune_host_allocator(); } // csrc/linprob.cu: big host blocks stay in the heap between solves

extern "C" int xfemm_age_cell_matrix(double ci, double co, double K, double *MG100)
{
    if (!MG100 || K == 0) { xb::set_error("invalid arguments"); return XFEMM_ERR_ARG; }
    xb::age_cell_matrix(ci, co, K, MG100);
    return XFEMM_OK;
}

struct xfemm_fsolver
{
    FSolver S;
    CBigLinProb *L = nullptr;
    CBigComplexLinProb *Lc = nullptr;
    bool verbose = false;
    ~xfemm_fsolver() { S.releaseDevice(); delete L; delete Lc; }
};

static int quiet(const char *, ...) { return 0; }

extern "C" {

int xfemm_fsolver_run(const char *basename, int verbose)
{
    if (!basename) return 1;
    FSolver S;
    S.PathName = basename;
    if (!verbose) { S.WarnMessage = quiet; S.PrintMessage = quiet; }
    if (!S.LoadProblemFile()) return 1;
    if (!S.runSolver(verbose != 0)) return 2;
    return 0;
}

int xfemm_fsolver_open(const char *basename, int verbose, xfemm_fsolver **out)
{
    xb::tune_host_allocator();
    if (!basename || !out) { xb::set_error("null argument"); return XFEMM_ERR_ARG; }
    xfemm_fsolver *s = new xfemm_fsolver;
    s->verbose = verbose != 0;
    s->S.PathName = basename;
    if (!verbose) { s->S.WarnMessage = quiet; s->S.PrintMessage = quiet; }
    if (!s->S.LoadProblemFile())
    {
        delete s;
        xb::set_error("problem loading %s.fem", basename);
        return XFEMM_ERR_IO;
    }
    *out = s;
    return XFEMM_OK;
}

int xfemm_fsolver_close(xfemm_fsolver *s) { delete s; return XFEMM_OK; }

int xfemm_fsolver_load_mesh(xfemm_fsolver *s, int delete_files)
{
    if (!s) return XFEMM_ERR_ARG;
    LoadMeshErr e = s->S.LoadMesh(delete_files != 0);
    if (e != NOERROR) { xb::set_error("%s", FSolver::getErrorString(e).c_str()); return XFEMM_ERR_IO; }
    return XFEMM_OK;
}

int xfemm_fsolver_renumber(xfemm_fsolver *s, int delete_files)
{
    if (!s) return XFEMM_ERR_ARG;
    if (s->S.meshLoadedFromPrevSolution) return XFEMM_OK; // the mesh of a [PrevSoln] file is in solver numbering already (fsolver.cpp:1224)
    if (!s->S.Cuthill(delete_files != 0)) { xb::set_error("problem renumbering node points"); return XFEMM_ERR_IO; }
    return XFEMM_OK;
}

int xfemm_fsolver_solve(xfemm_fsolver *s)
{
    if (!s) return XFEMM_ERR_ARG;
    FSolver &S = s->S;
    int ok;
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_start = now();
    if (S.Frequency == 0)
    {
        S.releaseDevice();
        delete s->L;
        if (timing) fprintf(stderr, "[xfemm timing] release of the previous device problem %.3f s\n", now() - t_start);
        s->L = new CBigLinProb;
        s->L->Precision = S.Precision;
        s->L->Create(S.NumNodes, S.BandWidth);
        ok = (S.ProblemType == femm::PLANAR) ? S.Static2D(*s->L) : S.StaticAxisymmetric(*s->L);
        if (timing) fprintf(stderr, "[xfemm timing] xfemm_fsolver_solve total %.3f s\n", now() - t_start);
    }
    else
    {
        S.releaseDevice();
        delete s->Lc;
        s->Lc = new CBigComplexLinProb;
        s->Lc->Precision = S.Precision;
        s->Lc->Create(S.NumNodes + S.NumCircProps, S.BandWidth, S.NumNodes);
        ok = (S.ProblemType == femm::PLANAR) ? S.Harmonic2D(*s->Lc, s->verbose) : S.HarmonicAxisymmetric(*s->Lc, s->verbose);
    }
    if (ok != 1)
    {
        std::string prev = xfemm_last_error();
        xb::set_error("solve failed (%s)", prev.c_str());
        return XFEMM_ERR_DIVERGED;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_gap_torque(xfemm_fsolver *s, int age_index, double *torque)
{
    if (!s || !torque) return XFEMM_ERR_ARG;
    if (!s->S.gapTorque(age_index, *torque))
    {
        std::string prev = xfemm_last_error();
        xb::set_error("gap torque failed (%s)", prev.c_str());
        return XFEMM_ERR_ARG;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_stored_energy(xfemm_fsolver *s, double *joules)
{
    if (!s || !joules) return XFEMM_ERR_ARG;
    if (!s->S.storedEnergy(*joules))
    {
        std::string prev = xfemm_last_error();
        xb::set_error("stored energy failed (%s)", prev.c_str());
        return XFEMM_ERR_ARG;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_flux_linkage(xfemm_fsolver *s, int circuit, double *webers)
{
    if (!s || !webers) return XFEMM_ERR_ARG;
    if (!s->S.fluxLinkage(circuit, *webers))
    {
        std::string prev = xfemm_last_error();
        xb::set_error("flux linkage failed (%s)", prev.c_str());
        return XFEMM_ERR_ARG;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_resolve(xfemm_fsolver *s)
{
    if (!s || !s->L) { xb::set_error("no resident static problem (call xfemm_fsolver_solve first)"); return XFEMM_ERR_ARG; }
    // the solution of the previous position is still on the device: it is the initial guess of this solve
    s->S.resolveStart = s->S.warmResolve;
    const int ok = s->S.staticNewton(*s->L);
    s->S.resolveStart = false;
    if (ok != 1)
    {
        std::string prev = xfemm_last_error();
        xb::set_error("solve failed (%s)", prev.c_str());
        return XFEMM_ERR_DIVERGED;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_set_rotor_angle(xfemm_fsolver *s, int age_index, double inner_angle_deg, double outer_angle_deg)
{
    if (!s) return XFEMM_ERR_ARG;
    if (!s->S.setAirGapAngles(age_index, inner_angle_deg, outer_angle_deg))
    {
        xb::set_error("cannot re-target air-gap element %d", age_index);
        return XFEMM_ERR_ARG;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_write(xfemm_fsolver *s)
{
    if (!s) return XFEMM_ERR_ARG;
    int ok = 0;
    // a sweep that kept the nodal values on the device (option fetch_solution = 0): bring them over now rather than
    // write the previous position's field
    if (s->L && !s->S.fetchDeviceSolution(*s->L)) { xb::set_error("could not fetch the solution from the device"); return XFEMM_ERR_CUDA; }
    if (s->L) ok = s->S.WriteStatic2D(*s->L);
    else if (s->Lc) ok = s->S.WriteHarmonic2D(*s->Lc);
    if (!ok) { xb::set_error("couldn't write results to disk"); return XFEMM_ERR_IO; }
    return XFEMM_OK;
}

int xfemm_fsolver_counts(xfemm_fsolver *s, int *n_nodes, int *n_elems, int *n_pbc, int *n_age, int *bandwidth,
                         int *n_unknowns, int *is_complex)
{
    if (!s) return XFEMM_ERR_ARG;
    const FSolver &S = s->S;
    if (n_nodes) *n_nodes = S.NumNodes;
    if (n_elems) *n_elems = S.NumEls;
    if (n_pbc) *n_pbc = S.NumPBCs;
    if (n_age) *n_age = S.NumAirGapElems;
    if (bandwidth) *bandwidth = S.BandWidth;
    if (is_complex) *is_complex = S.Frequency != 0;
    if (n_unknowns) *n_unknowns = S.Frequency != 0 ? S.NumNodes + S.NumCircProps : S.NumNodes;
    return XFEMM_OK;
}

xfemm_linprob *xfemm_fsolver_linprob(xfemm_fsolver *s) { return (s && s->L) ? s->L->device() : nullptr; }

long long xfemm_fsolver_nnz(xfemm_fsolver *s)
{
    if (s && s->Lc && s->Lc->device()) return xfemm_clinprob_nnz(s->Lc->device());
    if (!s || !s->L || !s->L->device()) return 0;
    return xfemm_linprob_nnz(s->L->device());
}

int xfemm_fsolver_amg_levels(xfemm_fsolver *s, int max_levels, int *n_rows, long long *nnz)
{
    if (s && s->L && s->L->device()) return xfemm_linprob_amg_levels(s->L->device(), max_levels, n_rows, nnz);
    if (s && s->Lc && s->Lc->device()) return xfemm_clinprob_amg_levels(s->Lc->device(), max_levels, n_rows, nnz);
    return 0;
}

int xfemm_fsolver_get_nodes(xfemm_fsolver *s, double *xy, int *bc)
{
    if (!s) return XFEMM_ERR_ARG;
    for (int i = 0; i < s->S.NumNodes; i++)
    {
        if (xy) { xy[2 * (size_t)i] = s->S.meshnode[i].x; xy[2 * (size_t)i + 1] = s->S.meshnode[i].y; }
        if (bc) bc[i] = s->S.meshnode[i].BoundaryMarker;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_get_elems(xfemm_fsolver *s, int *e8)
{
    if (!s || !e8) return XFEMM_ERR_ARG;
    for (int i = 0; i < s->S.NumEls; i++)
    {
        const femmsolver::CMElement &e = s->S.meshele[i];
        int *r = e8 + 8 * (size_t)i;
        r[0] = e.p[0]; r[1] = e.p[1]; r[2] = e.p[2]; r[3] = e.lbl; r[4] = e.blk; r[5] = e.e[0]; r[6] = e.e[1]; r[7] = e.e[2];
    }
    return XFEMM_OK;
}

int xfemm_fsolver_get_pbc(xfemm_fsolver *s, int *pbc3)
{
    if (!s || !pbc3) return XFEMM_ERR_ARG;
    for (int i = 0; i < s->S.NumPBCs; i++)
    {
        pbc3[3 * i] = s->S.pbclist[i].x; pbc3[3 * i + 1] = s->S.pbclist[i].y; pbc3[3 * i + 2] = s->S.pbclist[i].t;
    }
    return XFEMM_OK;
}

int xfemm_fsolver_get_permutation(xfemm_fsolver *s, int *newnum)
{
    if (!s || !newnum || (int)s->S.cmPermutation.size() != s->S.NumNodes) { xb::set_error("no permutation available"); return XFEMM_ERR_ARG; }
    memcpy(newnum, s->S.cmPermutation.data(), sizeof(int) * s->S.NumNodes);
    return XFEMM_OK;
}

int xfemm_fsolver_get_pattern(xfemm_fsolver *s, long long *nnz, int *rowptr, int *colidx)
{
    if (!s || s->S.sysRowptr.empty()) { xb::set_error("no system kept: set option keep_system=1 before solving"); return XFEMM_ERR_ARG; }
    if (nnz) *nnz = (long long)s->S.sysCol.size();
    if (rowptr) memcpy(rowptr, s->S.sysRowptr.data(), sizeof(int) * s->S.sysRowptr.size());
    if (colidx) memcpy(colidx, s->S.sysCol.data(), sizeof(int) * s->S.sysCol.size());
    return XFEMM_OK;
}

int xfemm_magdir_eval(const char *expr, double x, double y, double *out)
{
    if (!expr || !out) { xb::set_error("expression and output pointer"); return XFEMM_ERR_ARG; }
    xbhost::MagDirExpr e;
    std::string err;
    if (!e.compile(expr, err)) { xb::set_error("%s", err.c_str()); return XFEMM_ERR_ARG; }
    const double PI_ = 3.141592653589793238462643383;
    const xbhost::MagDirVars v = {x, y, atan2(y, x) * 180. / PI_, sqrt(x * x + y * y)};
    if (!e.eval(v, *out)) { xb::set_error("\"%s\" does not evaluate to a numerical value", expr); return XFEMM_ERR_ARG; }
    return XFEMM_OK;
}

int xfemm_fsolver_get_system(xfemm_fsolver *s, double *vals, double *rhs)
{
    if (!s || s->S.sysVals.empty()) { xb::set_error("no system kept: set option keep_system=1 before solving"); return XFEMM_ERR_ARG; }
    if (vals) memcpy(vals, s->S.sysVals.data(), sizeof(double) * s->S.sysVals.size());
    if (rhs) memcpy(rhs, s->S.sysRhs.data(), sizeof(double) * s->S.sysRhs.size());
    return XFEMM_OK;
}

int xfemm_fsolver_get_side_matrix(xfemm_fsolver *s, int k, double *vals)
{
    if (!s || k < 1 || k > 3 || !vals) { xb::set_error("handle, k in 1..3, values"); return XFEMM_ERR_ARG; }
    if (s->S.sysSide[k - 1].empty()) { xb::set_error("no side matrices kept: [ACSolver]=1 and option keep_system=1 are needed"); return XFEMM_ERR_ARG; }
    memcpy(vals, s->S.sysSide[k - 1].data(), sizeof(double) * s->S.sysSide[k - 1].size());
    return XFEMM_OK;
}

int xfemm_fsolver_get_solution(xfemm_fsolver *s, double *V, double *A)
{
    if (!s) return XFEMM_ERR_ARG;
    if (s->L)
    {
        if (!s->S.fetchDeviceSolution(*s->L)) { xb::set_error("could not fetch the solution from the device"); return XFEMM_ERR_CUDA; }
        if (V) memcpy(V, s->L->V, sizeof(double) * s->L->n);
        if (A) memcpy(A, s->L->b, sizeof(double) * s->L->n);
        return XFEMM_OK;
    }
    if (s->Lc)
    {
        if (V) memcpy(V, s->Lc->V, sizeof(CComplex) * s->Lc->n);
        if (A) memcpy(A, s->Lc->b, sizeof(CComplex) * s->Lc->n);
        return XFEMM_OK;
    }
    xb::set_error("no solution available");
    return XFEMM_ERR_ARG;
}

int xfemm_fsolver_get_mu(xfemm_fsolver *s, double *mu12)
{
    if (s && s->S.cmagdev) return xfemm_cmag_get_mu(s->S.cmagdev, mu12); // 4 doubles per element
    if (!s || !s->S.magdev) { xb::set_error("no device mesh"); return XFEMM_ERR_ARG; }
    return xfemm_mag_get_mu(s->S.magdev, mu12);
}

int xfemm_fsolver_get_stats(xfemm_fsolver *s, int *newton_steps, int *iters, double *solve_ms, double *assemble_ms,
                            double *relres)
{
    if (!s) return XFEMM_ERR_ARG;
    const FSolverStats &st = s->S.stats;
    if (newton_steps) *newton_steps = st.newton_iters;
    for (size_t k = 0; k < st.pcg_iters.size() && k < 64; k++)
    {
        if (iters) iters[k] = st.pcg_iters[k];
        if (solve_ms) solve_ms[k] = st.pcg_ms[k];
        if (assemble_ms && k < st.assemble_ms.size()) assemble_ms[k] = st.assemble_ms[k];
    }
    if (relres) *relres = st.final_relres;
    return XFEMM_OK;
}

int xfemm_fsolver_set_option(xfemm_fsolver *s, const char *key, double value)
{
    if (!s || !key) return XFEMM_ERR_ARG;
    if (!strcmp(key, "precond")) { s->S.precondOption = (int)value; return XFEMM_OK; }
    if (!strcmp(key, "max_iter")) { s->S.maxIterOption = (int)value; return XFEMM_OK; }
    if (!strcmp(key, "keep_system")) { s->S.keepSystem = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "locality_order")) { s->S.localityOrder = (int)value; return XFEMM_OK; }
    if (!strcmp(key, "extrapolate_resolve")) { s->S.extrapolateResolve = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "warm_resolve")) { s->S.warmResolve = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "fetch_solution")) { s->S.fetchSolution = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "cache_static_part")) { s->S.cacheStaticPart = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "assemble_only")) { s->S.assembleOnly = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "assemble_iter")) { s->S.assembleIter = (int)value; return XFEMM_OK; }
    if (!strcmp(key, "age_union_pattern")) { s->S.ageUnionPattern = value != 0; return XFEMM_OK; }
    if (!strcmp(key, "deterministic_assembly"))
    {
        s->S.deterministicAssembly = value != 0;
        if (s->S.magdev) xfemm_mag_set_deterministic(s->S.magdev, value != 0); // a resident problem switches at once
        return XFEMM_OK;
    }
    if (!strcmp(key, "dist_rank")) { s->S.distRank = (int)value; return XFEMM_OK; }
    if (!strcmp(key, "dist_world")) { s->S.distWorld = (int)value; return XFEMM_OK; }
    xb::set_error("unknown option %s", key);
    return XFEMM_ERR_ARG;
}

} // extern "C"
