// ref_harness — white-box driver around the UNMODIFIED reference objects (oracle/_ref/libref.a).
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md): used by tests/ to produce and check golden
// data, and by bench.py's cpu_baseline / --impl reference legs.  Never linked by the product.
//
// How it observes the reference without patching it: the link line uses GNU ld's --wrap on
//   CBigLinProb::PCGSolve(int)                 (cfemm/libfemm/spars.cpp:238)
//   CBigLinProb::Periodicity/AntiPeriodicity   (cfemm/libfemm/spars.cpp:421 / 366)
//   CBigComplexLinProb::PBCGSolveMod(int,bool) (cfemm/libfemm/cspars.cpp:1062)
// so the calls made from FSolver::Static2D etc. (cfemm/fsolver/static2d.cpp:934-948) land in the
// __wrap_ functions below, which snapshot the public members of L (pattern, values, RHS), time
// the phases and then forward to the real routine (__real_*).  Iteration counts are obtained by
// replaying the textbook PCG recurrence with the reference's own public MultA/MultPC/Dot.
//
// Modes
//   ref_harness dump <basename> <outdir> [maxsteps]
//        needs <basename>.fem/.node/.ele/.edge/.pbc (kept).  Writes .npy files + meta.json.
//   ref_harness time <basename> [pcg_iters] [pbc_pairs]
//        CPU baseline: full assembly + Dirichlet, a sample of `pbc_pairs` periodic pairs and
//        `pcg_iters` SSOR-PCG iterations (both <=0 → everything); prints one JSON line.
#include "fsolver.h"
#include "spars.h"
#include "cspars.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <sys/stat.h>
#include <unistd.h>

using Clock = std::chrono::steady_clock;
static double now_s() { return std::chrono::duration<double>(Clock::now().time_since_epoch()).count(); }

// ---------------------------------------------------------------- tiny .npy writer
static void npy_write(const std::string &path, const char *descr, size_t itemsize,
                      const void *data, size_t rows, size_t cols)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(9); }
    char hdr[256];
    int n;
    if (cols == 0)
        n = snprintf(hdr, sizeof hdr, "{'descr': '%s', 'fortran_order': False, 'shape': (%zu,), }", descr, rows);
    else
        n = snprintf(hdr, sizeof hdr, "{'descr': '%s', 'fortran_order': False, 'shape': (%zu, %zu), }", descr, rows, cols);
    size_t total = 10 + n + 1;
    size_t pad = (64 - total % 64) % 64;
    unsigned short hlen = (unsigned short)(n + pad + 1);
    fwrite("\x93NUMPY\x01\x00", 1, 8, f);
    fwrite(&hlen, 2, 1, f);
    fwrite(hdr, 1, n, f);
    for (size_t i = 0; i < pad; i++) fputc(' ', f);
    fputc('\n', f);
    fwrite(data, itemsize, rows * (cols ? cols : 1), f);
    fclose(f);
}
static void npy_f64(const std::string &p, const double *d, size_t r, size_t c = 0) { npy_write(p, "<f8", 8, d, r, c); }
static void npy_i32(const std::string &p, const int *d, size_t r, size_t c = 0) { npy_write(p, "<i4", 4, d, r, c); }
static void npy_c128(const std::string &p, const CComplex *d, size_t r, size_t c = 0) { npy_write(p, "<c16", 16, d, r, c); }

// ---------------------------------------------------------------- harness state
struct Harness
{
    bool dumping = false;
    bool timing = false;
    std::string outdir;
    int maxsteps = 8;
    int step = 0;
    int sample_iters = 0;   // time mode: PCG iterations to run (<=0: real solve)
    int sample_pbc = 0;     // time mode: periodic pairs to apply (<=0: all)
    int pbc_stride = 1;     // time mode: the sample is every pbc_stride-th pair of the list (spread over the whole boundary:
                            // what a pair costs in the reference depends on where its nodes sit in the band)
    int pbc_applied = 0;
    int pbc_seen = 0;
    double t_solve_begin = 0, t_pbc = 0, t_pcg = 0, t_first_pbc = -1, t_pcg_enter = 0;
    std::vector<int> iters;       // replayed iteration count per solve
    std::vector<double> pcg_secs; // seconds per solve (real or sampled)
    long pcg_sample_done = 0;
    FSolver *S = nullptr;
} H;

// replays the reference recurrence (spars.cpp:238-316) on copies of b and V using the public
// MultA / MultPC / Dot; returns the iteration count; optional cap for the timing sample.
static int replay_pcg(CBigLinProb &L, const std::vector<double> &b, std::vector<double> V,
                      bool warm, long cap)
{
    const int n = L.n;
    std::vector<double> Z(n), R(n), P(n), U(n);
    L.MultPC(b.data(), Z.data());
    double res_o = L.Dot(Z.data(), const_cast<double *>(b.data()));
    if (res_o == 0) return 0;
    if (!warm) std::fill(V.begin(), V.end(), 0.0);
    L.MultA(V.data(), R.data());
    for (int i = 0; i < n; i++) R[i] = b[i] - R[i];
    L.MultPC(R.data(), Z.data());
    P = Z;
    double res = L.Dot(Z.data(), R.data());
    int it = 0;
    double er;
    do
    {
        L.MultA(P.data(), U.data());
        double del = res / L.Dot(P.data(), U.data());
        for (int i = 0; i < n; i++) { V[i] += del * P[i]; R[i] -= del * U[i]; }
        L.MultPC(R.data(), Z.data());
        double res_new = L.Dot(Z.data(), R.data());
        double rho = res_new / res;
        res = res_new;
        for (int i = 0; i < n; i++) P[i] = Z[i] + rho * P[i];
        er = sqrt(res / res_o);
        it++;
        if (cap > 0 && it >= cap) break;
    } while (er > L.Precision);
    return it;
}

static void dump_real_system(CBigLinProb &L, int step)
{
    std::vector<int> rowptr(L.n + 1, 0), col;
    std::vector<double> val;
    for (int i = 0; i < L.n; i++)
    {
        for (CEntry *e = L.M[i]; e; e = e->next) { col.push_back(e->c); val.push_back(e->x); }
        rowptr[i + 1] = (int)col.size();
    }
    std::string p = H.outdir + "/step" + std::to_string(step) + "_";
    npy_i32(p + "rowptr.npy", rowptr.data(), rowptr.size());
    npy_i32(p + "col.npy", col.data(), col.size());
    npy_f64(p + "val.npy", val.data(), val.size());
    npy_f64(p + "rhs.npy", L.b, L.n);
    npy_f64(p + "vin.npy", L.V, L.n);
}

extern "C" {
bool __real__ZN11CBigLinProb8PCGSolveEi(CBigLinProb *, int);
void __real__ZN11CBigLinProb11PeriodicityEii(CBigLinProb *, int, int);
void __real__ZN11CBigLinProb15AntiPeriodicityEii(CBigLinProb *, int, int);
int __real__ZN18CBigComplexLinProb12PBCGSolveModEib(CBigComplexLinProb *, int, bool);

bool __wrap__ZN11CBigLinProb8PCGSolveEi(CBigLinProb *L, int flag)
{
    double t0 = now_s();
    H.t_pcg_enter = t0;
    if (H.dumping && H.step < H.maxsteps) dump_real_system(*L, H.step);
    bool ok = true;
    if (H.timing && H.sample_iters > 0)
    {
        // bounded sample: run `sample_iters` reference iterations on the assembled system
        for (int i = 0; i < L->n; i++) if (L->M[i]->x == 0) return false;
        std::vector<double> b(L->b, L->b + L->n), V(L->V, L->V + L->n);
        double t1 = now_s();
        int it = replay_pcg(*L, b, V, flag != 0, H.sample_iters);
        H.pcg_secs.push_back(now_s() - t1);
        H.iters.push_back(it);
        H.pcg_sample_done += it;
        L->Precision = 1e300; // lets the caller's Newton loop terminate after this step
    }
    else
    {
        std::vector<double> b, V;
        if (H.dumping) { b.assign(L->b, L->b + L->n); V.assign(L->V, L->V + L->n); }
        double t1 = now_s();
        ok = __real__ZN11CBigLinProb8PCGSolveEi(L, flag);
        H.pcg_secs.push_back(now_s() - t1);
        if (H.dumping && ok) H.iters.push_back(replay_pcg(*L, b, V, flag != 0, 0));
        if (H.dumping && H.step < H.maxsteps)
            npy_f64(H.outdir + "/step" + std::to_string(H.step) + "_vout.npy", L->V, L->n);
    }
    H.t_pcg += now_s() - t0;
    H.step++;
    return ok;
}

void __wrap__ZN11CBigLinProb11PeriodicityEii(CBigLinProb *L, int i, int j)
{
    if (H.t_first_pbc < 0) H.t_first_pbc = now_s();
    H.pbc_seen++;
    if (H.timing && H.sample_pbc > 0 && ((H.pbc_seen - 1) % H.pbc_stride != 0 || H.pbc_applied >= H.sample_pbc)) return;
    H.pbc_applied++;
    double t0 = now_s();
    __real__ZN11CBigLinProb11PeriodicityEii(L, i, j);
    H.t_pbc += now_s() - t0;
}
void __wrap__ZN11CBigLinProb15AntiPeriodicityEii(CBigLinProb *L, int i, int j)
{
    if (H.t_first_pbc < 0) H.t_first_pbc = now_s();
    H.pbc_seen++;
    if (H.timing && H.sample_pbc > 0 && ((H.pbc_seen - 1) % H.pbc_stride != 0 || H.pbc_applied >= H.sample_pbc)) return;
    H.pbc_applied++;
    double t0 = now_s();
    __real__ZN11CBigLinProb15AntiPeriodicityEii(L, i, j);
    H.t_pbc += now_s() - t0;
}

int __wrap__ZN18CBigComplexLinProb12PBCGSolveModEib(CBigComplexLinProb *L, int flag, bool verbose)
{
    double t0 = now_s();
    if (H.dumping && H.step < H.maxsteps)
    {
        std::vector<int> rowptr(L->n + 1, 0), col;
        std::vector<CComplex> val;
        for (int i = 0; i < L->n; i++)
        {
            for (CComplexEntry *e = L->M[i]; e; e = e->next) { col.push_back(e->c); val.push_back(e->x); }
            rowptr[i + 1] = (int)col.size();
        }
        std::string p = H.outdir + "/step" + std::to_string(H.step) + "_";
        npy_i32(p + "rowptr.npy", rowptr.data(), rowptr.size());
        npy_i32(p + "col.npy", col.data(), col.size());
        npy_c128(p + "val.npy", val.data(), val.size());
        npy_c128(p + "rhs.npy", L->b, L->n);
        npy_c128(p + "vin.npy", L->V, L->n);
        if (L->bNewton)
        {
            // side matrices of the Newton AC solver as stored (upper triangle, cspars.h:53-55)
            CComplexEntry **side[3] = {L->Mh, L->Ms, L->Ma};
            for (int h = 0; h < 3; h++)
            {
                std::vector<int> rp(L->n + 1, 0), cc;
                std::vector<CComplex> vv;
                for (int i = 0; i < L->n; i++)
                {
                    for (CComplexEntry *e = side[h][i]; e; e = e->next) { cc.push_back(e->c); vv.push_back(e->x); }
                    rp[i + 1] = (int)cc.size();
                }
                std::string q = p + "side" + std::to_string(h + 1) + "_";
                npy_i32(q + "rowptr.npy", rp.data(), rp.size());
                npy_i32(q + "col.npy", cc.data(), cc.size());
                npy_c128(q + "val.npy", vv.data(), vv.size());
            }
        }
    }
    int ok = __real__ZN18CBigComplexLinProb12PBCGSolveModEib(L, flag, verbose);
    if (H.dumping && H.step < H.maxsteps)
        npy_c128(H.outdir + "/step" + std::to_string(H.step) + "_vout.npy", L->V, L->n);
    H.pcg_secs.push_back(now_s() - t0);
    H.t_pcg += now_s() - t0;
    H.step++;
    return ok;
}
}

static int quiet(const char *, ...) { return 0; }

static void dump_mesh(FSolver &S)
{
    const std::string &o = H.outdir;
    std::vector<double> xy(2 * (size_t)S.NumNodes);
    std::vector<int> bc(S.NumNodes);
    for (int i = 0; i < S.NumNodes; i++)
    {
        xy[2 * i] = S.meshnode[i].x; xy[2 * i + 1] = S.meshnode[i].y; bc[i] = S.meshnode[i].BoundaryMarker;
    }
    npy_f64(o + "/nodes.npy", xy.data(), S.NumNodes, 2);
    npy_i32(o + "/node_bc.npy", bc.data(), S.NumNodes);
    std::vector<int> el(8 * (size_t)S.NumEls);
    for (int i = 0; i < S.NumEls; i++)
    {
        const auto &e = S.meshele[i];
        int *r = &el[8 * (size_t)i];
        r[0] = e.p[0]; r[1] = e.p[1]; r[2] = e.p[2]; r[3] = e.lbl; r[4] = e.blk; r[5] = e.e[0]; r[6] = e.e[1]; r[7] = e.e[2];
    }
    npy_i32(o + "/elems.npy", el.data(), S.NumEls, 8);
    std::vector<int> pbc(3 * (size_t)S.NumPBCs);
    for (int i = 0; i < S.NumPBCs; i++) { pbc[3 * i] = S.pbclist[i].x; pbc[3 * i + 1] = S.pbclist[i].y; pbc[3 * i + 2] = S.pbclist[i].t; }
    npy_i32(o + "/pbc.npy", pbc.data(), S.NumPBCs, 3);
    for (int a = 0; a < S.NumAirGapElems; a++)
    {
        const auto &g = S.agelist[a];
        std::vector<int> qn; std::vector<double> qw;
        for (int k = 0; k <= g.totalArcElements; k++)
        {
            const auto &q = g.quadNode[k];
            qn.insert(qn.end(), {q.n0, q.n1, q.n2, q.n3});
            qw.insert(qw.end(), {q.w0, q.w1, q.w2, q.w3});
        }
        double meta[11] = {(double)g.BdryFormat, g.InnerAngle, g.OuterAngle, g.ri, g.ro, g.totalArcLength,
                           g.agc.re, g.agc.im, (double)g.totalArcElements, g.InnerShift, g.OuterShift};
        npy_f64(o + "/age" + std::to_string(a) + "_meta.npy", meta, 11);
        npy_i32(o + "/age" + std::to_string(a) + "_qn.npy", qn.data(), g.totalArcElements + 1, 4);
        npy_f64(o + "/age" + std::to_string(a) + "_qw.npy", qw.data(), g.totalArcElements + 1, 4);
    }
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ref_harness dump|time <basename> ...\n"); return 64; }
    std::string mode = argv[1];
    FSolver S;
    H.S = &S;
    S.PathName = argv[2];
    S.WarnMessage = quiet;
    S.PrintMessage = quiet;
    if (mode == "dump")
    {
        if (argc < 4) return 64;
        H.dumping = true; H.outdir = argv[3];
        if (argc > 4) H.maxsteps = atoi(argv[4]);
        mkdir(H.outdir.c_str(), 0777);
    }
    else if (mode == "time")
    {
        H.timing = true;
        if (argc > 3) H.sample_iters = atoi(argv[3]);
        if (argc > 4) H.sample_pbc = atoi(argv[4]);
    }
    else return 64;

    // silence the reference's printf("Conjugate Gradient Solver") chatter
    FILE *realout = fdopen(dup(fileno(stdout)), "w");
    freopen("/dev/null", "w", stdout);

    double t0 = now_s();
    if (!S.LoadProblemFile()) { fprintf(stderr, "LoadProblemFile failed\n"); return 1; }
    double t_fem = now_s();
    if (S.LoadMesh(false) != NOERROR) { fprintf(stderr, "LoadMesh failed\n"); return 2; }
    double t_mesh = now_s();
    if (!S.Cuthill(false)) { fprintf(stderr, "Cuthill failed\n"); return 3; }
    double t_cm = now_s();
    if (H.dumping) dump_mesh(S);

    if (H.timing && H.sample_pbc > 0) H.pbc_stride = std::max(1, S.NumPBCs / H.sample_pbc);
    int ok = 0;
    double t_s0 = now_s(), t_s1;
    int n_unknowns = 0;
    if (S.Frequency == 0)
    {
        CBigLinProb L;
        L.Precision = S.Precision;
        L.Create(S.NumNodes, S.BandWidth);
        n_unknowns = L.n;
        H.t_solve_begin = t_s0 = now_s();
        ok = (S.ProblemType == 0) ? S.Static2D(L) : S.StaticAxisymmetric(L);
        t_s1 = now_s();
        if (H.dumping && ok)
        {
            npy_f64(H.outdir + "/V.npy", L.V, L.n);
            npy_f64(H.outdir + "/A.npy", L.b, L.n);
            // final system pattern + values (last Newton step), upper triangle incl. diagonal
            dump_real_system(L, 999);
        }
    }
    else
    {
        CBigComplexLinProb L;
        L.Precision = S.Precision;
        L.Create(S.NumNodes + S.NumCircProps, S.BandWidth, S.NumNodes);
        n_unknowns = L.n;
        H.t_solve_begin = t_s0 = now_s();
        ok = (S.ProblemType == 0) ? S.Harmonic2D(L, false) : S.HarmonicAxisymmetric(L, false);
        t_s1 = now_s();
        if (H.dumping && ok)
        {
            npy_c128(H.outdir + "/V.npy", L.V, L.n);
            npy_c128(H.outdir + "/A.npy", L.b, L.n);
        }
    }
    if (H.dumping && ok)
    {
        std::vector<double> mu(4 * (size_t)S.NumEls);
        for (int i = 0; i < S.NumEls; i++)
        {
            mu[4 * (size_t)i] = S.meshele[i].mu1.re; mu[4 * (size_t)i + 1] = S.meshele[i].mu1.im;
            mu[4 * (size_t)i + 2] = S.meshele[i].mu2.re; mu[4 * (size_t)i + 3] = S.meshele[i].mu2.im;
        }
        npy_f64(H.outdir + "/mu.npy", mu.data(), S.NumEls, 4);
        std::vector<double> circ;
        for (int k = 0; k < S.NumCircProps; k++)
        {
            const auto &c = S.circproplist[k];
            circ.insert(circ.end(), {(double)c.Case, c.J.re, c.J.im, c.dV.re, c.dV.im});
        }
        if (!circ.empty()) npy_f64(H.outdir + "/circ.npy", circ.data(), S.NumCircProps, 5);
    }

    // one JSON line with everything scalar
    std::string js = "{";
    char buf[512];
    snprintf(buf, sizeof buf,
             "\"ok\": %d, \"n_nodes\": %d, \"n_elems\": %d, \"n_unknowns\": %d, \"n_pbc\": %d, \"n_age\": %d, "
             "\"bandwidth\": %d, \"frequency\": %.17g, \"problem_type\": %d, \"precision\": %.17g, \"solves\": %d, ",
             ok, S.NumNodes, S.NumEls, n_unknowns, S.NumPBCs, S.NumAirGapElems, S.BandWidth, S.Frequency,
             (int)S.ProblemType, S.Precision, H.step);
    js += buf;
    double t_pre_solve = (H.t_first_pbc > 0 ? H.t_first_pbc : H.t_pcg_enter) - t_s0;
    snprintf(buf, sizeof buf,
             "\"t_load_fem\": %.6f, \"t_load_mesh\": %.6f, \"t_cuthill\": %.6f, \"t_formulation\": %.6f, "
             "\"t_assembly_first\": %.6f, \"t_pbc\": %.6f, \"t_pcg\": %.6f, \"pbc_applied\": %d, \"pbc_total_calls\": %d, ",
             t_fem - t0, t_mesh - t_fem, t_cm - t_mesh, t_s1 - t_s0, t_pre_solve, H.t_pbc, H.t_pcg,
             (H.timing && H.sample_pbc > 0) ? H.pbc_applied : H.pbc_seen, H.pbc_seen);
    js += buf;
    js += "\"iters\": [";
    for (size_t i = 0; i < H.iters.size(); i++) { js += (i ? ", " : "") + std::to_string(H.iters[i]); }
    js += "], \"pcg_secs\": [";
    for (size_t i = 0; i < H.pcg_secs.size(); i++) { snprintf(buf, sizeof buf, "%s%.6f", i ? ", " : "", H.pcg_secs[i]); js += buf; }
    js += "]}";
    fprintf(realout, "%s\n", js.c_str());
    if (H.dumping)
    {
        FILE *f = fopen((H.outdir + "/meta.json").c_str(), "w");
        fprintf(f, "%s\n", js.c_str());
        fclose(f);
    }
    fflush(realout);
    return ok ? 0 : 4;
}
