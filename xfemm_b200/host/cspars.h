// CBigComplexLinProb — drop-in for cfemm/libfemm/cspars.h:37-95 (complex-symmetric systems of the
// time-harmonic formulations), backed by the B200 COCG solver.
//
// Two ways of use, as for CBigLinProb (spars.h):
//  * FSolver::Harmonic2D / HarmonicAxisymmetric assemble on the device and bind the device problem with
//    attachDevice(); Put/SetValue/Periodicity are then refused (the device kernels do that work).
//  * a caller that assembles on the host through Put/Get/AddTo (selector k = 0: M, 1: Mh Hermitian,
//    2: Ms complex symmetric, 3: Ma anti-Hermitian), applies SetValue / Periodicity / AntiPeriodicity and
//    calls PBCGSolveMod: the entries live in sorted host rows (upper triangle, the reference's flip rules on
//    p>q), the frozen CSR goes to the GPU at the first solve, every matrix-vector product and solve runs there.
#ifndef XFEMM_B200_CSPARS_H
#define XFEMM_B200_CSPARS_H

#include "femmcomplex.h"
#include <vector>

struct xfemm_clinprob;

class CComplexEntry
{
public:
    CComplex x;
    int c;
    CComplexEntry *next;
    CComplexEntry() : x(0), c(0), next(nullptr) {}
};

class CBigComplexLinProb
{
public:
    CComplex *P, *U, *R, *V, *Z, *b, *uu, *vv;
    CComplexEntry **M, **Mh, **Ma, **Ms; // kept for source compatibility; not populated (entries live in `rows`)
    int n, bdw, bNewton, NumNodes;
    double Precision, Lambda;

    CBigComplexLinProb();
    ~CBigComplexLinProb();
    int Create(int d, int bw, int nodes);
    void Put(CComplex v, int p, int q, int k = 0);
    CComplex Get(int p, int q, int k = 0);
    void AddTo(CComplex v, int p, int q);
    void MultA(CComplex *X, CComplex *Y, int k = 0);     // k = -1: M + Mh + Ms conj(.) + Ma (cspars.cpp:303-312)
    void MultConjA(CComplex *X, CComplex *Y, int k = 0); // k = 0 or 2
    CComplex Dot(CComplex *x, CComplex *y);              // unconjugated
    CComplex ConjDot(CComplex *x, CComplex *y);
    void SetValue(int i, CComplex x);
    void Periodicity(int i, int j);
    void AntiPeriodicity(int i, int j);
    void Wipe();
    int PBCGSolveMod(int flag, bool verbose = false); // dispatcher: KludgeSolve when bNewton
    int PBCGSolve(int flag);                          // COCG on M from V (flag != 0) or from zero
    int KludgeSolve(int flag);
    // The remaining public routines of the reference class (cspars.h:37-95), for callers that drive the iteration
    // themselves.  MultPC applies THIS backend's stand-alone preconditioner — the diagonal (Jacobi) — in place of
    // the reference's serial SSOR sweep (cspars.cpp:439-480), which has no GPU-parallel equivalent; the solvers
    // above use the multigrid cycle inside their persistent kernel.  MultAPPA, PCGSQStart and BiCGSTAB (dead code in
    // the reference: its only call is commented out, cspars.cpp:1066) are built on MultA / MultPC exactly as there.
    void MultPC(CComplex *X, CComplex *Y);
    void MultAPPA(CComplex *X, CComplex *Y);
    int PCGSQStart();
    int BiCGSTAB(int flag);

    int precond, lastIterations;
    double lastSolveMs, lastRelRes;
    xfemm_clinprob *device() { return dev; }
    void attachDevice(xfemm_clinprob *d);

private:
    struct Ent { int c; CComplex x; };
    std::vector<std::vector<Ent>> rows[4];  // per selector: upper triangle, sorted by column
    std::vector<std::vector<int>> lower[4]; // lower[k][i] = rows r<i holding an entry (r,i)
    bool patternDirty, valuesDirty, deviceAssembled;
    xfemm_clinprob *dev;
    std::vector<int> csrRowptr, csrCol;
    void ensureRows();
    CComplex *findEntry(int p, int q, int k);
    void coupledRows(int i, std::vector<int> &out) const;
    void periodicImpl(int i, int j, double sgn);
    bool syncToDevice();
    int solve(int flag, bool sqStart);
};

#endif
