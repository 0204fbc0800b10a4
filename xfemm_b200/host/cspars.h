// CBigComplexLinProb — drop-in for cfemm/libfemm/cspars.h:37-95 (complex-symmetric systems of the
// time-harmonic formulations), backed by the B200 COCG solver.
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
    CComplexEntry **M, **Mh, **Ma, **Ms;
    int n, bdw, bNewton, NumNodes;
    double Precision, Lambda;

    CBigComplexLinProb();
    ~CBigComplexLinProb();
    int Create(int d, int bw, int nodes);
    void Put(CComplex v, int p, int q, int k = 0);
    CComplex Get(int p, int q, int k = 0);
    void AddTo(CComplex v, int p, int q);
    void SetValue(int i, CComplex x);
    void Periodicity(int i, int j);
    void AntiPeriodicity(int i, int j);
    void Wipe();
    int PBCGSolveMod(int flag, bool verbose = false);

    int precond, lastIterations;
    double lastSolveMs, lastRelRes;
    xfemm_clinprob *device() { return dev; }
    void attachDevice(xfemm_clinprob *d);

private:
    xfemm_clinprob *dev;
};

#endif
