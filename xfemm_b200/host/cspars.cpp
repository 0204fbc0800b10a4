// CBigComplexLinProb facade (reference: cfemm/libfemm/cspars.cpp:119-1081).
// The B200 FSolver assembles harmonic problems on the device and binds the device problem with
// attachDevice(); host-side assembly through Put/AddTo (not used by any caller of the reference
// outside fsolver itself) is not provided and fails loudly.
#include "cspars.h"
#include "../../include/xfemm_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

static int default_precond()
{
    // multigrid on |Re A| + |Im A| unless XFEMM_PRECOND says otherwise
    const char *e = getenv("XFEMM_PRECOND");
    if (!e) return XFEMM_PRECOND_AMG;
    if (!strcmp(e, "jacobi") || !strcmp(e, "0")) return XFEMM_PRECOND_JACOBI;
    if (!strcmp(e, "mcssor") || !strcmp(e, "1")) return XFEMM_PRECOND_MCSSOR;
    return XFEMM_PRECOND_AMG;
}

CBigComplexLinProb::CBigComplexLinProb()
    : P(nullptr), U(nullptr), R(nullptr), V(nullptr), Z(nullptr), b(nullptr), uu(nullptr), vv(nullptr), M(nullptr),
      Mh(nullptr), Ma(nullptr), Ms(nullptr), n(0), bdw(0), bNewton(0), NumNodes(0), Precision(1e-8), Lambda(1.5),
      precond(default_precond()), lastIterations(0), lastSolveMs(0), lastRelRes(0), dev(nullptr)
{
}

CBigComplexLinProb::~CBigComplexLinProb()
{
    if (dev) xfemm_clinprob_destroy(dev);
    free(P); free(U); free(R); free(V); free(Z); free(b); free(uu); free(vv);
}

int CBigComplexLinProb::Create(int d, int bw, int nodes)
{
    bdw = bw;
    NumNodes = nodes;
    n = d;
    b = (CComplex *)calloc(d, sizeof(CComplex));
    V = (CComplex *)calloc(d, sizeof(CComplex));
    P = (CComplex *)calloc(d, sizeof(CComplex));
    R = (CComplex *)calloc(d, sizeof(CComplex));
    U = (CComplex *)calloc(d, sizeof(CComplex));
    Z = (CComplex *)calloc(d, sizeof(CComplex));
    uu = (CComplex *)calloc(d, sizeof(CComplex));
    vv = (CComplex *)calloc(d, sizeof(CComplex));
    return 1;
}

void CBigComplexLinProb::attachDevice(xfemm_clinprob *d)
{
    if (dev && dev != d) xfemm_clinprob_destroy(dev);
    dev = d;
}
void CBigComplexLinProb::Put(CComplex, int, int, int) { fprintf(stderr, "xfemm_b200: CBigComplexLinProb::Put is not available in this build\n"); }
CComplex CBigComplexLinProb::Get(int, int, int) { return CComplex(0); }
void CBigComplexLinProb::AddTo(CComplex v, int p, int q) { Put(Get(p, q) + v, p, q); }
void CBigComplexLinProb::SetValue(int, CComplex) {}
void CBigComplexLinProb::Periodicity(int, int) {}
void CBigComplexLinProb::AntiPeriodicity(int, int) {}
void CBigComplexLinProb::Wipe() {}
int CBigComplexLinProb::PBCGSolveMod(int flag, bool)
{
    if (!dev)
    {
        fprintf(stderr, "xfemm_b200: CBigComplexLinProb has no device problem attached\n");
        return 0;
    }
    xfemm_solve_info info;
    // a multigrid attempt that does not converge within 3000 iterations is repeated with Jacobi (see FSolver)
    int rc = xfemm_clinprob_cocg(dev, Precision, flag, precond, precond == XFEMM_PRECOND_AMG ? 3000 : 0, &info);
    if (rc == XFEMM_ERR_DIVERGED && precond == XFEMM_PRECOND_AMG)
        rc = xfemm_clinprob_cocg(dev, Precision, 0, XFEMM_PRECOND_JACOBI, 0, &info);
    lastIterations = info.iterations; lastSolveMs = info.solve_ms; lastRelRes = info.relres;
    if (rc != XFEMM_OK) { fprintf(stderr, "%s\n", xfemm_last_error()); return 0; }
    return xfemm_clinprob_get_solution(dev, (double *)V) == XFEMM_OK;
}
