// B-H curve preparation and evaluation (CMaterialProp.cpp: GetSlopes, LaminatedBH, GetH, GetdHdB, Get_v, GetBHProps,
// IncrementalPermeability) on eight probe curves: static, stacking factor, a curve that needs smoothing passes, AC with
// hysteresis lag, AC with eddy currents in the laminations (two frequencies), both together, LamType 1.  Prints the
// bit patterns of every table entry and of evaluations over 0..3 T.  The SAME source is compiled twice: against this
// repo's host mirror (host/bh_curve.cpp) and against the unmodified reference class (oracle/Makefile ->
// oracle/_ref/bh_ref); tests/test_host_parity.py compares the outputs.
#ifdef BH_REF
#include "CMaterialProp.h"
#else
#include "fsolver.h"
#endif
#include <cstdio>
#include <cstring>
#include <vector>
using namespace femm;
static void hex(double v) { unsigned long long u; memcpy(&u, &v, 8); printf("%016llx ", u); }
static void run(const char *name, std::vector<double> B, std::vector<double> H, double omega, double lam_d, double cduct, double fill, int lamtype, double theta)
{
    CMSolverMaterialProp m;
    m.BHpoints = (int)B.size(); m.Bdata = B; m.Hdata.clear();
    for (double h : H) m.Hdata.push_back(CComplex(h));
    m.Lam_d = lam_d; m.Cduct = cduct; m.LamFill = fill; m.LamType = lamtype; m.Theta_hn = theta;
    m.GetSlopes(omega);
    printf("%s\n", name);
    for (int i = 0; i < m.BHpoints; i++) { hex(m.Bdata[i]); hex(m.Hdata[i].re); hex(m.Hdata[i].im); hex(m.slope[i].re); hex(m.slope[i].im); printf("\n"); }
    hex(m.mu_x); if (omega != 0) hex(m.MuMax); /* (only GetSlopes of an AC problem sets it) */ printf("\n");
    for (double b = 0.0; b < 3.0; b += 0.137)
    {
        CComplex v, dv; m.GetBHProps(b, v, dv); hex(v.re); hex(v.im); hex(dv.re); hex(dv.im);
        CComplex h = m.GetH(b), d = m.GetdHdB(b); hex(h.re); hex(h.im); hex(d.re); hex(d.im);
        if (b > 0) { CComplex q = m.Get_v(b); hex(q.re); hex(q.im); double a1, a2; m.IncrementalPermeability(b, a1, a2); hex(a1); hex(a2);
            if (omega != 0) { CComplex c1, c2; m.incrementalPermeability(b, omega, c1, c2); hex(c1.re); hex(c1.im); hex(c2.re); hex(c2.im); } }
        printf("\n");
    }
}
int main()
{
    std::vector<double> B = {0, 0.2, 0.4, 0.7, 1.0, 1.2, 1.3, 1.4, 1.5, 1.6, 1.7, 1.8, 1.9, 2.0, 2.1, 2.3};
    std::vector<double> H = {0, 40, 60, 90, 130, 180, 230, 320, 500, 900, 2000, 5000, 11000, 25000, 60000, 200000};
    std::vector<double> B2 = {0, 0.1, 0.5, 0.6, 1.5, 1.55, 1.9};         // wiggly: forces smoothing passes
    std::vector<double> H2 = {0, 100, 110, 400, 500, 3000, 3100};
    run("static", B, H, 0, 0, 0, 1, 0, 0);
    run("static lamfill", B, H, 0, 0.5, 0, 0.9, 0, 0);
    run("static wiggly", B2, H2, 0, 0, 0, 1, 0, 0);
    run("ac", B, H, 2 * 3.14159265358979323846 * 50, 0, 0, 1, 0, 15);
    run("ac lam", B, H, 2 * 3.14159265358979323846 * 50, 0.5, 2.0, 0.95, 0, 20);
    run("ac lam400", B, H, 2 * 3.14159265358979323846 * 400, 0.35, 3.0, 0.98, 0, 0);
    run("ac wiggly", B2, H2, 2 * 3.14159265358979323846 * 60, 0.5, 1.0, 0.9, 0, 10);
    run("ac lamtype1", B, H, 2 * 3.14159265358979323846 * 50, 0.5, 2.0, 0.9, 1, 5);
    return 0;
}
