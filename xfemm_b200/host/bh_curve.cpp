// B-H curve set-up and evaluation on the host (one-time work per material; the per-element
// lookups during assembly run on the device from the tables produced here).
// Behavioural mirror of cfemm/libfemm/CMaterialProp.cpp:
//   GetSlopes 127-348 · LaminatedBH 1062-1171 · GetdHdB 461-486 · GetH 960-987 ·
//   Get_v 899-904 · GetBHProps 997-1057 · IncrementalPermeability 906-934 / 846-897
#include "fsolver.h"
#include <cmath>
#include <complex>
#include <vector>

using namespace femm;
typedef std::complex<double> cplx;

namespace
{
const cplx J1(0.0, 1.0);

// dense complex Gauss elimination with partial pivoting (fullmatrix.cpp:183-218)
bool gauss_solve(std::vector<std::vector<cplx>> &M, std::vector<cplx> &b)
{
    const int n = (int)b.size();
    for (int i = 0; i < n; i++)
    {
        int q = i;
        cplx mx = 0;
        for (int j = i; j < n; j++)
            if (std::abs(M[j][i]) > std::abs(mx)) { mx = M[j][i]; q = j; }
        if (mx == cplx(0)) return false;
        std::swap(M[i], M[q]);
        std::swap(b[i], b[q]);
        for (int j = i + 1; j < n; j++)
        {
            cplx f = M[j][i] / M[i][i];
            b[j] = b[j] - f * b[i];
            for (int k = i; k < n; k++) M[j][k] -= f * M[i][k];
        }
    }
    for (int i = n - 1; i >= 0; i--)
    {
        cplx f = 0;
        for (int j = n - 1; j > i; j--) f += M[i][j] * b[j];
        b[i] = (b[i] - f) / M[i][i];
    }
    return true;
}
} // namespace

void CMSolverMaterialProp::GetSlopes(double omega)
{
    if (BHpoints == 0) return;
    if (!slope.empty()) return;
    const int n = BHpoints;
    std::vector<double> bn(n, 0.0);
    std::vector<cplx> hn(n, 0.0);
    std::vector<cplx> H(n);
    for (int i = 0; i < n; i++) H[i] = Hdata[i];
    std::vector<double> &B = Bdata;

    // initial linear guess for the first nonlinear pass
    mu_x = B[1] / (muo * std::abs(H[1]));
    mu_y = mu_x;
    Theta_hx = Theta_hn;
    Theta_hy = Theta_hn;

    if (omega != 0)
    {
        // effective fundamental-harmonic B(H) for sinusoidal H (CMaterialProp.cpp:156-196)
        double mumax = 0;
        for (int i = 1; i < n; i++)
        {
            hn[i] = H[i];
            bn[i] = 0;
            for (int k = 1; k <= i; k++)
            {
                cplx t = (4. * (H[k] * B[k - 1] - H[k - 1] * B[k]) * (-std::cos((H[k - 1] * PI) / (2. * H[i])) + std::cos((H[k] * PI) / (2. * H[i]))) +
                          (-B[k - 1] + B[k]) * ((H[k - 1] - H[k]) * PI + H[i] * (-std::sin((H[k - 1] * PI) / H[i]) + std::sin((H[k] * PI) / H[i])))) /
                         ((H[k - 1] - H[k]) * PI);
                bn[i] += t.real();
            }
        }
        for (int i = 1; i < n; i++)
        {
            B[i] = bn[i];
            H[i] = hn[i];
            double munow = (B[i] / H[i]).real();
            if (munow > mumax) mumax = munow;
        }
        // O'Kelly hysteresis lag proportional to permeability
        for (int i = 1; i < n; i++) H[i] *= std::exp(J1 * B[i] * Theta_hn * DEG / (H[i] * mumax));
        MuMax = mumax / muo;
    }

    bool CurveOK = false, ProcessedLams = false;
    std::vector<cplx> sl;
    while (!CurveOK)
    {
        // natural cubic spline in H(B): tridiagonal system for the slopes
        std::vector<std::vector<cplx>> M(n, std::vector<cplx>(n, 0.0));
        std::vector<cplx> rhs(n, 0.0);
        double l1 = B[1] - B[0], l2;
        M[0][0] = 4. / l1; M[0][1] = 2. / l1;
        rhs[0] = 6. * (H[1] - H[0]) / (l1 * l1);
        l1 = B[n - 1] - B[n - 2];
        M[n - 1][n - 1] = 4. / l1; M[n - 1][n - 2] = 2. / l1;
        rhs[n - 1] = 6. * (H[n - 1] - H[n - 2]) / (l1 * l1);
        for (int i = 1; i < n - 1; i++)
        {
            l1 = B[i] - B[i - 1];
            l2 = B[i + 1] - B[i];
            M[i][i - 1] = 2. / l1;
            M[i][i] = 4. * (l1 + l2) / (l1 * l2);
            M[i][i + 1] = 2. / l2;
            rhs[i] = 6. * (H[i] - H[i - 1]) / (l1 * l1) + 6. * (H[i + 1] - H[i]) / (l2 * l2);
        }
        gauss_solve(M, rhs);
        sl = rhs;

        // monotonicity test on the real part of every segment
        CurveOK = true;
        for (int i = 1; i < n; i++)
        {
            double d0 = sl[i - 1].real(), d1 = sl[i].real();
            double u0 = H[i - 1].real(), u1 = H[i].real();
            double L = B[i] - B[i - 1];
            double c0 = d0;
            double c1 = -(2. * (2. * d0 * L + d1 * L + 3. * u0 - 3. * u1)) / (L * L);
            double c2 = (3. * (d0 * L + d1 * L + 2. * u0 - 2. * u1)) / (L * L * L);
            double X0 = -1., X1 = -1.;
            double disc = c1 * c1 - 4. * c0 * c2;
            if (c2 == 0)
            {
                if (c1 != 0) X0 = -c0 / c1;
            }
            else if (disc > 0)
            {
                disc = std::sqrt(disc);
                X0 = -(c1 + disc) / (2. * c2);
                X1 = (-c1 + disc) / (2. * c2);
            }
            if (((X0 >= 0.) && (X0 <= L)) || ((X1 >= 0.) && (X1 <= L))) CurveOK = false;
        }
        if (!CurveOK)
        {
            // 3-point moving average of the interior points, then retry
            for (int i = 1; i < n - 1; i++)
            {
                bn[i] = (B[i - 1] + B[i] + B[i + 1]) / 3.;
                hn[i] = (H[i - 1] + H[i] + H[i + 1]) / 3.;
            }
            for (int i = 1; i < n - 1; i++) { H[i] = hn[i]; B[i] = bn[i]; }
        }
        if (CurveOK && !ProcessedLams)
        {
            // publish the current curve so LaminatedBH/GetH see it
            for (int i = 0; i < n; i++) Hdata[i] = H[i];
            slope.assign(sl.begin(), sl.end());
            if ((omega > 0) && (Lam_d != 0) && (Cduct != 0))
            {
                for (int i = 1; i < n; i++)
                {
                    cplx mu = LaminatedBH(omega, i);
                    bn[i] = std::abs(mu * H[i]);
                    hn[i] = bn[i] / mu;
                }
                for (int i = 1; i < n; i++) { B[i] = bn[i]; H[i] = hn[i]; }
                CurveOK = false;
            }
            if ((LamType == 0) && (LamFill != 1))
            {
                for (int i = 1; i < n; i++)
                {
                    cplx mu = LamFill * B[i] / H[i] + (1. - LamFill) * muo;
                    B[i] = std::abs(mu * H[i]);
                    H[i] = B[i] / mu;
                }
                CurveOK = false;
            }
            ProcessedLams = true;
            if (!CurveOK) slope.clear();
        }
    }
    for (int i = 0; i < n; i++) Hdata[i] = H[i];
    slope.assign(sl.begin(), sl.end());
}

CComplex CMSolverMaterialProp::GetH(double Bv)
{
    double b = fabs(Bv);
    if (BHpoints == 0) return CComplex(b / (mu_x * muo));
    if (b > Bdata[BHpoints - 1])
        return CComplex(cplx(Hdata[BHpoints - 1]) + cplx(slope[BHpoints - 1]) * (b - Bdata[BHpoints - 1]));
    for (int i = 0; i < BHpoints - 1; i++)
        if ((b >= Bdata[i]) && (b <= Bdata[i + 1]))
        {
            double l = Bdata[i + 1] - Bdata[i];
            double z = (b - Bdata[i]) / l, z2 = z * z;
            cplx h = (1. - 3. * z2 + 2. * z2 * z) * cplx(Hdata[i]) + z * (1. - 2. * z + z2) * l * cplx(slope[i]) +
                     z2 * (3. - 2. * z) * cplx(Hdata[i + 1]) + z2 * (z - 1.) * l * cplx(slope[i + 1]);
            return CComplex(h);
        }
    return CComplex(0);
}

CComplex CMSolverMaterialProp::GetdHdB(double Bv)
{
    double b = fabs(Bv);
    if (BHpoints == 0) return CComplex(b / (mu_x * muo));
    if (b > Bdata[BHpoints - 1]) return slope[BHpoints - 1];
    for (int i = 0; i < BHpoints - 1; i++)
        if ((b >= Bdata[i]) && (b <= Bdata[i + 1]))
        {
            double l = Bdata[i + 1] - Bdata[i];
            double z = (b - Bdata[i]) / l;
            cplx h = 6. * z * (z - 1.) * cplx(Hdata[i]) / l + (1. - 4. * z + 3. * z * z) * cplx(slope[i]) +
                     6. * z * (1. - z) * cplx(Hdata[i + 1]) / l + z * (3. * z - 2.) * cplx(slope[i + 1]);
            return CComplex(h);
        }
    return CComplex(0);
}

CComplex CMSolverMaterialProp::Get_v(double Bv)
{
    if (Bv == 0) return slope[0];
    return CComplex(cplx(GetH(Bv)) / Bv);
}

void CMSolverMaterialProp::GetBHProps(double Bv, CComplex &v, CComplex &dv)
{
    double b = fabs(Bv);
    if (BHpoints == 0) { v = mu_x; dv = 0; return; }
    if (b == 0) { v = slope[0]; dv = 0; return; }
    cplx h, dh;
    if (b > Bdata[BHpoints - 1])
    {
        h = cplx(Hdata[BHpoints - 1]) + cplx(slope[BHpoints - 1]) * (b - Bdata[BHpoints - 1]);
        dh = slope[BHpoints - 1];
    }
    else
    {
        bool found = false;
        for (int i = 0; i < BHpoints - 1 && !found; i++)
            if ((b >= Bdata[i]) && (b <= Bdata[i + 1]))
            {
                double l = Bdata[i + 1] - Bdata[i];
                double z = (b - Bdata[i]) / l, z2 = z * z;
                h = (1. - 3. * z2 + 2. * z2 * z) * cplx(Hdata[i]) + z * (1. - 2. * z + z2) * l * cplx(slope[i]) +
                    z2 * (3. - 2. * z) * cplx(Hdata[i + 1]) + z2 * (z - 1.) * l * cplx(slope[i + 1]);
                dh = 6. * z * (z - 1.) * cplx(Hdata[i]) / l + (1. - 4. * z + 3. * z * z) * cplx(slope[i]) +
                     6. * z * (1. - z) * cplx(Hdata[i + 1]) / l + z * (3. * z - 2.) * cplx(slope[i + 1]);
                found = true;
            }
        if (!found) return;
    }
    v = CComplex(h / b);
    dv = CComplex(0.5 * (dh / (b * b) - h / (b * b * b)));
}

void CMSolverMaterialProp::GetBHProps(double Bv, double &v, double &dv)
{
    CComplex vc, dvc;
    GetBHProps(Bv, vc, dvc);
    v = vc.re;
    dv = dvc.re;
}

// 1-D finite-element model of one lamination to get the apparent permeability at B-H point i
// for AC problems (CMaterialProp.cpp:1062-1171, solver flavour of GetH).
CComplex CMSolverMaterialProp::LaminatedBH(double w, int i)
{
    const int ElementsPerSkinDepth = 10;
    cplx mu = Bdata[i] / cplx(Hdata[i]);
    double o = Cduct * 1.e6;
    double d = (Lam_d * 0.001) / 2.;
    double ds = std::sqrt(2 / (w * o * std::abs(mu)));
    int n = ElementsPerSkinDepth * ((int)ceil(d / ds));
    double L = d / ((double)n);
    std::vector<cplx> x(n + 1, 0.0), b(n + 1, 0.0), m0(n + 1, 0.0), m1(n + 1, 0.0);
    int iter = 0;
    double res = 0, lastres, Relax = 1;
    bool Converged = false;
    do
    {
        for (int k = 0; k <= n; k++) { m0[k] = 0; m1[k] = 0; b[k] = 0; }
        for (int k = 0; k < n; k++)
        {
            cplx vo, vi;
            if (iter != 0)
            {
                double Bk = std::abs(x[k + 1] - x[k]) / L;
                vi = GetdHdB(Bk);
                vo = cplx(GetH(Bk)) / Bk;
            }
            else { vo = 1. / mu; vi = 1. / mu; }
            cplx Md = ((vi + vo) / (2. * L) + J1 * w * o * L / 4.);
            cplx Mo = (-(vi + vo) / (2. * L) + J1 * w * o * L / 4.);
            m0[k] += Md; m1[k] += Mo; m0[k + 1] += Md;
            Md = ((vi - vo) / (2. * L));
            Mo = (-(vi - vo) / (2. * L));
            b[k] += (Md * x[k] + Mo * x[k + 1]);
            b[k + 1] += (Mo * x[k] + Md * x[k + 1]);
        }
        m1[0] = 0; b[0] = 0;
        b[n] += cplx(Hdata[i]);
        for (int k = 0; k < n; k++)
        {
            cplx c = m1[k] / m0[k];
            m0[k + 1] -= m1[k] * c;
            b[k + 1] -= b[k] * c;
        }
        b[n] = b[n] / m0[n];
        for (int k = n - 1; k >= 0; k--) b[k] = (b[k] - m1[k] * b[k + 1]) / m0[k];
        iter++;
        lastres = res;
        res = std::abs(b[n] - x[n]) / d;
        if (res < 1.e-8) Converged = true;
        if (iter > 5)
        {
            if ((res > lastres) && (Relax > 0.1)) Relax /= 2.; else Relax += 0.1 * (1. - Relax);
        }
        for (int k = 0; k <= n; k++) x[k] = Relax * b[k] + (1.0 - Relax) * x[k];
    } while (!Converged);
    return CComplex(x[n] / (cplx(Hdata[i]) * d));
}

void CMSolverMaterialProp::IncrementalPermeability(double B, double &mu1, double &mu2)
{
    double muinc = 1. / (muo * GetdHdB(B).re);
    double murel = 1. / (muo * Get_v(B).re);
    if ((Lam_d == 0) || (LamFill == 0)) { mu1 = muinc; mu2 = murel; return; }
    mu1 = (muinc * LamFill + (1. - LamFill));
    mu2 = (murel * LamFill + (1. - LamFill));
}

void CMSolverMaterialProp::incrementalPermeability(double B, double w, CComplex &mu1, CComplex &mu2)
{
    double muinc = 1. / (muo * GetdHdB(B).re);
    double murel = 1. / (muo * Get_v(B).re);
    if ((Lam_d == 0) || (LamFill == 0))
    {
        mu1 = CComplex(muinc * std::exp(-J1 * Theta_hn * DEG * muinc / MuMax));
        mu2 = CComplex(murel * std::exp(-J1 * Theta_hn * DEG * murel / MuMax));
        return;
    }
    if (Cduct != 0)
    {
        const cplx deg45(1, 1);
        double mu = (muinc - (1. - LamFill)) / LamFill;
        cplx halflag = std::exp(-J1 * Theta_hn * DEG * mu / (2. * MuMax));
        double ds = std::sqrt(2. / (0.4 * PI * w * Cduct * mu));
        cplx K = halflag * deg45 * Lam_d * 0.001 / (2. * ds);
        mu1 = CComplex(LamFill * mu * std::tanh(K) / K + (1. - LamFill));
        mu = (murel - (1. - LamFill)) / LamFill;
        halflag = std::exp(-J1 * Theta_hn * DEG * mu / (2. * MuMax));
        ds = std::sqrt(2. / (0.4 * PI * w * Cduct * mu));
        K = halflag * deg45 * Lam_d * 0.001 / (2. * ds);
        mu2 = CComplex(LamFill * mu * std::tanh(K) / K + (1. - LamFill));
    }
    else
    {
        double mu = (muinc - (1. - LamFill)) / LamFill;
        mu1 = CComplex(mu * std::exp(-J1 * Theta_hn * DEG * mu / MuMax) * LamFill + (1. - LamFill));
        mu = (murel - (1. - LamFill)) / LamFill;
        mu2 = CComplex(mu * std::exp(-J1 * Theta_hn * DEG * mu / MuMax) * LamFill + (1. - LamFill));
    }
}
