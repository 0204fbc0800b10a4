// B-H curves on the host: one-time preparation per material (the per-element look-ups of the assembly run on the
// device from the tables left here) and the point evaluations the host-side callers need.
//
// What has to come out is fixed by the reference (cfemm/libfemm/CMaterialProp.cpp: GetSlopes 127-348, LaminatedBH
// 1062-1171, GetH 960-987, GetdHdB 461-486, Get_v 899-904, GetBHProps 997-1057, IncrementalPermeability 906-934 /
// 846-897) down to the rounding: Newton iterations amplify differences in the slopes.  How it is organised here:
//
//   Curve            working copy of the points (B real, H complex once a lag is applied)
//   BandSystem       the spline's tridiagonal system, eliminated with row exchanges in band storage; it performs the
//                    operations a dense elimination with partial pivoting performs on the non-zero entries, in the
//                    same order (the dense one only adds exact zeros)
//   monotone_slopes  spline slopes of a curve, smoothing it until every segment is monotone
//   HermiteSegment   one place for "which segment, which local coordinate" and the basis weights of H and dH/dB
//   LaminationSlab   1-D finite-element model of half a lamination sheet (apparent permeability at one curve point)
#include "fsolver.h"
#include <algorithm>
#include <cmath>
#include <complex>
#include <vector>

using namespace femm;

namespace
{
using cplx = std::complex<double>;
const cplx IMAG(0.0, 1.0);

struct Curve
{
    std::vector<double> B;
    std::vector<cplx> H;
    int size() const { return (int)B.size(); }
};

// Row r couples x[r-1], x[r], x[r+1] (and x[r+2] after a row exchange).
struct BandSystem
{
    std::vector<cplx> below, diag, above, above2, rhs;
    explicit BandSystem(int n) : below(n, 0.0), diag(n, 0.0), above(n, 0.0), above2(n, 0.0), rhs(n, 0.0) {}

    // Forward elimination with the larger of the two candidates as pivot (strictly larger to exchange), then back
    // substitution adding the far coupling first.  Returns false for a zero pivot column.  Solution in rhs.
    bool solve()
    {
        const int n = (int)rhs.size();
        for (int r = 0; r < n; r++)
        {
            double best = std::abs(diag[r]);
            bool exchange = false;
            if (r + 1 < n && std::abs(below[r + 1]) > best) { best = std::abs(below[r + 1]); exchange = true; }
            if (!(best > 0.0)) return false;
            if (exchange)
            {
                // row r <-> row r+1, expressed in the columns r, r+1, r+2
                const cplx p0 = below[r + 1], p1 = diag[r + 1], p2 = above[r + 1];
                below[r + 1] = diag[r]; diag[r + 1] = above[r]; above[r + 1] = above2[r];
                diag[r] = p0; above[r] = p1; above2[r] = p2;
                std::swap(rhs[r], rhs[r + 1]);
            }
            if (r + 1 < n)
            {
                const cplx factor = below[r + 1] / diag[r];
                rhs[r + 1] = rhs[r + 1] - factor * rhs[r];
                below[r + 1] -= factor * diag[r];
                diag[r + 1] -= factor * above[r];
                above[r + 1] -= factor * above2[r];
            }
        }
        for (int r = n - 1; r >= 0; r--)
        {
            cplx known = 0;
            if (r + 2 < n) known += above2[r] * rhs[r + 2];
            if (r + 1 < n) known += above[r] * rhs[r + 1];
            rhs[r] = (rhs[r] - known) / diag[r];
        }
        return true;
    }
};

// Slopes dH/dB at the points of `c` for the cubic spline with natural-type end conditions.
std::vector<cplx> spline_slopes(const Curve &c)
{
    const int n = c.size();
    BandSystem S(n);
    auto gap = [&](int k) { return c.B[k + 1] - c.B[k]; }; // width of segment k
    {
        const double w = gap(0);
        S.diag[0] = 4. / w; S.above[0] = 2. / w;
        S.rhs[0] = 6. * (c.H[1] - c.H[0]) / (w * w);
    }
    {
        const double w = gap(n - 2);
        S.diag[n - 1] = 4. / w; S.below[n - 1] = 2. / w;
        S.rhs[n - 1] = 6. * (c.H[n - 1] - c.H[n - 2]) / (w * w);
    }
    for (int k = 1; k < n - 1; k++)
    {
        const double wl = gap(k - 1), wr = gap(k);
        S.below[k] = 2. / wl;
        S.diag[k] = 4. * (wl + wr) / (wl * wr);
        S.above[k] = 2. / wr;
        S.rhs[k] = 6. * (c.H[k] - c.H[k - 1]) / (wl * wl) + 6. * (c.H[k + 1] - c.H[k]) / (wr * wr);
    }
    S.solve();
    return S.rhs;
}

// Does Re dH/dB change sign inside the segment that ends at point k?  (roots of the derivative's quadratic in [0, width])
bool segment_turns(const Curve &c, const std::vector<cplx> &s, int k)
{
    const double width = c.B[k] - c.B[k - 1];
    const double sa = s[k - 1].real(), sb = s[k].real();
    const double ha = c.H[k - 1].real(), hb = c.H[k].real();
    const double q0 = sa;
    const double q1 = -(2. * (2. * sa * width + sb * width + 3. * ha - 3. * hb)) / (width * width);
    const double q2 = (3. * (sa * width + sb * width + 2. * ha - 2. * hb)) / (width * width * width);
    double root_a = -1., root_b = -1.;
    if (q2 == 0)
    {
        if (q1 != 0) root_a = -q0 / q1;
    }
    else
    {
        double d = q1 * q1 - 4. * q0 * q2;
        if (d > 0)
        {
            d = std::sqrt(d);
            root_a = -(q1 + d) / (2. * q2);
            root_b = (-q1 + d) / (2. * q2);
        }
    }
    auto inside = [&](double x) { return (x >= 0.) && (x <= width); };
    return inside(root_a) || inside(root_b);
}

// Replaces every interior point by the mean of itself and its neighbours (all means taken from the old points).
void smooth(Curve &c)
{
    const int n = c.size();
    Curve old = c;
    for (int k = 1; k < n - 1; k++)
    {
        c.B[k] = (old.B[k - 1] + old.B[k] + old.B[k + 1]) / 3.;
        c.H[k] = (old.H[k - 1] + old.H[k] + old.H[k + 1]) / 3.;
    }
}

std::vector<cplx> monotone_slopes(Curve &c)
{
    for (;;)
    {
        std::vector<cplx> s = spline_slopes(c);
        bool monotone = true;
        for (int k = 1; k < c.size(); k++)
            if (segment_turns(c, s, k)) monotone = false;
        if (monotone) return s;
        smooth(c);
    }
}

// Sinusoidal H of amplitude H[i] through the piecewise-linear curve: B[i] becomes the amplitude of the fundamental of
// the flux density (closed-form Fourier integral over the segments up to i).  Then every H gets a lag proportional to
// the local permeability (hysteresis angle theta_deg at the largest one).  Returns the largest B/H of the new curve.
double fundamental_curve(Curve &c, double theta_deg)
{
    const int n = c.size();
    std::vector<double> amp(n, 0.0);
    for (int i = 1; i < n; i++)
    {
        const cplx Hp = c.H[i];
        for (int k = 1; k <= i; k++)
        {
            const cplx Ha = c.H[k - 1], Hb = c.H[k];
            const double Ba = c.B[k - 1], Bb = c.B[k];
            const cplx cross = 4. * (Hb * Ba - Ha * Bb);
            const cplx dcos = -std::cos((Ha * PI) / (2. * Hp)) + std::cos((Hb * PI) / (2. * Hp));
            const cplx dsin = -std::sin((Ha * PI) / Hp) + std::sin((Hb * PI) / Hp);
            const cplx piece = (cross * dcos + (-Ba + Bb) * ((Ha - Hb) * PI + Hp * dsin)) / ((Ha - Hb) * PI);
            amp[i] += piece.real();
        }
    }
    double steepest = 0;
    for (int i = 1; i < n; i++)
    {
        c.B[i] = amp[i];
        steepest = std::max(steepest, (c.B[i] / c.H[i]).real());
    }
    for (int i = 1; i < n; i++) c.H[i] *= std::exp(IMAG * c.B[i] * theta_deg * DEG / (c.H[i] * steepest));
    return steepest;
}

// Segment of a prepared curve that holds |B| = b, its width and the local coordinate; beyond the last point the curve
// continues as a straight line.
struct HermiteSegment
{
    enum Where { INSIDE, BEYOND, NOWHERE } where = NOWHERE;
    int k = 0;
    double width = 0, t = 0;

    HermiteSegment(const std::vector<double> &B, int npoints, double b)
    {
        if (b > B[npoints - 1]) { where = BEYOND; k = npoints - 1; return; }
        // first point (from the second on) that is not below b: b lies in the segment that ends there
        const auto first = B.begin() + 1, last = B.begin() + npoints;
        const auto hit = std::lower_bound(first, last, b);
        if (hit == last || b < *(hit - 1)) return;
        k = (int)(hit - B.begin()) - 1;
        width = B[k + 1] - B[k];
        t = (b - B[k]) / width;
        where = INSIDE;
    }
    cplx value(const std::vector<CComplex> &H, const std::vector<CComplex> &s) const
    {
        const double t2 = t * t;
        return (1. - 3. * t2 + 2. * t2 * t) * cplx(H[k]) + t * (1. - 2. * t + t2) * width * cplx(s[k]) +
               t2 * (3. - 2. * t) * cplx(H[k + 1]) + t2 * (t - 1.) * width * cplx(s[k + 1]);
    }
    cplx derivative(const std::vector<CComplex> &H, const std::vector<CComplex> &s) const
    {
        return 6. * t * (t - 1.) * cplx(H[k]) / width + (1. - 4. * t + 3. * t * t) * cplx(s[k]) +
               6. * t * (1. - t) * cplx(H[k + 1]) / width + t * (3. * t - 2.) * cplx(s[k + 1]);
    }
};

// Half a lamination sheet of thickness 2 d, discretised with first-order elements between the mid-plane (node 0,
// A = 0) and the surface (node n, tangential H prescribed).  Nonlinear through the material's own H(B); solved by
// successive substitution with the relaxation schedule of the reference.
struct LaminationSlab
{
    CMSolverMaterialProp &mat;
    double w, sigma, d, h;
    int n;
    std::vector<cplx> a, diag, off, load; // potential, system diagonal / first off-diagonal (symmetric), right-hand side

    LaminationSlab(CMSolverMaterialProp &m, double omega, cplx mu_start) : mat(m), w(omega)
    {
        const int per_skin_depth = 10;
        sigma = m.Cduct * 1.e6;
        d = (m.Lam_d * 0.001) / 2.;
        const double skin = std::sqrt(2 / (w * sigma * std::abs(mu_start)));
        n = per_skin_depth * ((int)ceil(d / skin));
        h = d / ((double)n);
        a.assign(n + 1, 0.0); diag.assign(n + 1, 0.0); off.assign(n + 1, 0.0); load.assign(n + 1, 0.0);
    }
    void assemble(bool first_pass, cplx mu_start)
    {
        for (int k = 0; k <= n; k++) { diag[k] = 0; off[k] = 0; load[k] = 0; }
        for (int e = 0; e < n; e++)
        {
            cplx nu_secant, nu_tangent;
            if (!first_pass)
            {
                const double Be = std::abs(a[e + 1] - a[e]) / h;
                nu_tangent = mat.GetdHdB(Be);
                nu_secant = cplx(mat.GetH(Be)) / Be;
            }
            else { nu_secant = 1. / mu_start; nu_tangent = 1. / mu_start; }
            cplx kd = ((nu_tangent + nu_secant) / (2. * h) + IMAG * w * sigma * h / 4.);
            cplx ko = (-(nu_tangent + nu_secant) / (2. * h) + IMAG * w * sigma * h / 4.);
            diag[e] += kd; off[e] += ko; diag[e + 1] += kd;
            kd = ((nu_tangent - nu_secant) / (2. * h));
            ko = (-(nu_tangent - nu_secant) / (2. * h));
            load[e] += (kd * a[e] + ko * a[e + 1]);
            load[e + 1] += (ko * a[e] + kd * a[e + 1]);
        }
    }
    // symmetric tridiagonal elimination; the solution replaces `load`
    void solve()
    {
        for (int k = 0; k < n; k++)
        {
            const cplx f = off[k] / diag[k];
            diag[k + 1] -= off[k] * f;
            load[k + 1] -= load[k] * f;
        }
        load[n] = load[n] / diag[n];
        for (int k = n - 1; k >= 0; k--) load[k] = (load[k] - off[k] * load[k + 1]) / diag[k];
    }
    // surface potential for the surface field H_surface
    cplx run(cplx H_surface, cplx mu_start)
    {
        double residual = 0, relax = 1;
        for (int pass = 0;; pass++)
        {
            assemble(pass == 0, mu_start);
            off[0] = 0; load[0] = 0; // A = 0 on the mid-plane
            load[n] += H_surface;
            solve();
            const double previous = residual;
            residual = std::abs(load[n] - a[n]) / d;
            const bool done = residual < 1.e-8;
            if (pass + 1 > 5)
            {
                if ((residual > previous) && (relax > 0.1)) relax /= 2.; else relax += 0.1 * (1. - relax);
            }
            for (int k = 0; k <= n; k++) a[k] = relax * load[k] + (1.0 - relax) * a[k];
            if (done) return a[n];
        }
    }
};
} // namespace

void CMSolverMaterialProp::GetSlopes(double omega)
{
    if (BHpoints == 0 || !slope.empty()) return;
    Curve c;
    c.B = Bdata;
    c.H.assign(Hdata.begin(), Hdata.end());
    auto publish = [&](const std::vector<cplx> &s) {
        Bdata = c.B;
        for (int i = 0; i < BHpoints; i++) Hdata[i] = c.H[i];
        slope.assign(s.begin(), s.end());
    };

    // linear stand-in for the first pass of a nonlinear solve
    mu_x = c.B[1] / (muo * std::abs(c.H[1]));
    mu_y = mu_x;
    Theta_hx = Theta_hn;
    Theta_hy = Theta_hn;

    if (omega != 0) MuMax = fundamental_curve(c, Theta_hn) / muo;

    std::vector<cplx> s = monotone_slopes(c);
    publish(s); // the lamination model below evaluates the curve through the members

    // the curve as the laminated bulk shows it: eddy currents in the sheets, then the stacking factor (once)
    bool reshaped = false;
    if ((omega > 0) && (Lam_d != 0) && (Cduct != 0))
    {
        Curve seen = c;
        for (int i = 1; i < BHpoints; i++)
        {
            const cplx mu = LaminatedBH(omega, i);
            seen.B[i] = std::abs(mu * c.H[i]);
            seen.H[i] = seen.B[i] / mu;
        }
        c = seen;
        reshaped = true;
    }
    if ((LamType == 0) && (LamFill != 1))
    {
        for (int i = 1; i < BHpoints; i++)
        {
            const cplx mu = LamFill * c.B[i] / c.H[i] + (1. - LamFill) * muo;
            c.B[i] = std::abs(mu * c.H[i]);
            c.H[i] = c.B[i] / mu;
        }
        reshaped = true;
    }
    if (reshaped)
    {
        slope.clear();
        s = monotone_slopes(c);
        publish(s);
    }
}

CComplex CMSolverMaterialProp::GetH(double Bv)
{
    const double b = fabs(Bv);
    if (BHpoints == 0) return CComplex(b / (mu_x * muo));
    const HermiteSegment seg(Bdata, BHpoints, b);
    switch (seg.where)
    {
    case HermiteSegment::BEYOND: return CComplex(cplx(Hdata[seg.k]) + cplx(slope[seg.k]) * (b - Bdata[seg.k]));
    case HermiteSegment::INSIDE: return CComplex(seg.value(Hdata, slope));
    default: return CComplex(0);
    }
}

CComplex CMSolverMaterialProp::GetdHdB(double Bv)
{
    const double b = fabs(Bv);
    if (BHpoints == 0) return CComplex(b / (mu_x * muo));
    const HermiteSegment seg(Bdata, BHpoints, b);
    switch (seg.where)
    {
    case HermiteSegment::BEYOND: return slope[seg.k];
    case HermiteSegment::INSIDE: return CComplex(seg.derivative(Hdata, slope));
    default: return CComplex(0);
    }
}

// The reference's Get_v lives in the base class and calls the base class's REAL-valued field (CMaterialProp.cpp:488-491,
// 899-904: Re of the signed complex interpolation, zero for a material without a curve), not the solver class's GetH:
// its result has no imaginary part even when the curve carries a hysteresis lag.
CComplex CMSolverMaterialProp::Get_v(double Bv)
{
    if (Bv == 0) return slope[0];
    const double b = fabs(Bv);
    double field = 0;
    if (BHpoints != 0)
    {
        const HermiteSegment seg(Bdata, BHpoints, b);
        cplx h = 0;
        if (seg.where == HermiteSegment::BEYOND) h = cplx(Hdata[seg.k]) + cplx(slope[seg.k]) * (b - Bdata[seg.k]);
        else if (seg.where == HermiteSegment::INSIDE) h = seg.value(Hdata, slope);
        field = (cplx(Bv / b) * h).real();
    }
    return CComplex(field / Bv);
}

void CMSolverMaterialProp::GetBHProps(double Bv, CComplex &v, CComplex &dv)
{
    const double b = fabs(Bv);
    if (BHpoints == 0) { v = mu_x; dv = 0; return; }
    if (b == 0) { v = slope[0]; dv = 0; return; }
    const HermiteSegment seg(Bdata, BHpoints, b);
    cplx field, rate;
    switch (seg.where)
    {
    case HermiteSegment::BEYOND:
        field = cplx(Hdata[seg.k]) + cplx(slope[seg.k]) * (b - Bdata[seg.k]);
        rate = slope[seg.k];
        break;
    case HermiteSegment::INSIDE:
        field = seg.value(Hdata, slope);
        rate = seg.derivative(Hdata, slope);
        break;
    default: return; // (outputs untouched, as the reference leaves them)
    }
    v = CComplex(field / b);
    dv = CComplex(0.5 * (rate / (b * b) - field / (b * b * b)));
}

void CMSolverMaterialProp::GetBHProps(double Bv, double &v, double &dv)
{
    CComplex vc, dvc;
    GetBHProps(Bv, vc, dvc);
    v = vc.re;
    dv = dvc.re;
}

// Apparent permeability of the laminated material at curve point i for AC problems
CComplex CMSolverMaterialProp::LaminatedBH(double w, int i)
{
    const cplx mu = Bdata[i] / cplx(Hdata[i]);
    LaminationSlab slab(*this, w, mu);
    const cplx a_surface = slab.run(cplx(Hdata[i]), mu);
    return CComplex(a_surface / (cplx(Hdata[i]) * slab.d));
}

void CMSolverMaterialProp::IncrementalPermeability(double B, double &mu1, double &mu2)
{
    const double incremental = 1. / (muo * GetdHdB(B).re);
    const double apparent = 1. / (muo * Get_v(B).re);
    const bool solid = (Lam_d == 0) || (LamFill == 0);
    auto stacked = [&](double mu) { return solid ? mu : (mu * LamFill + (1. - LamFill)); };
    mu1 = stacked(incremental);
    mu2 = stacked(apparent);
}

void CMSolverMaterialProp::incrementalPermeability(double B, double w, CComplex &mu1, CComplex &mu2)
{
    const double incremental = 1. / (muo * GetdHdB(B).re);
    const double apparent = 1. / (muo * Get_v(B).re);
    // relative permeability of the bulk -> complex permeability with hysteresis lag, sheet eddy currents and stacking
    auto effective = [&](double bulk) -> CComplex {
        if ((Lam_d == 0) || (LamFill == 0)) return CComplex(bulk * std::exp(-IMAG * Theta_hn * DEG * bulk / MuMax));
        const double mu = (bulk - (1. - LamFill)) / LamFill; // of the iron alone
        if (Cduct != 0)
        {
            const cplx deg45(1, 1);
            const cplx halflag = std::exp(-IMAG * Theta_hn * DEG * mu / (2. * MuMax));
            const double skin = std::sqrt(2. / (0.4 * PI * w * Cduct * mu));
            const cplx K = halflag * deg45 * Lam_d * 0.001 / (2. * skin);
            return CComplex(LamFill * mu * std::tanh(K) / K + (1. - LamFill));
        }
        return CComplex(mu * std::exp(-IMAG * Theta_hn * DEG * mu / MuMax) * LamFill + (1. - LamFill));
    };
    mu1 = effective(incremental);
    mu2 = effective(apparent);
}
