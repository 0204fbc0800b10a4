// CBigComplexLinProb facade (reference: cfemm/libfemm/cspars.cpp:119-1081), see cspars.h.
#include <cmath>
#include <vector>
#include "cspars.h"
#include "../../include/xfemm_b200.h"
#include <algorithm>
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
      precond(default_precond()), lastIterations(0), lastSolveMs(0), lastRelRes(0), patternDirty(true), valuesDirty(true),
      deviceAssembled(false), dev(nullptr)
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
    for (int k = 0; k < 4; k++)
    {
        std::vector<std::vector<Ent>>().swap(rows[k]);
        std::vector<std::vector<int>>().swap(lower[k]);
    }
    bNewton = 0;
    patternDirty = valuesDirty = true;
    return 1;
}

// host rows are created on first touch (the FSolver path never needs them)
void CBigComplexLinProb::ensureRows()
{
    if (!rows[0].empty() || n <= 0) return;
    for (int k = 0; k < 4; k++)
    {
        rows[k].assign(n, std::vector<Ent>());
        lower[k].assign(n, std::vector<int>());
    }
    for (int i = 0; i < n; i++) rows[0][i].push_back(Ent{i, CComplex(0)}); // M[i] starts with its diagonal (cspars.cpp:133-138)
}

void CBigComplexLinProb::attachDevice(xfemm_clinprob *d)
{
    if (dev && dev != d) xfemm_clinprob_destroy(dev);
    dev = d;
    deviceAssembled = true;
}

CComplex *CBigComplexLinProb::findEntry(int p, int q, int k)
{
    std::vector<Ent> &r = rows[k][p];
    auto it = std::lower_bound(r.begin(), r.end(), q, [](const Ent &e, int c) { return e.c < c; });
    return (it != r.end() && it->c == q) ? &it->x : nullptr;
}

// Put (cspars.cpp:146-237): stored on row min(p,q); a value given for the lower triangle is conjugated for the
// Hermitian matrix and negated-conjugated for the anti-Hermitian one
void CBigComplexLinProb::Put(CComplex v, int p, int q, int k)
{
    if (deviceAssembled) { fprintf(stderr, "CBigComplexLinProb::Put on a device-assembled problem is not supported\n"); return; }
    ensureRows();
    if (k < 0 || k > 3) k = 0;
    if (q < p)
    {
        std::swap(p, q);
        if (k == 1) v = conj(v);
        if (k == 3) v = -conj(v);
    }
    if (k > 0 && !bNewton)
    {
        // the reference allocates the side matrices with their diagonals here (cspars.cpp:162-186)
        bNewton = 1;
        for (int h = 1; h <= 3; h++)
            for (int i = 0; i < n; i++)
                if (rows[h][i].empty()) rows[h][i].push_back(Ent{i, CComplex(0)});
        patternDirty = true;
    }
    std::vector<Ent> &r = rows[k][p];
    auto it = std::lower_bound(r.begin(), r.end(), q, [](const Ent &e, int c) { return e.c < c; });
    if (it != r.end() && it->c == q) { it->x = v; valuesDirty = true; return; }
    r.insert(it, Ent{q, v});
    if (p != q) lower[k][q].push_back(p);
    patternDirty = valuesDirty = true;
}

CComplex CBigComplexLinProb::Get(int p, int q, int k)
{
    if (deviceAssembled) { fprintf(stderr, "CBigComplexLinProb::Get on a device-assembled problem is not supported\n"); return CComplex(0); }
    ensureRows();
    if (k < 0 || k > 3) k = 0;
    if (k > 0 && !bNewton) return CComplex(0);
    bool flip = false;
    if (q < p) { std::swap(p, q); flip = true; }
    CComplex *e = findEntry(p, q, k);
    if (!e) return CComplex(0);
    if (flip && k == 1) return conj(*e);
    if (flip && k == 3) return -conj(*e);
    return *e;
}

void CBigComplexLinProb::AddTo(CComplex v, int p, int q) { Put(Get(p, q) + v, p, q); }

CComplex CBigComplexLinProb::Dot(CComplex *x, CComplex *y)
{
    CComplex z(0);
    for (int i = 0; i < n; i++) z += x[i] * y[i];
    return z;
}
CComplex CBigComplexLinProb::ConjDot(CComplex *x, CComplex *y)
{
    CComplex z(0);
    for (int i = 0; i < n; i++) z += conj(x[i]) * y[i];
    return z;
}

void CBigComplexLinProb::Wipe()
{
    if (deviceAssembled) { if (dev) xfemm_clinprob_wipe(dev); }
    else
    {
        ensureRows();
        for (int k = 0; k < 4; k++)
            for (auto &r : rows[k])
                for (auto &e : r) e.x = 0;
    }
    for (int i = 0; i < n; i++) b[i] = 0;
    valuesDirty = true;
}

// rows structurally coupled to i in any of the matrices (the reference scans a band of rows; only these can
// hold a non-zero entry in column i)
void CBigComplexLinProb::coupledRows(int i, std::vector<int> &out) const
{
    out.clear();
    for (int k = 0; k < 4; k++)
    {
        if (rows[k].empty()) continue;
        out.insert(out.end(), lower[k][i].begin(), lower[k][i].end());
        for (auto &e : rows[k][i]) out.push_back(e.c);
    }
    out.push_back(i);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

// SetValue (cspars.cpp:482-537)
void CBigComplexLinProb::SetValue(int i, CComplex x)
{
    if (deviceAssembled) { fprintf(stderr, "CBigComplexLinProb::SetValue: use xfemm_clinprob_dirichlet on device problems\n"); return; }
    ensureRows();
    int fst = 0, lst = n;
    if (bdw != 0) { fst = std::max(0, i - bdw); lst = std::min(NumNodes, i + bdw); }
    std::vector<int> ks;
    coupledRows(i, ks);
    for (int k : ks)
    {
        // scanned rows: [fst,lst) and the circuit rows [NumNodes,n)
        if (!((k >= fst && k < lst) || (bdw != 0 && k >= NumNodes))) continue;
        CComplex z = Get(k, i);
        if (z != 0)
        {
            b[k] -= z * x;
            if (i != k) Put(CComplex(0), k, i);
        }
        if (bNewton)
        {
            z = Get(k, i, 1);
            if (z != 0) { if (i != k) b[k] = b[k] - z * x; Put(CComplex(0), k, i, 1); }
            z = Get(k, i, 2);
            if (z != 0) { if (i != k) b[k] = b[k] - z * conj(x); Put(CComplex(0), k, i, 2); }
            z = Get(k, i, 3);
            if (z != 0) { if (i != k) b[k] = b[k] - z * conj(x); Put(CComplex(0), k, i, 3); }
        }
    }
    b[i] = Get(i, i) * x;
    valuesDirty = true;
}

// Periodicity / AntiPeriodicity (cspars.cpp:592-758; KLUDGE is defined there, so every row is scanned)
void CBigComplexLinProb::periodicImpl(int i, int j, double sgn)
{
    if (deviceAssembled) { fprintf(stderr, "CBigComplexLinProb::Periodicity: use xfemm_clinprob_periodic on device problems\n"); return; }
    ensureRows();
    if (j < i) std::swap(i, j);
    std::vector<int> ks, kj;
    coupledRows(i, ks);
    coupledRows(j, kj);
    ks.insert(ks.end(), kj.begin(), kj.end());
    std::sort(ks.begin(), ks.end());
    ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
    for (int h = 0; h <= (bNewton ? 3 : 0); h++)
    {
        for (int k : ks)
        {
            if (k == i || k == j) continue;
            CComplex v1 = Get(k, i, h), v2 = Get(k, j, h);
            if (v1 != 0 || v2 != 0)
            {
                CComplex c = (v1 + sgn * v2) / 2.;
                Put(c, k, i, h);
                Put(sgn * c, k, j, h);
            }
        }
        if (h == 0)
        {
            CComplex c = 0.5 * (Get(i, i) + Get(j, j));
            Put(c, i, i);
            Put(c, j, j);
            c = 0.5 * (b[i] + sgn * b[j]);
            b[i] = c;
            b[j] = sgn * c;
        }
        else
        {
            CComplex c = (Get(i, i, h) + sgn * Get(i, j, h) + sgn * Get(j, i, h) + Get(j, j, h)) / 4.;
            Put(c, i, i, h);
            Put(sgn * c, i, j, h);
            Put(c, j, j, h);
        }
    }
}
void CBigComplexLinProb::Periodicity(int i, int j) { periodicImpl(i, j, 1.0); }
void CBigComplexLinProb::AntiPeriodicity(int i, int j) { periodicImpl(i, j, -1.0); }

// frozen pattern (union of the four matrices) + values → device problem
bool CBigComplexLinProb::syncToDevice()
{
    if (deviceAssembled) return dev != nullptr;
    ensureRows();
    const int nsel = bNewton ? 4 : 1;
    if (patternDirty || !dev)
    {
        std::vector<std::vector<int>> up(n);
        for (int k = 0; k < nsel; k++)
            for (int i = 0; i < n; i++)
                for (auto &e : rows[k][i]) up[i].push_back(e.c);
        std::vector<int> cnt(n + 1, 0);
        for (int i = 0; i < n; i++)
        {
            std::sort(up[i].begin(), up[i].end());
            up[i].erase(std::unique(up[i].begin(), up[i].end()), up[i].end());
            for (int c : up[i]) { cnt[i + 1]++; if (c != i) cnt[c + 1]++; }
        }
        for (int i = 0; i < n; i++) cnt[i + 1] += cnt[i];
        csrRowptr = cnt;
        csrCol.assign(cnt[n], 0);
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int r = 0; r < n; r++)
            for (int c : up[r]) if (c != r) csrCol[pos[c]++] = r; // lower part first: ascending rows = sorted columns
        for (int i = 0; i < n; i++)
            for (int c : up[i]) csrCol[pos[i]++] = c;
        if (dev) { xfemm_clinprob_destroy(dev); dev = nullptr; }
        if (xfemm_clinprob_create(n, NumNodes, csrRowptr.data(), csrCol.data(), &dev) != XFEMM_OK)
        {
            fprintf(stderr, "xfemm_b200: %s\n", xfemm_last_error());
            return false;
        }
        patternDirty = false;
        valuesDirty = true;
    }
    if (bNewton && xfemm_clinprob_enable_newton(dev) != XFEMM_OK) { fprintf(stderr, "xfemm_b200: %s\n", xfemm_last_error()); return false; }
    if (valuesDirty)
    {
        auto slot = [&](int r, int c) {
            const int *b0 = csrCol.data() + csrRowptr[r], *e0 = csrCol.data() + csrRowptr[r + 1];
            return (int)(std::lower_bound(b0, e0, c) - csrCol.data());
        };
        std::vector<CComplex> vals(csrCol.size());
        for (int k = 0; k < nsel; k++)
        {
            std::fill(vals.begin(), vals.end(), CComplex(0));
            for (int i = 0; i < n; i++)
                for (auto &e : rows[k][i])
                {
                    vals[slot(i, e.c)] = e.x;
                    if (e.c != i) vals[slot(e.c, i)] = k == 1 ? conj(e.x) : (k == 3 ? -conj(e.x) : e.x);
                }
            int rc = k == 0 ? xfemm_clinprob_set_values(dev, (const double *)vals.data())
                            : xfemm_clinprob_set_side_values(dev, k, (const double *)vals.data());
            if (rc != XFEMM_OK) { fprintf(stderr, "xfemm_b200: %s\n", xfemm_last_error()); return false; }
        }
        valuesDirty = false;
    }
    return true;
}

void CBigComplexLinProb::MultA(CComplex *X, CComplex *Y, int k)
{
    if (!syncToDevice()) return;
    if (!bNewton && k != 0) k = 0; // cspars.cpp:299
    int rc = k == -1 ? xfemm_clinprob_mult_newton(dev, (const double *)X, (double *)Y)
                     : xfemm_clinprob_mult(dev, k, 0, (const double *)X, (double *)Y);
    if (rc != XFEMM_OK) fprintf(stderr, "xfemm_b200: MultA failed: %s\n", xfemm_last_error());
}

void CBigComplexLinProb::MultConjA(CComplex *X, CComplex *Y, int k)
{
    if (!syncToDevice()) return;
    if (!bNewton && k != 0) k = 0; // cspars.cpp:369
    if (k != 0 && k != 2) { fprintf(stderr, "xfemm_b200: MultConjA is provided for k = 0 and 2 (the ones the solvers use)\n"); return; }
    if (xfemm_clinprob_mult(dev, k, 1, (const double *)X, (double *)Y) != XFEMM_OK)
        fprintf(stderr, "xfemm_b200: MultConjA failed: %s\n", xfemm_last_error());
}

// ---- stand-alone pieces of the reference's public surface ------------------------------------------------
void CBigComplexLinProb::MultPC(CComplex *X, CComplex *Y)
{
    // y = D^-1 x (see the header for why this is not the reference's SSOR sweep)
    for (int i = 0; i < n; i++)
    {
        const CComplex d = Get(i, i);
        Y[i] = (d.re != 0 || d.im != 0) ? X[i] / d : X[i];
    }
}

void CBigComplexLinProb::MultAPPA(CComplex *X, CComplex *Y)
{
    // Y = conj( A P conj( P A X ) ): the Hermitian positive operator PCGSQStart iterates on (cspars.cpp:406-415)
    MultA(X, Z);
    MultPC(Z, Y);
    for (int i = 0; i < n; i++) Y[i] = Y[i].Conj();
    MultPC(Y, Z);
    MultA(Z, Y);
    for (int i = 0; i < n; i++) Y[i] = Y[i].Conj();
}

int CBigComplexLinProb::PCGSQStart()
{
    // three conjugate-gradient steps on the squared problem to seed V (cspars.cpp:764-819)
    for (int i = 0; i < n; i++)
    {
        const CComplex d = Get(i, i);
        if (d.re == 0 && d.im == 0) { fprintf(stderr, "singular flag tripped."); return 0; }
    }
    // right-hand side of the squared problem: conj( A P conj( P b ) )
    MultPC(b, Z);
    for (int i = 0; i < n; i++) Z[i] = Z[i].Conj();
    MultPC(Z, P);
    MultA(P, Z);
    for (int i = 0; i < n; i++) { R[i] = Z[i].Conj(); V[i] = 0; P[i] = R[i]; }
    CComplex res = ConjDot(R, R);
    for (int k = 0; k < 3; k++)
    {
        if (res.re == 0 && res.im == 0) break;
        MultAPPA(P, U);
        const CComplex del = res / ConjDot(P, U);
        for (int i = 0; i < n; i++) { V[i] += del * P[i]; R[i] -= del * U[i]; }
        const CComplex resNew = ConjDot(R, R), rho = resNew / res;
        res = resNew;
        for (int i = 0; i < n; i++) P[i] = R[i] + rho * P[i];
    }
    return 1;
}

int CBigComplexLinProb::BiCGSTAB(int flag)
{
    // right-preconditioned BiCGSTAB on the full Newton operator (MultA(.,.,-1)), cspars.cpp:897-980
    auto norm2 = [&](const CComplex *x) { double s = 0; for (int i = 0; i < n; i++) s += x[i].re * x[i].re + x[i].im * x[i].im; return sqrt(s); };
    std::vector<CComplex> P2(n), Z2(n), R2(n), T(n);
    if (!flag) for (int i = 0; i < n; i++) V[i] = 0;
    MultA(V, R, -1);
    for (int i = 0; i < n; i++) { R[i] = b[i] - R[i]; R2[i] = R[i]; }
    const double normb = norm2(b);
    if (normb == 0) return 1;
    CComplex alf = 0, om = 0, rho1 = 0, rho2 = 0;
    double er = norm2(R) / normb;
    for (int k = 0; k < 1000000 && er >= Precision; k++)
    {
        rho1 = ConjDot(R2.data(), R).re;
        if (k == 0) for (int i = 0; i < n; i++) P[i] = R[i];
        else
        {
            const CComplex bta = (rho1 / rho2) * alf / om;
            for (int i = 0; i < n; i++) P[i] = R[i] + bta * (P[i] - om * U[i]);
        }
        MultPC(P, P2.data());
        MultA(P2.data(), U, -1);
        alf = rho1 / ConjDot(R2.data(), U).re;
        for (int i = 0; i < n; i++) Z[i] = R[i] - alf * U[i];
        MultPC(Z, Z2.data());
        MultA(Z2.data(), T.data(), -1);
        om = ConjDot(T.data(), Z).re / ConjDot(T.data(), T.data()).re;
        for (int i = 0; i < n; i++) { V[i] = V[i] + alf * P2[i] + om * Z2[i]; R[i] = Z[i] - om * T[i]; }
        rho2 = rho1;
        er = norm2(R) / normb;
        if (!(er == er)) return 0;
    }
    return er < Precision ? 1 : 0;
}

int CBigComplexLinProb::solve(int flag, bool sqStart)
{
    if (!dev && deviceAssembled)
    {
        fprintf(stderr, "xfemm_b200: CBigComplexLinProb has no device problem attached\n");
        return 0;
    }
    if (!syncToDevice()) return 0;
    if (!deviceAssembled)
    {
        if (xfemm_clinprob_set_rhs(dev, (const double *)b) != XFEMM_OK) return 0;
        if (!flag) for (int i = 0; i < n; i++) V[i] = 0;
        if ((flag || !sqStart) && xfemm_clinprob_set_solution(dev, (const double *)V) != XFEMM_OK) return 0;
    }
    xfemm_solve_info info;
    memset(&info, 0, sizeof info);
    int rc;
    if (bNewton && sqStart) // PBCGSolveMod on a Newton system → KludgeSolve (cspars.cpp:1065-1067)
        rc = xfemm_clinprob_kludge_solve(dev, Precision, flag, precond, 0, &info);
    else
    {
        // a multigrid attempt that does not converge within 3000 iterations is repeated with Jacobi (see FSolver)
        const int warm = (flag || !sqStart) ? 1 : 0; // PBCGSolve(false) = COCG from zero without the CGNR start
        rc = xfemm_clinprob_cocg(dev, Precision, warm, precond, precond == XFEMM_PRECOND_AMG ? 3000 : 0, &info);
        if (rc == XFEMM_ERR_DIVERGED && precond == XFEMM_PRECOND_AMG)
            rc = xfemm_clinprob_cocg(dev, Precision, 0, XFEMM_PRECOND_JACOBI, 0, &info);
    }
    lastIterations = info.iterations; lastSolveMs = info.solve_ms; lastRelRes = info.relres;
    if (rc != XFEMM_OK) { fprintf(stderr, "%s\n", xfemm_last_error()); return 0; }
    return xfemm_clinprob_get_solution(dev, (double *)V) == XFEMM_OK;
}

int CBigComplexLinProb::PBCGSolveMod(int flag, bool) { return solve(flag, true); }
int CBigComplexLinProb::PBCGSolve(int flag) { return solve(flag, false); }
int CBigComplexLinProb::KludgeSolve(int flag)
{
    if (!bNewton) return solve(flag, false);
    return solve(flag, true);
}
