// CBigLinProb facade over the device solver (reference semantics: cfemm/libfemm/spars.cpp:80-474).
#include "spars.h"
#include "../../include/xfemm_b200.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static int default_precond()
{
    // Three GPU-parallel replacements of the serial SSOR are available; the default is the one
    // with the lowest measured time to solution on B200.
    const char *e = getenv("XFEMM_PRECOND");
    if (!e) return XFEMM_PRECOND_AMG;
    if (!strcmp(e, "jacobi") || !strcmp(e, "0")) return XFEMM_PRECOND_JACOBI;
    if (!strcmp(e, "mcssor") || !strcmp(e, "1")) return XFEMM_PRECOND_MCSSOR;
    return XFEMM_PRECOND_AMG;
}

CBigLinProb::CBigLinProb()
    : V(nullptr), P(nullptr), R(nullptr), U(nullptr), Z(nullptr), b(nullptr), M(nullptr), n(0), bdw(0),
      Precision(1e-8), Lambda(1.5), Q(nullptr), precond(default_precond()), lastIterations(0), lastSolveMs(0),
      lastRelRes(0), patternDirty(true), valuesDirty(true), hostValuesValid(true), deviceAssembled(false), dev(nullptr)
{
}

void CBigLinProb::pinResultArrays()
{
    if (resultPinned || !V || !b || n <= 0) return;
    if (xfemm_host_register(V, sizeof(double) * (size_t)n) != XFEMM_OK) return;
    if (xfemm_host_register(b, sizeof(double) * (size_t)n) != XFEMM_OK) { xfemm_host_unregister(V); return; }
    resultPinned = true;
}

void CBigLinProb::unpinResultArrays()
{
    if (!resultPinned) return;
    xfemm_host_unregister(V);
    xfemm_host_unregister(b);
    resultPinned = false;
}

CBigLinProb::~CBigLinProb()
{
    unpinResultArrays();
    if (dev) xfemm_linprob_destroy(dev);
    free(V); free(P); free(R); free(U); free(Z); free(b); free(M); free(Q);
    n = 0;
}

int CBigLinProb::Create(int d, int bw)
{
    unpinResultArrays();
    bdw = bw;
    b = (double *)calloc(d, sizeof(double));
    V = (double *)calloc(d, sizeof(double));
    P = (double *)calloc(d, sizeof(double));
    R = (double *)calloc(d, sizeof(double));
    U = (double *)calloc(d, sizeof(double));
    Z = (double *)calloc(d, sizeof(double));
    M = (CEntry **)calloc(d, sizeof(CEntry *)); // not populated; see header
    Q = (int *)calloc(d, sizeof(int));
    n = d;
    // the host-side rows (used only when a caller assembles through Put(), i.e. the hsolver/esolver
    // style of use) are created on first touch: millions of tiny vectors are not free
    std::vector<std::vector<Ent>>().swap(rows);
    std::vector<std::vector<int>>().swap(lower);
    patternDirty = valuesDirty = true;
    return 1;
}

void CBigLinProb::ensureRows()
{
    if (!rows.empty() || n <= 0) return;
    rows.assign(n, std::vector<Ent>());
    lower.assign(n, std::vector<int>());
    for (int i = 0; i < n; i++) rows[i].push_back(Ent{i, 0.0});
}

void CBigLinProb::attachDevice(xfemm_linprob *d, const int *rowptr, const int *col)
{
    if (dev && dev != d) xfemm_linprob_destroy(dev);
    dev = d;
    // the pattern stays with the device problem (it keeps a host copy); Get()/patternRowptr() fetch it on first use
    (void)rowptr; (void)col;
    std::vector<int>().swap(csrRowptr);
    std::vector<int>().swap(csrCol);
    deviceAssembled = true;
    hostValuesValid = false;
    // the host rows are not used in this mode
    std::vector<std::vector<Ent>>().swap(rows);
    std::vector<std::vector<int>>().swap(lower);
}

void CBigLinProb::ensurePattern()
{
    if (!deviceAssembled || !dev || !csrRowptr.empty()) return;
    csrRowptr.resize((size_t)n + 1);
    csrCol.resize((size_t)xfemm_linprob_nnz(dev));
    xfemm_linprob_get_pattern(dev, csrRowptr.data(), csrCol.data());
}

double *CBigLinProb::findEntry(int p, int q)
{
    if (!deviceAssembled) ensureRows();
    if (q < p) std::swap(p, q);
    std::vector<Ent> &r = rows[p];
    auto it = std::lower_bound(r.begin(), r.end(), q, [](const Ent &e, int c) { return e.c < c; });
    if (it != r.end() && it->c == q) return &it->x;
    return nullptr;
}

void CBigLinProb::fetchDeviceValues()
{
    // device-assembled problem queried through Get(): pull pattern values once
    if (hostValuesValid || !dev) return;
    long long nnz = xfemm_linprob_nnz(dev);
    csrVals.resize((size_t)nnz);
    xfemm_linprob_get_values(dev, csrVals.data());
    hostValuesValid = true;
}

void CBigLinProb::Put(double v, int p, int q)
{
    if (!deviceAssembled) ensureRows();
    if (deviceAssembled)
    {
        fprintf(stderr, "CBigLinProb::Put on a device-assembled problem is not supported\n");
        return;
    }
    if (q < p) std::swap(p, q);
    std::vector<Ent> &r = rows[p];
    auto it = std::lower_bound(r.begin(), r.end(), q, [](const Ent &e, int c) { return e.c < c; });
    if (it != r.end() && it->c == q) { it->x = v; valuesDirty = true; return; }
    r.insert(it, Ent{q, v});
    if (p != q) lower[q].push_back(p);
    patternDirty = valuesDirty = true;
}

double CBigLinProb::Get(int p, int q)
{
    if (deviceAssembled)
    {
        fetchDeviceValues();
        ensurePattern();
        if (csrRowptr.empty()) return 0;
        const int *b0 = csrCol.data() + csrRowptr[p], *e0 = csrCol.data() + csrRowptr[p + 1];
        const int *it = std::lower_bound(b0, e0, q);
        return (it != e0 && *it == q) ? csrVals[it - csrCol.data()] : 0.0;
    }
    double *e = findEntry(p, q);
    return e ? *e : 0.0;
}

void CBigLinProb::AddTo(double v, int p, int q) { Put(Get(p, q) + v, p, q); }

CEntry **CBigLinProb::Entries()
{
    if (!M || n <= 0) return M;
    entryPool.clear();
    std::vector<size_t> first((size_t)n + 1, 0);
    if (deviceAssembled)
    {
        fetchDeviceValues();
        ensurePattern();
        if (csrRowptr.empty()) return M;
        for (int i = 0; i < n; i++)
        {
            first[i] = entryPool.size();
            for (int s = csrRowptr[i]; s < csrRowptr[i + 1]; s++)
                if (csrCol[s] >= i) { CEntry e; e.c = csrCol[s]; e.x = csrVals[s]; entryPool.push_back(e); }
        }
    }
    else
    {
        ensureRows();
        for (int i = 0; i < n; i++)
        {
            first[i] = entryPool.size();
            for (const Ent &r : rows[i]) { CEntry e; e.c = r.c; e.x = r.x; entryPool.push_back(e); }
        }
    }
    first[n] = entryPool.size();
    for (int i = 0; i < n; i++)
    {
        M[i] = first[i + 1] > first[i] ? &entryPool[first[i]] : nullptr;
        for (size_t k = first[i]; k + 1 < first[i + 1]; k++) entryPool[k].next = &entryPool[k + 1];
    }
    return M;
}

double CBigLinProb::Dot(double *X, double *Y)
{
    double z = 0;
    for (int i = 0; i < n; i++) z += X[i] * Y[i];
    return z;
}

void CBigLinProb::Wipe()
{
    if (!deviceAssembled) ensureRows();
    if (deviceAssembled) { if (dev) xfemm_linprob_wipe(dev); hostValuesValid = false; }
    else
        for (auto &r : rows)
            for (auto &e : r) e.x = 0;
    for (int i = 0; i < n; i++) b[i] = 0.;
    valuesDirty = true;
}

// SetValue with the reference's band restriction, visiting only the structurally coupled rows
void CBigLinProb::SetValue(int i, double x)
{
    if (!deviceAssembled) ensureRows();
    if (deviceAssembled) { fprintf(stderr, "CBigLinProb::SetValue: use xfemm_linprob_dirichlet on device problems\n"); return; }
    int fst = 0, lst = n;
    if (bdw != 0) { fst = std::max(0, i - bdw); lst = std::min(n, i + bdw); }
    for (int k : lower[i])
    {
        if (k < fst || k >= lst) continue;
        double *e = findEntry(k, i);
        if (e && *e != 0) { b[k] -= *e * x; *e = 0; }
    }
    for (auto &e : rows[i])
    {
        if (e.c == i || e.c < fst || e.c >= lst) continue;
        if (e.x != 0) { b[e.c] -= e.x * x; e.x = 0; }
    }
    if (i >= fst && i < lst) { double z = Get(i, i); if (z != 0) b[i] -= z * x; }
    b[i] = Get(i, i) * x;
    valuesDirty = true;
}

void CBigLinProb::periodicImpl(int i, int j, double sgn)
{
    if (!deviceAssembled) ensureRows();
    if (deviceAssembled) { fprintf(stderr, "CBigLinProb::Periodicity: use xfemm_linprob_periodic on device problems\n"); return; }
    if (j < i) std::swap(i, j);
    // the reference scans every row k (KLUDGE sets bdw=0, spars.cpp:371-374); only rows
    // structurally coupled to i or j can have a non-zero entry
    std::vector<int> ks;
    for (int r : {i, j})
    {
        ks.insert(ks.end(), lower[r].begin(), lower[r].end());
        for (auto &e : rows[r]) ks.push_back(e.c);
    }
    std::sort(ks.begin(), ks.end());
    ks.erase(std::unique(ks.begin(), ks.end()), ks.end());
    for (int k : ks)
    {
        if (k == i || k == j) continue;
        double v1 = Get(k, i), v2 = Get(k, j);
        if (v1 != 0 || v2 != 0)
        {
            double c = (v1 + sgn * v2) / 2.;
            Put(c, k, i);
            Put(sgn * c, k, j);
        }
    }
    double c = 0.5 * (Get(i, i) + Get(j, j));
    Put(c, i, i);
    Put(c, j, j);
    c = 0.5 * (b[i] + sgn * b[j]);
    b[i] = c;
    b[j] = sgn * c;
}

void CBigLinProb::Periodicity(int i, int j) { periodicImpl(i, j, 1.0); }
void CBigLinProb::AntiPeriodicity(int i, int j) { periodicImpl(i, j, -1.0); }

bool CBigLinProb::syncToDevice()
{
    if (!deviceAssembled) ensureRows();
    if (deviceAssembled) return dev != nullptr;
    if (patternDirty || !dev)
    {
        // upper rows → full CSR
        std::vector<int> cnt(n + 1, 0);
        for (int i = 0; i < n; i++)
            for (auto &e : rows[i])
            {
                cnt[i + 1]++;
                if (e.c != i) cnt[e.c + 1]++;
            }
        for (int i = 0; i < n; i++) cnt[i + 1] += cnt[i];
        csrRowptr = cnt;
        csrCol.assign(cnt[n], 0);
        std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        // lower part first (rows k<i ascending gives sorted columns), then the upper part
        for (int k = 0; k < n; k++)
            for (auto &e : rows[k])
                if (e.c != k) csrCol[pos[e.c]++] = k;
        for (int i = 0; i < n; i++)
            for (auto &e : rows[i]) csrCol[pos[i]++] = e.c;
        if (dev) { xfemm_linprob_destroy(dev); dev = nullptr; }
        if (xfemm_linprob_create(n, csrRowptr.data(), csrCol.data(), &dev) != XFEMM_OK)
        {
            fprintf(stderr, "xfemm_b200: %s\n", xfemm_last_error());
            return false;
        }
        patternDirty = false;
        valuesDirty = true;
    }
    if (valuesDirty)
    {
        csrVals.assign(csrCol.size(), 0.0);
        std::vector<int> pos(csrRowptr.begin(), csrRowptr.end() - 1);
        for (int k = 0; k < n; k++)
            for (auto &e : rows[k])
                if (e.c != k) csrVals[pos[e.c]++] = e.x;
        for (int i = 0; i < n; i++)
            for (auto &e : rows[i]) csrVals[pos[i]++] = e.x;
        if (xfemm_linprob_set_values(dev, csrVals.data()) != XFEMM_OK)
        {
            fprintf(stderr, "xfemm_b200: %s\n", xfemm_last_error());
            return false;
        }
        valuesDirty = false;
    }
    return true;
}

bool CBigLinProb::PCGSolve(int flag)
{
    if (!syncToDevice()) return false;
    if (!deviceAssembled)
    {
        if (xfemm_linprob_set_rhs(dev, b) != XFEMM_OK) return false;
        if (flag && xfemm_linprob_set_solution(dev, V) != XFEMM_OK) return false;
    }
    xfemm_solve_info info;
    int rc = xfemm_linprob_pcg(dev, Precision, flag, precond, 0, &info);
    if (rc == XFEMM_ERR_DIVERGED && precond == XFEMM_PRECOND_AMG) // safety net, see FSolver::staticNewton
        rc = xfemm_linprob_pcg(dev, Precision, 0, XFEMM_PRECOND_JACOBI, 0, &info);
    lastIterations = info.iterations;
    lastSolveMs = info.solve_ms;
    lastRelRes = info.relres;
    if (rc != XFEMM_OK)
    {
        fprintf(stderr, "%s\n", xfemm_last_error());
        return false;
    }
    return xfemm_linprob_get_solution(dev, V) == XFEMM_OK;
}

void CBigLinProb::MultA(double *X, double *Y)
{
    if (!syncToDevice() || xfemm_linprob_multa(dev, X, Y) != XFEMM_OK)
        fprintf(stderr, "xfemm_b200: MultA failed: %s\n", xfemm_last_error());
}

void CBigLinProb::MultPC(const double *X, double *Y)
{
    if (!syncToDevice() || xfemm_linprob_multpc(dev, precond, X, Y) != XFEMM_OK)
        fprintf(stderr, "xfemm_b200: MultPC failed: %s\n", xfemm_last_error());
}

void CBigLinProb::ComputeBandwidth()
{
    int maxbw = 0;
    for (int k = 0; k < n && !rows.empty(); k++)
        if (!rows[k].empty()) maxbw = std::max(maxbw, rows[k].back().c - k);
    printf("Assumed Bandwidth = %i\nActual Bandwidth = %i", bdw, maxbw);
}
