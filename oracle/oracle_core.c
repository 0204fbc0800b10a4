/* oracle_core.c — CPU restatement (plain C) of the reference algorithms on the fsolver hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the
 * product (xfemm_b200/); allowed users are tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg.  Parity status: PINNED — every routine here is checked in
 * tests/test_oracle.py against outputs of the unmodified reference (oracle/_ref, golden vectors
 * under tests/golden/: renumbering, element order, assembled system, iteration counts 116/117
 * and 220/149/1, solution).
 *
 * Each function cites the reference lines it follows (paths relative to the xfemm tree).
 * The sparse matrix keeps the reference's storage idea — per row, the diagonal first and then the
 * upper-triangular entries in ascending column order (cfemm/libfemm/spars.h:25-35) — in growable
 * arrays instead of linked lists.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.141592653589793238462643383
#define ORC_MUO 1.2566370614359173e-6

typedef struct
{
    int n, bdw;
    int *len, *cap;
    int **col;
    double **val;
    double *b, *V, *P, *R, *U, *Z;
    double lambda, precision;
} orc_mat;

/* ---- spars.cpp:80-103 Create */
orc_mat *orc_create(int n, int bdw)
{
    orc_mat *m = (orc_mat *)calloc(1, sizeof(orc_mat));
    m->n = n; m->bdw = bdw; m->lambda = 1.5; m->precision = 1e-8;
    m->len = (int *)calloc(n, sizeof(int));
    m->cap = (int *)calloc(n, sizeof(int));
    m->col = (int **)calloc(n, sizeof(int *));
    m->val = (double **)calloc(n, sizeof(double *));
    m->b = (double *)calloc(n, sizeof(double)); m->V = (double *)calloc(n, sizeof(double));
    m->P = (double *)calloc(n, sizeof(double)); m->R = (double *)calloc(n, sizeof(double));
    m->U = (double *)calloc(n, sizeof(double)); m->Z = (double *)calloc(n, sizeof(double));
    for (int i = 0; i < n; i++)
    {
        m->cap[i] = 8;
        m->col[i] = (int *)malloc(8 * sizeof(int));
        m->val[i] = (double *)malloc(8 * sizeof(double));
        m->len[i] = 1;
        m->col[i][0] = i;
        m->val[i][0] = 0.0;
    }
    return m;
}

void orc_destroy(orc_mat *m)
{
    if (!m) return;
    for (int i = 0; i < m->n; i++) { free(m->col[i]); free(m->val[i]); }
    free(m->len); free(m->cap); free(m->col); free(m->val);
    free(m->b); free(m->V); free(m->P); free(m->R); free(m->U); free(m->Z);
    free(m);
}

double *orc_b(orc_mat *m) { return m->b; }
double *orc_V(orc_mat *m) { return m->V; }
void orc_set_precision(orc_mat *m, double p) { m->precision = p; }

/* ---- spars.cpp:105-142 Put: symmetric, stored in row min(p,q), sorted insert */
void orc_put(orc_mat *m, double v, int p, int q)
{
    if (q < p) { int t = p; p = q; q = t; }
    int *c = m->col[p];
    int len = m->len[p], k = 0;
    while (k < len && c[k] < q) k++;
    if (k < len && c[k] == q) { m->val[p][k] = v; return; }
    if (len == m->cap[p])
    {
        m->cap[p] *= 2;
        m->col[p] = (int *)realloc(m->col[p], m->cap[p] * sizeof(int));
        m->val[p] = (double *)realloc(m->val[p], m->cap[p] * sizeof(double));
    }
    memmove(m->col[p] + k + 1, m->col[p] + k, (len - k) * sizeof(int));
    memmove(m->val[p] + k + 1, m->val[p] + k, (len - k) * sizeof(double));
    m->col[p][k] = q;
    m->val[p][k] = v;
    m->len[p] = len + 1;
}

/* ---- spars.cpp:144-160 Get */
double orc_get(orc_mat *m, int p, int q)
{
    if (q < p) { int t = p; p = q; q = t; }
    const int *c = m->col[p];
    for (int k = 0; k < m->len[p] && c[k] <= q; k++)
        if (c[k] == q) return m->val[p][k];
    return 0.0;
}

/* ---- spars.cpp:162-165 AddTo */
void orc_addto(orc_mat *m, double v, int p, int q) { orc_put(m, orc_get(m, p, q) + v, p, q); }

/* ---- spars.cpp:348-364 Wipe */
void orc_wipe(orc_mat *m)
{
    for (int i = 0; i < m->n; i++)
    {
        m->b[i] = 0.0;
        for (int k = 0; k < m->len[i]; k++) m->val[i][k] = 0.0;
    }
}

/* ---- spars.cpp:318-346 SetValue (band limited scan) */
void orc_setvalue(orc_mat *m, int i, double x)
{
    int fst = 0, lst = m->n;
    if (m->bdw != 0)
    {
        fst = i - m->bdw; if (fst < 0) fst = 0;
        lst = i + m->bdw; if (lst > m->n) lst = m->n;
    }
    for (int k = fst; k < lst; k++)
    {
        double z = orc_get(m, k, i);
        if (z != 0)
        {
            m->b[k] = m->b[k] - (z * x);
            if (i != k) orc_put(m, 0., k, i);
        }
    }
    m->b[i] = orc_get(m, i, i) * x;
}

/* ---- spars.cpp:421-474 Periodicity / 366-419 AntiPeriodicity (KLUDGE: whole-matrix scan) */
void orc_periodicity(orc_mat *m, int i, int j, int anti)
{
    const double s = anti ? -1.0 : 1.0;
    if (j < i) { int t = i; i = j; j = t; }
    for (int k = 0; k < m->n; k++)
    {
        if (k == i || k == j) continue;
        double v1 = orc_get(m, k, i), v2 = orc_get(m, k, j);
        if (v1 != 0 || v2 != 0)
        {
            double c = (v1 + s * v2) / 2.;
            orc_put(m, c, k, i);
            orc_put(m, s * c, k, j);
        }
    }
    double c = 0.5 * (orc_get(m, i, i) + orc_get(m, j, j));
    orc_put(m, c, i, i);
    orc_put(m, c, j, j);
    c = 0.5 * (m->b[i] + s * m->b[j]);
    m->b[i] = c;
    m->b[j] = s * c;
}

/* ---- spars.cpp:167-185 MultA (upper storage, scatter form) */
void orc_multa(orc_mat *m, const double *X, double *Y)
{
    for (int i = 0; i < m->n; i++) Y[i] = 0;
    for (int i = 0; i < m->n; i++)
    {
        Y[i] += m->val[i][0] * X[i];
        for (int k = 1; k < m->len[i]; k++)
        {
            int c = m->col[i][k];
            double a = m->val[i][k];
            Y[i] += a * X[c];
            Y[c] += a * X[i];
        }
    }
}

/* ---- spars.cpp:187-195 Dot */
double orc_dot(const orc_mat *m, const double *X, const double *Y)
{
    double z = 0;
    for (int i = 0; i < m->n; i++) z += X[i] * Y[i];
    return z;
}

/* ---- spars.cpp:197-236 MultPC: SSOR, y = L(2-L) (D+LU)^-1 D (D+LL)^-1 x */
void orc_multpc(orc_mat *m, const double *X, double *Y)
{
    const double lam = m->lambda, c = lam * (2. - lam);
    for (int i = 0; i < m->n; i++) Y[i] = X[i] * c;
    for (int i = 0; i < m->n; i++)
    {
        Y[i] /= m->val[i][0];
        for (int k = 1; k < m->len[i]; k++) Y[m->col[i][k]] -= m->val[i][k] * Y[i] * lam;
    }
    for (int i = 0; i < m->n; i++) Y[i] *= m->val[i][0];
    for (int i = m->n - 1; i >= 0; i--)
    {
        for (int k = 1; k < m->len[i]; k++) Y[i] -= m->val[i][k] * Y[m->col[i][k]] * lam;
        Y[i] /= m->val[i][0];
    }
}

/* ---- spars.cpp:238-316 PCGSolve; returns iterations (>=0) or -1 on a zero diagonal.
 * max_iter <= 0: run to convergence like the reference */
int orc_pcgsolve(orc_mat *m, int flag, int max_iter)
{
    const int n = m->n;
    for (int i = 0; i < n; i++) if (m->val[i][0] == 0) return -1;
    orc_multpc(m, m->b, m->Z);
    double res_o = orc_dot(m, m->Z, m->b);
    if (res_o == 0) return 0;
    if (flag == 0) for (int i = 0; i < n; i++) m->V[i] = 0;
    orc_multa(m, m->V, m->R);
    for (int i = 0; i < n; i++) m->R[i] = m->b[i] - m->R[i];
    orc_multpc(m, m->R, m->Z);
    for (int i = 0; i < n; i++) m->P[i] = m->Z[i];
    double res = orc_dot(m, m->Z, m->R), er;
    int it = 0;
    do
    {
        orc_multa(m, m->P, m->U);
        double del = res / orc_dot(m, m->P, m->U);
        for (int i = 0; i < n; i++) { m->V[i] += del * m->P[i]; m->R[i] -= del * m->U[i]; }
        orc_multpc(m, m->R, m->Z);
        double res_new = orc_dot(m, m->Z, m->R);
        double rho = res_new / res;
        res = res_new;
        for (int i = 0; i < n; i++) m->P[i] = m->Z[i] + rho * m->P[i];
        er = sqrt(res / res_o);
        it++;
        if (max_iter > 0 && it >= max_iter) break;
    } while (er > m->precision);
    return it;
}

/* export: upper-triangular CSR in the reference's list order */
long orc_nnz(const orc_mat *m)
{
    long t = 0;
    for (int i = 0; i < m->n; i++) t += m->len[i];
    return t;
}
void orc_export(const orc_mat *m, int *rowptr, int *col, double *val)
{
    long w = 0;
    rowptr[0] = 0;
    for (int i = 0; i < m->n; i++)
    {
        for (int k = 0; k < m->len[i]; k++) { col[w] = m->col[i][k]; val[w] = m->val[i][k]; w++; }
        rowptr[i + 1] = (int)w;
    }
}

/* ---- cuthill.cpp:40-85 SortElements: comb sort that also stops after the first pass without a swap */
void orc_sort_elements(int n, const int *tri3, int *order)
{
    int *score = (int *)malloc((n > 0 ? n : 1) * sizeof(int));
    for (int k = 0; k < n; k++) { score[k] = tri3[3 * k] + tri3[3 * k + 1] + tri3[3 * k + 2]; order[k] = k; }
    int gap = n, i;
    do
    {
        if (gap > 1)
        {
            gap = (gap * 10) / 13;
            if (gap == 10 || gap == 9) gap = 11;
        }
        i = 0;
        for (int j = 0; (j + gap) < n; j++)
        {
            if (score[j] > score[j + gap])
            {
                int k = j + gap, t = score[k];
                score[k] = score[j]; score[j] = t;
                t = order[k]; order[k] = order[j]; order[j] = t;
                i = 1;
            }
        }
    } while (gap > 1 && i > 0);
    free(score);
}

/* ---- cuthill.cpp:94-350 Cuthill: neighbour lists in .edge order, bubble-sorted by degree,
 * FIFO numbering from the first minimum-degree node, restart on new components. */
int orc_cuthill(int n, int n_edges, const int *edges, int *newnum, int *bandwidth)
{
    int *numcon = (int *)calloc(n, sizeof(int)), *nxt = (int *)calloc(n + 1, sizeof(int));
    int **ocon = (int **)calloc(n, sizeof(int *));
    for (int e = 0; e < n_edges; e++) { numcon[edges[2 * e]]++; numcon[edges[2 * e + 1]]++; }
    for (int i = 0; i < n; i++) ocon[i] = (int *)malloc((numcon[i] > 0 ? numcon[i] : 1) * sizeof(int));
    for (int e = 0; e < n_edges; e++)
    {
        int a = edges[2 * e], b = edges[2 * e + 1];
        ocon[a][nxt[a]++] = b;
        ocon[b][nxt[b]++] = a;
    }
    for (int a = 0; a < n; a++) /* cuthill.cpp:220-230 */
        for (int i = 1; i < numcon[a]; i++)
            for (int j = 1; j < numcon[a]; j++)
                if (numcon[ocon[a][j]] < numcon[ocon[a][j - 1]])
                {
                    int t = ocon[a][j]; ocon[a][j] = ocon[a][j - 1]; ocon[a][j - 1] = t;
                }
    long j = numcon[0];
    int n0 = 0;
    for (int i = 1; i < n; i++) /* cuthill.cpp:233-245 */
    {
        if (numcon[i] < j) { j = numcon[i]; n0 = i; }
        if (j == 2) break;
    }
    for (int i = 0; i < n; i++) { nxt[i] = -1; newnum[i] = -1; }
    nxt[n] = -1;
    newnum[n0] = 0;
    int cnt = 1;
    nxt[0] = n0;
    if (n > 1)
        do /* cuthill.cpp:254-304 */
        {
            for (int i = 0; i < numcon[n0]; i++)
                if (newnum[ocon[n0][i]] < 0)
                {
                    newnum[ocon[n0][i]] = cnt;
                    nxt[cnt] = ocon[n0][i];
                    cnt++;
                }
            if (nxt[newnum[n0] + 1] < 0)
            {
                if (cnt >= n) break;
                for (int i = 0; i < n; i++)
                    if (newnum[i] < 0) { j = numcon[i]; n0 = i; break; }
                for (int i = 0; i < n; i++)
                {
                    if (newnum[i] < 0 && numcon[i] < j) { j = numcon[i]; n0 = i; }
                    if (j == 2) break;
                }
                newnum[n0] = cnt;
                nxt[cnt] = n0;
                cnt++;
            }
            else
                n0 = nxt[newnum[n0] + 1];
        } while (cnt < n);
    int wide = 0;
    for (int a = 0; a < n; a++)
        for (int i = 0; i < numcon[a]; i++)
        {
            int d = abs(newnum[a] - newnum[ocon[a][i]]);
            if (d > wide) wide = d;
        }
    *bandwidth = wide + 1;
    for (int i = 0; i < n; i++) free(ocon[i]);
    free(ocon); free(numcon); free(nxt);
    return 0;
}

/* ---- CMaterialProp.cpp:1007-1057 GetBHProps (real part): nu = H/B, dnu/dB^2 */
void orc_bhprops(int npts, const double *Bd, const double *Hd, const double *Sd, double mu_x, double B,
                 double *v, double *dv)
{
    double b = fabs(B);
    if (npts == 0) { *v = mu_x; *dv = 0; return; }
    if (b == 0) { *v = Sd[0]; *dv = 0; return; }
    if (b > Bd[npts - 1])
    {
        double h = Hd[npts - 1] + Sd[npts - 1] * (b - Bd[npts - 1]), dh = Sd[npts - 1];
        *v = h / b;
        *dv = 0.5 * (dh / (b * b) - h / (b * b * b));
        return;
    }
    for (int i = 0; i < npts - 1; i++)
        if (b >= Bd[i] && b <= Bd[i + 1])
        {
            double l = Bd[i + 1] - Bd[i], z = (b - Bd[i]) / l, z2 = z * z;
            double h = (1. - 3. * z2 + 2. * z2 * z) * Hd[i] + z * (1. - 2. * z + z2) * l * Sd[i] +
                       z2 * (3. - 2. * z) * Hd[i + 1] + z2 * (z - 1.) * l * Sd[i + 1];
            double dh = 6. * z * (z - 1.) * Hd[i] / l + (1. - 4. * z + 3. * z * z) * Sd[i] +
                        6. * z * (1. - z) * Hd[i + 1] / l + z * (3. * z - 2.) * Sd[i + 1];
            *v = h / b;
            *dv = 0.5 * (dh / (b * b) - h / (b * b * b));
            return;
        }
}

/* ---- static2d.cpp:352-816: the planar element loop.
 * Per label (8 doubles): mu1_0, mu2_0 (Iter 0 values), Jsrc (= J_block + circuit term), H_c,
 * MagDir, LamFill, LamType, BHpoints ; bh_off[label] = offset into the flat B/H/slope tables.
 * bdry (4 doubles per boundary property): BdryFormat, c0, c1, unused.
 * mu12 (2 per element) carries meshele[i].mu1/mu2 between Newton steps. */
void orc_static2d_elements(orc_mat *L, int iter, int n_el, const int *el8, const double *xy,
                           const double *lab, const int *bh_off, const double *Bd, const double *Hd,
                           const double *Sd, const double *bdry, double *mu12)
{
    const double c = ORC_PI * 4.e-05;
    for (int i = 0; i < n_el; i++)
    {
        const int *E = el8 + 8 * i;
        const int n[3] = {E[0], E[1], E[2]};
        const int lbl = E[3];
        const double *T = lab + 8 * lbl;
        double Me[3][3], Mx[3][3], My[3][3], Mn[3][3], be[3], l[3], p[3], q[3];
        for (int j = 0; j < 3; j++)
        {
            for (int k = 0; k < 3; k++) Me[j][k] = Mx[j][k] = My[j][k] = Mn[j][k] = 0;
            be[j] = 0;
        }
        const double X[3] = {xy[2 * n[0]], xy[2 * n[1]], xy[2 * n[2]]};
        const double Y[3] = {xy[2 * n[0] + 1], xy[2 * n[1] + 1], xy[2 * n[2] + 1]};
        p[0] = Y[1] - Y[2]; p[1] = Y[2] - Y[0]; p[2] = Y[0] - Y[1];
        q[0] = X[2] - X[1]; q[1] = X[0] - X[2]; q[2] = X[1] - X[0];
        for (int j = 0; j < 3; j++)
        {
            int k = (j + 1) % 3;
            l[j] = sqrt(pow(X[k] - X[j], 2.) + pow(Y[k] - Y[j], 2.));
        }
        const double a = (p[0] * q[1] - p[1] * q[0]) / 2.;
        double K = (-1. / (4. * a));
        for (int j = 0; j < 3; j++)
            for (int k = j; k < 3; k++)
            {
                Mx[j][k] += K * p[j] * p[k];
                My[j][k] += K * q[j] * q[k];
                if (j != k) { Mx[k][j] += K * p[j] * p[k]; My[k][j] += K * q[j] * q[k]; }
            }
        /* mixed boundary condition on a side, static2d.cpp:460-480 */
        for (int j = 0; j < 3; j++)
            if (E[5 + j] >= 0 && bdry[4 * E[5 + j]] == 2)
            {
                int k = (j + 1) % 3;
                K = -0.0001 * c * bdry[4 * E[5 + j] + 1] * l[j] / 6.;
                Me[j][j] += K * 2.; Me[k][k] += K * 2.; Me[j][k] += K; Me[k][j] += K;
                K = (bdry[4 * E[5 + j] + 2] * l[j] / 2.) * 0.0001;
                be[j] += K; be[k] += K;
            }
        /* source current, static2d.cpp:483-507 */
        for (int j = 0; j < 3; j++) be[j] += -(T[2]) * a / 3.;
        /* permanent magnet, static2d.cpp:584-598 */
        for (int j = 0; j < 3; j++)
        {
            int k = (j + 1) % 3;
            K = 0.0001 * T[3] * (cos(T[4] * ORC_PI / 180.) * (X[k] - X[j]) + sin(T[4] * ORC_PI / 180.) * (Y[k] - Y[j])) / 2.;
            be[j] += K; be[k] += K;
        }
        const int lamtype = (int)T[6], bhpts = (int)T[7];
        const double t = T[5];
        double *mu1 = mu12 + 2 * i, *mu2 = mu12 + 2 * i + 1;
        if (iter == 0) { *mu1 = T[0]; *mu2 = T[1]; }
        else if (bhpts > 0)
        {
            const double *B0 = Bd + bh_off[lbl], *H0 = Hd + bh_off[lbl], *S0 = Sd + bh_off[lbl];
            const double V0[3] = {L->V[n[0]], L->V[n[1]], L->V[n[2]]};
            double B1 = 0, B2 = 0, B, mu, dv = 0, v[3], u[3];
            if (lamtype == 0 && *mu1 == *mu2) /* static2d.cpp:687-717 */
            {
                for (int j = 0; j < 3; j++) { B1 += V0[j] * q[j]; B2 += V0[j] * p[j]; }
                B = c * sqrt(B1 * B1 + B2 * B2) / (0.02 * a);
                orc_bhprops(bhpts, B0, H0, S0, T[0], B, &mu, &dv);
                mu = 1. / (ORC_MUO * mu);
                *mu1 = mu; *mu2 = mu;
                for (int j = 0; j < 3; j++)
                {
                    v[j] = 0;
                    for (int w = 0; w < 3; w++) v[j] += (Mx[j][w] + My[j][w]) * V0[w];
                }
                K = -200. * c * c * c * dv / a;
                for (int j = 0; j < 3; j++)
                    for (int w = 0; w < 3; w++) Mn[j][w] = K * v[j] * v[w];
            }
            if (lamtype == 1) /* static2d.cpp:719-757 */
            {
                for (int j = 0; j < 3; j++) { B1 += V0[j] * q[j]; B2 += V0[j] * p[j] / t; }
                B = c * sqrt(B1 * B1 + B2 * B2) / (0.02 * a);
                orc_bhprops(bhpts, B0, H0, S0, T[0], B, &mu, &dv);
                mu = 1. / (ORC_MUO * mu);
                *mu1 = mu * t; *mu2 = mu / (t + mu * (1. - t));
                for (int j = 0; j < 3; j++)
                {
                    v[j] = 0; u[j] = 0;
                    for (int w = 0; w < 3; w++)
                    {
                        v[j] += (My[j][w] / t + Mx[j][w]) * V0[w];
                        u[j] += (My[j][w] / t + t * Mx[j][w]) * V0[w];
                    }
                }
                K = -100. * c * c * c * dv / (a);
                for (int j = 0; j < 3; j++)
                    for (int w = 0; w < 3; w++) Mn[j][w] = K * (v[j] * u[w] + v[w] * u[j]);
            }
            if (lamtype == 2) /* static2d.cpp:758-796 */
            {
                for (int j = 0; j < 3; j++) { B1 += (V0[j] * q[j]) / t; B2 += V0[j] * p[j]; }
                B = c * sqrt(B1 * B1 + B2 * B2) / (0.02 * a);
                orc_bhprops(bhpts, B0, H0, S0, T[0], B, &mu, &dv);
                mu = 1. / (ORC_MUO * mu);
                *mu2 = mu * t; *mu1 = mu / (t + mu * (1. - t));
                for (int j = 0; j < 3; j++)
                {
                    v[j] = 0; u[j] = 0;
                    for (int w = 0; w < 3; w++)
                    {
                        v[j] += (Mx[j][w] / t + My[j][w]) * V0[w];
                        u[j] += (Mx[j][w] / t + t * My[j][w]) * V0[w];
                    }
                }
                K = -100. * c * c * c * dv / (a);
                for (int j = 0; j < 3; j++)
                    for (int w = 0; w < 3; w++) Mn[j][w] = K * (v[j] * u[w] + v[w] * u[j]);
            }
        }
        /* static2d.cpp:799-815 */
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++)
            {
                Me[j][k] += (Mx[j][k] / (*mu2) + My[j][k] / (*mu1) + Mn[j][k]);
                be[j] += Mn[j][k] * L->V[n[k]];
            }
        for (int j = 0; j < 3; j++)
        {
            for (int k = j; k < 3; k++) orc_addto(L, -Me[j][k], n[j], n[k]);
            L->b[n[j]] -= be[j];
        }
    }
}
