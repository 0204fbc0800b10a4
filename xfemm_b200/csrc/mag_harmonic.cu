// xfemm_cmag_* — complex (time-harmonic) element assembly on the GPU.
// Replaces the element loops of FSolver::Harmonic2D (cfemm/fsolver/harmonic2d.cpp:386-725) and
// FSolver::HarmonicAxisymmetric (cfemm/fsolver/harmonicaxi.cpp:216-646): complex permeability,
// eddy-current mass term, mixed / small-skin-depth boundary sides, circuits with unknown voltage
// gradient as extra rows, successive-approximation update of the permeability from the B-H curve
// ([ACSolver]=0).  The harmonic formulation stores +Me and b += be (opposite sign to the static one).
#include "clinprob.cuh"
#include <vector>
#include <algorithm>
#include <cmath>

#define XB_PI 3.141592653589793238462643383
#define XB_MUO 1.2566370614359173e-6

struct xfemm_cmag
{
    xfemm_clinprob *lin = nullptr;
    int axisymmetric = 0;
    int n_nodes = 0, n_elems = 0, n_labels = 0, n_bh = 0;
    double omega = 0;
    xb::DevBuf<double> xy;
    xb::DevBuf<int> tri, emap;
    xb::DevBuf<xfemm_cmag_label> labels;
    xb::DevBuf<double> bhB, bhHr, bhHi, bhSr, bhSi;
    xb::DevBuf<double> mu; // n_elems*4 (mu1 re,im, mu2 re,im)
    int n_side = 0;
    xb::DevBuf<int> side_es;
    xb::DevBuf<double> side_k; // per side: type, c0re, c0im, c1re, c1im, Kssd re, Kssd im
    int n_age_cells = 0;
    xb::DevBuf<int> age_nodes, age_mid;
    xb::DevBuf<double> age_w, age_mats;
    int n_psrc = 0;
    xb::DevBuf<int> psrc_node;
    xb::DevBuf<double> psrc_val; // complex
    int n_circ_rhs = 0;
    xb::DevBuf<int> circ_row;
    xb::DevBuf<double> circ_val; // complex
    int n_fix = 0;
    xb::DevBuf<int> fix_rows;
    double ext_zo = 0, ext_ro = 0, ext_ri = 0;
    xb::DevBuf<int> err_flag;
    int newton = 0; // [ACSolver]=1: passes with iter>0 build the Newton matrices
};

namespace xb {

__device__ __forceinline__ cd hmul(cd a, cd b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd hdiv(cd a, cd b)
{
    double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ cd hadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd hsc(cd a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ void hatomic(cd *p, cd v)
{
    atomicAdd(&p->x, v.x);
    atomicAdd(&p->y, v.y);
}
__device__ __forceinline__ int hfind(const int *__restrict__ rowptr, const int *__restrict__ col, int row, int c)
{
    int lo = rowptr[row], hi = rowptr[row + 1] - 1;
    while (lo <= hi)
    {
        int mid = (lo + hi) >> 1, v = col[mid];
        if (v == c) return mid;
        if (v < c) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

__global__ void kh_build_emap(int ne, const int4 *__restrict__ tri, const int *__restrict__ rowptr,
                              const int *__restrict__ col, int *__restrict__ emap, int *err)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int4 t = tri[e];
    int n[3] = {t.x, t.y, t.z};
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++)
        {
            int s = hfind(rowptr, col, n[j], n[k]);
            if (s < 0) atomicExch(err, 1);
            emap[(size_t)(3 * j + k) * ne + e] = s;
        }
}

// Re(H(B))/B ("Get_v": CMaterialProp.cpp:899-904 calls the real-valued GetH overload) and the
// complex dH/dB (CMaterialProp.cpp:461-486) from the cubic-Hermite tables.
__device__ __forceinline__ void bh_harmonic(const double *__restrict__ Bd, const double *__restrict__ Hr,
                                            const double *__restrict__ Hi, const double *__restrict__ Sr,
                                            const double *__restrict__ Si, int npts, double B, cd &v, cd &dh)
{
    const double b = fabs(B);
    if (b > Bd[npts - 1])
    {
        double h = Hr[npts - 1] + Sr[npts - 1] * (b - Bd[npts - 1]);
        v = make_double2(h / b, 0.);
        dh = make_double2(Sr[npts - 1], Si[npts - 1]);
        return;
    }
    v = make_double2(0., 0.);
    dh = make_double2(0., 0.);
    for (int i = 0; i < npts - 1; i++)
        if (b >= Bd[i] && b <= Bd[i + 1])
        {
            const double l = Bd[i + 1] - Bd[i], z = (b - Bd[i]) / l, z2 = z * z;
            const double c0 = (1. - 3. * z2 + 2. * z2 * z), c1 = z * (1. - 2. * z + z2) * l, c2 = z2 * (3. - 2. * z), c3 = z2 * (z - 1.) * l;
            const double h = c0 * Hr[i] + c1 * Sr[i] + c2 * Hr[i + 1] + c3 * Sr[i + 1];
            const double d0 = 6. * z * (z - 1.) / l, d1 = (1. - 4. * z + 3. * z * z), d2 = 6. * z * (1. - z) / l, d3 = z * (3. * z - 2.);
            dh = make_double2(d0 * Hr[i] + d1 * Sr[i] + d2 * Hr[i + 1] + d3 * Sr[i + 1],
                              d0 * Hi[i] + d1 * Si[i] + d2 * Hi[i + 1] + d3 * Si[i + 1]);
            v = (b == 0) ? make_double2(Sr[0], Si[0]) : make_double2(h / b, 0.);
            return;
        }
}

// complex nu = H(B)/B and d nu/d(B^2) = (H'/B^2 - H/B^3)/2 (GetBHProps, CMaterialProp.cpp:1008-1057) for the Newton AC solver
__device__ __forceinline__ void bh_props_complex(const double *__restrict__ Bd, const double *__restrict__ Hr,
                                                 const double *__restrict__ Hi, const double *__restrict__ Sr,
                                                 const double *__restrict__ Si, int npts, double B, cd &v, cd &dv)
{
    const double b = fabs(B);
    if (b == 0) { v = make_double2(Sr[0], Si[0]); dv = make_double2(0., 0.); return; }
    cd h, dh;
    if (b > Bd[npts - 1])
    {
        const double d = b - Bd[npts - 1];
        h = make_double2(Hr[npts - 1] + Sr[npts - 1] * d, Hi[npts - 1] + Si[npts - 1] * d);
        dh = make_double2(Sr[npts - 1], Si[npts - 1]);
    }
    else
    {
        int i = 0;
        while (i < npts - 2 && !(b >= Bd[i] && b <= Bd[i + 1])) i++;
        const double l = Bd[i + 1] - Bd[i], z = (b - Bd[i]) / l, z2 = z * z;
        const double c0 = (1. - 3. * z2 + 2. * z2 * z), c1 = z * (1. - 2. * z + z2) * l, c2 = z2 * (3. - 2. * z), c3 = z2 * (z - 1.) * l;
        const double d0 = 6. * z * (z - 1.) / l, d1 = (1. - 4. * z + 3. * z * z), d2 = 6. * z * (1. - z) / l, d3 = z * (3. * z - 2.);
        h = make_double2(c0 * Hr[i] + c1 * Sr[i] + c2 * Hr[i + 1] + c3 * Sr[i + 1], c0 * Hi[i] + c1 * Si[i] + c2 * Hi[i + 1] + c3 * Si[i + 1]);
        dh = make_double2(d0 * Hr[i] + d1 * Sr[i] + d2 * Hr[i + 1] + d3 * Sr[i + 1], d0 * Hi[i] + d1 * Si[i] + d2 * Hi[i + 1] + d3 * Si[i + 1]);
    }
    v = hsc(h, 1. / b);
    dv = make_double2(0.5 * (dh.x / (b * b) - h.x / (b * b * b)), 0.5 * (dh.y / (b * b) - h.y / (b * b * b)));
}

template <bool AXI>
__global__ void __launch_bounds__(256)
kh_assemble(int ne, int iter, int n_nodes, double omega, const int4 *__restrict__ tri, const double2 *__restrict__ xy,
            const int *__restrict__ emap, const xfemm_cmag_label *__restrict__ labels, const double *__restrict__ bhB,
            const double *__restrict__ bhHr, const double *__restrict__ bhHi, const double *__restrict__ bhSr,
            const double *__restrict__ bhSi, const cd *__restrict__ V, const int *__restrict__ rowptr,
            const int *__restrict__ col, cd *vals, cd *b, double4 *mu_out, double ext_zo, double ext_ro, double ext_ri,
            int *err, int newton, cd *vals_h, cd *vals_s, cd *vals_a)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const double c = XB_PI * 4.e-05;
    const int4 t4 = tri[e];
    const int n[3] = {t4.x, t4.y, t4.z};
    const double2 P0 = xy[n[0]], P1 = xy[n[1]], P2 = xy[n[2]];
    const double X[3] = {P0.x, P1.x, P2.x}, Y[3] = {P0.y, P1.y, P2.y};
    const double p[3] = {P1.y - P2.y, P2.y - P0.y, P0.y - P1.y};
    const double q[3] = {P2.x - P1.x, P0.x - P2.x, P1.x - P0.x};
    const double a = (p[0] * q[1] - p[1] * q[0]) / 2.;
    const xfemm_cmag_label L = labels[t4.w];
    double Mx[3][3], My[3][3];
    double R = 0., vol = 0.;
    if (!AXI)
    {
        const double K = -1. / (4. * a);
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) { Mx[j][k] = K * p[j] * p[k]; My[j][k] = K * q[j] * q[k]; }
    }
    else
    {
        // harmonicaxi.cpp:262-345 (same r-weighted matrices as staticaxi.cpp:189-305, Mxy never filled)
        const double rn[3] = {X[0], X[1], X[2]};
        const double g[3] = {(X[2] + X[1]) / 2., (X[0] + X[2]) / 2., (X[1] + X[0]) / 2.};
        R = (X[0] + X[1] + X[2]) / 3.;
        double a_hat = 0;
        for (int j = 0; j < 3; j++) a_hat += (rn[j] * rn[j] * p[j] / (4. * R));
        vol = 2. * R * a_hat;
        int flag = 0;
        for (int j = 0; j < 3; j++) if (rn[j] < 1.e-06) flag++;
        double R_hat = 0.;
        if (flag == 2) R_hat = R;
        else if (flag == 1)
        {
            if (rn[0] < 1.e-06) R_hat = (fabs(rn[1] - rn[2]) < 1.e-06) ? rn[2] / 2. : (rn[1] - rn[2]) / (2. * log(rn[1]) - 2. * log(rn[2]));
            if (rn[1] < 1.e-06) R_hat = (fabs(rn[2] - rn[0]) < 1.e-06) ? rn[0] / 2. : (rn[2] - rn[0]) / (2. * log(rn[2]) - 2. * log(rn[0]));
            if (rn[2] < 1.e-06) R_hat = (fabs(rn[0] - rn[1]) < 1.e-06) ? rn[1] / 2. : (rn[0] - rn[1]) / (2. * log(rn[0]) - 2. * log(rn[1]));
        }
        else
        {
            if (fabs(q[0]) < 1.e-06) R_hat = (q[1] * q[1]) / (2. * (-q[1] + rn[0] * log(rn[0] / rn[2])));
            else if (fabs(q[1]) < 1.e-06) R_hat = (q[2] * q[2]) / (2. * (-q[2] + rn[1] * log(rn[1] / rn[0])));
            else if (fabs(q[2]) < 1.e-06) R_hat = (q[0] * q[0]) / (2. * (-q[0] + rn[2] * log(rn[2] / rn[1])));
            else R_hat = -(q[0] * q[1] * q[2]) / (2. * (q[0] * rn[0] * log(rn[0]) + q[1] * rn[1] * log(rn[1]) + q[2] * rn[2] * log(rn[2])));
        }
        const double Kx = (-1. / (2. * a_hat * R)), Ky = (-1. / (2. * a_hat * R_hat));
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++)
            {
                Mx[j][k] = Kx * p[j] * rn[j] * p[k] * rn[k];
                My[j][k] = Ky * (q[j] * rn[j]) * (q[k] * rn[k]) * (g[j] / R) * (g[k] / R);
            }
        for (int j = 0; j < 3; j++)
            if (rn[j] < 1.e-06) Mx[j][j] += Mx[0][0] + Mx[1][1] + Mx[2][2];
    }
    cd Me[3][3], be[3];
    for (int j = 0; j < 3; j++)
    {
        for (int k = 0; k < 3; k++) Me[j][k] = make_double2(0., 0.);
        be[j] = make_double2(0., 0.);
    }
    // eddy currents: K = -j w sigma a c /12, off-diagonal K, diagonal 2K (harmonic2d.cpp:459-479);
    // axisymmetric: every entry gets (-j R a w sigma c /6)*4/3 (harmonicaxi.cpp:347-359)
    if (L.sigma_eddy != 0.)
    {
        if (!AXI)
        {
            const cd K = make_double2(0., -a * omega * L.sigma_eddy * c / 12.);
            for (int j = 0; j < 3; j++)
                for (int k = j; k < 3; k++) { Me[j][k] = hadd(Me[j][k], K); Me[k][j] = hadd(Me[k][j], K); }
        }
        else
        {
            const cd K = make_double2(0., -R * a * omega * L.sigma_eddy * c / 6. * 4. / 3.);
            for (int j = 0; j < 3; j++)
                for (int k = 0; k < 3; k++) Me[j][k] = hadd(Me[j][k], K);
        }
    }
    // source current density (harmonic2d.cpp:518-537 / harmonicaxi.cpp:404-425)
    cd Jel = make_double2(L.J[0], L.J[1]);
    if (AXI) Jel = hadd(Jel, make_double2(L.dVs[0] / R, L.dVs[1] / R));
    const cd Kj = AXI ? hsc(Jel, -2. * R * a / 3.) : hsc(Jel, -a / 3.);
    for (int j = 0; j < 3; j++) be[j] = hadd(be[j], Kj);
    if (L.circuit >= 0)
    {
        // Case 2: voltage gradient of the circuit is an unknown (row/column NumNodes+k)
        const int crow = n_nodes + L.circuit;
        const cd Kb = AXI ? hsc(Kj, 3. / R) : hsc(Kj, 3.);
        hatomic(b + crow, Kb);
        const cd K = AXI ? make_double2(0., -2. * a * omega * L.sigma_circ * c) : make_double2(0., -a * omega * L.sigma_circ * c);
        for (int j = 0; j < 3; j++)
        {
            const int s1 = hfind(rowptr, col, n[j], crow), s2 = hfind(rowptr, col, crow, n[j]);
            if (s1 < 0 || s2 < 0) { atomicExch(err, 3); return; }
            hatomic(vals + s1, hsc(K, 1. / 3.));
            hatomic(vals + s2, hsc(K, 1. / 3.));
        }
        const int sd = hfind(rowptr, col, crow, crow);
        if (sd < 0) { atomicExch(err, 3); return; }
        hatomic(vals + sd, AXI ? hsc(K, 1. / R) : K);
    }
    // permeability
    cd mu1 = make_double2(L.mu1[0], L.mu1[1]), mu2 = make_double2(L.mu2[0], L.mu2[1]);
    if (AXI && L.is_external)
    {
        const double Z = (Y[0] + Y[1] + Y[2]) / 3. - ext_zo;
        const double kludge = (R * R + Z * Z) * ext_ri / (ext_ro * ext_ro * ext_ro);
        mu1 = hsc(mu1, 1. / kludge);
        mu2 = hsc(mu2, 1. / kludge);
    }
    cd Kn = make_double2(0., 0.);
    cd Vn[3] = {make_double2(0., 0.), make_double2(0., 0.), make_double2(0., 0.)};
    if (iter > 0 && L.bh_points > 0 && L.lamtype == 0)
    {
        Vn[0] = V[n[0]]; Vn[1] = V[n[1]]; Vn[2] = V[n[2]];
        double B;
        if (!AXI)
        {
            // harmonic2d.cpp:602-607
            cd B1 = make_double2(0., 0.), B2 = make_double2(0., 0.);
            for (int j = 0; j < 3; j++) { B1 = hadd(B1, hsc(Vn[j], q[j])); B2 = hadd(B2, hsc(Vn[j], p[j])); }
            B = c * sqrt((B1.x * B1.x + B1.y * B1.y) + (B2.x * B2.x + B2.y * B2.y)) / (0.02 * a);
        }
        else
        {
            // harmonicaxi.cpp: B from the element energy, conj(V).(Mx+My)V
            cd dv = make_double2(0., 0.);
            for (int j = 0; j < 3; j++)
            {
                cd vj = make_double2(0., 0.);
                for (int w = 0; w < 3; w++) vj = hadd(vj, hsc(Vn[w], Mx[j][w] + My[j][w]));
                dv = hadd(dv, hmul(make_double2(Vn[j].x, -Vn[j].y), vj));
            }
            dv = hsc(dv, 10000. * c * c / vol);
            B = sqrt(sqrt(dv.x * dv.x + dv.y * dv.y));
        }
        cd v, dh;
        const int o = L.bh_offset;
        const cd one = make_double2(1., 0.);
        if (newton)
        {
            // Newton AC (harmonic2d.cpp:611-636, harmonicaxi.cpp:520-548): mu from nu(B); with v = (Mx+My) V and
            // K = -200 c^3 (d nu/dB^2) / area:  Mn = K Re(v v^H) goes into the complex-symmetric matrix,
            // Mnh = Re(K) v v^H /2 - Re(Mn) (Hermitian), Mna = j Im(K) v v^H /2 - j Im(Mn) (anti-Hermitian) and
            // Mns = K v v^T /2 (acts on conj(V)) into the side matrices; all four also feed the right-hand side.
            cd dv;
            bh_props_complex(bhB + o, bhHr + o, bhHi + o, bhSr + o, bhSi + o, L.bh_points, B, v, dv);
            const cd mu = hdiv(one, hsc(v, XB_MUO));
            mu1 = mu; mu2 = mu;
            cd vv[3];
            for (int j = 0; j < 3; j++)
            {
                vv[j] = make_double2(0., 0.);
                for (int w = 0; w < 3; w++) vv[j] = hadd(vv[j], hsc(Vn[w], Mx[j][w] + My[j][w]));
            }
            const cd K = hsc(dv, -200. * c * c * c / (AXI ? vol : a));
            if (mu_out) mu_out[e] = make_double4(mu1.x, mu1.y, mu2.x, mu2.y);
            const cd i1 = hdiv(one, mu1), i2 = hdiv(one, mu2);
            for (int j = 0; j < 3; j++)
            {
                cd bj = be[j];
                for (int k = 0; k < 3; k++)
                {
                    const cd vvh = hmul(vv[j], make_double2(vv[k].x, -vv[k].y)); // v_j conj(v_k)
                    const cd mn = hsc(K, vvh.x);
                    const cd mnh = make_double2(0.5 * K.x * vvh.x - mn.x, 0.5 * K.x * vvh.y);
                    const cd mna = make_double2(-0.5 * K.y * vvh.y, 0.5 * K.y * vvh.x - mn.y);
                    const cd mns = hsc(hmul(K, hmul(vv[j], vv[k])), 0.5);
                    const cd me = hadd(hadd(Me[j][k], mn), hadd(hsc(i2, Mx[j][k]), hsc(i1, My[j][k])));
                    bj = hadd(bj, hmul(hadd(hadd(mnh, mna), mn), Vn[k]));
                    bj = hadd(bj, hmul(mns, make_double2(Vn[k].x, -Vn[k].y)));
                    const int slot = emap[(size_t)(3 * j + k) * ne + e];
                    hatomic(vals + slot, me);
                    hatomic(vals_h + slot, mnh);
                    hatomic(vals_s + slot, mns);
                    hatomic(vals_a + slot, mna);
                }
                hatomic(b + n[j], bj);
            }
            return;
        }
        bh_harmonic(bhB + o, bhHr + o, bhHi + o, bhSr + o, bhSi + o, L.bh_points, B, v, dh);
        // successive approximation with the averaged permeability (harmonic2d.cpp:641-657)
        const cd murel = hdiv(one, hsc(v, XB_MUO)), muinc = hdiv(one, hsc(dh, XB_MUO));
        const cd K = hdiv(hsc(hmul(murel, muinc), 2.), hadd(murel, muinc));
        mu1 = K; mu2 = K;
        const cd a1 = hdiv(one, murel), a2 = hdiv(one, K);
        Kn = make_double2(-(a1.x - a2.x), -(a1.y - a2.y));
    }
    if (mu_out) mu_out[e] = make_double4(mu1.x, mu1.y, mu2.x, mu2.y);
    const cd i1 = hdiv(make_double2(1., 0.), mu1), i2 = hdiv(make_double2(1., 0.), mu2);
    for (int j = 0; j < 3; j++)
    {
        cd bj = be[j];
        for (int k = 0; k < 3; k++)
        {
            cd me = hadd(Me[j][k], hadd(hsc(i2, Mx[j][k]), hsc(i1, My[j][k])));
            bj = hadd(bj, hmul(hsc(Kn, Mx[j][k] + My[j][k]), Vn[k]));
            hatomic(vals + emap[(size_t)(3 * j + k) * ne + e], me);
        }
        hatomic(b + n[j], bj);
    }
}

// boundary sides: type 2 mixed (complex c0, c1) and type 1 small skin depth
// (harmonic2d.cpp:481-516 ; harmonicaxi.cpp:361-402 carries the extra 2r)
__global__ void kh_sides(int ns, int ne, const int *__restrict__ es, const double *__restrict__ sk, const int4 *__restrict__ tri,
                         const double2 *__restrict__ xy, const int *__restrict__ emap, cd *vals, cd *b, int axi)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= ns) return;
    const double c = XB_PI * 4.e-05;
    const int e = es[2 * m], j = es[2 * m + 1], k = (j + 1) % 3;
    const int4 t = tri[e];
    const int n[3] = {t.x, t.y, t.z};
    const double2 A = xy[n[j]], B = xy[n[k]];
    const double l = sqrt((B.x - A.x) * (B.x - A.x) + (B.y - A.y) * (B.y - A.y));
    const double rf = axi ? 2. * ((A.x + B.x) / 2.) : 1.;
    const double *S = sk + 7 * m;
    cd K;
    if (S[0] == 2.)
    {
        K = hsc(make_double2(S[1], S[2]), -0.0001 * c * rf * l / 6.);
        const cd K2 = hsc(make_double2(S[3], S[4]), (l / 2.) * rf * 0.0001);
        hatomic(b + n[j], K2);
        hatomic(b + n[k], K2);
    }
    else
        K = hsc(make_double2(S[5], S[6]), rf * l / 6.);
    hatomic(vals + emap[(size_t)(3 * j + j) * ne + e], hsc(K, 2.));
    hatomic(vals + emap[(size_t)(3 * k + k) * ne + e], hsc(K, 2.));
    hatomic(vals + emap[(size_t)(3 * j + k) * ne + e], K);
    hatomic(vals + emap[(size_t)(3 * k + j) * ne + e], K);
}

__global__ void kh_age_scatter(int ncells, const int *__restrict__ nodes, const double *__restrict__ w,
                               const int *__restrict__ mid, const double *__restrict__ mats,
                               const int *__restrict__ rowptr, const int *__restrict__ col, cd *vals, int *err)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * 55) return;
    const int cell = t / 55;
    int q = t % 55, ii = 0;
    while (q >= 10 - ii) { q -= 10 - ii; ii++; }
    const int jj = ii + q;
    const int a = nodes[cell * 10 + ii], bnode = nodes[cell * 10 + jj];
    // harmonic2d.cpp:382 — opposite sign to the static formulation
    const double v = -(mats[mid[cell] * 100 + ii * 10 + jj] * w[cell * 10 + ii] * w[cell * 10 + jj]);
    const int s1 = hfind(rowptr, col, a, bnode);
    if (s1 < 0) { atomicExch(err, 2); return; }
    atomicAdd(&vals[s1].x, v);
    if (a != bnode)
    {
        const int s2 = hfind(rowptr, col, bnode, a);
        if (s2 < 0) { atomicExch(err, 2); return; }
        atomicAdd(&vals[s2].x, v);
    }
}
__global__ void kh_add_rhs(int np, const int *__restrict__ node, const cd *__restrict__ val, cd *b)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) hatomic(b + node[i], val[i]);
}
// L.Put(L.Get(0,0), row, row) for circuits with a priori known current/voltage (harmonic2d.cpp:803-807)
__global__ void kh_fix_diag(int nfix, const int *__restrict__ rows, const int *__restrict__ rowptr,
                            const int *__restrict__ col, cd *vals)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nfix) return;
    const int s0 = hfind(rowptr, col, 0, 0), s = hfind(rowptr, col, rows[i], rows[i]);
    if (s0 >= 0 && s >= 0) vals[s] = vals[s0];
}
// planar: b = V c, circuit rows j c w V ; axisymmetric: b = V c 2 pi r 0.01, circuit rows j w c 0.01 V
__global__ void kh_scale_to_A(int n, int n_nodes, const cd *__restrict__ V, const double2 *__restrict__ xy, int axi,
                              double omega, cd *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c = XB_PI * 4.e-05;
    cd v = V[i];
    if (i < n_nodes) out[i] = axi ? hsc(v, c * 2. * XB_PI * xy[i].x * 0.01) : hsc(v, c);
    else
    {
        const double f = axi ? omega * c * 0.01 : c * omega;
        out[i] = make_double2(-f * v.y, f * v.x);
    }
}

} // namespace xb

using namespace xb;

extern "C" {

int xfemm_cmag_set_labels(xfemm_cmag *m, int n_labels, const xfemm_cmag_label *labels)
{
    XB_ARG(m && labels && n_labels > 0, "labels");
    XB_CUDA(m->labels.upload(labels, n_labels, m->lin->stream));
    XB_CUDA(cudaStreamSynchronize(m->lin->stream));
    m->n_labels = n_labels;
    return XFEMM_OK;
}

int xfemm_cmag_create(const xfemm_cmag_desc *d, xfemm_clinprob *lin, xfemm_cmag **out)
{
    XB_ARG(d && lin && out, "null pointer");
    XB_ARG(d->n_nodes > 0 && d->n_elems > 0 && d->xy && d->tri && d->labels && d->n_labels > 0, "mesh arrays");
    XB_ARG(d->n_nodes <= lin->n, "linear problem smaller than the mesh");
    for (int e = 0; e < d->n_elems; e++)
    {
        const int *t = d->tri + 4 * (size_t)e;
        if (t[0] < 0 || t[0] >= d->n_nodes || t[1] < 0 || t[1] >= d->n_nodes || t[2] < 0 || t[2] >= d->n_nodes || t[3] < 0 || t[3] >= d->n_labels)
        {
            set_error("element %d has node or label index out of range", e);
            return XFEMM_ERR_ARG;
        }
    }
    xfemm_cmag *m = new xfemm_cmag;
    m->lin = lin;
    m->axisymmetric = d->axisymmetric;
    m->n_nodes = d->n_nodes; m->n_elems = d->n_elems; m->n_bh = d->n_bh; m->omega = d->omega;
    m->ext_zo = d->ext_zo; m->ext_ro = d->ext_ro; m->ext_ri = d->ext_ri;
    cudaStream_t st = lin->stream;
    auto body = [&]() -> int {
        XB_CUDA(m->xy.upload(d->xy, (size_t)d->n_nodes * 2, st));
        XB_CUDA(m->tri.upload(d->tri, (size_t)d->n_elems * 4, st));
        if (d->n_bh > 0)
        {
            XB_ARG(d->bh_B && d->bh_H && d->bh_H_im && d->bh_slope && d->bh_slope_im, "B-H tables");
            XB_CUDA(m->bhB.upload(d->bh_B, d->n_bh, st));
            XB_CUDA(m->bhHr.upload(d->bh_H, d->n_bh, st));
            XB_CUDA(m->bhHi.upload(d->bh_H_im, d->n_bh, st));
            XB_CUDA(m->bhSr.upload(d->bh_slope, d->n_bh, st));
            XB_CUDA(m->bhSi.upload(d->bh_slope_im, d->n_bh, st));
        }
        m->n_side = d->n_sides;
        if (d->n_sides > 0)
        {
            XB_ARG(d->side_elem_side && d->side_coeffs, "boundary side arrays");
            XB_CUDA(m->side_es.upload(d->side_elem_side, (size_t)d->n_sides * 2, st));
            XB_CUDA(m->side_k.upload(d->side_coeffs, (size_t)d->n_sides * 7, st));
        }
        m->n_age_cells = d->n_age_cells;
        if (d->n_age_cells > 0)
        {
            XB_ARG(d->age_nodes && d->age_weights && d->age_matrix_id && d->age_matrices && d->n_age_matrices > 0, "AGE tables");
            XB_CUDA(m->age_nodes.upload(d->age_nodes, (size_t)d->n_age_cells * 10, st));
            XB_CUDA(m->age_w.upload(d->age_weights, (size_t)d->n_age_cells * 10, st));
            XB_CUDA(m->age_mid.upload(d->age_matrix_id, d->n_age_cells, st));
            XB_CUDA(m->age_mats.upload(d->age_matrices, (size_t)d->n_age_matrices * 100, st));
        }
        m->n_psrc = d->n_rhs_add;
        if (d->n_rhs_add > 0)
        {
            XB_ARG(d->rhs_add_row && d->rhs_add_val, "rhs arrays");
            for (int i = 0; i < d->n_rhs_add; i++) XB_ARG(d->rhs_add_row[i] >= 0 && d->rhs_add_row[i] < lin->n, "rhs row out of range");
            XB_CUDA(m->psrc_node.upload(d->rhs_add_row, d->n_rhs_add, st));
            XB_CUDA(m->psrc_val.upload(d->rhs_add_val, (size_t)d->n_rhs_add * 2, st));
        }
        m->n_fix = d->n_fix_diag;
        if (d->n_fix_diag > 0)
        {
            XB_ARG(d->fix_diag_rows, "fix rows");
            XB_CUDA(m->fix_rows.upload(d->fix_diag_rows, d->n_fix_diag, st));
        }
        XB_CUDA(m->mu.alloc((size_t)d->n_elems * 4));
        XB_CUDA(m->emap.alloc((size_t)d->n_elems * 9));
        XB_CUDA(m->err_flag.alloc(1));
        XB_CUDA(m->err_flag.zero(st));
        ++g_launch, kh_build_emap<<<(d->n_elems + 255) / 256, 256, 0, st>>>(d->n_elems, (const int4 *)m->tri.p, lin->rowptr.p,
                                                                              lin->col.p, m->emap.p, m->err_flag.p);
        int flag = 0;
        XB_CUDA(cudaMemcpyAsync(&flag, m->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        XB_CUDA(cudaStreamSynchronize(st));
        XB_CUDA(cudaGetLastError());
        if (flag) { set_error("the linear problem's pattern does not contain the mesh graph"); return XFEMM_ERR_ARG; }
        return XFEMM_OK;
    };
    int rc = body();
    if (rc == XFEMM_OK) rc = xfemm_cmag_set_labels(m, d->n_labels, d->labels);
    if (rc != XFEMM_OK) { delete m; return rc; }
    *out = m;
    return XFEMM_OK;
}

int xfemm_cmag_set_newton(xfemm_cmag *m, int on)
{
    XB_ARG(m, "null handle");
    m->newton = on ? 1 : 0;
    return XFEMM_OK;
}

int xfemm_cmag_destroy(xfemm_cmag *m)
{
    delete m;
    return XFEMM_OK;
}

int xfemm_cmag_assemble(xfemm_cmag *m, int iter, double *elapsed_ms)
{
    XB_ARG(m && m->lin, "null handle");
    xfemm_clinprob *L = m->lin;
    cudaStream_t st = L->stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (elapsed_ms) { XB_CUDA(cudaEventCreate(&e0)); XB_CUDA(cudaEventCreate(&e1)); XB_CUDA(cudaEventRecord(e0, st)); }
    XB_CUDA(L->vals.zero(st));
    XB_CUDA(L->b.zero(st));
    XB_CUDA(m->err_flag.zero(st));
    const int newton = (m->newton && iter > 0) ? 1 : 0;
    if (m->newton)
    {
        XB_ARG(L->newton, "xfemm_cmag_set_newton needs xfemm_clinprob_enable_newton on the linear problem");
        for (auto *v : {&L->vals_h, &L->vals_s, &L->vals_a}) XB_CUDA(v->zero(st));
    }
    if (m->n_age_cells > 0 && !m->axisymmetric)
    {
        int nt = m->n_age_cells * 55;
        ++g_launch, kh_age_scatter<<<(nt + 127) / 128, 128, 0, st>>>(m->n_age_cells, m->age_nodes.p, m->age_w.p, m->age_mid.p,
                                                                     m->age_mats.p, L->rowptr.p, L->col.p, L->vals.p, m->err_flag.p);
    }
    const int nb = (m->n_elems + 255) / 256;
    if (m->axisymmetric)
        ++g_launch, kh_assemble<true><<<nb, 256, 0, st>>>(m->n_elems, iter, m->n_nodes, m->omega, (const int4 *)m->tri.p,
            (const double2 *)m->xy.p, m->emap.p, m->labels.p, m->bhB.p, m->bhHr.p, m->bhHi.p, m->bhSr.p, m->bhSi.p, L->V.p,
            L->rowptr.p, L->col.p, L->vals.p, L->b.p, (double4 *)m->mu.p, m->ext_zo, m->ext_ro, m->ext_ri, m->err_flag.p,
            newton, L->vals_h.p, L->vals_s.p, L->vals_a.p);
    else
        ++g_launch, kh_assemble<false><<<nb, 256, 0, st>>>(m->n_elems, iter, m->n_nodes, m->omega, (const int4 *)m->tri.p,
            (const double2 *)m->xy.p, m->emap.p, m->labels.p, m->bhB.p, m->bhHr.p, m->bhHi.p, m->bhSr.p, m->bhSi.p, L->V.p,
            L->rowptr.p, L->col.p, L->vals.p, L->b.p, (double4 *)m->mu.p, m->ext_zo, m->ext_ro, m->ext_ri, m->err_flag.p,
            newton, L->vals_h.p, L->vals_s.p, L->vals_a.p);
    if (m->n_side > 0)
        ++g_launch, kh_sides<<<(m->n_side + 127) / 128, 128, 0, st>>>(m->n_side, m->n_elems, m->side_es.p, m->side_k.p,
                                                                     (const int4 *)m->tri.p, (const double2 *)m->xy.p, m->emap.p,
                                                                     L->vals.p, L->b.p, m->axisymmetric);
    if (m->n_psrc > 0)
        ++g_launch, kh_add_rhs<<<(m->n_psrc + 127) / 128, 128, 0, st>>>(m->n_psrc, m->psrc_node.p, (const cd *)m->psrc_val.p, L->b.p);
    int flag = 0;
    XB_CUDA(cudaMemcpyAsync(&flag, m->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (elapsed_ms) XB_CUDA(cudaEventRecord(e1, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    if (elapsed_ms)
    {
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *elapsed_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    if (flag) { set_error("assembly found couplings missing from the pattern (code %d: 2 air gap, 3 circuit column)", flag); return XFEMM_ERR_ARG; }
    return XFEMM_OK;
}

// after the Dirichlet conditions: copy A(0,0) onto the diagonals of circuits whose unknown is not used
int xfemm_cmag_fix_circuit_diagonals(xfemm_cmag *m)
{
    XB_ARG(m && m->lin, "null handle");
    if (m->n_fix <= 0) return XFEMM_OK;
    xfemm_clinprob *L = m->lin;
    ++g_launch, kh_fix_diag<<<(m->n_fix + 127) / 128, 128, 0, L->stream>>>(m->n_fix, m->fix_rows.p, L->rowptr.p, L->col.p, L->vals.p);
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

int xfemm_cmag_get_mu(xfemm_cmag *m, double *mu4)
{
    XB_ARG(m && mu4, "null pointer");
    XB_CUDA(cudaMemcpyAsync(mu4, m->mu.p, (size_t)m->n_elems * 4 * sizeof(double), cudaMemcpyDeviceToHost, m->lin->stream));
    XB_CUDA(cudaStreamSynchronize(m->lin->stream));
    return XFEMM_OK;
}

int xfemm_cmag_solution_to_A(xfemm_cmag *m, double *A_host)
{
    XB_ARG(m && A_host, "null pointer");
    xfemm_clinprob *L = m->lin;
    cudaStream_t st = L->stream;
    ++g_launch, kh_scale_to_A<<<(L->n + 255) / 256, 256, 0, st>>>(L->n, m->n_nodes, L->V.p, (const double2 *)m->xy.p,
                                                                m->axisymmetric, m->omega, L->b.p);
    xb::g_d2h_bytes += (long long)L->n * 16;
    XB_CUDA(cudaMemcpyAsync(A_host, L->b.p, (size_t)L->n * sizeof(cd), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

} // extern "C"
