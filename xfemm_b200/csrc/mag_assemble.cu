// xfemm_mag_* — per-triangle stiffness / source assembly on the GPU.
// Replaces the element loops of FSolver::Static2D (cfemm/fsolver/static2d.cpp:352-816) and
// FSolver::StaticAxisymmetric (cfemm/fsolver/staticaxi.cpp:189-640): one thread per element,
// B-H reluctivity lookup + Newton d(nu)/dB^2 term on the device, FP64 atomics into the fixed
// CSR pattern of the bound linear problem through a per-element slot map built once per mesh.
#include "linprob.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <vector>
#include <algorithm>
#include <cmath>

#define XB_PI 3.141592653589793238462643383
#define XB_MUO 1.2566370614359173e-6

struct xfemm_mag
{
    xfemm_linprob *lin = nullptr;
    int axisymmetric = 0;
    int n_nodes = 0, n_elems = 0, n_labels = 0, n_bh = 0;
    xb::DevBuf<double> xy;        // n_nodes*2 (double2)
    xb::DevBuf<int> tri;          // n_elems*4 (int4: p0,p1,p2,label)
    xb::DevBuf<int> emap;         // 9*n_elems, SoA: emap[q*n_elems+e] = CSR slot of (p_j,p_k), q=3j+k
    xb::DevBuf<double> magdir;    // optional per element
    xb::DevBuf<xfemm_mag_label> labels;
    xb::DevBuf<double> bhB, bhH, bhS;
    xb::DevBuf<double> mu;        // n_elems*2
    // mixed-BC edges
    int n_mixed = 0;
    xb::DevBuf<int> mixed_es;
    xb::DevBuf<double> mixed_c;
    // air-gap elements
    int n_age_cells = 0, n_age_mats = 0;
    xb::DevBuf<int> age_nodes, age_mid;
    xb::DevBuf<double> age_w, age_mats;
    xb::DevBuf<double> gap_out; // torque + per-cell flux densities of the last gap integral
    // point sources
    int n_psrc = 0;
    xb::DevBuf<int> psrc_node;
    xb::DevBuf<double> psrc_val;
    double ext_zo = 0, ext_ro = 0, ext_ri = 0, r_eps = 0;
    xb::DevBuf<int> err_flag;
    // deterministic ("row-owner gather") assembly: element kernels store their 9 + 3 contributions in det_scratch
    // (slot q*ne+e, right-hand side 9*ne + j*ne + e), a gather pass adds, per CSR slot / row, its contributors in
    // ascending order of that index (lists built once per mesh: stable radix sort by target)
    int deterministic = 0;
    xb::DevBuf<double> det_scratch; // 12 * n_elems
    xb::DevBuf<int> det_src;        // 12 * n_elems contributor indices, sorted by target
    xb::DevBuf<int> det_ptr;        // nnz + n_nodes + 1
};

namespace xb {

__device__ __forceinline__ int csr_find(const int *__restrict__ rowptr, const int *__restrict__ col, int row, int c)
{
    int lo = rowptr[row], hi = rowptr[row + 1] - 1;
    while (lo <= hi)
    {
        int mid = (lo + hi) >> 1;
        int v = col[mid];
        if (v == c) return mid;
        if (v < c) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

__global__ void k_build_emap(int ne, const int4 *__restrict__ tri, const int *__restrict__ rowptr,
                             const int *__restrict__ col, int *__restrict__ emap, int *err)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    int4 t = tri[e];
    int n[3] = {t.x, t.y, t.z};
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            int s = csr_find(rowptr, col, n[j], n[k]);
            if (s < 0) atomicExch(err, 1);
            emap[(size_t)(3 * j + k) * ne + e] = s;
        }
}

// Cubic-Hermite H(B) lookup: nu = H/B and dnu/dB^2 (reference: CMaterialProp.cpp:1007-1057).
__device__ __forceinline__ void bh_props(const double *__restrict__ Bd, const double *__restrict__ Hd,
                                         const double *__restrict__ Sd, int npts, double mu_x, double B,
                                         double &v, double &dv)
{
    double b = fabs(B);
    if (npts == 0) { v = mu_x; dv = 0; return; }
    if (b == 0) { v = Sd[0]; dv = 0; return; }
    if (b > Bd[npts - 1])
    {
        double h = Hd[npts - 1] + Sd[npts - 1] * (b - Bd[npts - 1]);
        double dh = Sd[npts - 1];
        v = h / b;
        dv = 0.5 * (dh / (b * b) - h / (b * b * b));
        return;
    }
    v = 0; dv = 0; // (the reference leaves v,dv untouched if no interval matches; cannot happen for b>=Bd[0]=0)
    for (int i = 0; i < npts - 1; i++)
    {
        if (b >= Bd[i] && b <= Bd[i + 1])
        {
            double l = Bd[i + 1] - Bd[i];
            double z = (b - Bd[i]) / l, z2 = z * z;
            double h = (1. - 3. * z2 + 2. * z2 * z) * Hd[i] + z * (1. - 2. * z + z2) * l * Sd[i] +
                       z2 * (3. - 2. * z) * Hd[i + 1] + z2 * (z - 1.) * l * Sd[i + 1];
            double dh = 6. * z * (z - 1.) * Hd[i] / l + (1. - 4. * z + 3. * z * z) * Sd[i] +
                        6. * z * (1. - z) * Hd[i + 1] / l + z * (3. * z - 2.) * Sd[i + 1];
            v = h / b;
            dv = 0.5 * (dh / (b * b) - h / (b * b * b));
            return;
        }
    }
}

// ---------------------------------------------------------------------------------------
// planar magnetostatic element (static2d.cpp:368-815)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_assemble_static2d(int ne, int iter, const int4 *__restrict__ tri, const double2 *__restrict__ xy,
                    const int *__restrict__ emap, const xfemm_mag_label *__restrict__ labels,
                    const double *__restrict__ magdir_elem, const double *__restrict__ bhB,
                    const double *__restrict__ bhH, const double *__restrict__ bhS,
                    const double *__restrict__ V, double *vals, double *b, double2 *mu_out, int det)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const double c = XB_PI * 4.e-05;
    const int4 t = tri[e];
    const int n[3] = {t.x, t.y, t.z};
    const double2 P0 = xy[n[0]], P1 = xy[n[1]], P2 = xy[n[2]];
    const double p[3] = {P1.y - P2.y, P2.y - P0.y, P0.y - P1.y};
    const double q[3] = {P2.x - P1.x, P0.x - P2.x, P1.x - P0.x};
    const double a = (p[0] * q[1] - p[1] * q[0]) / 2.;
    const double K = -1. / (4. * a);
    const xfemm_mag_label L = labels[t.w];

    double Mx[3][3], My[3][3], Mn[3][3], be[3];
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            Mx[j][k] = K * p[j] * p[k];
            My[j][k] = K * q[j] * q[k];
            Mn[j][k] = 0.;
        }
        // current density incl. circuit term (static2d.cpp:483-504)
        be[j] = -(L.Jsrc) * a / 3.;
    }
    // permanent magnet (static2d.cpp:584-598); dx,dy of side j→k are -q, p of the opposite node
    if (L.H_c != 0.)
    {
        const double th = (magdir_elem ? magdir_elem[e] : L.magdir) * XB_PI / 180.;
        const double cs = cos(th), sn = sin(th);
        const double X[3] = {P0.x, P1.x, P2.x}, Y[3] = {P0.y, P1.y, P2.y};
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            const int k = (j + 1) % 3;
            const double Km = 0.0001 * L.H_c * (cs * (X[k] - X[j]) + sn * (Y[k] - Y[j])) / 2.;
            be[j] += Km;
            be[k] += Km;
        }
    }

    double mu1 = L.mu1, mu2 = L.mu2;
    double Vn[3] = {0., 0., 0.};
    if (iter > 0 && L.bh_points > 0)
    {
        Vn[0] = V[n[0]]; Vn[1] = V[n[1]]; Vn[2] = V[n[2]];
        const double *Bd = bhB + L.bh_offset, *Hd = bhH + L.bh_offset, *Sd = bhS + L.bh_offset;
        double v[3], u[3], nu, dv;
        if (L.lamtype == 0)
        {
            // static2d.cpp:687-717 (mu1==mu2 always holds for B-H materials here: GetSlopes sets mu_y=mu_x)
            double B1 = 0., B2 = 0.;
#pragma unroll
            for (int j = 0; j < 3; j++) { B1 += Vn[j] * q[j]; B2 += Vn[j] * p[j]; }
            const double B = c * sqrt(B1 * B1 + B2 * B2) / (0.02 * a);
            bh_props(Bd, Hd, Sd, L.bh_points, L.mu1, B, nu, dv);
            const double mu = 1. / (XB_MUO * nu);
            mu1 = mu; mu2 = mu;
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
                v[j] = 0;
#pragma unroll
                for (int w = 0; w < 3; w++) v[j] += (Mx[j][w] + My[j][w]) * Vn[w];
            }
            const double Kn = -200. * c * c * c * dv / a;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int w = 0; w < 3; w++) Mn[j][w] = Kn * v[j] * v[w];
        }
        else if (L.lamtype == 1 || L.lamtype == 2)
        {
            // laminated in-plane variants (static2d.cpp:719-796)
            const double tt = L.lamfill;
            double B1 = 0., B2 = 0.;
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
                if (L.lamtype == 1) { B1 += Vn[j] * q[j]; B2 += Vn[j] * p[j] / tt; }
                else { B1 += (Vn[j] * q[j]) / tt; B2 += Vn[j] * p[j]; }
            }
            const double B = c * sqrt(B1 * B1 + B2 * B2) / (0.02 * a);
            bh_props(Bd, Hd, Sd, L.bh_points, L.mu1, B, nu, dv);
            const double mu = 1. / (XB_MUO * nu);
            if (L.lamtype == 1) { mu1 = mu * tt; mu2 = mu / (tt + mu * (1. - tt)); }
            else { mu2 = mu * tt; mu1 = mu / (tt + mu * (1. - tt)); }
#pragma unroll
            for (int j = 0; j < 3; j++)
            {
                v[j] = 0; u[j] = 0;
#pragma unroll
                for (int w = 0; w < 3; w++)
                {
                    if (L.lamtype == 1)
                    {
                        v[j] += (My[j][w] / tt + Mx[j][w]) * Vn[w];
                        u[j] += (My[j][w] / tt + tt * Mx[j][w]) * Vn[w];
                    }
                    else
                    {
                        v[j] += (Mx[j][w] / tt + My[j][w]) * Vn[w];
                        u[j] += (Mx[j][w] / tt + tt * My[j][w]) * Vn[w];
                    }
                }
            }
            const double Kn = -100. * c * c * c * dv / a;
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int w = 0; w < 3; w++) Mn[j][w] = Kn * (v[j] * u[w] + v[w] * u[j]);
        }
    }
    if (mu_out) mu_out[e] = make_double2(mu1, mu2);

    // Me = Mx/mu2 + My/mu1 + Mn ; be += Mn V ; scatter -Me, -be (static2d.cpp:799-815)
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        double bj = be[j];
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            const double me = Mx[j][k] / mu2 + My[j][k] / mu1 + Mn[j][k];
            bj += Mn[j][k] * Vn[k];
            // det: own scratch cell of this contribution (vals = b = the scratch array): a plain, coalesced store;
            // k_det_gather adds the cells up per slot
            const size_t cell = (size_t)(3 * j + k) * ne + e;
            if (det) vals[cell] = -me;
            else atomicAdd(vals + emap[cell], -me);
        }
        if (det) b[(size_t)(9 + j) * ne + e] = -bj;
        else atomicAdd(b + n[j], -bj);
    }
}

// ---------------------------------------------------------------------------------------
// axisymmetric magnetostatic element (staticaxi.cpp:169-633): r-weighted matrices with the
// a_hat / R_hat (log-mean radius) terms, special cases for nodes on the axis, B from the element
// energy, conformally mapped external region.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_assemble_staticaxi(int ne, int iter, const int4 *__restrict__ tri, const double2 *__restrict__ xy,
                     const int *__restrict__ emap, const xfemm_mag_label *__restrict__ labels,
                     const double *__restrict__ magdir_elem, const double *__restrict__ bhB,
                     const double *__restrict__ bhH, const double *__restrict__ bhS,
                     const double *__restrict__ V, double *vals, double *b, double2 *mu_out,
                     double ext_zo, double ext_ro, double ext_ri, int det)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const double c = XB_PI * 4.e-05;
    const int4 t4 = tri[e];
    const int n[3] = {t4.x, t4.y, t4.z};
    const double2 P0 = xy[n[0]], P1 = xy[n[1]], P2 = xy[n[2]];
    const double X[3] = {P0.x, P1.x, P2.x}, Y[3] = {P0.y, P1.y, P2.y};
    const double rn[3] = {P0.x, P1.x, P2.x};
    const double p[3] = {P1.y - P2.y, P2.y - P0.y, P0.y - P1.y};
    const double q[3] = {P2.x - P1.x, P0.x - P2.x, P1.x - P0.x};
    const double g[3] = {(P2.x + P1.x) / 2., (P0.x + P2.x) / 2., (P1.x + P0.x) / 2.};
    const double a = (p[0] * q[1] - p[1] * q[0]) / 2.;
    const double R = (P0.x + P1.x + P2.x) / 3.;
    double a_hat = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) a_hat += (rn[j] * rn[j] * p[j] / (4. * R));
    const double vol = 2. * R * a_hat;
    int flag = 0;
#pragma unroll
    for (int j = 0; j < 3; j++) if (rn[j] < 1.e-06) flag++;
    double R_hat = 0.;
    if (flag == 2) R_hat = R;
    else if (flag == 1)
    {
        // staticaxi.cpp:226-244 (one node on the axis)
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
    const xfemm_mag_label L = labels[t4.w];
    double Mx[3][3], My[3][3], Mn[3][3], be[3];
    const double Kx = (-1. / (2. * a_hat * R)), Ky = (-1. / (2. * a_hat * R_hat));
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            Mx[j][k] = Kx * p[j] * rn[j] * p[k] * rn[k];
            My[j][k] = Ky * (q[j] * rn[j]) * (q[k] * rn[k]) * (g[j] / R) * (g[k] / R);
            Mn[j][k] = 0.;
        }
    {
        // nodes on the r=0 line get the trace on their diagonal (staticaxi.cpp:276-277; the loop
        // runs in order, so a later node sees the already patched entries)
#pragma unroll
        for (int j = 0; j < 3; j++)
            if (rn[j] < 1.e-06) Mx[j][j] += Mx[0][0] + Mx[1][1] + Mx[2][2];
    }
    // source current density (staticaxi.cpp:332-348): circuit voltage term carries 1/R
    const double Jel = L.Jsrc + L.dVs / R;
#pragma unroll
    for (int j = 0; j < 3; j++) be[j] = -2. * R * Jel * a / 3.;
    if (L.H_c != 0.)
    {
        const double th = (magdir_elem ? magdir_elem[e] : L.magdir) * XB_PI / 180.;
        const double cs = cos(th), sn = sin(th);
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            const int k = (j + 1) % 3;
            const double r = (X[j] + X[k]) / 2.;
            const double Km = -0.0001 * r * L.H_c * (cs * (X[k] - X[j]) + sn * (Y[k] - Y[j]));
            be[j] += Km;
            be[k] += Km;
        }
    }
    double mu1 = L.mu1, mu2 = L.mu2;
    if (L.is_external)
    {
        // conformally mapped exterior (staticaxi.cpp:611-617): applied at Iter 0 and kept
        const double Z = (Y[0] + Y[1] + Y[2]) / 3. - ext_zo;
        const double kludge = (R * R + Z * Z) * ext_ri / (ext_ro * ext_ro * ext_ro);
        mu1 /= kludge;
        mu2 /= kludge;
    }
    double Vn[3] = {0., 0., 0.};
    if (iter > 0 && L.bh_points > 0)
    {
        Vn[0] = V[n[0]]; Vn[1] = V[n[1]]; Vn[2] = V[n[2]];
        const double *Bd = bhB + L.bh_offset, *Hd = bhH + L.bh_offset, *Sd = bhS + L.bh_offset;
        const double tt = L.lamfill;
        double v[3], u[3], nu, dv = 0;
        // B from the element energy (staticaxi.cpp:511-520 / 542-552 / 575-585)
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            v[j] = 0;
#pragma unroll
            for (int w = 0; w < 3; w++)
            {
                double m = Mx[j][w] + My[j][w];
                if (L.lamtype == 1) m = Mx[j][w] + My[j][w] / (tt * tt);
                if (L.lamtype == 2) m = Mx[j][w] / (tt * tt) + My[j][w];
                v[j] += m * Vn[w];
            }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) dv += Vn[j] * v[j];
        dv *= (10000. * c * c / vol);
        const double B = sqrt(fabs(dv));
        if (L.lamtype <= 2)
        {
            bh_props(Bd, Hd, Sd, L.bh_points, L.mu1, B, nu, dv);
            const double mu = 1. / (XB_MUO * nu);
            if (L.lamtype == 0)
            {
                mu1 = mu; mu2 = mu;
                const double Kn = -200. * c * c * c * dv / vol;
#pragma unroll
                for (int j = 0; j < 3; j++)
                {
                    v[j] = 0;
#pragma unroll
                    for (int w = 0; w < 3; w++) v[j] += (Mx[j][w] + My[j][w]) * Vn[w];
                }
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int w = 0; w < 3; w++) Mn[j][w] = Kn * v[j] * v[w];
            }
            else
            {
                if (L.lamtype == 1) { mu1 = mu * tt; mu2 = mu / (tt + mu * (1. - tt)); }
                else { mu2 = mu * tt; mu1 = mu / (tt + mu * (1. - tt)); }
#pragma unroll
                for (int j = 0; j < 3; j++)
                {
                    v[j] = 0; u[j] = 0;
#pragma unroll
                    for (int w = 0; w < 3; w++)
                    {
                        if (L.lamtype == 1)
                        {
                            v[j] += (My[j][w] / tt + Mx[j][w]) * Vn[w];
                            u[j] += (My[j][w] / tt + tt * Mx[j][w]) * Vn[w];
                        }
                        else
                        {
                            v[j] += (Mx[j][w] / tt + My[j][w]) * Vn[w];
                            u[j] += (Mx[j][w] / tt + tt * My[j][w]) * Vn[w];
                        }
                    }
                }
                const double Kn = -100. * c * c * c * dv / vol;
#pragma unroll
                for (int j = 0; j < 3; j++)
#pragma unroll
                    for (int w = 0; w < 3; w++) Mn[j][w] = Kn * (v[j] * u[w] + v[w] * u[j]);
            }
        }
    }
    if (mu_out) mu_out[e] = make_double2(mu1, mu2);
#pragma unroll
    for (int j = 0; j < 3; j++)
    {
        double bj = be[j];
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            const double me = Mx[j][k] / mu2 + My[j][k] / mu1 + Mn[j][k];
            bj += Mn[j][k] * Vn[k];
            // det: own scratch cell of this contribution (vals = b = the scratch array): a plain, coalesced store;
            // k_det_gather adds the cells up per slot
            const size_t cell = (size_t)(3 * j + k) * ne + e;
            if (det) vals[cell] = -me;
            else atomicAdd(vals + emap[cell], -me);
        }
        if (det) b[(size_t)(9 + j) * ne + e] = -bj;
        else atomicAdd(b + n[j], -bj);
    }
}

// mixed (Robin) boundary edges, BdryFormat==2 (static2d.cpp:460-480; staticaxi.cpp:308-329 adds 2r)
__global__ void k_mixed_edges(int nm, int ne, const int *__restrict__ es, const double *__restrict__ c01,
                              const int4 *__restrict__ tri, const double2 *__restrict__ xy,
                              const int *__restrict__ emap, double *vals, double *b, int axi, int det)
{
    int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nm) return;
    const double c = XB_PI * 4.e-05;
    const int e = es[2 * m], j = es[2 * m + 1], k = (j + 1) % 3;
    const int4 t = tri[e];
    const int n[3] = {t.x, t.y, t.z};
    const double2 A = xy[n[j]], B = xy[n[k]];
    const double l = sqrt((B.x - A.x) * (B.x - A.x) + (B.y - A.y) * (B.y - A.y));
    const double rfac = axi ? 2. * ((A.x + B.x) / 2.) : 1.;
    const double K = axi ? -0.0001 * c * rfac * c01[2 * m] * l / 6. : -0.0001 * c * c01[2 * m] * l / 6.;
    auto slot = [&](int a, int bq) { const size_t cell = (size_t)(3 * a + bq) * ne + e; return det ? cell : (size_t)emap[cell]; };
    atomicAdd(vals + slot(j, j), -(K * 2.));
    atomicAdd(vals + slot(k, k), -(K * 2.));
    atomicAdd(vals + slot(j, k), -K);
    atomicAdd(vals + slot(k, j), -K);
    const double K2 = axi ? (c01[2 * m + 1] * l / 2.) * 0.0001 * rfac : (c01[2 * m + 1] * l / 2.) * 0.0001;
    atomicAdd(b + (det ? (size_t)(9 + j) * ne + e : (size_t)n[j]), -K2);
    atomicAdd(b + (det ? (size_t)(9 + k) * ne + e : (size_t)n[k]), -K2);
}

// air-gap element ring cells: L.AddTo(MG[ii][jj]*ww[ii]*ww[jj], nn[ii], nn[jj]) for jj>=ii
// (static2d.cpp:345-347).  sign=-1 for the harmonic formulation (harmonic2d.cpp:382).
__global__ void k_age_scatter(int ncells, const int *__restrict__ nodes, const double *__restrict__ w,
                              const int *__restrict__ mid, const double *__restrict__ mats,
                              const int *__restrict__ rowptr, const int *__restrict__ col, double *vals,
                              double sign, int *err)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncells * 55) return;
    const int cell = t / 55;
    int q = t % 55, ii = 0;
    while (q >= 10 - ii) { q -= 10 - ii; ii++; }
    const int jj = ii + q;
    const int a = nodes[cell * 10 + ii], bnode = nodes[cell * 10 + jj];
    const double v = sign * (mats[mid[cell] * 100 + ii * 10 + jj] * w[cell * 10 + ii] * w[cell * 10 + jj]);
    const int s1 = csr_find(rowptr, col, a, bnode);
    if (s1 < 0) { atomicExch(err, 2); return; }
    atomicAdd(vals + s1, v);
    if (a != bnode)
    {
        const int s2 = csr_find(rowptr, col, bnode, a);
        if (s2 < 0) { atomicExch(err, 2); return; }
        atomicAdd(vals + s2, v);
    }
}

// ---- deterministic assembly (see xfemm_mag::deterministic) -------------------------------------------
// target of contribution idx: a CSR slot (idx < 9 ne) or nnz + row (right-hand side)
__global__ void k_det_keys(int ne, long long nnz, const int *__restrict__ emap, const int4 *__restrict__ tri, int *keys, int *vals_idx)
{
    const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= 12ll * ne) return;
    vals_idx[idx] = (int)idx;
    if (idx < 9ll * ne) { keys[idx] = emap[idx]; return; }
    const int j = (int)((idx - 9ll * ne) / ne), e = (int)((idx - 9ll * ne) % ne);
    const int4 t = tri[e];
    keys[idx] = (int)nnz + (j == 0 ? t.x : j == 1 ? t.y : t.z);
}
__global__ void k_det_count(long long len, const int *__restrict__ keys, int *cnt)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < len) atomicAdd(cnt + keys[i], 1);
}
// vals[s] += sum of the contributions of slot s, b[r] += those of row r, each list in ascending contributor order
__global__ void k_det_gather(long long nnz, int n, const int *__restrict__ ptr, const int *__restrict__ src,
                             const double *__restrict__ scratch, double *vals, double *b)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= nnz + n) return;
    const int k0 = ptr[t], k1 = ptr[t + 1];
    if (k0 == k1) return;
    double sum = 0.0;
    for (int k = k0; k < k1; k += 4)
    {
        // four at a time, loads first (lists are short: ~3 contributors per slot, ~6 per row)
        int i0 = src[k], i1 = k + 1 < k1 ? src[k + 1] : -1, i2 = k + 2 < k1 ? src[k + 2] : -1, i3 = k + 3 < k1 ? src[k + 3] : -1;
        const double v0 = scratch[i0], v1 = i1 >= 0 ? scratch[i1] : 0.0, v2 = i2 >= 0 ? scratch[i2] : 0.0, v3 = i3 >= 0 ? scratch[i3] : 0.0;
        sum += v0; sum += v1; sum += v2; sum += v3;
    }
    if (t < nnz) vals[t] += sum;
    else b[t - nnz] += sum;
}
// air-gap blocks in a fixed order: one block, the cells one after the other, the 55 pairs of a cell in parallel
__global__ void k_age_scatter_ordered(int ncells, const int *__restrict__ nodes, const double *__restrict__ w,
                                      const int *__restrict__ mid, const double *__restrict__ mats,
                                      const int *__restrict__ rowptr, const int *__restrict__ col, double *vals, double sign, int *err)
{
    for (int cell = 0; cell < ncells; cell++)
    {
        const int t = threadIdx.x;
        if (t < 55)
        {
            int q = t, ii = 0;
            while (q >= 10 - ii) { q -= 10 - ii; ii++; }
            const int jj = ii + q;
            const int a = nodes[cell * 10 + ii], bnode = nodes[cell * 10 + jj];
            const double v = sign * (mats[mid[cell] * 100 + ii * 10 + jj] * w[cell * 10 + ii] * w[cell * 10 + jj]);
            const int s1 = csr_find(rowptr, col, a, bnode);
            if (s1 < 0) atomicExch(err, 2);
            else
            {
                atomicAdd(vals + s1, v); // within a cell a slot is hit at most twice (repeated ring node): commutative
                if (a != bnode)
                {
                    const int s2 = csr_find(rowptr, col, bnode, a);
                    if (s2 < 0) atomicExch(err, 2);
                    else atomicAdd(vals + s2, v);
                }
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// DC torque through an air-gap element from the resident solution (reference post-processor:
// cfemm/fpproc/fpproc.cpp:1391-1615 flux density at the centre of every ring cell from the 10 ring
// nodes and its harmonics, fpproc.cpp:5418-5446 torque = pi R^2 Depth/mu0 * sum_k Re(brc btc* + brs bts*)).
// One block; ncells cells starting at cell0 of the flattened ring tables; A = cscale * V.
// out[0] = torque [N m]; out[1 .. ncells] = Br, out[1+ncells .. 2 ncells] = Bt per cell [T].
__global__ void k_gap_torque(int cell0, int ncells, const int *__restrict__ nodes, const double *__restrict__ w,
                             const double *__restrict__ V, double cscale, double ci, double co, double dt, double R,
                             double dr, int bdry_format, int m, int nharm, double scale, double *out)
{
    extern __shared__ double sh[]; // br[ncells], bt[ncells], partial[blockDim/32]
    double *br = sh, *bt = sh + ncells, *part = sh + 2 * ncells;
    for (int k = threadIdx.x; k < ncells; k += blockDim.x)
    {
        double a[10];
        for (int q = 0; q < 10; q++) a[q] = cscale * V[nodes[(size_t)(cell0 + k) * 10 + q]] * w[(size_t)(cell0 + k) * 10 + q];
        const double ci2 = ci * ci, ci3 = ci2 * ci, co2 = co * co, co3 = co2 * co;
        const double r = (-(ci * a[1]) - 2 * a[2] + 2 * a[3] + ci * (a[2] + a[3] - a[4]) - ci3 * (a[0] - 4 * a[1] + 6 * a[2] - 4 * a[3] + a[4]) +
                          ci2 * (a[0] - 5 * a[1] + 9 * a[2] - 7 * a[3] + 2 * a[4]) - 2 * a[7] + 2 * a[8] + co * (-a[6] + a[7] + a[8] - a[9]) -
                          co3 * (a[5] - 4 * a[6] + 6 * a[7] - 4 * a[8] + a[9]) + co2 * (a[5] - 5 * a[6] + 9 * a[7] - 7 * a[8] + 2 * a[9])) / (4 * dt * R);
        const double t = (ci * a[1] + 2 * a[2] + 2 * a[3] - ci2 * (a[0] - 3 * a[1] + a[2] + 3 * a[3] - 2 * a[4]) + ci * (a[2] - a[3] - a[4]) +
                          ci3 * (a[0] - 2 * a[1] + 2 * a[3] - a[4]) - co * a[6] + (-2 + co) * (1 + co) * a[7] - 2 * a[8] +
                          co * (a[8] + co * (a[5] - 3 * a[6] + 3 * a[8] - 2 * a[9]) + a[9] + co2 * (-a[5] + 2 * a[6] - 2 * a[8] + a[9]))) / (4 * dr);
        br[k] = r; bt[k] = t;
        out[1 + k] = r; out[1 + ncells + k] = t;
    }
    __syncthreads();
    double tq = 0.0;
    for (int j = threadIdx.x; j < nharm; j += blockDim.x)
    {
        const int nh = bdry_format == 0 ? m * j : m * (2 * j + 1);
        double brc = 0, brs = 0, btc = 0, bts = 0;
        for (int k = 0; k < ncells; k++)
        {
            const double tta = ((double)k + 0.5) * dt * nh;
            const double c = cos(tta), s = sin(tta);
            brc += br[k] * c; brs += br[k] * s; btc += bt[k] * c; bts += bt[k] * s;
        }
        const bool whole = (nh == 0) || ((j == nharm - 1) && bdry_format == 0 && (ncells % 2) == 0);
        const double div = whole ? (double)ncells : (double)ncells / 2.;
        tq += (brc / div) * (btc / div) + (brs / div) * (bts / div);
    }
    tq = xb::warp_sum(tq);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = tq;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0;
        for (int q = 0; q < (int)(blockDim.x >> 5); q++) t += part[q];
        out[0] = t * scale;
    }
}

// ---------------------------------------------------------------------------------------------------
// Block integrals of the reference post-processor from the resident solution (no .ans file, no nodal D2H):
// per block label l,  E_l = sum_e a_e w_e  (stored energy density, fpproc.cpp:3762-3830 case 2 with
// CMMaterialProp::DoEnergy / GetEnergy, CMaterialProp.cpp:536-650) and  AJ_l = sum_e a_e mean(A_e)  (the A.J
// integrand of fpproc.cpp:3678-3686 / GetFluxLinkage 5157-5196 for a current density that is constant over the
// label).  Planar magnetostatics; a_e in m^2, A in Wb/m, B in T from the three nodal A as FPProc::GetElementB
// (fpproc.cpp:2970-3004).  The host multiplies by the depth, the label's current density and 1/I.
// ---------------------------------------------------------------------------------------------------
constexpr int INTEGRAL_MAX_LABELS = 32;
struct IntegralLabels
{
    int n;
    xfemm_integral_label l[INTEGRAL_MAX_LABELS];
};

// stored energy density of a B-H material at |B| = x: integral of H dB along the cubic Hermite curve
// (CMMaterialProp::GetEnergy, CMaterialProp.cpp:536-585)
__device__ double bh_energy(double x, int np, const double *__restrict__ Bd, const double *__restrict__ Hd, const double *__restrict__ Sd)
{
    const double b = fabs(x);
    double nrg = 0.0;
    if (np == 0) return 0.0;
    for (int i = 0; i < np - 1; i++)
    {
        const double b0 = Bd[i], h0 = Hd[i], b1 = Bd[i + 1], h1 = Hd[i + 1], dh0 = Sd[i], dh1 = Sd[i + 1];
        if (b >= b0 && b <= b1)
        {
            const double l = b1 - b0, z = (b - b0) / l, z2 = z * z;
            nrg += (dh0 * l * l * (6. + z * (-8. + 3. * z)) * z2) / 12. + (h0 * l * z * (2. + (-2. + z) * z2)) / 2. -
                   (h1 * l * (-2. + z) * z2 * z) / 2. + (dh1 * l * l * (-4. + 3. * z) * z2 * z) / 12;
            return nrg;
        }
        nrg += ((b0 - b1) * ((b0 - b1) * (dh0 - dh1) - 6. * (h0 + h1))) / 12.;
    }
    const double h0 = Hd[np - 1], dh0 = Sd[np - 1], b0 = Bd[np - 1];
    nrg += ((b - b0) * (b * dh0 - b0 * dh0 + 2 * h0)) / 2.;
    return nrg;
}

__global__ void __launch_bounds__(256)
k_label_integrals(int ne, const int4 *__restrict__ tri, const double2 *__restrict__ xy, const double *__restrict__ V, double cscale,
                  const double *__restrict__ magdir_elem, const double *__restrict__ bhB, const double *__restrict__ bhH,
                  const double *__restrict__ bhS, IntegralLabels T, double *part /* [gridDim][2][n] */, int *unsupported)
{
    __shared__ double sm[33];
    double accE[INTEGRAL_MAX_LABELS], accA[INTEGRAL_MAX_LABELS];
#pragma unroll
    for (int l = 0; l < INTEGRAL_MAX_LABELS; l++) { accE[l] = 0.0; accA[l] = 0.0; }
    const double muo = 4.e-7 * XB_PI;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x)
    {
        const int4 t = tri[e];
        const int lb = t.w;
        if (lb < 0 || lb >= T.n) continue;
        const xfemm_integral_label &M = T.l[lb];
        const double2 p0 = xy[t.x], p1 = xy[t.y], p2 = xy[t.z];
        const double b0 = p1.y - p2.y, b1 = p2.y - p0.y, b2 = p0.y - p1.y;
        const double c0 = p2.x - p1.x, c1 = p0.x - p2.x, c2 = p1.x - p0.x;
        const double da = b0 * c1 - b1 * c0;         // cm^2 (twice the area)
        const double a = fabs(da) * 0.5e-4;          // m^2
        const double A0 = cscale * V[t.x], A1 = cscale * V[t.y], A2 = cscale * V[t.z];
        const double B1 = (A0 * c0 + A1 * c1 + A2 * c2) / (da * 0.01);
        const double B2 = -(A0 * b0 + A1 * b1 + A2 * b2) / (da * 0.01);
        double w = 0.0;
        if (M.lamtype > 2) { atomicExch(unsupported, 1); }
        else if (M.H_c != 0.0)
        {
            if (M.bh_points != 0) atomicExch(unsupported, 2);
            else
            {
                // second-quadrant representation of a linear permanent magnet (fpproc.cpp:3786-3800)
                const double dir = (magdir_elem ? magdir_elem[e] : M.magdir) * XB_PI / 180.;
                const double H1 = B1 / (M.mu_x * muo) - M.H_c * cos(dir), H2 = B2 / (M.mu_y * muo) - M.H_c * sin(dir);
                w = 0.5 * muo * (M.mu_x * H1 * H1 + M.mu_y * H2 * H2);
            }
        }
        else if (M.bh_points == 0)
        {
            // CMMaterialProp::DoEnergy, linear branch, literally (CMaterialProp.cpp:606-632; types 1 and 2 use b1 twice)
            double h1 = 0, h2 = 0;
            if (M.lamtype == 0) { h1 = B1 / ((1. + M.lamfill * (M.mu_x - 1.)) * muo); h2 = B2 / ((1. + M.lamfill * (M.mu_y - 1.)) * muo); }
            if (M.lamtype == 1) { h1 = B1 / ((1. + M.lamfill * (M.mu_x - 1.)) * muo); h2 = B1 * (M.lamfill / (M.mu_y * muo) + (1. - M.lamfill) / muo); }
            if (M.lamtype == 2) { h2 = B1 / ((1. + M.lamfill * (M.mu_y - 1.)) * muo); h1 = B1 * (M.lamfill / (M.mu_x * muo) + (1. - M.lamfill) / muo); }
            w = (h1 * B1 + h2 * B2) / 2.;
        }
        else
        {
            const double *Bd = bhB + M.bh_offset, *Hd = bhH + M.bh_offset, *Sd = bhS + M.bh_offset;
            if (M.lamtype == 0) w = bh_energy(sqrt(B1 * B1 + B2 * B2), M.bh_points, Bd, Hd, Sd);
            if (M.lamtype == 1)
            {
                const double biron = sqrt((B1 / M.lamfill) * (B1 / M.lamfill) + B2 * B2), bair = B2;
                w = M.lamfill * bh_energy(biron, M.bh_points, Bd, Hd, Sd) + (1 - M.lamfill) * bair * bair / (2. * muo);
            }
            if (M.lamtype == 2)
            {
                const double biron = sqrt((B2 / M.lamfill) * (B2 / M.lamfill) + B1 * B1), bair = B1;
                w = M.lamfill * bh_energy(biron, M.bh_points, Bd, Hd, Sd) + (1 - M.lamfill) * bair * bair / (2. * muo);
            }
        }
        const double mA = a * (A0 + A1 + A2) / 3.;
        // the label index is data: select the accumulator without dynamic indexing of the register arrays
#pragma unroll
        for (int l = 0; l < INTEGRAL_MAX_LABELS; l++)
            if (l == lb) { accE[l] += a * w; accA[l] += mA; }
    }
    for (int l = 0; l < T.n; l++)
    {
        double ve = 0.0, va = 0.0;
#pragma unroll
        for (int q = 0; q < INTEGRAL_MAX_LABELS; q++) if (q == l) { ve = accE[q]; va = accA[q]; }
        ve = block_sum(ve, sm);
        va = block_sum(va, sm);
        if (threadIdx.x == 0) { part[((size_t)blockIdx.x * 2 + 0) * T.n + l] = ve; part[((size_t)blockIdx.x * 2 + 1) * T.n + l] = va; }
    }
}
__global__ void k_label_integrals_final(int nblk, int n2, const double *__restrict__ part, double *out)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x; // one thread per output, blocks added in order: deterministic
    if (q >= n2) return;
    double s = 0.0;
    for (int b = 0; b < nblk; b++) s += part[(size_t)b * n2 + q];
    out[q] = s;
}

__global__ void k_point_sources(int np, const int *__restrict__ node, const double *__restrict__ val, double *b)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) atomicAdd(b + node[i], val[i]);
}

__global__ void k_scale_to_A(int n, const double *__restrict__ V, const double2 *__restrict__ xy, int axi, double *out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c = XB_PI * 4.e-05;
    double v = V[i] * c;
    if (axi) v = V[i] * c * 2. * XB_PI * xy[i].x * 0.01; // staticaxi.cpp:779-783
    out[i] = v;
}

} // namespace xb

using namespace xb;

extern "C" {

int xfemm_mag_set_labels(xfemm_mag *m, int n_labels, const xfemm_mag_label *labels)
{
    XB_ARG(m && labels && n_labels > 0, "labels");
    XB_CUDA(m->labels.upload(labels, n_labels, m->lin->stream));
    XB_CUDA(cudaStreamSynchronize(m->lin->stream));
    m->n_labels = n_labels;
    return XFEMM_OK;
}

int xfemm_mag_set_age(xfemm_mag *m, int n_age_cells, const int *age_nodes, const double *age_weights,
                      const int *age_matrix_id, int n_age_matrices, const double *age_matrices)
{
    XB_ARG(m, "null handle");
    m->n_age_cells = n_age_cells;
    m->n_age_mats = n_age_matrices;
    if (n_age_cells <= 0) return XFEMM_OK;
    XB_ARG(age_nodes && age_weights && age_matrix_id && age_matrices && n_age_matrices > 0, "AGE tables");
    for (int i = 0; i < n_age_cells * 10; i++) XB_ARG(age_nodes[i] >= 0 && age_nodes[i] < m->n_nodes, "AGE node out of range");
    for (int i = 0; i < n_age_cells; i++) XB_ARG(age_matrix_id[i] >= 0 && age_matrix_id[i] < n_age_matrices, "AGE matrix id");
    cudaStream_t st = m->lin->stream;
    XB_CUDA(m->age_nodes.upload(age_nodes, (size_t)n_age_cells * 10, st));
    XB_CUDA(m->age_w.upload(age_weights, (size_t)n_age_cells * 10, st));
    XB_CUDA(m->age_mid.upload(age_matrix_id, n_age_cells, st));
    XB_CUDA(m->age_mats.upload(age_matrices, (size_t)n_age_matrices * 100, st));
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

int xfemm_mag_create(const xfemm_mag_desc *d, xfemm_linprob *lin, xfemm_mag **out)
{
    XB_ARG(d && lin && out, "null pointer");
    XB_ARG(d->n_nodes > 0 && d->n_elems > 0 && d->xy && d->tri && d->labels && d->n_labels > 0, "mesh arrays");
    XB_ARG(d->n_nodes <= lin->n, "linear problem smaller than the mesh");
    for (int e = 0; e < d->n_elems; e++)
    {
        const int *t = d->tri + 4 * (size_t)e;
        if (t[0] < 0 || t[0] >= d->n_nodes || t[1] < 0 || t[1] >= d->n_nodes || t[2] < 0 || t[2] >= d->n_nodes ||
            t[3] < 0 || t[3] >= d->n_labels)
        {
            set_error("element %d has node or label index out of range", e);
            return XFEMM_ERR_ARG;
        }
    }
    xfemm_mag *m = new xfemm_mag;
    m->lin = lin;
    m->axisymmetric = d->axisymmetric;
    m->n_nodes = d->n_nodes; m->n_elems = d->n_elems; m->n_bh = d->n_bh;
    m->ext_zo = d->ext_zo; m->ext_ro = d->ext_ro; m->ext_ri = d->ext_ri; m->r_eps = d->r_eps;
    cudaStream_t st = lin->stream;
    auto body = [&]() -> int {
        XB_CUDA(m->xy.upload(d->xy, (size_t)d->n_nodes * 2, st));
        XB_CUDA(m->tri.upload(d->tri, (size_t)d->n_elems * 4, st));
        if (d->magdir_elem) XB_CUDA(m->magdir.upload(d->magdir_elem, d->n_elems, st));
        if (d->n_bh > 0)
        {
            XB_ARG(d->bh_B && d->bh_H && d->bh_slope, "B-H tables");
            XB_CUDA(m->bhB.upload(d->bh_B, d->n_bh, st));
            XB_CUDA(m->bhH.upload(d->bh_H, d->n_bh, st));
            XB_CUDA(m->bhS.upload(d->bh_slope, d->n_bh, st));
        }
        m->n_mixed = d->n_mixed;
        if (d->n_mixed > 0)
        {
            XB_ARG(d->mixed_elem_side && d->mixed_c0c1, "mixed-BC arrays");
            XB_CUDA(m->mixed_es.upload(d->mixed_elem_side, (size_t)d->n_mixed * 2, st));
            XB_CUDA(m->mixed_c.upload(d->mixed_c0c1, (size_t)d->n_mixed * 2, st));
        }
        m->n_psrc = d->n_point_src;
        if (d->n_point_src > 0)
        {
            XB_ARG(d->point_src_node && d->point_src_val, "point source arrays");
            XB_CUDA(m->psrc_node.upload(d->point_src_node, d->n_point_src, st));
            XB_CUDA(m->psrc_val.upload(d->point_src_val, d->n_point_src, st));
        }
        XB_CUDA(m->mu.alloc((size_t)d->n_elems * 2));
        XB_CUDA(m->emap.alloc((size_t)d->n_elems * 9));
        XB_CUDA(m->err_flag.alloc(1));
        XB_CUDA(m->err_flag.zero(st));
        ++xb::g_launch, k_build_emap<<<(d->n_elems + 255) / 256, 256, 0, st>>>(d->n_elems, (const int4 *)m->tri.p, lin->rowptr.p,
                                                                lin->col.p, m->emap.p, m->err_flag.p);
        int flag = 0;
        XB_CUDA(cudaMemcpyAsync(&flag, m->err_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        XB_CUDA(cudaStreamSynchronize(st));
        XB_CUDA(cudaGetLastError());
        if (flag) { set_error("the linear problem's pattern does not contain the mesh graph"); return XFEMM_ERR_ARG; }
        return XFEMM_OK;
    };
    int rc = body();
    if (rc == XFEMM_OK) rc = xfemm_mag_set_labels(m, d->n_labels, d->labels);
    if (rc == XFEMM_OK)
        rc = xfemm_mag_set_age(m, d->n_age_cells, d->age_nodes, d->age_weights, d->age_matrix_id, d->n_age_matrices,
                               d->age_matrices);
    if (rc != XFEMM_OK) { delete m; return rc; }
    *out = m;
    return XFEMM_OK;
}

int xfemm_mag_destroy(xfemm_mag *m)
{
    if (!m) return XFEMM_OK;
    // the bound linear problem may already be gone (it is owned by the caller's CBigLinProb);
    // cudaFree in the DevBuf destructors synchronises on its own
    delete m;
    return XFEMM_OK;
}

int xfemm_mag_assemble_static(xfemm_mag *m, int iter, double *elapsed_ms)
{
    return xfemm_mag_assemble_static_parts(m, iter, 3, elapsed_ms);
}

// parts: bit 0 = wipe + elements + mixed sides + point sources (everything that does not depend on the rotor
// position), bit 1 = the air-gap element blocks, added to whatever the value array holds.  A sweep on a linear
// problem assembles and boundary-conditions bit 0 once, keeps a snapshot (xfemm_linprob_snapshot) and afterwards
// only restores it and adds bit 1 for each position.
// contributor lists of the deterministic assembly: once per mesh
static int det_prepare(xfemm_mag *m)
{
    if (m->det_ptr.n) return XFEMM_OK;
    xfemm_linprob *L = m->lin;
    cudaStream_t st = L->stream;
    const long long len = 12ll * m->n_elems, targets = L->nnz + m->n_nodes;
    XB_ARG(len < (1ll << 31) && targets < (1ll << 31), "mesh too large for the deterministic assembly lists");
    xb::DevBuf<int> keys, keys2, idx, cnt;
    xb::DevBuf<unsigned char> tmp;
    XB_CUDA(keys.alloc(len)); XB_CUDA(keys2.alloc(len)); XB_CUDA(idx.alloc(len));
    XB_CUDA(m->det_src.alloc(len));
    XB_CUDA(cnt.alloc(targets + 1)); XB_CUDA(cnt.zero(st));
    XB_CUDA(m->det_ptr.alloc(targets + 1));
    XB_CUDA(m->det_scratch.alloc(len));
    ++xb::g_launch, xb::k_det_keys<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(m->n_elems, L->nnz, m->emap.p, (const int4 *)m->tri.p, keys.p, idx.p);
    int end_bit = 1;
    while ((1ll << end_bit) <= targets && end_bit < 31) end_bit++;
    size_t bytes = 0;
    XB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys2.p, idx.p, m->det_src.p, (int)len, 0, end_bit, st));
    XB_CUDA(tmp.alloc(bytes));
    XB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, keys.p, keys2.p, idx.p, m->det_src.p, (int)len, 0, end_bit, st)); // stable
    ++xb::g_launch, xb::k_det_count<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(len, keys.p, cnt.p);
    size_t bytes2 = 0;
    XB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes2, cnt.p, m->det_ptr.p, (int)(targets + 1), st));
    if (bytes2 > tmp.n) XB_CUDA(tmp.alloc(bytes2));
    XB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes2, cnt.p, m->det_ptr.p, (int)(targets + 1), st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

int xfemm_mag_set_deterministic(xfemm_mag *m, int on)
{
    XB_ARG(m, "null handle");
    m->deterministic = on ? 1 : 0;
    if (m->lin) m->lin->deterministic = m->deterministic; // the multigrid cycle then keeps its ordered restriction pass
    return XFEMM_OK;
}

int xfemm_mag_assemble_static_parts(xfemm_mag *m, int iter, int parts, double *elapsed_ms)
{
    XB_ARG(m && m->lin, "null handle");
    XB_ARG(parts >= 1 && parts <= 3, "parts is a mask of 1 (position-independent part) and 2 (air-gap blocks)");
    xfemm_linprob *L = m->lin;
    cudaStream_t st = L->stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (elapsed_ms) { XB_CUDA(cudaEventCreate(&e0)); XB_CUDA(cudaEventCreate(&e1)); XB_CUDA(cudaEventRecord(e0, st)); }
    // Wipe (static2d.cpp:186-189); on iter 0 the caller hands over a fresh problem, wiping is harmless
    if (parts & 1)
    {
        XB_CUDA(L->vals.zero(st));
        XB_CUDA(L->b.zero(st));
    }
    XB_CUDA(m->err_flag.zero(st));
    const int det = m->deterministic;
    if (det && (parts & 1))
    {
        int rc = det_prepare(m);
        if (rc != XFEMM_OK) return rc; // (no zeroing: every cell of the scratch array is stored by the element kernel)
    }
    // element contributions: straight into the CSR slots (FP64 atomics), or into the scratch cells of the
    // deterministic mode (vals = b = scratch, see the kernels)
    double *ev = det ? m->det_scratch.p : L->vals.p, *eb = det ? m->det_scratch.p : L->b.p;
    if (!det && (parts & 2) && m->n_age_cells > 0)
    {
        int nt = m->n_age_cells * 55;
        ++xb::g_launch, k_age_scatter<<<(nt + 127) / 128, 128, 0, st>>>(m->n_age_cells, m->age_nodes.p, m->age_w.p, m->age_mid.p,
                                                        m->age_mats.p, L->rowptr.p, L->col.p, L->vals.p, 1.0, m->err_flag.p);
    }
    if (!(parts & 1)) { /* air-gap blocks only */ }
    else if (m->axisymmetric)
        ++xb::g_launch, k_assemble_staticaxi<<<(m->n_elems + 255) / 256, 256, 0, st>>>(
            m->n_elems, iter, (const int4 *)m->tri.p, (const double2 *)m->xy.p, m->emap.p, m->labels.p,
            m->magdir.n ? m->magdir.p : nullptr, m->bhB.p, m->bhH.p, m->bhS.p, L->V.p, ev, eb,
            (double2 *)m->mu.p, m->ext_zo, m->ext_ro, m->ext_ri, det);
    else
        ++xb::g_launch, k_assemble_static2d<<<(m->n_elems + 255) / 256, 256, 0, st>>>(
            m->n_elems, iter, (const int4 *)m->tri.p, (const double2 *)m->xy.p, m->emap.p, m->labels.p,
            m->magdir.n ? m->magdir.p : nullptr, m->bhB.p, m->bhH.p, m->bhS.p, L->V.p, ev, eb, (double2 *)m->mu.p, det);
    if ((parts & 1) && m->n_mixed > 0)
        ++xb::g_launch, k_mixed_edges<<<(m->n_mixed + 127) / 128, 128, 0, st>>>(m->n_mixed, m->n_elems, m->mixed_es.p, m->mixed_c.p,
                                                                (const int4 *)m->tri.p, (const double2 *)m->xy.p,
                                                                m->emap.p, ev, eb, m->axisymmetric, det);
    if (det && (parts & 1))
    {
        const long long targets = L->nnz + m->n_nodes;
        ++xb::g_launch, k_det_gather<<<(unsigned)((targets + 255) / 256), 256, 0, st>>>(L->nnz, m->n_nodes, m->det_ptr.p, m->det_src.p,
                                                                                       m->det_scratch.p, L->vals.p, L->b.p);
    }
    if (det && (parts & 2) && m->n_age_cells > 0)
        ++xb::g_launch, k_age_scatter_ordered<<<1, 64, 0, st>>>(m->n_age_cells, m->age_nodes.p, m->age_w.p, m->age_mid.p, m->age_mats.p,
                                                                 L->rowptr.p, L->col.p, L->vals.p, 1.0, m->err_flag.p);
    if ((parts & 1) && m->n_psrc > 0)
        ++xb::g_launch, k_point_sources<<<(m->n_psrc + 127) / 128, 128, 0, st>>>(m->n_psrc, m->psrc_node.p, m->psrc_val.p, L->b.p);
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
    if (flag) { set_error("air-gap element couples nodes that are missing from the pattern"); return XFEMM_ERR_ARG; }
    return XFEMM_OK;
}

int xfemm_mag_get_mu(xfemm_mag *m, double *mu12)
{
    XB_ARG(m && mu12, "null pointer");
    XB_CUDA(xb::copy_d2h(mu12, m->mu.p, (size_t)m->n_elems * 2 * sizeof(double), m->lin->stream));
    XB_CUDA(cudaStreamSynchronize(m->lin->stream));
    return XFEMM_OK;
}

int xfemm_mag_gap_torque(xfemm_mag *m, int cell0, int ncells, double inner_shift, double outer_shift, double arc_step_rad,
                         double r_mid_m, double dr_m, int bdry_format, int m_harm, double depth_m, double *torque,
                         double *br_bt /* optional, 2*ncells */)
{
    XB_ARG(m && torque && ncells > 0 && cell0 >= 0 && cell0 + ncells <= m->n_age_cells, "air-gap cell range");
    XB_ARG(!m->axisymmetric, "gap torque is defined for planar problems");
    xfemm_linprob *L = m->lin;
    cudaStream_t st = L->stream;
    const int nharm = bdry_format == 0 ? ncells / 2 + 1 : (ncells + 1) / 2;
    const double muo = 1.2566370614359173e-6, PIc = 3.141592653589793238462643383;
    const double scale = PIc * r_mid_m * r_mid_m * depth_m / muo;
    if (m->gap_out.n != (size_t)1 + 2 * ncells) XB_CUDA(m->gap_out.alloc((size_t)1 + 2 * ncells));
    const int threads = 128;
    const size_t smem = ((size_t)2 * ncells + threads / 32) * sizeof(double);
    XB_ARG(smem <= 48 * 1024, "air-gap element with too many cells");
    ++xb::g_launch, k_gap_torque<<<1, threads, smem, st>>>(cell0, ncells, m->age_nodes.p, m->age_w.p, L->V.p, PIc * 4.e-05, inner_shift,
                                                          outer_shift, arc_step_rad, r_mid_m, dr_m, bdry_format, m_harm, nharm, scale,
                                                          m->gap_out.p);
    std::vector<double> h((size_t)1 + 2 * ncells);
    xb::g_d2h_bytes += (long long)h.size() * 8;
    XB_CUDA(cudaMemcpyAsync(h.data(), m->gap_out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    *torque = h[0];
    if (br_bt) std::copy(h.begin() + 1, h.end(), br_bt);
    return XFEMM_OK;
}

int xfemm_mag_label_integrals(xfemm_mag *m, int n_labels, const xfemm_integral_label *labels, double *energy_area, double *a_area)
{
    XB_ARG(m && labels && energy_area && a_area, "null pointer");
    XB_ARG(!m->axisymmetric, "block integrals on the device are provided for planar problems (axisymmetric: write the .ans and use the post-processor)");
    XB_ARG(n_labels >= 1 && n_labels <= xb::INTEGRAL_MAX_LABELS, "between 1 and 32 block labels");
    xfemm_linprob *L = m->lin;
    cudaStream_t st = L->stream;
    xb::IntegralLabels T;
    T.n = n_labels;
    for (int l = 0; l < n_labels; l++) T.l[l] = labels[l];
    const int nblk = std::max(1, std::min((m->n_elems + 255) / 256, L->num_sms * 4));
    xb::DevBuf<double> part, out;
    xb::DevBuf<int> flag;
    XB_CUDA(part.alloc((size_t)nblk * 2 * n_labels));
    XB_CUDA(out.alloc((size_t)2 * n_labels));
    XB_CUDA(flag.alloc(1));
    XB_CUDA(flag.zero(st));
    const double PIc = 3.141592653589793238462643383;
    ++xb::g_launch;
    xb::k_label_integrals<<<nblk, 256, 0, st>>>(m->n_elems, (const int4 *)m->tri.p, (const double2 *)m->xy.p, L->V.p, PIc * 4.e-05,
                                                m->magdir.n ? m->magdir.p : nullptr, m->bhB.p, m->bhH.p, m->bhS.p, T, part.p, flag.p);
    ++xb::g_launch;
    xb::k_label_integrals_final<<<1, 64, 0, st>>>(nblk, 2 * n_labels, part.p, out.p);
    std::vector<double> h((size_t)2 * n_labels);
    int unsupported = 0;
    xb::g_d2h_bytes += (long long)h.size() * 8 + 4;
    XB_CUDA(cudaMemcpyAsync(h.data(), out.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaMemcpyAsync(&unsupported, flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    if (unsupported)
    {
        xb::set_error(unsupported == 1 ? "stored energy of wound regions (LamType > 2) is not evaluated on the device"
                                       : "stored energy of non-linear permanent magnets is not evaluated on the device");
        return XFEMM_ERR_ARG;
    }
    for (int l = 0; l < n_labels; l++) { energy_area[l] = h[l]; a_area[l] = h[(size_t)n_labels + l]; }
    return XFEMM_OK;
}

int xfemm_mag_solution_to_A(xfemm_mag *m, double *A_host)
{
    XB_ARG(m && A_host, "null pointer");
    xfemm_linprob *L = m->lin;
    cudaStream_t st = L->stream;
    // reference: L.b[i] = L.V[i]*c — the RHS array doubles as the result array (static2d.cpp:1018-1021)
    ++xb::g_launch;
    xb::g_d2h_bytes += (long long)m->n_nodes * 8;
    k_scale_to_A<<<(m->n_nodes + 255) / 256, 256, 0, st>>>(m->n_nodes, L->V.p, (const double2 *)m->xy.p, m->axisymmetric, L->b.p);
    XB_CUDA(xb::copy_d2h(A_host, L->b.p, (size_t)m->n_nodes * sizeof(double), st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    return XFEMM_OK;
}

} // extern "C"
