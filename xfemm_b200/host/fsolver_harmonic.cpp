// FSolver: host orchestration of the time-harmonic formulations (complex-symmetric systems).
// Reference behaviour mirrored: FSolver::Harmonic2D (cfemm/fsolver/harmonic2d.cpp:38-890) and
// FSolver::HarmonicAxisymmetric (cfemm/fsolver/harmonicaxi.cpp:34-821), with [ACSolver]=0
// (successive approximation) or 1 (Newton: side matrices + KludgeSolve, csrc/cnewton.cu).  All O(elements)/O(nnz) arithmetic runs on the GPU (xfemm_cmag_*,
// xfemm_clinprob_*); this file prepares the per-label / per-side tables and drives the outer loop.
#include "fsolver.h"
#include "../../include/xfemm_b200.h"
#include "../csrc/layout.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <set>

using namespace femm;
using namespace femmsolver;
typedef std::complex<double> cplx;

namespace xbhost
{
void flatten_age(const std::vector<CAirGapElement> &agelist, std::vector<int> &nodes, std::vector<double> &w,
                 std::vector<int> &mid, std::vector<double> &mats, bool harmonicWeights);
}

static double hwall() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int FSolver::harmonicSolve(CBigComplexLinProb &L, bool axi, bool verbose)
{
    const double tEnter = hwall();
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    const double c = PI * 4.e-05;
    static const double units[] = {2.54, 0.1, 1., 100., 0.00254, 1.e-04};
    const cplx I1(0, 1), deg45(1, 1);
    const double w = Frequency * 2. * PI;
    stats = FSolverStats();
    releaseDevice();
    if (xfemm_device_count() <= 0)
    {
        WarnMessage("xfemm_b200: no CUDA device available (this backend has no CPU solve path)\n");
        return false;
    }
    // one pass of the host threads over the element records: on-edge laminations (not supported in AC analyses) and the
    // elements that sit in a circuit — the serial sums below visit those alone, in element order
    std::vector<int> inCircuit;
    {
        std::vector<std::vector<int>> part(64);
        std::vector<char> onEdge(64, 0);
        xb::parallel_for(NumEls, [&](int i0, int i1, int tid) {
            for (int i = i0; i < i1; i++)
            {
                const int lt = blockproplist[meshele[i].blk].LamType;
                if (lt == 1 || lt == 2) onEdge[tid] = 1;
                if (NumCircProps > 0 && meshele[i].lbl >= 0 && labellist[meshele[i].lbl].InCircuit != -1) part[tid].push_back(i);
            }
        });
        for (char f : onEdge)
            if (f) { WarnMessage("On-edge lamination not supported in AC analyses"); return false; }
        for (auto &v : part) inCircuit.insert(inCircuit.end(), v.begin(), v.end()); // (chunks are contiguous and ascending)
    }
    for (int i = 0; i < NumBlockLabels; i++) GetFillFactor(i);
    for (int i = 0; i < NumBlockLabels; i++)
        if (!labellist[i].MagDirFctn.empty()) { /* magnetisation does not enter the harmonic formulation */ }

    // ---- circuits (harmonic2d.cpp:95-170 / harmonicaxi.cpp:83-140)
    if (NumCircProps > 0)
    {
        std::vector<cplx> C1(NumCircProps, 0.), C2(NumCircProps, 0.), C3(NumCircProps, 0.);
        for (int i : inCircuit)
        {
            const CMElement &El = meshele[i];
            const int ic = labellist[El.lbl].InCircuit;
            const int *n = El.p;
            double p0 = meshnode[n[1]].y - meshnode[n[2]].y, p1 = meshnode[n[2]].y - meshnode[n[0]].y;
            double q0 = meshnode[n[2]].x - meshnode[n[1]].x, q1 = meshnode[n[0]].x - meshnode[n[2]].x;
            double a = (p0 * q1 - p1 * q0) / 2.;
            double r = (meshnode[n[0]].x + meshnode[n[1]].x + meshnode[n[2]].x) / 3.;
            double Cduct = blockproplist[El.blk].Cduct;
            if (labellist[El.lbl].bIsWound) Cduct = 0;
            C1[ic] += a;
            C2[ic] += axi ? a * Cduct / (0.01 * r) : a * Cduct;
            C3[ic] += cplx(blockproplist[El.blk].J.re, blockproplist[El.blk].J.im) * a * 100.;
        }
        for (int i = 0; i < NumCircProps; i++)
        {
            CMCircuit &cp = circproplist[i];
            if (cp.CircType == 0)
            {
                if (C2[i] == 0.)
                {
                    cp.Case = 1;
                    cp.J = (C1[i] == 0.) ? CComplex(0.) : CComplex(0.01 * (cplx(cp.Amps.re, cp.Amps.im) - C3[i]) / C1[i]);
                }
                else cp.Case = 2;
            }
            else
            {
                cp.Case = 0;
                cp.dV = CComplex(cp.dVolts.re, cp.dVolts.im);
            }
        }
    }

    // ---- effective complex permeability per block (harmonic2d.cpp:172-215)
    std::vector<cplx> Mu0(NumBlockProps), Mu1(NumBlockProps);
    for (int k = 0; k < NumBlockProps; k++)
    {
        const CMSolverMaterialProp &bp = blockproplist[k];
        if (bp.LamType == 0)
        {
            Mu0[k] = bp.mu_x * std::exp(-I1 * bp.Theta_hx * DEG);
            Mu1[k] = bp.mu_y * std::exp(-I1 * bp.Theta_hy * DEG);
            if (bp.Lam_d != 0)
            {
                if (bp.Cduct != 0)
                {
                    cplx halflag = std::exp(-I1 * bp.Theta_hx * DEG / 2.);
                    double ds = sqrt(2. / (0.4 * PI * w * bp.Cduct * bp.mu_x));
                    cplx K = halflag * deg45 * bp.Lam_d * 0.001 / (2. * ds);
                    Mu0[k] = ((Mu0[k] * std::tanh(K)) / K) * bp.LamFill + (1. - bp.LamFill);
                    halflag = std::exp(-I1 * bp.Theta_hy * DEG / 2.);
                    ds = sqrt(2. / (0.4 * PI * w * bp.Cduct * bp.mu_y));
                    K = halflag * deg45 * bp.Lam_d * 0.001 / (2. * ds);
                    Mu1[k] = ((Mu1[k] * std::tanh(K)) / K) * bp.LamFill + (1. - bp.LamFill);
                }
                else
                {
                    Mu0[k] = Mu0[k] * bp.LamFill + (1. - bp.LamFill);
                    Mu1[k] = Mu1[k] * bp.LamFill + (1. - bp.LamFill);
                }
            }
        }
        else { Mu0[k] = 1; Mu1[k] = 1; }
    }

    // ---- per-label table + B-H tables
    std::vector<xfemm_cmag_label> labels(NumBlockLabels);
    std::vector<double> bhB, bhHr, bhHi, bhSr, bhSi;
    std::vector<int> bhOffset(NumBlockProps, -1);
    for (int k = 0; k < NumBlockProps; k++)
    {
        const CMSolverMaterialProp &bp = blockproplist[k];
        if (bp.BHpoints <= 0) continue;
        bhOffset[k] = (int)bhB.size();
        for (int i = 0; i < bp.BHpoints; i++)
        {
            bhB.push_back(bp.Bdata[i]);
            bhHr.push_back(bp.Hdata[i].re); bhHi.push_back(bp.Hdata[i].im);
            bhSr.push_back(bp.slope[i].re); bhSi.push_back(bp.slope[i].im);
        }
    }
    std::vector<char> labelNonlinear(NumBlockLabels, 0);
    for (int l = 0; l < NumBlockLabels; l++)
    {
        xfemm_cmag_label &T = labels[l];
        memset(&T, 0, sizeof T);
        T.circuit = -1; T.bh_offset = -1;
        T.mu1[0] = T.mu2[0] = 1;
        const CMBlockLabel &bl = labellist[l];
        if (bl.BlockType < 0 || bl.BlockType >= NumBlockProps) continue;
        const CMSolverMaterialProp &bp = blockproplist[bl.BlockType];
        cplx m1 = Mu0[bl.BlockType], m2 = Mu1[bl.BlockType];
        if (bp.LamType > 2) { m1 = cplx(bl.ProximityMu.re, bl.ProximityMu.im); m2 = m1; } // harmonic2d.cpp:663-668
        T.mu1[0] = m1.real(); T.mu1[1] = m1.imag(); T.mu2[0] = m2.real(); T.mu2[1] = m2.imag();
        cplx J(bp.J.re, bp.J.im), dVs(0, 0);
        if (bl.InCircuit >= 0)
        {
            const CMCircuit &cp = circproplist[bl.InCircuit];
            if (cp.Case == 1) J += cplx(cp.J.re, cp.J.im);
            if (cp.Case == 0)
            {
                if (axi) dVs = -100. * cplx(cp.dV.re, cp.dV.im) * bp.Cduct; // harmonicaxi.cpp: Jv = -100 dV sigma / R
                else J += -cplx(cp.dV.re, cp.dV.im) * bp.Cduct;
            }
            if (cp.Case == 2) T.circuit = bl.InCircuit;
        }
        T.J[0] = J.real(); T.J[1] = J.imag();
        T.dVs[0] = dVs.real(); T.dVs[1] = dVs.imag();
        T.sigma_eddy = bp.Cduct;
        if ((bp.LamType == 0) && (bp.Lam_d > 0)) T.sigma_eddy = 0; // in-plane laminations: eddy loss is in mu
        if (bl.bIsWound) T.sigma_eddy = 0;                          // wound coils have no bulk conductivity
        T.sigma_circ = bp.Cduct;
        T.bh_points = bp.BHpoints;
        T.bh_offset = bhOffset[bl.BlockType];
        T.lamtype = bp.LamType;
        T.is_external = bl.IsExternal ? 1 : 0;
        if (bp.BHpoints != 0) labelNonlinear[l] = 1;
    }
    bool LinearFlag = true; // (decided in the element pass below)

    // ---- element table, boundary sides, Dirichlet list, RHS terms, circuit couplings
    const int nUnknowns = NumNodes + NumCircProps;
    std::vector<int> tri(4 * (size_t)NumEls), sideES;
    std::vector<double> sideK;
    std::vector<int> dirNodes;
    std::vector<double> dirFirst, dirLast; // interleaved complex
    std::vector<int> dirSlot(NumNodes, -1);
    auto setValue = [&](int node, cplx x) {
        int &s = dirSlot[node];
        if (s < 0)
        {
            s = (int)dirNodes.size();
            dirNodes.push_back(node);
            dirFirst.push_back(x.real()); dirFirst.push_back(x.imag());
            dirLast.push_back(x.real()); dirLast.push_back(x.imag());
        }
        else { dirLast[2 * s] = x.real(); dirLast[2 * s + 1] = x.imag(); }
    };
    std::vector<int> rhsRow;
    std::vector<double> rhsVal;
    // nodes with a point property or (axisymmetric) on the axis: a handful among millions — found by the host threads,
    // visited in node order by the two passes below
    std::vector<int> special;
    {
        const double axisTol = units[LengthUnits] * 1.e-06;
        std::vector<std::vector<int>> part(64);
        xb::parallel_for(NumNodes, [&](int i0, int i1, int tid) {
            for (int i = i0; i < i1; i++)
                if (meshnode[i].BoundaryMarker >= 0 || (axi && meshnode[i].x < axisTol)) part[tid].push_back(i);
        });
        for (auto &v : part) special.insert(special.end(), v.begin(), v.end());
    }
    for (int i : special)
        if (meshnode[i].BoundaryMarker >= 0)
        {
            const CMPointProp &pp = nodeproplist[meshnode[i].BoundaryMarker];
            cplx K = axi ? (2. * meshnode[i].x * 0.01) * cplx(pp.J.re, pp.J.im) : 0.01 * cplx(pp.J.re, pp.J.im);
            rhsRow.push_back(i);
            rhsVal.push_back(-K.real()); rhsVal.push_back(-K.imag());
        }
    std::vector<int> fixRows;
    for (int i = 0; i < NumCircProps; i++)
    {
        if (circproplist[i].Case == 2)
        {
            cplx K = (axi ? 2. * 0.01 : 0.01) * cplx(circproplist[i].Amps.re, circproplist[i].Amps.im);
            rhsRow.push_back(NumNodes + i);
            rhsVal.push_back(K.real()); rhsVal.push_back(K.imag());
        }
        else fixRows.push_back(NumNodes + i);
    }
    for (int i : special)
    {
        if (axi && meshnode[i].x < (units[LengthUnits] * 1.e-06)) { setValue(i, 0.); continue; } // harmonicaxi.cpp:648-653
        if (meshnode[i].BoundaryMarker >= 0)
        {
            const CMPointProp &pp = nodeproplist[meshnode[i].BoundaryMarker];
            if ((pp.J.re == 0) && (pp.J.im == 0)) setValue(i, cplx(pp.A.re, pp.A.im) / c);
        }
    }
    // second pass of the host threads: the flat element table for the device, the elements with a tagged side, the
    // (node, circuit unknown) couplings of the elements whose circuit carries an unknown, and whether any element sits
    // in a B-H material.  The tagged elements are then visited in element order (first/last semantics of SetValue).
    std::vector<std::pair<int, int>> circPairs; // sorted, unique
    std::vector<int> tagged;
    {
        std::vector<std::vector<int>> partTag(64);
        std::vector<std::vector<std::pair<int, int>>> partPairs(64);
        std::vector<char> nonlinear(64, 0);
        xb::parallel_for(NumEls, [&](int i0, int i1, int tid) {
            for (int i = i0; i < i1; i++)
            {
                const CMElement &El = meshele[i];
                int *t4 = &tri[4 * (size_t)i];
                t4[0] = El.p[0]; t4[1] = El.p[1]; t4[2] = El.p[2]; t4[3] = El.lbl;
                if (labelNonlinear[El.lbl]) nonlinear[tid] = 1;
                if (labels[El.lbl].circuit >= 0)
                    for (int j = 0; j < 3; j++) partPairs[tid].emplace_back(El.p[j], NumNodes + labels[El.lbl].circuit);
                if (El.e[0] >= 0 || El.e[1] >= 0 || El.e[2] >= 0) partTag[tid].push_back(i);
            }
            // a node is shared by ~6 elements: thin the thread's list before the merge
            std::sort(partPairs[tid].begin(), partPairs[tid].end());
            partPairs[tid].erase(std::unique(partPairs[tid].begin(), partPairs[tid].end()), partPairs[tid].end());
        });
        for (char f : nonlinear) if (f) LinearFlag = false;
        for (auto &v : partTag) tagged.insert(tagged.end(), v.begin(), v.end()); // (chunks are contiguous and ascending)
        for (auto &v : partPairs) circPairs.insert(circPairs.end(), v.begin(), v.end());
        std::sort(circPairs.begin(), circPairs.end());
        circPairs.erase(std::unique(circPairs.begin(), circPairs.end()), circPairs.end());
    }
    for (int i : tagged)
    {
        const CMElement &El = meshele[i];
        for (int j = 0; j < 3; j++)
        {
            if (El.e[j] < 0) continue;
            const CMBoundaryProp &bp = lineproplist[El.e[j]];
            const int k = (j + 1) % 3;
            if (bp.BdryFormat == 2 || bp.BdryFormat == 1)
            {
                sideES.push_back(i); sideES.push_back(j);
                double row[7] = {(double)bp.BdryFormat, bp.c0.re, bp.c0.im, bp.c1.re, bp.c1.im, 0, 0};
                if (bp.BdryFormat == 1)
                {
                    double ds = sqrt(2. / (0.4 * PI * w * bp.Sig * bp.Mu));
                    cplx K = deg45 / (-ds * bp.Mu * 100.);
                    row[5] = K.real(); row[6] = K.imag();
                }
                sideK.insert(sideK.end(), row, row + 7);
            }
            if (bp.BdryFormat == 0)
            {
                for (int node : {El.p[j], El.p[k]})
                {
                    double x = meshnode[node].x, y = meshnode[node].y, a;
                    if (Coords == CART)
                    {
                        x /= units[LengthUnits];
                        y /= units[LengthUnits];
                        a = bp.A0 + x * bp.A1 + y * bp.A2;
                    }
                    else
                    {
                        double r = sqrt(x * x + y * y), th;
                        if ((x == 0) && (y == 0)) th = 0; else th = atan2(y, x) / DEG;
                        r /= units[LengthUnits];
                        a = bp.A0 + r * bp.A1 + th * bp.A2;
                    }
                    setValue(node, (a / c) * std::exp(I1 * bp.phi * DEG));
                }
            }
        }
    }
    std::vector<int> ageNodes, ageMid;
    std::vector<double> ageW, ageMats;
    if (!axi) xbhost::flatten_age(agelist, ageNodes, ageW, ageMid, ageMats, true);
    std::vector<int> pbcI(NumPBCs), pbcJ(NumPBCs), pbcT(NumPBCs);
    for (int k = 0; k < NumPBCs; k++) { pbcI[k] = pbclist[k].x; pbcJ[k] = pbclist[k].y; pbcT[k] = pbclist[k].t; }
    std::vector<int> exI, exJ;
    for (auto &pr : circPairs) { exI.push_back(pr.first); exJ.push_back(pr.second); }
    // Newton AC: Periodicity/AntiPeriodicity on the side matrices write the entry coupling the two nodes of a pair
    // (cspars.cpp:664-667, 748-751)
    if (ACSolver == 1)
        for (int k = 0; k < NumPBCs; k++) { exI.push_back(pbcI[k]); exJ.push_back(pbcJ[k]); }

    // ---- pattern → device problem
    double t0 = hwall();
    if (timing) fprintf(stderr, "[xfemm timing]   host tables before the pattern %.3f s\n", t0 - tEnter);
    int *rowptr = nullptr, *colidx = nullptr;
    unsigned char *fillmask = nullptr;
    if (xfemm_pattern_build_device(nUnknowns, NumEls, tri.data(), 4, (int)ageMid.size(), ageNodes.data(), (int)exI.size(), exI.data(),
                            exJ.data(), NumPBCs, pbcI.data(), pbcJ.data(), &rowptr, &colidx, keepSystem ? &fillmask : nullptr) != XFEMM_OK)
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    stats.t_pattern = hwall() - t0;
    t0 = hwall();
    xfemm_clinprob *lin = nullptr;
    int rc = xfemm_clinprob_create(nUnknowns, NumNodes, rowptr, colidx, &lin);
    std::vector<int> hRowptr, hCol;
    if (rc == XFEMM_OK)
    {
        L.attachDevice(lin);
        if (keepSystem) { hRowptr.assign(rowptr, rowptr + nUnknowns + 1); hCol.assign(colidx, colidx + rowptr[nUnknowns]); }
    }
    if (rc == XFEMM_OK && fillmask) rc = xfemm_clinprob_set_fill_mask(lin, fillmask);
    xfemm_free(fillmask);
    xfemm_free(rowptr);
    xfemm_free(colidx);
    if (rc != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
    // the solver structure (layout, real companion problem, multigrid hierarchy) only needs the pattern: a host
    // thread builds it while the tables are uploaded and the first system is assembled
    if (!getenv("XFEMM_NO_PREBUILD"))
    {
        joinPrebuild();
        const int pc = precondOption >= 0 ? precondOption : L.precond;
        prebuild = std::thread([lin, pc]() { xfemm_clinprob_prebuild(lin, pc); });
    }
    struct JoinOnExit { FSolver *s; ~JoinOnExit() { s->joinPrebuild(); } } joinOnExit{this}; // never outlive `lin`

    std::vector<double> xy(2 * (size_t)NumNodes);
    xb::parallel_for(NumNodes, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++) { xy[2 * (size_t)i] = meshnode[i].x; xy[2 * (size_t)i + 1] = meshnode[i].y; }
    });
    xfemm_cmag_desc D;
    memset(&D, 0, sizeof D);
    D.axisymmetric = axi ? 1 : 0;
    D.n_nodes = NumNodes; D.n_elems = NumEls; D.omega = w;
    D.xy = xy.data(); D.tri = tri.data();
    D.n_labels = NumBlockLabels; D.labels = labels.data();
    D.n_bh = (int)bhB.size(); D.bh_B = bhB.data(); D.bh_H = bhHr.data(); D.bh_H_im = bhHi.data();
    D.bh_slope = bhSr.data(); D.bh_slope_im = bhSi.data();
    D.n_sides = (int)(sideES.size() / 2); D.side_elem_side = sideES.data(); D.side_coeffs = sideK.data();
    D.n_age_cells = (int)ageMid.size(); D.age_nodes = ageNodes.data(); D.age_weights = ageW.data();
    D.age_matrix_id = ageMid.data(); D.n_age_matrices = (int)(ageMats.size() / 100); D.age_matrices = ageMats.data();
    D.n_rhs_add = (int)rhsRow.size(); D.rhs_add_row = rhsRow.data(); D.rhs_add_val = rhsVal.data();
    D.n_fix_diag = (int)fixRows.size(); D.fix_diag_rows = fixRows.data();
    D.ext_zo = extZo * units[LengthUnits]; D.ext_ro = extRo * units[LengthUnits]; D.ext_ri = extRi * units[LengthUnits];
    if (xfemm_cmag_create(&D, lin, &cmagdev) != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
    stats.t_upload = hwall() - t0;

    const int precond = precondOption >= 0 ? precondOption : L.precond;
    bool anyBH = false; // an element whose permeability follows a B-H curve (LamType 0): the only source of side-matrix entries
    for (int i = 0; i < NumEls && !anyBH; i++)
        anyBH = blockproplist[meshele[i].blk].LamType == 0 && blockproplist[meshele[i].blk].BHpoints > 0;
    if (ACSolver == 1 && (xfemm_clinprob_enable_newton(lin) != XFEMM_OK || xfemm_cmag_set_newton(cmagdev, 1) != XFEMM_OK))
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    int Iter = 0;
    double res = 0, lastres = 0;
    Relax = 1.;
    const double tLoop = hwall();
    do
    {
        if (verbose) PrintMessage("Matrix Construction\n");
        double ms_asm = 0;
        if (xfemm_cmag_assemble(cmagdev, Iter, &ms_asm) != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
        double tb = hwall();
        if (xfemm_clinprob_dirichlet(lin, (int)dirNodes.size(), dirNodes.data(), dirFirst.data(), dirLast.data(), L.bdw) != XFEMM_OK ||
            xfemm_cmag_fix_circuit_diagonals(cmagdev) != XFEMM_OK ||
            xfemm_clinprob_periodic(lin, NumPBCs, pbcI.data(), pbcJ.data(), pbcT.data()) != XFEMM_OK)
        {
            WarnMessage(xfemm_last_error());
            return false;
        }
        stats.assemble_ms.push_back(ms_asm + (hwall() - tb) * 1e3);
        if (keepSystem)
        {
            const size_t nnz0 = hCol.size();
            std::vector<double> v0(2 * nnz0);
            std::vector<unsigned char> ex(nnz0);
            sysRhs.resize(2 * (size_t)nUnknowns);
            xfemm_clinprob_get_values(lin, v0.data());
            xfemm_clinprob_get_rhs(lin, sysRhs.data());
            xfemm_clinprob_get_exists_mask(lin, ex.data());
            sysRowptr.assign(nUnknowns + 1, 0);
            sysCol.clear();
            sysVals.clear();
            for (int i = 0; i < nUnknowns; i++)
            {
                for (int q = hRowptr[i]; q < hRowptr[i + 1]; q++)
                    if (ex[q]) { sysCol.push_back(hCol[q]); sysVals.push_back(v0[2 * (size_t)q]); sysVals.push_back(v0[2 * (size_t)q + 1]); }
                sysRowptr[i + 1] = (int)sysCol.size();
            }
            for (int h = 0; h < 3; h++)
            {
                sysSide[h].clear();
                if (ACSolver != 1) continue;
                xfemm_clinprob_get_side_values(lin, h + 1, v0.data());
                for (size_t q = 0; q < nnz0; q++)
                    if (ex[q]) { sysSide[h].push_back(v0[2 * q]); sysSide[h].push_back(v0[2 * q + 1]); }
            }
        }
        xfemm_clinprob_save_solution(lin);
        joinPrebuild(); // the solver structure is needed from here on (a failed pre-build is simply redone by the solve)
        xfemm_solve_info info;
        // XFEMM_PRECOND_AMG here = multigrid V-cycle of the real matrix |Re A| + |Im A| (stiffness + omega*sigma*mass)
        // applied to the real and the imaginary part of the residual; as a safety net an attempt that does not
        // converge within 3000 iterations is repeated with the Jacobi preconditioner
        // Newton AC ([ACSolver]=1): from the second pass on the system carries the side matrices (L.bNewton), the
        // linear tolerance follows the nonlinear residual (harmonic2d.cpp:821-825) and PBCGSolveMod dispatches to
        // KludgeSolve (cspars.cpp:1065-1067)
        const bool newtonPass = ACSolver == 1 && Iter > 0 && anyBH;
        if (newtonPass)
        {
            L.bNewton = 1;
            L.Precision = std::min(1.e-4, 0.001 * res);
            if (L.Precision < Precision) L.Precision = Precision;
            rc = xfemm_clinprob_kludge_solve(lin, L.Precision, Iter, precond, maxIterOption, &info);
        }
        else
        {
            rc = xfemm_clinprob_cocg(lin, L.Precision, Iter, precond,
                                     (precond == XFEMM_PRECOND_AMG && maxIterOption <= 0) ? 3000 : maxIterOption, &info);
            if (rc == XFEMM_ERR_DIVERGED && precond == XFEMM_PRECOND_AMG)
                rc = xfemm_clinprob_cocg(lin, L.Precision, 0, XFEMM_PRECOND_JACOBI, maxIterOption, &info);
        }
        stats.pcg_iters.push_back(info.iterations);
        stats.pcg_ms.push_back(info.solve_ms);
        stats.final_relres = info.relres;
        L.lastIterations = info.iterations; L.lastSolveMs = info.solve_ms; L.lastRelRes = info.relres;
        if (rc != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
        if (LinearFlag == false)
        {
            double x = 0, y = 0;
            xfemm_clinprob_newton_norms(lin, NumNodes, &x, &y);
            if (y == 0) LinearFlag = true;
            else { lastres = res; res = sqrt(x / y); }
            if (Iter > 5)
            {
                if ((res > lastres) && (Relax > 0.1)) Relax /= 2.;
                else Relax += 0.1 * (1. - Relax);
                xfemm_clinprob_relax(lin, Relax);
            }
            if (verbose)
            {
                char outstr[256];
                snprintf(outstr, sizeof outstr, ACSolver == 1 ? "Newton Iteration(%i) Relax=%.4g\n" : "Successive Approx(%i) Relax=%.4g\n", Iter, Relax);
                PrintMessage(outstr);
            }
        }
        if ((res < 100. * Precision) && Iter > 0) LinearFlag = true;
        Iter++;
    } while (LinearFlag == false);
    stats.newton_iters = Iter;
    const double tFetch = hwall();
    if (timing)
        fprintf(stderr, "[xfemm timing] pattern %.3f s, device create+upload %.3f s, %d pass(es) %.3f s (krylov %.1f ms)\n",
                stats.t_pattern, stats.t_upload, Iter, tFetch - tLoop, stats.pcg_ms.empty() ? 0. : stats.pcg_ms[0]);
    if (xfemm_clinprob_get_solution(lin, (double *)L.V) != XFEMM_OK || xfemm_cmag_solution_to_A(cmagdev, (double *)L.b) != XFEMM_OK)
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    if (timing) fprintf(stderr, "[xfemm timing]   result copies %.3f s, harmonic solve total %.3f s\n", hwall() - tFetch, hwall() - tEnter);
    // element permeabilities back to the mesh records (read by nothing in the writer, kept for callers)
    return true;
}
