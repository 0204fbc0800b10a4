// FSolver: renumbering, problem set-up and the Newton loop of the magnetostatic formulations.
// Host orchestration only — every O(elements) / O(nnz) arithmetic step runs on the GPU through
// the C ABI (include/xfemm_b200.h).  Reference behaviour mirrored:
//   FEASolver::Cuthill/SortElements   cfemm/libfemm/cuthill.cpp:40-391
//   FSolver::SortNodes                cfemm/fsolver/fsolver.cpp:1341-1353
//   FSolver::GetFillFactor, ElmArea   cfemm/fsolver/fsolver.cpp:1083-1211
//   FSolver::runSolver                cfemm/fsolver/fsolver.cpp:1213-1338
//   FSolver::Static2D                 cfemm/fsolver/static2d.cpp:53-1033
//   FSolver::StaticAxisymmetric       cfemm/fsolver/staticaxi.cpp:45-794
#include "fsolver.h"
#include <type_traits>
#include "magdir_expr.h"
#include "../../include/xfemm_b200.h"
#include "../csrc/layout.h"
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

using namespace femm;
using namespace femmsolver;
typedef std::complex<double> cplx;

namespace xb { void age_cell_matrix(double ci, double co, double K, double *MG); }

// debugging aid: XFEMM_DEBUG_DUMP=<dir> writes raw little-endian arrays per Newton step
static void debug_dump(const char *name, int step, const void *data, size_t bytes)
{
    const char *dir = getenv("XFEMM_DEBUG_DUMP");
    if (!dir) return;
    char path[1024];
    snprintf(path, sizeof path, "%s/step%d_%s.bin", dir, step, name);
    FILE *f = fopen(path, "wb");
    if (!f) return;
    fwrite(data, 1, bytes, f);
    fclose(f);
}

static double wall() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void FSolver::releaseDevice()
{
    joinPrebuild();
    solutionFetches = 0;
    if (dpcg) { xfemm_dpcg_destroy(dpcg); dpcg = nullptr; }
    if (magdev) { xfemm_mag_destroy(magdev); magdev = nullptr; }
    if (cmagdev) { xfemm_cmag_destroy(cmagdev); cmagdev = nullptr; }
}

// ---------------------------------------------------------------------------------------
// renumbering
// ---------------------------------------------------------------------------------------
void FSolver::SortNodes(const std::vector<int> &newnum)
{
    // the reference cycles swaps until every node sits at its new index; the result is
    // "node i moves to position newnum[i]"
    std::vector<CNode> out(meshnode.size());
    for (int i = 0; i < NumNodes; i++) out[newnum[i]] = meshnode[i];
    meshnode.swap(out);
}

int FSolver::SortElements()
{
    std::vector<int> tri(3 * (size_t)NumEls), order(NumEls);
    xb::parallel_for(NumEls, [&](int k0, int k1, int) {
        for (int k = k0; k < k1; k++)
            for (int j = 0; j < 3; j++) tri[3 * (size_t)k + j] = meshele[k].p[j];
    });
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    const double t0 = wall();
    if (xfemm_renumber_sort_elements(NumEls, tri.data(), order.data()) != XFEMM_OK) return false;
    if (timing) fprintf(stderr, "[xfemm timing]     SortElements: comb passes %.3f s, sizeof(CMElement) %zu\n", wall() - t0, sizeof(CMElement));
    if (std::is_trivially_copyable<CMElement>::value)
    {
        // gather into raw scratch memory (its pages are first touched by the threads that fill them; a value-initialised
        // vector of 4 M elements would be faulted in by one thread), then copy back in place
        CMElement *tmp = (CMElement *)malloc(sizeof(CMElement) * (size_t)std::max(NumEls, 1));
        if (!tmp) return false;
        xb::parallel_for(NumEls, [&](int k0, int k1, int) {
            for (int k = k0; k < k1; k++) memcpy((void *)(tmp + k), (const void *)&meshele[order[k]], sizeof(CMElement));
        });
        xb::parallel_for(NumEls, [&](int k0, int k1, int) {
            if (k1 > k0) memcpy((void *)&meshele[k0], (const void *)(tmp + k0), sizeof(CMElement) * (size_t)(k1 - k0));
        });
        free(tmp);
        if (timing) fprintf(stderr, "[xfemm timing]     SortElements: total %.3f s\n", wall() - t0);
        return true;
    }
    std::vector<CMElement> out(NumEls);
    xb::parallel_for(NumEls, [&](int k0, int k1, int) {
        for (int k = k0; k < k1; k++) out[k] = meshele[order[k]];
    });
    meshele.swap(out);
    return true;
}

int FSolver::Cuthill(bool deleteFiles)
{
    std::string edgeFile = PathName + ".edge";
    if (edgeList.empty())
    {
        // LoadMesh normally caches the edge list; fall back to the file like the reference
        FILE *fp = fopen(edgeFile.c_str(), "rt");
        if (!fp) { printf("Couldn't open %s", edgeFile.c_str()); return false; }
        long nl = 0, flag = 0;
        if (fscanf(fp, "%li", &nl) != 1 || fscanf(fp, "%li", &flag) != 1) { fclose(fp); return false; }
        edgeList.resize(2 * (size_t)nl);
        for (long i = 0; i < nl; i++)
        {
            long idx, mk;
            if (fscanf(fp, "%li %i %i %li", &idx, &edgeList[2 * i], &edgeList[2 * i + 1], &mk) != 4) { fclose(fp); return false; }
        }
        fclose(fp);
    }
    if (deleteFiles) remove(edgeFile.c_str());
    std::vector<int> newnum(NumNodes);
    int bw = 0;
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    double tt = wall();
    auto tick = [&](const char *what) { if (timing) { fprintf(stderr, "[xfemm timing]   Cuthill %s %.3f s\n", what, wall() - tt); tt = wall(); } };
    if (xfemm_renumber_cuthill(NumNodes, (int)(edgeList.size() / 2), edgeList.data(), newnum.data(), &bw) != XFEMM_OK)
        return false;
    tick("numbering");
    BandWidth = bw;
    for (int i = 0; i < NumPBCs; i++)
    {
        pbclist[i].x = newnum[pbclist[i].x];
        pbclist[i].y = newnum[pbclist[i].y];
    }
    for (auto &age : agelist)
        for (auto &q : age.quadNode)
        {
            q.n0 = newnum[q.n0]; q.n1 = newnum[q.n1]; q.n2 = newnum[q.n2]; q.n3 = newnum[q.n3];
        }
    xb::parallel_for(NumEls, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++)
            for (int j = 0; j < 3; j++) meshele[i].p[j] = newnum[meshele[i].p[j]];
    });
    tick("remap");
    SortNodes(newnum);
    cmPermutation = newnum;
    tick("SortNodes");
    SortElements();
    tick("SortElements");
    std::vector<int>().swap(edgeList);
    return true;
}

// ---------------------------------------------------------------------------------------
double FSolver::ElmArea(int i)
{
    const int *n = meshele[i].p;
    double b0 = meshnode[n[1]].y - meshnode[n[2]].y;
    double b1 = meshnode[n[2]].y - meshnode[n[0]].y;
    double c0 = meshnode[n[2]].x - meshnode[n[1]].x;
    double c1 = meshnode[n[0]].x - meshnode[n[2]].x;
    return 0.0001 * (b0 * c1 - b1 * c0) / 2.;
}

void FSolver::GetFillFactor(int lbl)
{
    CMSolverMaterialProp *bp = &blockproplist[labellist[lbl].BlockType];
    CMBlockLabel *bl = &labellist[lbl];
    bl->bIsWound = (abs(bl->Turns) > 1) || (bp->LamType > 2);
    if ((Frequency == 0) || (bp->LamType < 3)) { bl->ProximityMu = 1.; return; }

    double atot = 0;
    for (int i = 0; i < NumEls; i++)
        if (meshele[i].lbl == lbl) atot += ElmArea(i);
    if (atot == 0) return;
    if (bp->Cduct == 0) { bl->ProximityMu = 1; return; }
    const cplx I1(0, 1);
    int wiretype = bp->LamType - 3;
    double W, d, fill, dd, o, R = 0, awire = 0, c1 = 0, c2 = 0;
    if (wiretype == 3) // rectangular wire → equivalent foil
    {
        W = 2. * PI * Frequency;
        d = bp->WireD * 0.001;
        fill = fabs(d * d * ((double)bl->Turns) / atot);
        dd = d / sqrt(fill);
        fill = d / dd;
        o = bp->Cduct * (d / dd) * 1.e6;
        cplx ufd = muo * std::tanh(std::sqrt(I1 * W * o * muo) * d / 2.) / (std::sqrt(I1 * W * o * muo) * d / 2.);
        bl->ProximityMu = CComplex((fill * ufd + (1. - fill) * muo) / muo);
        return;
    }
    switch (wiretype)
    {
    case 0: R = bp->WireD * 0.0005; awire = PI * R * R * ((double)bp->NStrands) * ((double)bl->Turns); break;
    case 1: R = bp->WireD * 0.0005 * sqrt((double)bp->NStrands); awire = PI * R * R * ((double)bl->Turns); break;
    case 2: R = bp->WireD * 0.0005; awire = PI * R * R * ((double)bp->NStrands) * ((double)bl->Turns); break;
    }
    fill = fabs(awire / atot);
    o = bp->Cduct * 1.e6;
    W = 2. * PI * Frequency * o * muo * R * R / 2.;
    switch (wiretype)
    {
    case 0: case 1: case 2:
        c1 = 0.7756067409818643 + fill * (0.6873854335408803 + fill * (0.06841584481674128 - 0.07143732702512284 * fill));
        c2 = 1.5 * fill / c1;
        break;
    case 4:
        c1 = 0.7270741505617485 + 0.8902950067721367 * fill + 0.11894736885885195 * fill * fill - 0.12247276254503957 * fill * fill * fill;
        c2 = 0.006784920229549677 + 1.8942880489198526 * fill - 1.3631438759519217 * fill * fill + 0.504431701685587 * fill * fill * fill;
        break;
    case 5:
        c1 = 0.7486913529860821 + 0.9042845510838825 * fill + 0.1361040321433224 * fill * fill - 0.10652380745682069 * fill * fill * fill;
        c2 = 0.006790468527313965 + 1.8945509985370095 * fill - 1.3643501010185972 * fill * fill + 0.5036765577982594 * fill * fill * fill;
        break;
    }
    cplx ufd = c2 * (std::tanh(std::sqrt(c1 * I1 * W)) / std::sqrt(c1 * I1 * W)) + (1. - c2);
    bl->ProximityMu = CComplex(ufd);
}

// ---------------------------------------------------------------------------------------
bool FSolver::runSolver(bool verbose)
{
    double t0 = wall();
    LoadMeshErr err = LoadMesh();
    if (err != NOERROR) { WarnMessage(getErrorString(err).c_str()); return false; }
    stats.t_load_mesh = wall() - t0;
    if (previousSolutionFile.empty())
    {
        if (verbose) PrintMessage("renumbering nodes using Cuthill-McKee method\n");
        t0 = wall();
        if (!Cuthill()) { WarnMessage("problem renumbering node points\n"); return false; }
        stats.t_renumber = wall() - t0;
    }
    if (verbose)
    {
        PrintMessage("solving...\n");
        std::string st = "Problem Statistics:\n" + std::to_string(NumNodes) + " nodes\n" + std::to_string(NumEls) +
                         " elements\nPrecision: " + std::to_string(Precision) + "\n";
        PrintMessage(st.c_str());
    }
    if (Frequency == 0)
    {
        CBigLinProb L;
        L.Precision = Precision;
        if (L.Create(NumNodes, BandWidth) == false) { WarnMessage("couldn't allocate enough space for matrices\n"); return false; }
        if (ProblemType == PLANAR)
        {
            if (Static2D(L) != true) { WarnMessage("Couldn't solve the problem\n"); return false; }
            if (verbose) PrintMessage("Static 2-D problem solved\n");
        }
        else
        {
            if (StaticAxisymmetric(L) != true) { WarnMessage("Couldn't solve the problem"); return false; }
            if (verbose) PrintMessage("Static axisymmetric problem solved\n");
        }
        t0 = wall();
        if (WriteStatic2D(L) == false) { WarnMessage("couldn't write results to disk\n"); return false; }
        stats.t_write = wall() - t0;
        if (verbose) PrintMessage("results written to disk\n");
    }
    else
    {
        CBigComplexLinProb L;
        L.Precision = Precision;
        if (!L.Create(NumNodes + NumCircProps, BandWidth, NumNodes)) { WarnMessage("couldn't allocate enough space for matrices\n"); return false; }
        if (ProblemType == PLANAR)
        {
            if (!Harmonic2D(L, verbose)) { WarnMessage("Couldn't solve the problem\n"); return false; }
            if (verbose) PrintMessage("Harmonic 2-D problem solved\n");
        }
        else
        {
            if (!HarmonicAxisymmetric(L, verbose)) { WarnMessage("Couldn't solve the problem\n"); return false; }
            if (verbose) PrintMessage("Harmonic axisymmetric problem solved\n");
        }
        t0 = wall();
        if (!WriteHarmonic2D(L)) { WarnMessage("couldn't write results to disk\n"); return false; }
        stats.t_write = wall() - t0;
        if (verbose) PrintMessage("results written to disk.\n");
    }
    return true;
}

int FSolver::Static2D(CBigLinProb &L) { return staticSolve(L, false); }
int FSolver::StaticAxisymmetric(CBigLinProb &L) { return staticSolve(L, true); }

// ---------------------------------------------------------------------------------------
// shared set-up pieces
// ---------------------------------------------------------------------------------------
namespace xbhost
{
// Flattens the air-gap elements into ring cells: 10 nodes + 10 weights per cell and one 10x10
// matrix per AGE (static2d.cpp:192-343).
void flatten_age(const std::vector<CAirGapElement> &agelist, std::vector<int> &nodes, std::vector<double> &w,
                 std::vector<int> &mid, std::vector<double> &mats, bool harmonicWeights)
{
    nodes.clear(); w.clear(); mid.clear(); mats.clear();
    for (size_t a = 0; a < agelist.size(); a++)
    {
        const CAirGapElement &g = agelist[a];
        const int N = g.totalArcElements;
        double dt = (PI / 180.) * (g.totalArcLength / N);
        double K = 2. * (g.ro - g.ri) / (dt * (g.ro + g.ri));
        double ci = g.InnerShift, co = g.OuterShift;
        if (ci > co) { ci = ci - co; co = 0; }
        else { ci = 1 - co + ci; co = 1; }
        mats.resize(mats.size() + 100);
        xb::age_cell_matrix(ci, co, K, &mats[mats.size() - 100]);
        for (int k = 0; k < N; k++)
        {
            int nn[10];
            double ww[10];
            const CQuadPoint &qm = g.quadNode[(k - 1) < 0 ? N - 1 : k - 1];
            const CQuadPoint &q0 = g.quadNode[k];
            const CQuadPoint &q1 = g.quadNode[k + 1];
            const CQuadPoint &q2 = g.quadNode[(k + 2) > N ? 1 : k + 2];
            nn[0] = qm.n0; ww[0] = qm.w0;
            nn[1] = q0.n0; ww[1] = q0.w0;
            nn[2] = q0.n1; ww[2] = q0.w1;
            nn[3] = q1.n1; ww[3] = q1.w1;
            nn[4] = q2.n1; ww[4] = q2.w1;
            nn[5] = qm.n2; ww[5] = qm.w2;
            nn[6] = q0.n2; ww[6] = q0.w2;
            nn[7] = q0.n3; ww[7] = q0.w3;
            nn[8] = q1.n3; ww[8] = q1.w3;
            nn[9] = q2.n3; ww[9] = q2.w3;
            if (g.BdryFormat == 1)
            {
                if (k == 0) { ww[0] = -ww[0]; ww[5] = -ww[5]; }
                // static: last cell flips the far end (static2d.cpp:338); the harmonic code tests
                // k==totalArcElements, which never holds inside the loop (harmonic2d.cpp:373)
                if (!harmonicWeights && (k + 1) == N) { ww[4] = -ww[4]; ww[9] = -ww[9]; }
            }
            nodes.insert(nodes.end(), nn, nn + 10);
            w.insert(w.end(), ww, ww + 10);
            mid.push_back((int)a);
        }
    }
}
} // namespace xbhost
using xbhost::flatten_age;

// ---------------------------------------------------------------------------------------
// magnetostatics
// ---------------------------------------------------------------------------------------
int FSolver::staticSolve(CBigLinProb &L, bool axi)
{
    const double tEnter = wall();
    const double c = PI * 4.e-05;
    static const double units[] = {2.54, 0.1, 1., 100., 0.00254, 1.e-04};
    stats = FSolverStats();
    releaseDevice();
    if (xfemm_device_count() <= 0)
    {
        WarnMessage("xfemm_b200: no CUDA device available (this backend has no CPU solve path)\n");
        return false;
    }
    for (int i = 0; i < NumBlockLabels; i++) GetFillFactor(i);
    // functional magnetisation directions: compiled once per label, evaluated per element below
    // (the reference runs its Lua interpreter per element, static2d.cpp:513-583 / staticaxi.cpp:354-420)
    std::vector<xbhost::MagDirExpr> magExpr(NumBlockLabels);
    bool anyMagDirFctn = false;
    for (int i = 0; i < NumBlockLabels; i++)
        if (!labellist[i].MagDirFctn.empty())
        {
            std::string err;
            if (!magExpr[i].compile(labellist[i].MagDirFctn, err))
            {
                WarnMessage(("xfemm_b200: " + err + "\n").c_str());
                return -7;
            }
            anyMagDirFctn = true;
        }

    // ---- one pass of the host threads over the element records (96 B each, 380 MB at 4 M elements): the flat
    // element table for the device, the elements with a tagged side, the elements that sit in a circuit.  The serial
    // passes below visit only those lists, in element order.
    std::vector<int> tri(4 * (size_t)NumEls);
    std::vector<int> tagged, inCircuit;
    {
        std::vector<std::vector<int>> partTag(64), partCirc(64);
        xb::parallel_for(NumEls, [&](int i0, int i1, int tid) {
            for (int i = i0; i < i1; i++)
            {
                const CMElement &El = meshele[i];
                int *t4 = &tri[4 * (size_t)i];
                t4[0] = El.p[0]; t4[1] = El.p[1]; t4[2] = El.p[2]; t4[3] = El.lbl;
                if (El.e[0] >= 0 || El.e[1] >= 0 || El.e[2] >= 0) partTag[tid].push_back(i);
                if (NumCircProps > 0 && El.lbl >= 0 && labellist[El.lbl].InCircuit != -1) partCirc[tid].push_back(i);
            }
        });
        for (auto &v : partTag) tagged.insert(tagged.end(), v.begin(), v.end()); // (chunks are contiguous and ascending)
        for (auto &v : partCirc) inCircuit.insert(inCircuit.end(), v.begin(), v.end());
    }

    // ---- circuits (static2d.cpp:84-167): area, conductance and source integrals per circuit
    if (NumCircProps > 0)
    {
        std::vector<double> I1(NumCircProps, 0.), I2(NumCircProps, 0.), I3(NumCircProps, 0.);
        // (over the elements that belong to a circuit, in element order: the order fixes the rounding of the integrals)
        for (int i : inCircuit)
        {
            const CMElement &El = meshele[i];
            const int ic = labellist[El.lbl].InCircuit;
            const int *n = El.p;
            double p0 = meshnode[n[1]].y - meshnode[n[2]].y, p1 = meshnode[n[2]].y - meshnode[n[0]].y;
            double q0 = meshnode[n[2]].x - meshnode[n[1]].x, q1 = meshnode[n[0]].x - meshnode[n[2]].x;
            double a = (p0 * q1 - p1 * q0) / 2.;
            double Cduct = blockproplist[El.blk].Cduct;
            if (labellist[El.lbl].bIsWound) Cduct = 0;
            I1[ic] += a;
            // staticaxi.cpp:99-106: the conductance integral carries 100/r of the element centroid
            if (axi) I2[ic] += 100. * a * Cduct / ((meshnode[n[0]].x + meshnode[n[1]].x + meshnode[n[2]].x) / 3.);
            else I2[ic] += a * Cduct;
            I3[ic] += blockproplist[El.blk].J.re * a * 100.;
        }
        for (int i = 0; i < NumCircProps; i++)
        {
            CMCircuit &cp = circproplist[i];
            if (cp.CircType == 0)
            {
                if (I2[i] == 0)
                {
                    cp.Case = 1;
                    cp.J = (I1[i] == 0.) ? 0. : 0.01 * (cp.Amps.re - I3[i]) / I1[i];
                }
                else
                {
                    cp.Case = 0;
                    cp.dV = -0.01 * (cp.Amps.re - I3[i]) / I2[i];
                }
            }
            else
            {
                cp.Case = 0;
                cp.dV = cp.dVolts.re;
            }
        }
    }

    // ---- per-label table and flat B-H tables
    std::vector<xfemm_mag_label> labels(NumBlockLabels);
    std::vector<double> bhB, bhH, bhS;
    std::vector<int> bhOffset(NumBlockProps, -1);
    for (int k = 0; k < NumBlockProps; k++)
    {
        const CMSolverMaterialProp &bp = blockproplist[k];
        if (bp.BHpoints <= 0) continue;
        bhOffset[k] = (int)bhB.size();
        for (int i = 0; i < bp.BHpoints; i++)
        {
            bhB.push_back(bp.Bdata[i]);
            bhH.push_back(bp.Hdata[i].re);
            bhS.push_back(bp.slope[i].re);
        }
    }
    bool anyNonlinearLabel = false;
    std::vector<char> labelNonlinear(NumBlockLabels, 0);
    for (int l = 0; l < NumBlockLabels; l++)
    {
        xfemm_mag_label &T = labels[l];
        memset(&T, 0, sizeof T);
        const CMBlockLabel &bl = labellist[l];
        T.bh_offset = -1;
        T.circuit = -1;
        T.mu1 = T.mu2 = 1;
        if (bl.BlockType < 0 || bl.BlockType >= NumBlockProps) continue;
        const CMSolverMaterialProp &bp = blockproplist[bl.BlockType];
        // linear permeabilities of the first pass (static2d.cpp:603-631)
        double t = bp.LamFill, mu;
        T.mu1 = -1; T.mu2 = -1;
        if (!axi)
        {
            if (bp.LamType == 0) { T.mu1 = bp.mu_x * t + (1. - t); T.mu2 = bp.mu_y * t + (1. - t); }
            if (bp.LamType == 1) { mu = bp.mu_x; T.mu1 = mu * t + (1. - t); T.mu2 = mu / (t + mu * (1. - t)); }
            if (bp.LamType == 2) { mu = bp.mu_y; T.mu2 = mu * t + (1. - t); T.mu1 = mu / (t + mu * (1. - t)); }
        }
        else
        {
            // staticaxi.cpp:431-447 — literally: type 0 drops the (1-t) term, type 2 assigns like type 1
            if (bp.LamType == 0) { T.mu1 = bp.mu_x * t; T.mu2 = bp.mu_y * t; }
            if (bp.LamType == 1) { mu = bp.mu_x; T.mu1 = mu * t + (1. - t); T.mu2 = mu / (t + mu * (1. - t)); }
            if (bp.LamType == 2) { mu = bp.mu_y; T.mu1 = mu * t + (1. - t); T.mu2 = mu / (t + mu * (1. - t)); }
        }
        if (bp.LamType > 2) { T.mu1 = 1; T.mu2 = 1; }
        // source current density incl. the circuit contribution (static2d.cpp:483-504)
        double tc = 0;
        if (bl.InCircuit >= 0)
        {
            const CMCircuit &cp = circproplist[bl.InCircuit];
            if (cp.Case == 1) tc = cp.J.Re();
            if (cp.Case == 0)
            {
                if (axi) T.dVs = -100. * cp.dV.Re() * bp.Cduct; // divided by the element's R on the device
                else tc = -cp.dV.Re() * bp.Cduct;
            }
        }
        T.Jsrc = bp.J.re + tc;
        T.H_c = bp.H_c;
        T.magdir = bl.MagDir;
        T.lamfill = bp.LamFill;
        T.lamtype = bp.LamType;
        T.bh_points = bp.BHpoints;
        T.bh_offset = bhOffset[bl.BlockType];
        T.is_external = bl.IsExternal ? 1 : 0;
        T.sigma = bp.Cduct;
        if (bp.BHpoints != 0) { labelNonlinear[l] = 1; anyNonlinearLabel = true; }
    }
    // LinearFlag turns false as soon as one element sits in a B-H material (static2d.cpp:633-639)
    bool LinearFlag = true;
    if (anyNonlinearLabel)
        for (int i = 0; i < NumEls && LinearFlag; i++)
            if (labelNonlinear[meshele[i].lbl]) LinearFlag = false;

    // ---- element table, mixed-BC edges, Dirichlet list, point sources
    std::vector<int> mixedES;
    std::vector<double> mixedC;
    std::vector<int> dirNodes;
    std::vector<double> dirFirst, dirLast;
    std::vector<int> dirSlot(NumNodes, -1);
    auto setValue = [&](int node, double x) {
        int &s = dirSlot[node];
        if (s < 0) { s = (int)dirNodes.size(); dirNodes.push_back(node); dirFirst.push_back(x); dirLast.push_back(x); }
        else dirLast[s] = x;
    };
    std::vector<int> psNode;
    std::vector<double> psVal;
    // nodes with a point property or (axisymmetric) on the axis: a handful among millions — the host threads find
    // them, the two passes below visit them in node order
    std::vector<int> special;
    {
        const double axisTol = units[LengthUnits] * 1.e-06;
        std::vector<std::vector<int>> part(64);
        xb::parallel_for(NumNodes, [&](int i0, int i1, int tid) {
            for (int i = i0; i < i1; i++)
                if (meshnode[i].BoundaryMarker >= 0 || (axi && fabs(meshnode[i].x) < axisTol)) part[tid].push_back(i);
        });
        for (auto &v : part) special.insert(special.end(), v.begin(), v.end());
    }
    for (int i : special)
        if (meshnode[i].BoundaryMarker >= 0)
        {
            psNode.push_back(i);
            double pv = 0.01 * nodeproplist[meshnode[i].BoundaryMarker].J.re;
            if (axi) pv = (0.01 * nodeproplist[meshnode[i].BoundaryMarker].J.re * 2. * meshnode[i].x); // staticaxi.cpp:636-641
            psVal.push_back(pv);
        }
    // fixed potentials at points (static2d.cpp:827-838) come first, then along segments
    for (int i : special)
    {
        // staticaxi.cpp:644-651: nodes on the axis are pinned to zero
        if (axi && fabs(meshnode[i].x) < (units[LengthUnits] * 1.e-06)) { setValue(i, 0.); continue; }
        if (meshnode[i].BoundaryMarker >= 0)
        {
            const CMPointProp &pp = nodeproplist[meshnode[i].BoundaryMarker];
            if ((pp.J.re == 0) && (pp.J.im == 0)) setValue(i, pp.A.re / c);
        }
    }
    // the (few) elements with a tagged side, visited in element order, which is what fixes the first/last semantics
    // of repeated SetValue calls (static2d.cpp:854-921)
    for (int i : tagged)
    {
        const CMElement &El = meshele[i];
        for (int j = 0; j < 3; j++)
        {
            if (El.e[j] < 0) continue;
            const CMBoundaryProp &bp = lineproplist[El.e[j]];
            const int k = (j + 1) % 3;
            if (bp.BdryFormat == 2)
            {
                mixedES.push_back(i); mixedES.push_back(j);
                mixedC.push_back(bp.c0.re); mixedC.push_back(bp.c1.re);
            }
            if (bp.BdryFormat == 0)
            {
                // static2d.cpp:854-921
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
                    a *= cos(bp.phi * DEG);
                    // staticaxi.cpp:674,686: nodes with x == 0 keep the axis condition
                    if (axi && (meshnode[node].x / units[LengthUnits]) == 0) continue;
                    setValue(node, a / c);
                }
            }
        }
    }
    std::vector<int> ageNodes, ageMid;
    std::vector<double> ageW, ageMats;
    if (!axi) flatten_age(agelist, ageNodes, ageW, ageMid, ageMats, false); // StaticAxisymmetric has no AGE loop
    std::vector<int> pbcI(NumPBCs), pbcJ(NumPBCs), pbcT(NumPBCs);
    for (int k = 0; k < NumPBCs; k++) { pbcI[k] = pbclist[k].x; pbcJ[k] = pbclist[k].y; pbcT[k] = pbclist[k].t; }

    // ---- sparsity pattern → device problem
    double t0 = wall();
    if (getenv("XFEMM_TIMING")) fprintf(stderr, "[xfemm timing]   host tables before the pattern %.3f s\n", t0 - tEnter);
    int *rowptr = nullptr, *colidx = nullptr;
    unsigned char *fillmask = nullptr;
    // rotor-angle sweeps re-target the air-gap tables on the resident problem: couple every pair of
    // ring nodes structurally so the pattern does not depend on the angle (2n ring nodes, n = 96)
    std::vector<int> exI, exJ;
    if (ageUnionPattern)
        for (const auto &g : agelist)
        {
            std::vector<int> ring;
            for (const auto &q : g.quadNode) { ring.push_back(q.n0); ring.push_back(q.n1); ring.push_back(q.n2); ring.push_back(q.n3); }
            std::sort(ring.begin(), ring.end());
            ring.erase(std::unique(ring.begin(), ring.end()), ring.end());
            for (size_t a = 0; a < ring.size(); a++)
                for (size_t b = a; b < ring.size(); b++) { exI.push_back(ring[a]); exJ.push_back(ring[b]); }
        }
    if (xfemm_pattern_build_device(NumNodes, NumEls, tri.data(), 4, (int)ageMid.size(), ageNodes.data(), (int)exI.size(), exI.data(), exJ.data(),
                            NumPBCs, pbcI.data(), pbcJ.data(), &rowptr, &colidx, keepSystem ? &fillmask : nullptr) != XFEMM_OK)
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    stats.t_pattern = wall() - t0;
    t0 = wall();
    xfemm_linprob *lin = nullptr;
    const bool timing = getenv("XFEMM_TIMING") != nullptr;
    double tt = wall();
    auto tick = [&](const char *what) { if (timing) { fprintf(stderr, "[xfemm timing]   %s %.3f s\n", what, wall() - tt); tt = wall(); } };
    int rc = xfemm_linprob_create(NumNodes, rowptr, colidx, &lin);
    tick("linprob_create");
    if (rc == XFEMM_OK) L.attachDevice(lin, rowptr, colidx);
    tick("attachDevice");
    if (rc == XFEMM_OK && fillmask) rc = xfemm_linprob_set_fill_mask(lin, fillmask);
    xfemm_free(fillmask);
    xfemm_free(rowptr);
    xfemm_free(colidx);
    if (rc != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }

    std::vector<double> xy(2 * (size_t)NumNodes);
    xb::parallel_for(NumNodes, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++) { xy[2 * (size_t)i] = meshnode[i].x; xy[2 * (size_t)i + 1] = meshnode[i].y; }
    });
    // Locality hint for the solver's internal row order (Z-curve through the node coordinates): the finer
    // layout costs ~0.08 s of host time at 2 M nodes and saves ~2 ms per solve, so by default it is used
    // when the device problem is going to be solved repeatedly (Newton iteration, rotor sweeps).
    // (a partitioned solve always: a rank's share of the rows is then a compact patch of the mesh, its halo small)
    if (localityOrder > 0 || distWorld > 1 || (localityOrder < 0 && (ageUnionPattern || !LinearFlag))) xfemm_linprob_set_coords(lin, xy.data());
    // The solver structure (SELL layout, multigrid hierarchy) depends on the pattern only: a host thread
    // builds it while this thread uploads the mesh and runs the first assembly and boundary conditions.
    if (!assembleOnly && !getenv("XFEMM_NO_PREBUILD"))
    {
        joinPrebuild();
        const int pc = precondOption >= 0 ? precondOption : L.precond;
        prebuild = std::thread([lin, pc]() { xfemm_linprob_prebuild(lin, pc); });
    }
    struct JoinOnExit { FSolver *s; ~JoinOnExit() { s->joinPrebuild(); } } joinOnExit{this}; // never outlive `lin`
    std::vector<double> magdirElem;
    if (anyMagDirFctn)
    {
        magdirElem.resize(NumEls);
        for (int i = 0; i < NumEls; i++)
        {
            const int lbl = meshele[i].lbl;
            double t = labellist[lbl].MagDir;
            if (!labellist[lbl].MagDirFctn.empty())
            {
                double cx = 0, cy = 0;
                for (int j = 0; j < 3; j++) { cx += meshnode[meshele[i].p[j]].x; cy += meshnode[meshele[i].p[j]].y; }
                cx = cx / units[LengthUnits] / 3.; cy = cy / units[LengthUnits] / 3.;
                const xbhost::MagDirVars vars = {cx, cy, atan2(cy, cx) * 180. / PI, sqrt(cx * cx + cy * cy)};
                if (!magExpr[lbl].eval(vars, t))
                {
                    WarnMessage(("\"" + labellist[lbl].MagDirFctn + "\" does not evaluate to a numerical value\n").c_str());
                    return -7;
                }
            }
            magdirElem[i] = t;
        }
    }
    xfemm_mag_desc D;
    memset(&D, 0, sizeof D);
    D.axisymmetric = axi ? 1 : 0;
    D.magdir_elem = anyMagDirFctn ? magdirElem.data() : nullptr;
    D.n_nodes = NumNodes; D.n_elems = NumEls;
    D.xy = xy.data(); D.tri = tri.data();
    D.n_labels = NumBlockLabels; D.labels = labels.data();
    D.n_bh = (int)bhB.size(); D.bh_B = bhB.data(); D.bh_H = bhH.data(); D.bh_slope = bhS.data();
    D.n_mixed = (int)(mixedES.size() / 2); D.mixed_elem_side = mixedES.data(); D.mixed_c0c1 = mixedC.data();
    D.n_age_cells = (int)ageMid.size(); D.age_nodes = ageNodes.data(); D.age_weights = ageW.data();
    D.age_matrix_id = ageMid.data(); D.n_age_matrices = (int)(ageMats.size() / 100); D.age_matrices = ageMats.data();
    D.n_point_src = (int)psNode.size(); D.point_src_node = psNode.data(); D.point_src_val = psVal.data();
    D.ext_zo = extZo * units[LengthUnits]; D.ext_ro = extRo * units[LengthUnits]; D.ext_ri = extRi * units[LengthUnits];
    D.r_eps = 1.e-6 * units[LengthUnits];
    tick("host tables");
    if (xfemm_mag_create(&D, lin, &magdev) != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
    if (deterministicAssembly || getenv("XFEMM_DETERMINISTIC")) xfemm_mag_set_deterministic(magdev, 1);
    tick("mag_create");
    stats.t_upload = wall() - t0;

    plan.dirNodes = dirNodes; plan.dirFirst = dirFirst; plan.dirLast = dirLast;
    plan.pbcI = pbcI; plan.pbcJ = pbcJ; plan.pbcT = pbcT;
    plan.labelJsrc.resize(NumBlockLabels);
    plan.labelBhOffset.resize(NumBlockLabels);
    plan.labelBhPoints.resize(NumBlockLabels);
    for (int l = 0; l < NumBlockLabels; l++)
    {
        plan.labelJsrc[l] = labels[l].Jsrc;
        plan.labelBhOffset[l] = labels[l].bh_offset;
        plan.labelBhPoints[l] = labels[l].bh_points;
    }
    plan.linear = LinearFlag;
    plan.ready = true;
    // The boundary-condition operators are linear in the matrix and only touch entries in the rows/columns of
    // fixed and periodic nodes.  When no air-gap node is one of those, BC(A_static + A_gap) = BC(A_static) + A_gap:
    // the static part can be kept (option age_union_pattern = the problem will be re-solved for other positions).
    plan.staticPartCached = false;
    plan.staticPartUsable = false;
    if (LinearFlag && ageUnionPattern && !ageMid.empty() && !keepSystem && !assembleOnly && !getenv("XFEMM_NO_STATIC_CACHE"))
    {
        std::vector<unsigned char> bcNode(NumNodes, 0);
        for (int nd : dirNodes) bcNode[nd] = 1;
        for (int k = 0; k < NumPBCs; k++) { bcNode[pbcI[k]] = 1; bcNode[pbcJ[k]] = 1; }
        bool disjoint = true;
        for (int nd : ageNodes) // 10 per air-gap cell; every ring node appears in some cell at any rotor position
            if (nd >= 0 && nd < NumNodes && bcNode[nd]) { disjoint = false; break; }
        plan.staticPartUsable = disjoint;
    }
    t0 = wall();
    const int ok = staticNewton(L);
    if (getenv("XFEMM_TIMING"))
        fprintf(stderr, "[xfemm timing] pattern %.3f s, device create+upload %.3f s, first newton loop %.3f s (pcg %.1f ms)\n",
                stats.t_pattern, stats.t_upload, wall() - t0, stats.pcg_ms.empty() ? 0. : stats.pcg_ms[0]);
    return ok;
}

// Newton / "LinearFlag" loop of Static2D (static2d.cpp:177-1016) on the resident device problem:
// assemble → Dirichlet → periodic → PCG → relaxation bookkeeping.  Can be re-run on the same
// device objects (bench.py's device-resident timing, rotor-angle sweeps).
int FSolver::staticNewton(CBigLinProb &L)
{
    if (!plan.ready || !magdev || !L.device()) { WarnMessage("xfemm_b200: no prepared device problem\n"); return false; }
    xfemm_linprob *lin = L.device();
    bool LinearFlag = plan.linear;
    int rc;
    Relax = 1.;
    stats.pcg_iters.clear(); stats.pcg_ms.clear(); stats.assemble_ms.clear();
    if (plan.ageDirty)
    {
        std::vector<int> ageNodes, ageMid;
        std::vector<double> ageW, ageMats;
        flatten_age(agelist, ageNodes, ageW, ageMid, ageMats, false);
        if (xfemm_mag_set_age(magdev, (int)ageMid.size(), ageNodes.data(), ageW.data(), ageMid.data(),
                              (int)(ageMats.size() / 100), ageMats.data()) != XFEMM_OK)
        {
            WarnMessage(xfemm_last_error());
            return false;
        }
        plan.ageDirty = false;
    }
    const int precond = precondOption >= 0 ? precondOption : L.precond;
    int Iter = (assembleOnly && assembleIter > 0) ? assembleIter : 0;
    if (Iter > 0) LinearFlag = false; // a later pass of a nonlinear problem: B-H terms from the iterate on the device
    double res = 0, lastres = 0;
    const bool stepTiming = getenv("XFEMM_STEP_TIMING") != nullptr;
    const double tStep0 = wall();
    double tAsm0 = 0, tAsm1 = 0, tSolve0 = 0, tSolve1 = 0;
    // re-solve on the resident problem (next position of a sweep): initial guess from the positions solved
    // before — the last one, or the linear extrapolation of the last two (equidistant positions)
    if (resolveStart)
    {
        const char *eo = getenv("XFEMM_EXTRAPOLATION_ORDER");
        xfemm_linprob_history_guess(lin, extrapolateResolve ? (eo ? atoi(eo) : 2) : 0);
    }
    do
    {
        PrintMessage("Matrix Construction\n");
        double ms_asm = 0, ms_gap = 0;
        double tb = wall();
        tAsm0 = tb;
        const bool cacheStatic = plan.staticPartUsable && cacheStaticPart;
        if (cacheStatic && plan.staticPartCached)
        {
            // next position of a sweep: boundary-conditioned static part back from its device copy + this position's
            // air-gap blocks
            if (xfemm_linprob_restore(lin) != XFEMM_OK || xfemm_mag_assemble_static_parts(magdev, Iter, 2, &ms_gap) != XFEMM_OK)
            {
                WarnMessage(xfemm_last_error());
                return false;
            }
        }
        else
        {
            const int parts = cacheStatic ? 1 : 3;
            if (xfemm_mag_assemble_static_parts(magdev, Iter, parts, &ms_asm) != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
            tb = wall();
            if (xfemm_linprob_dirichlet(lin, (int)plan.dirNodes.size(), plan.dirNodes.data(), plan.dirFirst.data(), plan.dirLast.data(), L.bdw) != XFEMM_OK ||
                xfemm_linprob_periodic(lin, NumPBCs, plan.pbcI.data(), plan.pbcJ.data(), plan.pbcT.data()) != XFEMM_OK)
            {
                WarnMessage(xfemm_last_error());
                return false;
            }
            if (cacheStatic)
            {
                if (xfemm_linprob_snapshot(lin) != XFEMM_OK || xfemm_mag_assemble_static_parts(magdev, Iter, 2, &ms_gap) != XFEMM_OK)
                {
                    WarnMessage(xfemm_last_error());
                    return false;
                }
                plan.staticPartCached = true;
            }
        }
        stats.assemble_ms.push_back(ms_asm + ms_gap + (wall() - tb) * 1e3);
        tAsm1 = wall();
        if (keepSystem)
        {
            // snapshot for the parity tests, pruned to the entries the reference's linked lists hold
            const int *rp0 = L.patternRowptr(), *col0 = L.patternCol();
            const size_t nnz0 = (size_t)rp0[NumNodes];
            std::vector<double> v0(nnz0);
            std::vector<unsigned char> ex(nnz0);
            sysRhs.resize(NumNodes);
            xfemm_linprob_get_values(lin, v0.data());
            xfemm_linprob_get_rhs(lin, sysRhs.data());
            xfemm_linprob_get_exists_mask(lin, ex.data());
            sysRowptr.assign(NumNodes + 1, 0);
            sysCol.clear();
            sysVals.clear();
            for (int i = 0; i < NumNodes; i++)
            {
                for (int q = rp0[i]; q < rp0[i + 1]; q++)
                    if (ex[q]) { sysCol.push_back(col0[q]); sysVals.push_back(v0[q]); }
                sysRowptr[i + 1] = (int)sysCol.size();
            }
        }
        if (assembleOnly) { stats.newton_iters = 0; return true; } // the caller solves (row-partitioned PCG)
        xfemm_linprob_save_solution(lin); // V_old = V
        joinPrebuild(); // the solver structure is needed from here on (a failed pre-build is simply redone by the solve)
        PrintMessage("Conjugate Gradient Solver\n");
        tSolve0 = wall();
        xfemm_solve_info info;
        // Newton steps after the first continue from the last iterate (static2d.cpp:929); so does the first
        // solve of a re-solve on the resident problem (next rotor position: the previous field is close)
        if (distWorld > 1)
        {
            // one mesh over several GPUs: every process has assembled the same system; the solve is shared
            memset(&info, 0, sizeof info);
            if (!dpcg)
            {
                const char *dir = getenv("XFEMM_DIST_DIR"), *job = getenv("XFEMM_DIST_JOB");
                rc = xfemm_dpcg_create(lin, distRank, distWorld, &dpcg);
                if (rc == XFEMM_OK) rc = xfemm_dpcg_connect_files(dpcg, dir ? dir : "/dev/shm", job ? job : "xfemm_dpcg", 300);
            }
            else rc = XFEMM_OK;
            if (rc == XFEMM_OK) rc = xfemm_dpcg_solve(dpcg, L.Precision, (Iter > 0 || resolveStart) ? 1 : 0, maxIterOption, &info);
            if (rc == XFEMM_OK) rc = xfemm_dpcg_share_solution(dpcg);
            if (rc == XFEMM_OK) rc = xfemm_dpcg_finish(dpcg);
        }
        else
        rc = xfemm_linprob_pcg(lin, L.Precision, (Iter > 0 || resolveStart) ? 1 : 0, precond, maxIterOption, &info);
        if (rc == XFEMM_ERR_DIVERGED && precond == XFEMM_PRECOND_AMG && distWorld <= 1)
        {
            // safety net: the multigrid hierarchy is built from the sparsity pattern alone; should a problem
            // defeat it, the (slower, always applicable) Jacobi-preconditioned kernel takes over on the GPU
            PrintMessage("multigrid-preconditioned CG did not converge; retrying with the Jacobi preconditioner\n");
            // V may hold the failed iterate: go back to the saved one, which is the warm start the first attempt had
            const int warm = (Iter > 0 || resolveStart) ? 1 : 0;
            if (warm) xfemm_linprob_restore_solution(lin);
            rc = xfemm_linprob_pcg(lin, L.Precision, warm, XFEMM_PRECOND_JACOBI, maxIterOption, &info);
        }
        tSolve1 = wall();
        stats.pcg_iters.push_back(info.iterations);
        stats.pcg_ms.push_back(info.solve_ms);
        stats.final_relres = info.relres;
        L.lastIterations = info.iterations; L.lastSolveMs = info.solve_ms; L.lastRelRes = info.relres;
        if (rc != XFEMM_OK) { WarnMessage(xfemm_last_error()); return false; }
        if (getenv("XFEMM_DEBUG_DUMP"))
        {
            std::vector<double> tmp(NumNodes);
            xfemm_linprob_get_solution(lin, tmp.data());
            debug_dump("vout", Iter, tmp.data(), tmp.size() * sizeof(double));
            if (keepSystem)
            {
                debug_dump("vals", Iter, sysVals.data(), sysVals.size() * sizeof(double));
                debug_dump("rhs", Iter, sysRhs.data(), sysRhs.size() * sizeof(double));
            }
        }

        if (LinearFlag == false)
        {
            double x = 0, y = 0;
            xfemm_linprob_newton_norms(lin, &x, &y);
            if (y == 0) LinearFlag = true;
            else { lastres = res; res = sqrt(x / y); }
            if (Iter > 5)
            {
                if ((res > lastres) && (Relax > 0.125)) Relax /= 2.;
                else Relax += 0.1 * (1. - Relax);
                xfemm_linprob_relax(lin, Relax);
            }
            char outstr[256];
            snprintf(outstr, sizeof outstr, "Newton Iteration(%i) Relax=%.4g\n", Iter, Relax);
            PrintMessage(outstr);
        }
        if ((res < 100. * Precision) && (Iter > 0)) LinearFlag = true;
        Iter++;
    } while (LinearFlag == false);
    stats.newton_iters = Iter;
    xfemm_linprob_history_push(lin); // remembered as a starting point for the next problem on this handle

    // results back to the caller's arrays: L.V internal units, L.b = A (static2d.cpp:1018-1021)
    // (sweeps that only read derived quantities — gapTorque — leave the nodal values on the device)
    hostSolutionStale = true;
    const double tFetch0 = wall();
    if (fetchSolution && !fetchDeviceSolution(L)) return false;
    if (stepTiming)
        fprintf(stderr, "[xfemm step] before assembly %.3f ms, assembly + BC %.3f ms (wall), to solver %.3f ms, solve call %.3f ms (kernel %.3f ms), "
                        "after solve %.3f ms, fetch %.3f ms, total %.3f ms\n",
                (tAsm0 - tStep0) * 1e3, (tAsm1 - tAsm0) * 1e3, (tSolve0 - tAsm1) * 1e3, (tSolve1 - tSolve0) * 1e3,
                stats.pcg_ms.empty() ? 0. : stats.pcg_ms.back(), (tFetch0 - tSolve1) * 1e3, (wall() - tFetch0) * 1e3, (wall() - tStep0) * 1e3);
    // The reference also accumulates meshele[i].Jprev += Jsrc on every pass of a non-incremental problem
    // (static2d.cpp:506); nothing reads it afterwards — WriteStatic2D does not print it and WriteHarmonic2D prints the
    // value loaded from a previous solution (harmonic2d.cpp:947) — so it is not kept here.
    return true;
}

bool FSolver::fetchDeviceSolution(CBigLinProb &L)
{
    if (!hostSolutionStale) return true;
    if (!L.device() || !magdev) { WarnMessage("xfemm_b200: no device solution to fetch\n"); return false; }
    // a problem whose results are fetched again and again (positions of a sweep): from the second time on into
    // page-locked arrays
    if (++solutionFetches == 2) L.pinResultArrays();
    if (xfemm_linprob_get_solution(L.device(), L.V) != XFEMM_OK || xfemm_mag_solution_to_A(magdev, L.b) != XFEMM_OK)
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    hostSolutionStale = false;
    return true;
}

int FSolver::Harmonic2D(CBigComplexLinProb &L, bool verbose) { return harmonicSolve(L, false, verbose); }
int FSolver::HarmonicAxisymmetric(CBigComplexLinProb &L, bool verbose) { return harmonicSolve(L, true, verbose); }

// harmonicSolve lives in fsolver_harmonic.cpp

// ---------------------------------------------------------------------------------------
// Rotor position without re-meshing.  fmesher rebuilds the ring tables of an
// air-gap element from the boundary nodes, their angular positions and the inner/outer angle
// (cfemm/fmesher/writepoly.cpp:1853-1966); the mesh itself does not depend on the angle.  The same
// table is regenerated here from the table already loaded: ring nodes are mapped onto the annulus
// grid for every periodic copy, sorted by angle, and bracketed per annulus cell.
// ---------------------------------------------------------------------------------------
// Torque through an air-gap element from the solution resident on the device: what the reference
// obtains by writing the .ans file and evaluating mo_gapintegral(name,0) in fpproc
// (cfemm/fpproc/fpproc.cpp:1391-1615 and 5418-5446).  Planar magnetostatics.
bool FSolver::gapTorque(int ageIndex, double &torque)
{
    static const double lengthConv[] = {0.0254, 0.001, 0.01, 1., 2.54e-05, 1.e-06}; // metres per unit (fpproc.cpp LengthConv)
    if (ageIndex < 0 || ageIndex >= (int)agelist.size() || !magdev || Frequency != 0 || ProblemType != PLANAR)
    {
        WarnMessage("xfemm_b200: gap torque needs a solved planar magnetostatic problem with that air-gap element\n");
        return false;
    }
    int cell0 = 0;
    for (int a = 0; a < ageIndex; a++) cell0 += agelist[a].totalArcElements;
    const CAirGapElement &g = agelist[ageIndex];
    const double conv = lengthConv[LengthUnits];
    const double ri = g.ri * conv, ro = g.ro * conv;
    const double R = (ri + ro) / 2., dr = ro - ri;
    const double dt = (PI / 180.) * g.totalArcLength / ((double)g.totalArcElements);
    const int m = (g.BdryFormat == 0) ? (int)std::lround(360. / g.totalArcLength) : (int)std::lround(180. / g.totalArcLength);
    const double depth = (Depth == -1) ? 1. : Depth * conv;
    if (xfemm_mag_gap_torque(magdev, cell0, g.totalArcElements, g.InnerShift, g.OuterShift, dt, R, dr, g.BdryFormat, m, depth,
                             &torque, nullptr) != XFEMM_OK)
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    return true;
}

// Block integrals from the resident solution: what the reference obtains from fpproc after writing
// and re-reading the .ans file.  cfemm/fpproc/fpproc.cpp:3762-3830 (stored energy), 5157-5196 (flux linkage).
bool FSolver::labelIntegrals(std::vector<double> &energyArea, std::vector<double> &aArea)
{
    if (!magdev || !plan.ready || Frequency != 0 || ProblemType != PLANAR)
    {
        WarnMessage("xfemm_b200: block integrals on the device need a solved planar magnetostatic problem\n");
        return false;
    }
    std::vector<xfemm_integral_label> T(NumBlockLabels);
    for (int l = 0; l < NumBlockLabels; l++)
    {
        const CMSolverMaterialProp &M = blockproplist[labellist[l].BlockType];
        T[l].mu_x = M.mu_x; T[l].mu_y = M.mu_y; T[l].lamfill = M.LamFill; T[l].H_c = M.H_c; T[l].magdir = labellist[l].MagDir;
        T[l].lamtype = M.LamType; T[l].bh_offset = plan.labelBhOffset[l]; T[l].bh_points = plan.labelBhPoints[l]; T[l].pad_ = 0;
    }
    energyArea.assign(NumBlockLabels, 0.);
    aArea.assign(NumBlockLabels, 0.);
    if (xfemm_mag_label_integrals(magdev, NumBlockLabels, T.data(), energyArea.data(), aArea.data()) != XFEMM_OK)
    {
        WarnMessage(xfemm_last_error());
        return false;
    }
    return true;
}

bool FSolver::storedEnergy(double &joules)
{
    std::vector<double> e, a;
    if (!labelIntegrals(e, a)) return false;
    const double depth = (Depth == -1) ? 1. : Depth * LengthConvMeters[LengthUnits];
    joules = 0;
    for (double v : e) joules += v * depth;
    return true;
}

bool FSolver::fluxLinkage(int circ, double &webers)
{
    // circuits of the .fem file; series circuits were expanded into one parallel circuit per label by
    // LoadProblemFile (fsolver.cpp:278-317): OrigCirc leads back
    if (circ < 0 || circ >= NumCircPropsOrig) { WarnMessage("xfemm_b200: no such circuit\n"); return false; }
    const CComplex amps = circproplist[circ].Amps;
    if (amps.re == 0 && amps.im == 0)
    {
        WarnMessage("xfemm_b200: the flux linkage of a circuit without current is not evaluated on the device\n");
        return false;
    }
    std::vector<double> e, a;
    if (!labelIntegrals(e, a)) return false;
    const double depth = (Depth == -1) ? 1. : Depth * LengthConvMeters[LengthUnits];
    double aj = 0;
    for (int l = 0; l < NumBlockLabels; l++)
    {
        int c = labellist[l].InCircuit;
        if (c < 0) continue;
        if (c >= NumCircPropsOrig) c = circproplist[c].OrigCirc;
        if (c == circ) aj += a[l] * plan.labelJsrc[l] * 1.e6 * depth; // MA/m^2 -> A/m^2
    }
    webers = aj / amps.re;
    return true;
}

bool FSolver::setAirGapAngles(int ageIndex, double innerAngle, double outerAngle)
{
    if (ageIndex < 0 || ageIndex >= (int)agelist.size()) return false;
    CAirGapElement &g = agelist[ageIndex];
    const int n = g.totalArcElements;
    if (n <= 0 || (int)g.quadNode.size() != n + 1) return false;
    const double dtta = g.totalArcLength / n;
    const int n0 = (int)std::lround(360. / dtta);
    const int n1 = (int)std::lround(360. / g.totalArcLength);
    // the n distinct boundary nodes of each ring, in the order they first appear in the table
    std::vector<int> inner, outer;
    auto push_unique = [](std::vector<int> &v, int x) { if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x); };
    for (const auto &q : g.quadNode) { push_unique(inner, q.n0); push_unique(inner, q.n1); push_unique(outer, q.n2); push_unique(outer, q.n3); }
    if ((int)inner.size() != n || (int)outer.size() != n) return false;
    const double cm = 100 * LengthConvMeters[LengthUnits];
    const cplx agc(g.agc.re * cm, g.agc.im * cm);
    struct RingPt { int node; double z, sgn; };
    auto build = [&](const std::vector<int> &nodes, double angle) {
        std::vector<RingPt> ring;
        ring.reserve(n0);
        for (int j = 0; j < n1; j++)
        {
            const double dL = ((g.BdryFormat == 1) && (j % 2 != 0)) ? -1 : 1;
            const cplx rot = std::exp(cplx(0, 1) * ((j * g.totalArcLength + angle) * 0.017453292519943295));
            for (int node : nodes)
            {
                cplx a0 = rot * (cplx(meshnode[node].x, meshnode[node].y) - agc);
                double ang = (a0.imag() >= 0) ? std::arg(a0) : (std::arg(a0) + 2. * PI);
                ring.push_back({node, ang * (180. / PI) / dtta, dL});
            }
        }
        std::stable_sort(ring.begin(), ring.end(), [](const RingPt &a, const RingPt &b) { return a.z < b.z; });
        return ring;
    };
    std::vector<RingPt> I = build(inner, innerAngle), O = build(outer, outerAngle);
    if ((int)I.size() != n0 || (int)O.size() != n0) return false;
    g.InnerAngle = innerAngle;
    g.OuterAngle = outerAngle;
    g.InnerShift = I[0].z;
    g.OuterShift = O[0].z;
    for (int i = 0; i <= n; i++)
    {
        int p1 = i; if (p1 == n0) p1 = 0;
        int p0 = p1 - 1; if (p0 < 0) p0 = n0 + p0;
        CQuadPoint &q = g.quadNode[i];
        q.n0 = I[p0].node; q.w0 = I[p0].sgn; q.n1 = I[p1].node; q.w1 = I[p1].sgn;
        q.n2 = O[p0].node; q.w2 = O[p0].sgn; q.n3 = O[p1].node; q.w3 = O[p1].sgn;
    }
    plan.ageDirty = true;
    return true;
}
