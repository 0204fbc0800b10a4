// Drop-in host-side interface of the fsolver hot path (B200 backend).
//
// This header mirrors the PUBLIC face of the reference classes so that the reference's callers
// (cfemm/fsolver/main.cpp:39-84, mfemm/mexfsolver.cpp:34-135, cfemm/femmcli/LuaMagneticsCommands.cpp:817-843)
// compile against it unchanged:
//   class FSolver            — cfemm/fsolver/fsolver.h:50-150 over cfemm/libfemm/feasolver.h:76-209
//   femm::CNode, femmsolver::CMElement, femm::CM*Prop, femmsolver::CAirGapElement … (data records)
// Everything numerical behind Static2D/StaticAxisymmetric/Harmonic2D/HarmonicAxisymmetric and
// CBigLinProb/CBigComplexLinProb runs on the GPU through the C ABI in include/xfemm_b200.h.
// There is no CPU solve path: without a CUDA device these calls fail loudly.
#ifndef XFEMM_B200_FSOLVER_H
#define XFEMM_B200_FSOLVER_H

#include "femmcomplex.h"
#include "spars.h"
#include "cspars.h"

#include <string>
#include <thread>
#include <vector>

#ifndef PI
#define PI 3.141592653589793238462643383
#endif
#ifndef DEG
#define DEG 0.01745329251994329576923690768
#endif
#ifndef muo
#define muo 1.2566370614359173e-6
#endif

enum LoadMeshErr
{
    NOERROR,
    BADFEMFILE,
    BADNODEFILE,
    BADPBCFILE,
    BADELEMENTFILE,
    BADEDGEFILE,
    MISSINGMATPROPS,
    ELMLABELTOOBIG
};

namespace femm
{
enum LengthUnit { LengthInches = 0, LengthMillimeters = 1, LengthCentimeters = 2, LengthMeters = 3, LengthMils = 4, LengthMicrometers = 5 };
enum CoordsType { CART = 0, POLAR = 1 };
enum ProblemType { PLANAR = 0, AXISYMMETRIC = 1 };
extern const double LengthConvMeters[6];

int PrintWarningMsg(const char *message, ...);

struct CNode
{
    double x = 0, y = 0;
    int BoundaryMarker = -1;
};

struct CCommonPoint
{
    int x = 0, y = 0, t = 0;
};

struct CQuadPoint
{
    int n0, n1, n2, n3;
    double w0, w1, w2, w3;
};

struct CMPointProp
{
    std::string PointName = "New Point Property";
    CComplex J, A;
};

struct CMBoundaryProp
{
    std::string BdryName = "New Boundary";
    int BdryFormat = 0;
    double InnerAngle = 0, OuterAngle = 0;
    double A0 = 0, A1 = 0, A2 = 0, phi = 0;
    double Mu = 0, Sig = 0;
    CComplex c0, c1;
};

// Magnetic material incl. the cubic-Hermite H(B) table (reference: cfemm/libfemm/CMaterialProp.h,
// CMMaterialProp + CMSolverMaterialProp).  Spline set-up stays on the host; evaluation during
// assembly happens on the device from the uploaded tables.
struct CMSolverMaterialProp
{
    std::string BlockName = "New Material";
    double mu_x = 1, mu_y = 1;
    int BHpoints = 0;
    std::vector<double> Bdata;
    std::vector<CComplex> Hdata;
    std::vector<CComplex> slope;
    int LamType = 0;
    double LamFill = 1;
    double H_c = 0, Nrg = 0;
    CComplex J;
    double Cduct = 0, Lam_d = 0;
    double Theta_hn = 0, Theta_hx = 0, Theta_hy = 0;
    int NStrands = 0;
    double WireD = 0;
    double MuMax = 0;
    double Theta_m = 0;
    CComplex mu_fdx, mu_fdy;

    void GetSlopes(double omega);                        // CMaterialProp.cpp:127-348
    void clearSlopes() { slope.clear(); }
    CComplex LaminatedBH(double w, int i);               // CMaterialProp.cpp:1062-1171
    CComplex GetH(double B);                             // CMaterialProp.cpp:960-987
    CComplex GetdHdB(double B);                          // CMaterialProp.cpp:461-486
    CComplex Get_v(double B);                            // CMaterialProp.cpp:899-904
    void GetBHProps(double B, CComplex &v, CComplex &dv); // CMaterialProp.cpp:1007-1057
    void GetBHProps(double B, double &v, double &dv);
    void IncrementalPermeability(double B, double &mu1, double &mu2);
    void incrementalPermeability(double B, double w, CComplex &mu1, CComplex &mu2);
};

struct CMCircuit
{
    std::string CircName = "New Circuit";
    int CircType = 0;
    CComplex Amps, dVolts;
    int OrigCirc = 0;
    CComplex J, dV;
    int Case = 0;
};

struct CMBlockLabel
{
    double x = 0, y = 0, MaxArea = 0;
    int InGroup = 0;
    bool IsExternal = false, IsDefault = false;
    int BlockType = -1, InCircuit = -1;
    CComplex ProximityMu;
    int Turns = 1;
    bool bIsWound = false;
    double MagDir = 0;
    std::string MagDirFctn;
    int Case = 0;
    CComplex J, dVolts;
    double FillFactor = 1;
    CComplex o, mu;
};
} // namespace femm

namespace femmsolver
{
struct CMElement
{
    int p[3] = {0, 0, 0};
    int e[3] = {-1, -1, -1};
    int blk = 0, lbl = 0;
    CComplex mu1, mu2, v12;
    double magdir = 0;
    double Jprev = 0;
};

struct CAirGapElement
{
    std::string BdryName;
    int BdryFormat = 0;
    int totalArcElements = 0;
    double totalArcLength = 0;
    double ri = 0, ro = 0;
    double InnerAngle = 0, OuterAngle = 0;
    double InnerShift = 0, OuterShift = 0;
    CComplex agc;
    std::vector<femm::CQuadPoint> quadNode;
};
} // namespace femmsolver

struct xfemm_mag; // opaque device problem (include/xfemm_b200.h)

// Per-run metrics the B200 backend records (not in the reference).
struct FSolverStats
{
    int newton_iters = 0;
    std::vector<int> pcg_iters;      // linear iterations of each (Newton) solve
    std::vector<double> pcg_ms;      // device-timed
    std::vector<double> assemble_ms; // device-timed (assembly + boundary conditions)
    double final_relres = 0;         // true ||b-Ax||/||b|| of the last solve
    double t_load_mesh = 0, t_renumber = 0, t_pattern = 0, t_upload = 0, t_write = 0;
};

class FSolver
{
public:
    FSolver();
    virtual ~FSolver();

    // ---- general problem attributes (feasolver.h:98-120)
    double FileFormat, Precision, MinAngle, Depth;
    femm::LengthUnit LengthUnits;
    femm::CoordsType Coords;
    femm::ProblemType ProblemType;
    double extZo, extRo, extRi;
    std::string comment;
    int ACSolver;
    bool DoForceMaxMeshArea, DoSmartMesh, bMultiplyDefinedLabels;

    int BandWidth;
    std::vector<femmsolver::CMElement> meshele;
    int NumNodes, NumEls;
    int NumBlockProps, NumPBCs, NumLineProps, NumPointProps, NumCircProps, NumBlockLabels, NumAirGapElems;
    std::vector<femm::CCommonPoint> pbclist;
    std::string PathName;
    int PrevType;
    std::string previousSolutionFile;

    std::vector<femm::CMPointProp> nodeproplist;
    std::vector<femm::CMBoundaryProp> lineproplist;
    std::vector<femm::CMSolverMaterialProp> blockproplist;
    std::vector<femm::CMCircuit> circproplist;
    std::vector<femm::CMBlockLabel> labellist;
    std::vector<femmsolver::CAirGapElement> agelist;

    // ---- fsolver.h:65-75
    double Frequency;
    double Relax;
    std::vector<femm::CNode> meshnode;
    int NumCircPropsOrig;

    int (*WarnMessage)(const char *, ...);
    int (*PrintMessage)(const char *, ...);

    // ---- operations (same names, argument meaning and return conventions as the reference)
    bool LoadProblemFile();                        // fsolver.cpp:202-348
    // [PrevSoln] with [PrevType] = 0: the mesh (nodes, elements, periodic pairs, air-gap tables) comes from an earlier
    // .ans file in its solver numbering; LoadMesh and Cuthill are then no-ops (fsolver.cpp:224-239, 357, 990-1081)
    bool loadPreviousSolution(bool loadAprev);
    bool meshLoadedFromPrevSolution = false;
    LoadMeshErr LoadMesh(bool deleteFiles = true); // fsolver.cpp:350-718
    int Cuthill(bool deleteFiles = true);          // cuthill.cpp:94-391
    int SortElements();                            // cuthill.cpp:40-85
    int Static2D(CBigLinProb &L);                  // static2d.cpp:53-1033
    int WriteStatic2D(CBigLinProb &L);             // static2d.cpp:1038-1195
    int StaticAxisymmetric(CBigLinProb &L);        // staticaxi.cpp:45-794
    int Harmonic2D(CBigComplexLinProb &L, bool verbose = false);          // harmonic2d.cpp:38-890
    int WriteHarmonic2D(CBigComplexLinProb &L);                           // harmonic2d.cpp:892-1036
    int HarmonicAxisymmetric(CBigComplexLinProb &L, bool verbose = false); // harmonicaxi.cpp:34-821
    void GetFillFactor(int lbl);                   // fsolver.cpp:1083-1194
    double ElmArea(int i);                         // fsolver.cpp:1196-1211
    virtual bool runSolver(bool verbose = false);  // fsolver.cpp:1213-1338
    virtual void CleanUp();
    static std::string getErrorString(LoadMeshErr err);

    // ---- B200 backend extras
    FSolverStats stats;
    std::vector<double> Aprev;
    // newnum[i] = solver index of mesh-file node i (the Cuthill–McKee permutation), kept for parity checks
    std::vector<int> cmPermutation;
    std::vector<int> edgeList;   // .edge node pairs cached by LoadMesh for Cuthill
    // One mesh over several GPUs (one process per GPU, every process runs the same FSolver calls): rank / size of
    // this process.  distWorld > 1 routes the linear solves of Static2D / StaticAxisymmetric through xfemm_dpcg_*
    // (csrc/dist_amg.cu); the processes find each other through files XFEMM_DIST_DIR/XFEMM_DIST_JOB.<rank>
    // (defaults: /dev/shm, "xfemm_dpcg").  Options "dist_rank" / "dist_world" of xfemm_fsolver_set_option; the
    // fsolver command line reads XFEMM_DIST_RANK / XFEMM_DIST_WORLD.
    int distRank = 0, distWorld = 1;
    struct xfemm_dpcg *dpcg = nullptr; // owned, tied to the resident device problem
    int precondOption;           // XFEMM_PRECOND_* or -1 = backend default
    int maxIterOption;           // <=0 = backend default
    xfemm_mag *magdev;           // device mesh of the last solve (kept for queries), owned
    struct xfemm_cmag *cmagdev;  // same for the time-harmonic formulations
    void releaseDevice();
    // host-side inputs of the Newton loop that stay valid while the device problem is resident
    struct StaticPlan
    {
        std::vector<int> dirNodes, pbcI, pbcJ, pbcT;
        std::vector<double> dirFirst, dirLast, labelJsrc;
        std::vector<int> labelBhOffset, labelBhPoints; // B-H table of each label in the device arrays
        bool linear = true, ready = false, ageDirty = false;
        // sweeps on a linear problem: the position-independent part (elements, sides, sources, all boundary
        // conditions) is assembled once and kept on the device; a position restores it and adds its air-gap blocks
        bool staticPartUsable = false, staticPartCached = false;
    } plan;
    bool warmResolve;            // a re-solve on the resident problem starts from the solution it finds there (default true)
    bool resolveStart;           // set for the duration of such a re-solve
    bool extrapolateResolve;     // ... from the linear extrapolation of the last two solutions (equidistant sweeps; default true)
    bool cacheStaticPart;        // sweeps on linear problems keep the position-independent, boundary-conditioned part on the device (default true)
    bool fetchSolution;          // copy V and A back to the host after a solve (default true; sweeps that only need torque switch it off)
    bool hostSolutionStale = false; // the last solve left V and A on the device (fetchSolution was off): see fetchDeviceSolution
    int solutionFetches = 0; // fetches from the current device problem (reset with it)
    // copies V and A of the last static solve from the device into L.V / L.b if they were left there; false on a CUDA error
    bool fetchDeviceSolution(CBigLinProb &L);
    bool gapTorque(int ageIndex, double &torque); // from the resident solution, = fpproc's gap integral 0
    // from the resident solution (planar magnetostatics): stored energy of all blocks [J] = fpproc's block integral 2 over
    // every block; flux linkage [Wb] of circuit `circ` of the .fem file (a circuit that carries a current)
    bool storedEnergy(double &joules);
    bool fluxLinkage(int circ, double &webers);
    bool labelIntegrals(std::vector<double> &energyArea, std::vector<double> &aArea);
    std::thread prebuild;        // host thread building the solver structure while the first assembly runs
    void joinPrebuild() { if (prebuild.joinable()) prebuild.join(); }
    int localityOrder;           // 1/0: order the solver's rows along a Z-curve through the nodes; -1 = when the problem is re-solved
    bool assembleOnly;           // stop after assembly + boundary conditions of the first pass
    int assembleIter = 0;        // ... which then counts as Newton pass number `assembleIter` (> 0: the Jacobian of the
                                 // iterate that is on the device, xfemm_linprob_set_solution; parity tests of the Newton terms)
    bool deterministicAssembly = false; // gather-based assembly in a fixed order instead of FP64 atomics (option "deterministic_assembly", env XFEMM_DETERMINISTIC)
    bool ageUnionPattern;        // couple all ring nodes of an air-gap element structurally (sweeps)
    // regenerates the ring tables of air-gap element `ageIndex` for new rotor/stator angles [deg]
    bool setAirGapAngles(int ageIndex, double innerAngle, double outerAngle);
    int staticNewton(CBigLinProb &L); // re-runnable on the resident device problem
    // optional host snapshot of the last assembled + boundary-conditioned system (parity tests)
    bool keepSystem;
    std::vector<int> sysRowptr, sysCol;
    std::vector<double> sysVals, sysRhs;
    std::vector<double> sysSide[3]; // Newton AC: Mh, Ms, Ma on the slots of sysCol (keep_system)

protected:
    void SortNodes(const std::vector<int> &newnum); // fsolver.cpp:1341-1353
    bool parseFem(const std::string &file);
    int staticSolve(CBigLinProb &L, bool axisymmetric);
    int harmonicSolve(CBigComplexLinProb &L, bool axisymmetric, bool verbose);
};

double Power(double x, int n);

#endif
