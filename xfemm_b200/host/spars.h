// CBigLinProb — drop-in for cfemm/libfemm/spars.h:38-83, backed by the B200 solver.
//
// Same public data members and methods as the reference class, so FSolver/HSolver/ESolver-style
// callers compile unchanged.  Two ways to fill it:
//   * host assembly through Put/AddTo/SetValue/Periodicity (what hsolver/esolver do): entries are
//     kept in sorted per-row arrays (upper triangle, like the reference's linked lists, but O(deg)
//     for the boundary-condition scans) and frozen into CSR + uploaded when PCGSolve is called;
//   * device assembly (what the B200 FSolver does): attachDevice() binds a device problem whose
//     values never visit the host; Get() then reads from a lazily downloaded copy.
// PCGSolve always runs on the GPU; there is no CPU solver behind this class.
#ifndef XFEMM_B200_SPARS_H
#define XFEMM_B200_SPARS_H

#include <vector>

struct xfemm_linprob;

class CEntry
{
public:
    double x; // value stored in the entry
    int c;    // column that the entry lives in
    CEntry *next;
    CEntry() : x(0), c(0), next(nullptr) {}
};

class CBigLinProb
{
public:
    // data members (reference names)
    double *V; // solution
    double *P; // search direction (host scratch of length n for callers, as in hsolver.cpp:996)
    double *R;
    double *U;
    double *Z;
    double *b;  // RHS of linear equation
    CEntry **M; // row lists of the reference (upper triangle, ascending columns, diagonal first): filled ON DEMAND by
                // Entries() — the solver itself keeps the matrix on the device; walking M without that call finds nulls
    int n;
    int bdw;
    double Precision;
    double Lambda; // relaxation factor of the SSOR preconditioner
    int *Q;        // used by esolver and hsolver

    CBigLinProb();
    virtual ~CBigLinProb();
    virtual int Create(int d, int bw);
    void Put(double v, int p, int q);
    double Get(int p, int q);
    bool PCGSolve(int flag); // flag != 0: V holds an initial guess
    void MultPC(const double *X, double *Y);
    void AddTo(double v, int p, int q);
    void MultA(double *X, double *Y);
    void SetValue(int i, double x);
    void Periodicity(int i, int j);
    void AntiPeriodicity(int i, int j);
    void Wipe();
    double Dot(double *X, double *Y);
    void ComputeBandwidth();
    // Materialises the reference's storage — M[i] = singly linked list of the stored entries of row i, columns >= i
    // ascending, first node the diagonal (spars.h:25-35, spars.cpp:95-99) — from the current matrix, wherever it lives
    // (host rows of a Put()-assembled problem, or the device values of a device-assembled one), and returns M.  For
    // callers that walk the lists (white-box harnesses); the lists are a snapshot, later Put()s do not update them.
    CEntry **Entries();

    // ---- B200 backend
    int precond;       // XFEMM_PRECOND_*, default from env XFEMM_PRECOND or multicolour SSOR
    int lastIterations;
    double lastSolveMs, lastRelRes;
    xfemm_linprob *device() { return dev; }
    // binds an externally built device problem (ownership passes to this object)
    void attachDevice(xfemm_linprob *d, const int *rowptr, const int *col);
    void invalidateHostCopy() { hostValuesValid = false; }
    // Page-locks V and b (cudaHostRegister) so that results copied back from the device land there by DMA at full
    // PCIe speed: worth its one-time cost (milliseconds) when a resident problem is solved again and again (sweeps).
    // Undone by the destructor / the next Create().
    void pinResultArrays();
    const int *patternRowptr() { ensurePattern(); return csrRowptr.data(); }
    const int *patternCol() { ensurePattern(); return csrCol.data(); }

private:
    struct Ent { int c; double x; };
    void ensureRows();                   // lazily creates the host rows below
    void ensurePattern();                // device-assembled problems: the host copy of the pattern is fetched on first use
    std::vector<std::vector<Ent>> rows;  // upper triangle, sorted by column, first entry = diagonal
    std::vector<std::vector<int>> lower; // lower[i] = rows k<i holding an entry (k,i)
    bool patternDirty, valuesDirty, hostValuesValid, deviceAssembled;
    bool resultPinned = false;
    void unpinResultArrays();
    xfemm_linprob *dev;
    std::vector<int> csrRowptr, csrCol;
    std::vector<double> csrVals;
    std::vector<CEntry> entryPool; // nodes of the lists handed out by Entries()
    double *findEntry(int p, int q);
    bool syncToDevice();
    void periodicImpl(int i, int j, double sgn);
    void fetchDeviceValues();
};

#endif
