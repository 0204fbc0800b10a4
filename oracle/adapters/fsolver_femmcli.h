// Binding a maintainer of the reference would add to run `mi_analyze` of femmcli on the B200 backend
// (INTEGRATION.md): a class with the name and the handful of members cfemm/femmcli/LuaMagneticsCommands.cpp:817-842
// touches on `FSolver` — PathName, WarnMessage, PrintMessage, previousSolutionFile, LoadProblemFile(), runSolver() —
// forwarding to the C ABI of libxfemm_b200.so (include/xfemm_b200.h).  oracle/Makefile force-includes this header
// in place of cfemm/fsolver/fsolver.h (masked through its include guard) when it recompiles that ONE reference
// source, unmodified, into oracle/_ref/femmcli_b200.  Test infrastructure: lets the reference's own Lua scripts
// (mi_analyze + mo_* post-processing) drive the GPU solver end to end.
#ifndef XFEMM_B200_FSOLVER_FEMMCLI_ADAPTER_H
#define XFEMM_B200_FSOLVER_FEMMCLI_ADAPTER_H
#ifndef FSOLVER_H
#define FSOLVER_H
#endif
#include <cstddef>
#include <string>

struct xfemm_fsolver;

class FSolver
{
public:
    struct Sized { std::size_t count = 0; std::size_t size() const { return count; } }; // only read inside assert()
    std::string PathName, previousSolutionFile;
    int (*WarnMessage)(const char *, ...);
    int (*PrintMessage)(const char *, ...);
    int ACSolver = 0;
    double Frequency = 0;
    Sized lineproplist, nodeproplist, blockproplist, circproplist, labellist;

    FSolver();
    ~FSolver();
    bool LoadProblemFile();          // xfemm_fsolver_open
    bool runSolver(bool verbose = false); // load_mesh, renumber, solve, write (fsolver.cpp:1196-1339)

private:
    xfemm_fsolver *handle;
};
#endif
