// See fsolver_femmcli.h.
#include "fsolver_femmcli.h"
#include "../../include/xfemm_b200.h"
#include <cstdio>

static int quiet(const char *, ...) { return 0; }

FSolver::FSolver() : WarnMessage(&quiet), PrintMessage(&quiet), handle(nullptr) {}
FSolver::~FSolver() { if (handle) xfemm_fsolver_close(handle); }

bool FSolver::LoadProblemFile()
{
    if (handle) { xfemm_fsolver_close(handle); handle = nullptr; }
    if (!previousSolutionFile.empty())
    {
        WarnMessage("xfemm_b200: [PrevSoln] problems are not supported\n");
        return false;
    }
    if (xfemm_fsolver_open(PathName.c_str(), 0, &handle) != XFEMM_OK)
    {
        WarnMessage("%s\n", xfemm_last_error());
        handle = nullptr;
        return false;
    }
    return true;
}

bool FSolver::runSolver(bool)
{
    if (!handle) return false;
    const bool ok = xfemm_fsolver_load_mesh(handle, 1) == XFEMM_OK && xfemm_fsolver_renumber(handle, 1) == XFEMM_OK &&
                    xfemm_fsolver_solve(handle) == XFEMM_OK && xfemm_fsolver_write(handle) == XFEMM_OK;
    if (!ok) WarnMessage("%s\n", xfemm_last_error());
    return ok;
}
