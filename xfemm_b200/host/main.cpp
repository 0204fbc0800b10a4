// fsolver command line: `fsolver <basename-without-.fem>`; exit code 0 ok, 1 the .fem could not
// be loaded, 2 the solve failed (reference: cfemm/fsolver/main.cpp:39-84).
#include "fsolver.h"
#include <cstdio>
#include <cstring>

int main(int argc, char **argv)
{
    FSolver theFSolver;
    char PathName[512];
    if (argc < 2)
    {
        printf("Enter fem file name without extension:\n");
        if (!fgets(PathName, 512, stdin)) return 1;
        char *pos;
        if ((pos = strchr(PathName, '\n')) != NULL) *pos = '\0';
    }
    else if (argc > 2)
    {
        printf("Too many arguments");
        return 1;
    }
    else
    {
        strncpy(PathName, argv[1], 511);
        PathName[511] = 0;
    }
    theFSolver.PathName = PathName;
    if (theFSolver.LoadProblemFile() != true)
    {
        theFSolver.WarnMessage("problem loading .fem file\n");
        return 1;
    }
    if (!theFSolver.runSolver(true)) return 2;
    return 0;
}
