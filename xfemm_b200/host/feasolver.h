// Source-compatibility shim: the reference's callers include "feasolver.h" for the FEASolver<> base
// (cfemm/libfemm/feasolver.h:76-209).  In this backend the base's public members live directly in
// class FSolver (fsolver.h); there is no template base to expose.
#ifndef XFEMM_B200_FEASOLVER_H
#define XFEMM_B200_FEASOLVER_H
#include "fsolver.h"
#endif
