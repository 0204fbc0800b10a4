// Drives CBigComplexLinProb through its host-assembly interface (cfemm/libfemm/cspars.h:37-95): Create, AddTo / Put with
// the selector k (M and the Newton side matrices Mh, Ms, Ma), SetValue, Periodicity / AntiPeriodicity, MultA /
// MultConjA, PBCGSolveMod (which dispatches to KludgeSolve once side matrices exist).  The SAME source is compiled
// twice: against this repo's host facade (solves on the GPU) and against the unmodified reference class
// (oracle/Makefile -> oracle/_ref/cfacade_ref, CPU); tests/test_gpu_facade.py compares the printed vectors.
#include "femmcomplex.h"
#include "cspars.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

static void dump(const char *tag, CComplex *v, int n)
{
    printf("# %s\n", tag);
    for (int k = 0; k < n; k++) printf("%.17g %.17g\n", v[k].re, v[k].im);
}

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 16;     // N x N grid, 5-point stencil
    const int newton = argc > 2 ? atoi(argv[2]) : 0; // 1: add side matrices; a third argument: dump the matrices, do not solve
    const int nodes = N * N, n = nodes + 2;          // two appended circuit unknowns
    CBigComplexLinProb L;
    L.Precision = 1e-10;
    L.Create(n, N + 1, nodes);
    for (int j = 0; j < N; j++)
        for (int i = 0; i < N; i++)
        {
            const int k = j * N + i;
            L.AddTo(CComplex(-(4.0 + 0.01 * (k % 7)), -(0.3 + 0.02 * (k % 3))), k, k);
            if (i + 1 < N) L.AddTo(CComplex(1.0, 0.05), k, k + 1);
            if (j + 1 < N) L.AddTo(CComplex(1.0 + 0.001 * (k % 5), 0.02), k, k + N);
            L.b[k] = CComplex(sin(0.37 * k) + 0.1, cos(0.11 * k));
        }
    // circuit unknown coupled to a block of nodes (Case 2 of harmonic2d.cpp:540-549), second one decoupled
    for (int j = N / 2; j < N / 2 + 3; j++)
        for (int i = N / 2; i < N / 2 + 3; i++)
        {
            const int k = j * N + i;
            L.Put(L.Get(k, nodes) + CComplex(0, -0.07), k, nodes);
            L.Put(L.Get(nodes, nodes) + CComplex(0, -0.21), nodes, nodes);
        }
    L.b[nodes] = CComplex(1.5, 0);
    L.Put(L.Get(0, 0), nodes + 1, nodes + 1);
    if (newton)
    {
        // small side matrices on a patch that touches fixed and periodic nodes; both (p,q) orders exercise the
        // conjugation rules of Put/Get (cspars.cpp:152-158, 285-296)
        for (int j = 0; j < N - 1; j++)
            for (int i = 0; i < 5; i++)
            {
                const int p = j * N + i, q = p + 1, r = p + N;
                const double s = 0.02 + 0.001 * ((p * 7) % 11);
                L.Put(L.Get(p, p, 1) + CComplex(-2 * s, 0), p, p, 1);
                L.Put(L.Get(q, p, 1) + CComplex(s, 0.5 * s), q, p, 1);
                L.Put(L.Get(p, r, 1) + CComplex(0.7 * s, -0.3 * s), p, r, 1);
                L.Put(L.Get(p, p, 2) + CComplex(-s, 0.4 * s), p, p, 2);
                L.Put(L.Get(r, p, 2) + CComplex(0.5 * s, 0.2 * s), r, p, 2);
                L.Put(L.Get(p, p, 3) + CComplex(0, 0.3 * s), p, p, 3);
                L.Put(L.Get(q, p, 3) + CComplex(0.2 * s, 0.1 * s), q, p, 3);
            }
    }
    // fixed values on the left column (SetValue twice on one node: first/last value semantics)
    for (int j = 0; j < N; j++) L.SetValue(j * N, CComplex(0.1 * j, -0.05 * j));
    L.SetValue(0, CComplex(0.5, 0.25));
    // bottom row tied to the top row: periodic for even columns, antiperiodic for odd ones
    for (int i = 1; i < N - 1; i++)
    {
        if (i % 2 == 0) L.Periodicity(i, (N - 1) * N + i);
        else L.AntiPeriodicity(i, (N - 1) * N + i);
    }
    if (argc > 3)
    {
        // host logic only (no device needed): every entry of the four matrices through Get() — both index orders,
        // i.e. including the conjugation rules — and the right-hand side after the boundary conditions
        for (int h = 0; h <= 3; h++)
        {
            printf("# Get k=%d\n", h);
            for (int p = 0; p < n; p++)
                for (int q = 0; q < n; q++)
                {
                    const CComplex v = L.Get(p, q, h);
                    printf("%.17g %.17g\n", v.re, v.im);
                }
        }
        dump("b", L.b, n);
        return 0;
    }
    // matrix-vector products after the boundary conditions
    CComplex *x = (CComplex *)calloc(n, sizeof(CComplex)), *y = (CComplex *)calloc(n, sizeof(CComplex));
    for (int k = 0; k < n; k++) x[k] = CComplex(cos(0.3 * k), sin(0.7 * k) - 0.2);
    L.MultA(x, y); dump("MultA", y, n);
    L.MultConjA(x, y); dump("MultConjA", y, n);
    if (newton)
    {
        for (int h = 1; h <= 3; h++) { L.MultA(x, y, h); dump("MultA side", y, n); }
        L.MultConjA(x, y, 2); dump("MultConjA 2", y, n);
        L.MultA(x, y, -1); dump("MultA full", y, n);
    }
    if (!L.PBCGSolveMod(0)) { fprintf(stderr, "PBCGSolveMod failed\n"); return 2; }
    dump("V cold", L.V, n);
    // the Newton solver leaves a corrected right-hand side in b (cspars.cpp:1029): set it again
    for (int k = 0; k < nodes; k++) L.b[k] = CComplex(1.5 * (sin(0.37 * k) + 0.1), 1.5 * cos(0.11 * k));
    for (int j = 0; j < N; j++) L.b[j * N] = L.Get(j * N, j * N) * CComplex(0.1 * j, -0.05 * j);
    L.b[0] = L.Get(0, 0) * CComplex(0.5, 0.25);
    L.b[nodes] = CComplex(2.0, 0.5);
    if (!L.PBCGSolveMod(1)) { fprintf(stderr, "PBCGSolveMod (warm) failed\n"); return 2; }
    dump("V warm", L.V, n);
    return 0;
}
