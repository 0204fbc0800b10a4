// Drives the host facade of CBigLinProb the way hsolver/esolver drive the reference class
// (cfemm/hsolver/hsolver.cpp:458-860: Create, AddTo/Put per element, SetValue, Periodicity /
// AntiPeriodicity, PCGSolve) — the element loop stays on the host, the solve runs on the GPU.
// Prints the solution, one value per line; tests/test_gpu_facade.py replays the same calls on the
// pinned CPU restatement (oracle/oracle_core.c) and compares.
#include "spars.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>

int main(int argc, char **argv)
{
    const int N = argc > 1 ? atoi(argv[1]) : 24; // N x N grid, 5-point stencil
    const int n = N * N;
    CBigLinProb L;
    L.Precision = 1e-10;
    L.Create(n, N + 1);
    for (int j = 0; j < N; j++)
        for (int i = 0; i < N; i++)
        {
            const int k = j * N + i;
            L.AddTo(4.0 + 0.01 * (k % 7), k, k);
            if (i + 1 < N) L.AddTo(-1.0, k, k + 1);
            if (j + 1 < N) L.AddTo(-1.0 - 0.001 * (k % 5), k, k + N);
            L.b[k] = sin(0.37 * k) + 0.1;
        }
    // fixed values on the left column (SetValue twice on one node exercises first/last semantics)
    for (int j = 0; j < N; j++) L.SetValue(j * N, 0.1 * j);
    L.SetValue(0, 0.5);
    // bottom row tied to the top row: periodic for even columns, antiperiodic for odd ones
    for (int i = 1; i < N - 1; i++)
    {
        if (i % 2 == 0) L.Periodicity(i, (N - 1) * N + i);
        else L.AntiPeriodicity(i, (N - 1) * N + i);
    }
    if (argc > 2)
    {
        // host logic only (no device needed): every entry through Get() and the right-hand side after the boundary
        // conditions; tests/test_host_parity.py compares this build with the one against the unmodified reference class
        for (int p = 0; p < n; p++)
            for (int q = 0; q < n; q++) printf("%.17g\n", L.Get(p, q));
        printf("#\n");
        for (int k = 0; k < n; k++) printf("%.17g\n", L.b[k]);
        return 0;
    }
    if (!L.PCGSolve(0)) { fprintf(stderr, "PCGSolve failed\n"); return 2; }
    for (int k = 0; k < n; k++) printf("%.17g\n", L.V[k]);
    // second solve with a changed right-hand side and a warm start (Newton-style reuse)
    for (int k = 0; k < n; k++) L.b[k] *= 1.5;
    if (!L.PCGSolve(1)) { fprintf(stderr, "PCGSolve (warm) failed\n"); return 2; }
    printf("#\n");
    for (int k = 0; k < n; k++) printf("%.17g\n", L.V[k]);
#ifdef XFEMM_B200_SPARS_H /* statistics only this repo's class keeps */
    fprintf(stderr, "iterations %d relres %.3e\n", L.lastIterations, L.lastRelRes);
#endif
    return 0;
}
