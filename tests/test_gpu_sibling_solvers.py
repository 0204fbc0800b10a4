"""The reference's heat-flow and electrostatics solvers on the B200 CBigLinProb.

hsolver and esolver share `CBigLinProb` with fsolver (cfemm/hsolver/hsolver.cpp:458-860, cfemm/esolver/esolver.cpp:389-640:
Create, Put/AddTo per element, L.Q flags, SetValue, Periodicity/AntiPeriodicity, PCGSolve).  oracle/Makefile compiles
their UNMODIFIED sources twice: against the reference's own spars.cpp (`hsolver`, `esolver`: the CPU oracle) and against
this repo's host/spars.h + libxfemm_b200.so (`hsolver_b200`, `esolver_b200`: every PCGSolve on the GPU).  Both are run
on the reference's own test inputs (cfemm/hsolver/test/Temp0.feh, Temp1.feh, cfemm/esolver/test/test.fee, copied to
tests/golden/siblings/) and the nodal solutions of the answer files are compared."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _nodal_solution(path):
    lines = open(path).read().splitlines()
    i = next(k for k, ln in enumerate(lines) if ln.strip() == "[Solution]")
    n = int(lines[i + 1])
    return np.array([[float(x) for x in ln.split()[:3]] for ln in lines[i + 2:i + 2 + n]])


@pytest.mark.parametrize("solver,bases,ext,out", [("hsolver", ["Temp0", "Temp1"], ".feh", ".anh"), ("esolver", ["test"], ".fee", ".res")])
def test_reference_element_loops_on_the_gpu_backend(solver, bases, ext, out, tmp_path):
    for exe in (solver, solver + "_b200", "fmesher"):
        if not os.path.exists(os.path.join(REF, exe)):
            pytest.skip(f"oracle/_ref/{exe} not built (run make -C oracle where /root/reference is available)")
    sols = {}
    for exe in (solver, solver + "_b200"):
        d = tmp_path / exe
        d.mkdir()
        # Temp1.feh is a time step that starts from Temp0.anh ([PrevSoln], [dT]): solved in sequence in one directory
        for base in bases:
            shutil.copy(os.path.join(ROOT, "tests", "golden", "siblings", base + ext), d)
            subprocess.run([os.path.join(REF, "fmesher"), base + ext], cwd=d, check=True, stdout=subprocess.DEVNULL)
            r = subprocess.run([os.path.join(REF, exe), base], cwd=d, capture_output=True, text=True)
            assert os.path.exists(d / (base + out)), r.stdout[-500:] + r.stderr[-500:]
            sols[exe, base] = _nodal_solution(d / (base + out))
    for base in bases:
        a, b = sols[solver + "_b200", base], sols[solver, base]
        assert a.shape == b.shape and np.array_equal(a[:, :2], b[:, :2])  # same nodes in the same order
        err = np.linalg.norm(a[:, 2] - b[:, 2]) / np.linalg.norm(b[:, 2])
        print(solver, base, len(a), "nodes, relative L2 difference of the nodal solution", err)
        assert err <= 1e-6  # the bar of the path (nodal solution within 1e-6 relative L2); both solves stop at Precision 1e-8
