"""Newton iteration on a saturating problem at a size the fixtures do not reach (SURVEY §8d config 3,
scaled down to keep the suite short): the authored sat2d case re-meshed by the reference's fmesher to
~190 000 nodes, solved with the multigrid and with the Jacobi preconditioner.  The two runs share
nothing but the assembly, so agreement of the fields pins the multigrid path on strongly varying
coefficients (mu_r from 1 to a few thousand inside one problem)."""
import os
import subprocess

import numpy as np
import pytest

import xfemm_b200 as X

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FMESHER = os.path.join(ROOT, "oracle", "_ref", "fmesher")


def _set_mesh_size(text, size):
    out, n = [], 0
    for ln in text.replace("\r\n", "\n").split("\n"):
        if n > 0 and ln.strip():
            f = ln.split("\t")
            f[3] = repr(size)
            ln = "\t".join(f)
            n -= 1
        elif ln.strip().lower().startswith("[numblocklabels]"):
            n = int(ln.split("=")[1])
        out.append(ln)
    return "\n".join(out)


@pytest.mark.skipif(not os.path.exists(FMESHER), reason="oracle/_ref/fmesher not built")
def test_newton_at_scale_multigrid_vs_jacobi(tmp_path):
    text = open(os.path.join(ROOT, "tests", "golden", "sat2d", "sat2d.fem"), newline="").read()
    res = {}
    for name, precond in (("amg", X.PRECOND_AMG), ("jacobi", X.PRECOND_JACOBI)):
        d = tmp_path / name
        d.mkdir()
        with open(d / "big.fem", "w", newline="") as f:
            f.write(_set_mesh_size(text, 0.03))
        subprocess.run([FMESHER, "big.fem"], cwd=str(d), check=True, stdout=subprocess.DEVNULL)
        S = X.FSolver(str(d / "big"))
        S.set_option("precond", precond)
        S.load_mesh(False)
        S.renumber(False)
        S.solve()
        st = S.stats()
        res[name] = (S.solution()[1], st)
        print(name, "nodes", S.counts()["n_nodes"], "newton", st["newton_steps"], "iterations", st["iters"],
              "pcg ms", [round(x, 2) for x in st["solve_ms"]], "relres", st["relres"])
        S.close()
    a, sa = res["amg"]
    j, sj = res["jacobi"]
    assert len(a) > 150000
    assert sa["newton_steps"] == sj["newton_steps"]
    assert np.linalg.norm(a - j) <= 1e-6 * np.linalg.norm(j)
    # the point of the preconditioner: an order of magnitude fewer iterations
    assert sum(sa["iters"]) * 8 < sum(sj["iters"])
