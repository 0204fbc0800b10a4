"""GPU parity: time-harmonic formulations (complex-symmetric COCG) through the C ABI against dumps of
the unmodified reference on authored fixtures (the reference ships no harmonic fixture).

tbh     TorqueBenchmark at 50 Hz: air-gap element (harmonic sign), 360 periodic pairs, eddy currents
harm2d  planar solenoid section at 50 Hz: B-H steel with hysteresis angle (successive approximation,
        65 passes in the reference), conducting magnet, wound coil (Case 1), solid bar with
        prescribed total current (Case 2: circuit unknown appended to the system)
harmaxi the same axisymmetric (nodes on the axis, 1/R circuit terms)
"""
import numpy as np
import pytest

import xfemm_b200 as X
from util import load_ref, rel_l2, stage_case, upper_from_full

pytestmark = pytest.mark.gpu


def _solve(name, tmp_path, precond):
    S = X.FSolver(stage_case(name, tmp_path))
    S.set_option("keep_system", 1)
    S.set_option("precond", precond)
    S.load_mesh(False)
    S.renumber(False)
    S.solve()
    return S


@pytest.mark.parametrize("precond", [X.PRECOND_JACOBI, X.PRECOND_MCSSOR])
def test_harmonic_linear_age_periodic(precond, tmp_path):
    ref, meta = load_ref("tbh")
    S = _solve("tbh", tmp_path, precond)
    rowptr, col, vals, rhs = S.system()
    urp, ucol, uval = upper_from_full(rowptr, col, vals)
    assert np.array_equal(urp, ref["step0_rowptr"]) and np.array_equal(ucol, ref["step0_col"])
    scale = np.abs(ref["step0_val"]).max()
    assert np.abs(uval - ref["step0_val"]).max() <= 1e-11 * scale
    assert np.abs(rhs - ref["step0_rhs"]).max() <= 1e-11 * np.abs(ref["step0_rhs"]).max()
    V, A = S.solution()
    st = S.stats()
    print("tbh precond", precond, "iters", st["iters"], "relres", st["relres"], "relL2(A)", rel_l2(A, ref["A"]))
    assert st["newton_steps"] == 1
    assert rel_l2(V, ref["V"]) <= 1e-6 and rel_l2(A, ref["A"]) <= 1e-6


@pytest.mark.parametrize("name", ["harm2d", "harmaxi"])
def test_harmonic_nonlinear_circuits(name, tmp_path):
    ref, meta = load_ref(name)
    S = _solve(name, tmp_path, X.PRECOND_JACOBI)
    st = S.stats()
    V, A = S.solution()
    n_nodes = meta["n_nodes"]
    print(name, "passes", st["newton_steps"], "reference", meta["solves"], "COCG iters (first, last)", st["iters"][0], st["iters"][-1],
          "relL2(A nodes)", rel_l2(A[:n_nodes], ref["A"][:n_nodes]), "circuit unknowns", A[n_nodes:], ref["A"][n_nodes:])
    rowptr, col, vals, rhs = S.system()
    urp, ucol, uval = upper_from_full(rowptr, col, vals)
    last = meta["solves"] - 1
    assert np.array_equal(urp, ref[f"step{last}_rowptr"]) and np.array_equal(ucol, ref[f"step{last}_col"])
    # successive approximation stops when the relative change drops below 100*Precision = 1e-6; the
    # number of passes may differ by one when the change sits at the threshold
    assert abs(st["newton_steps"] - meta["solves"]) <= 1
    scale = np.abs(ref[f"step{last}_val"]).max()
    assert np.abs(uval - ref[f"step{last}_val"]).max() <= 1e-4 * scale
    # both runs are converged to ~1e-6/(1-rho) of the fixed point of the iteration (rho = contraction
    # factor of the successive approximation), so that is the meaningful bar for the difference
    assert rel_l2(A[:n_nodes], ref["A"][:n_nodes]) <= 1e-6
    assert np.allclose(A[n_nodes:], ref["A"][n_nodes:], rtol=1e-4, atol=1e-12)
    mu = S.mu()
    assert np.abs(mu - ref["mu"]).max() <= 1e-3 * np.abs(ref["mu"]).max()


def test_multigrid_preconditioner_on_a_problem_without_eddy_currents(tmp_path):
    """XFEMM_PRECOND_AMG for the complex solver: the V-cycle of the real matrix |Re A| + |Im A|, applied to the real
    and the imaginary part of the residual.  Same answer, an order of magnitude fewer COCG iterations on tbh."""
    ref, meta = load_ref("tbh")
    its = {}
    for pc in (X.PRECOND_JACOBI, X.PRECOND_AMG):
        d = tmp_path / str(pc)
        d.mkdir()
        S = X.FSolver(stage_case("tbh", str(d)))
        S.set_option("precond", pc)
        S.load_mesh(False)
        S.renumber(False)
        S.solve()
        V, A = S.solution()
        assert rel_l2(A, ref["A"]) <= 1e-6
        its[pc] = sum(S.stats()["iters"])
        S.close()
    print("COCG iterations: jacobi", its[X.PRECOND_JACOBI], "multigrid", its[X.PRECOND_AMG])
    assert its[X.PRECOND_AMG] * 4 < its[X.PRECOND_JACOBI]


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["harm2d", "harmaxi"])
def test_multigrid_with_eddy_currents_and_circuit_unknowns(case, tmp_path):
    """Conductors add j*omega*sigma*mass to the matrix: the companion matrix of the V-cycle is stiffness + omega*sigma*mass.
    Circuit unknowns (rows appended after the mesh nodes) are left out of it and keep the Jacobi step, so the
    preconditioner stays symmetric.  COCG converges without the Jacobi fallback, in several times fewer iterations
    than the Jacobi run, to the same answer."""
    ref, meta = load_ref(case)
    n_nodes = len(ref["nodes"])
    its = {}
    for pc in (X.PRECOND_JACOBI, X.PRECOND_AMG):
        d = tmp_path / str(pc)
        d.mkdir()
        S = X.FSolver(stage_case(case, str(d)))
        S.set_option("precond", pc)
        S.load_mesh(False)
        S.renumber(False)
        S.solve()
        V, A = S.solution()
        assert rel_l2(A[:n_nodes], ref["A"][:n_nodes]) <= 1e-6
        assert np.allclose(A[n_nodes:], ref["A"][n_nodes:], rtol=1e-4, atol=1e-12)
        its[pc] = sum(S.stats()["iters"])
        S.close()
    print(case, "COCG iterations: jacobi", its[X.PRECOND_JACOBI], "multigrid", its[X.PRECOND_AMG])
    assert its[X.PRECOND_AMG] * 4 < its[X.PRECOND_JACOBI]


def _dense_upper(n, rowptr, col, val):
    import scipy.sparse as sp
    return sp.csr_matrix((val, col, rowptr), shape=(n, n))


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["harm2dn", "harmaxin", "tbhn"])
def test_newton_ac_solver(case, tmp_path):
    """[ACSolver]=1: Newton passes with the side matrices Mh/Ms/Ma (assembled on the device next to M) and KludgeSolve.
    Same number of passes as the reference, its last system (M and the three side matrices, upper triangle as the
    reference stores them) and its answer."""
    import scipy.sparse as sp
    ref, meta = load_ref(case)
    n_nodes, n = meta["n_nodes"], meta["n_unknowns"]
    S = _solve(case, tmp_path, X.PRECOND_AMG)
    st = S.stats()
    V, A = S.solution()
    print(case, "Newton passes", st["newton_steps"], "reference", meta["solves"], "COCG iterations per pass", st["iters"],
          "relL2(A nodes)", rel_l2(A[:n_nodes], ref["A"][:n_nodes]))
    assert abs(st["newton_steps"] - meta["solves"]) <= 1
    assert rel_l2(A[:n_nodes], ref["A"][:n_nodes]) <= 1e-6
    assert np.allclose(A[n_nodes:], ref["A"][n_nodes:], rtol=1e-4, atol=1e-12)
    if st["newton_steps"] != meta["solves"]:
        return
    last = meta["solves"] - 1
    rowptr, col, vals, rhs = S.system()
    full = sp.csr_matrix((vals, col, rowptr), shape=(n, n))
    refM = sp.csr_matrix((ref[f"step{last}_val"], ref[f"step{last}_col"], ref[f"step{last}_rowptr"]), shape=(n, n))
    scale = np.abs(ref[f"step{last}_val"]).max()
    assert abs(sp.triu(full) - refM).max() <= 1e-4 * scale
    for k in (1, 2, 3):
        mine = sp.triu(sp.csr_matrix((S.side_matrix(k), col, rowptr), shape=(n, n)))
        theirs = sp.csr_matrix((ref[f"step{last}_side{k}_val"], ref[f"step{last}_side{k}_col"], ref[f"step{last}_side{k}_rowptr"]), shape=(n, n))
        sc = max(np.abs(ref[f"step{last}_side{k}_val"]).max(), 1e-300)
        err = abs(mine - theirs).max() / sc
        print("   side matrix", k, "max |ref|", sc, "rel err", err)
        assert err <= 1e-4
        if len(ref["pbc"]) and theirs.nnz:
            # rows of the periodic nodes (entries there are small: judged on their own scale), including the entry
            # coupling the two nodes of a pair that Periodicity/AntiPeriodicity creates in the side matrices
            rows = np.unique(ref["pbc"][:, :2])
            tfull = theirs + {1: sp.triu(theirs, 1).conj().T, 2: sp.triu(theirs, 1).T, 3: -sp.triu(theirs, 1).conj().T}[k]
            mfull = sp.csr_matrix((S.side_matrix(k), col, rowptr), shape=(n, n))
            scp = abs(tfull[rows]).max()
            if scp == 0:
                assert abs(mfull[rows]).max() == 0
                continue
            errp = abs(mfull[rows] - tfull[rows]).max() / scp
            pair = max(abs(mfull[i, j] - tfull[i, j]) for i, j, _ in ref["pbc"]) / scp
            print("      periodic rows: max |ref|", scp, "rel err", errp, "pair entries rel err", pair)
            assert errp <= 1e-3 and pair <= 1e-3
        # symmetry class on the full storage: Hermitian / complex symmetric / anti-Hermitian
        G = sp.csr_matrix((S.side_matrix(k), col, rowptr), shape=(n, n))
        other = {1: G.conj().T, 2: G.T, 3: -G.conj().T}[k]
        assert abs(G - other).max() <= 1e-12 * sc


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["harm2d", "harm2dn", "harmaxin"])
def test_harmonic_answer_file_through_the_reference_postprocessor(case, tmp_path):
    """The .ans written by this repo's command line (WriteHarmonic2D) is loaded by the REFERENCE post-processor
    (oracle/_ref/femmcli): circuit properties (current, voltage drop, flux linkage — complex) and block integrals
    (A.J, stored energy, hysteresis / lamination losses) agree with those of the reference solver's own .ans within 1e-4."""
    import os, re, shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    femmcli, ref_solver = os.path.join(root, "oracle", "_ref", "femmcli"), os.path.join(root, "oracle", "_ref", "fsolver")
    ours = os.path.join(root, "xfemm_b200", "fsolver_b200")
    if not (os.path.exists(femmcli) and os.path.exists(ref_solver)):
        pytest.skip("oracle/_ref not built")
    lua = ('open("%s.fem")\nmi_loadsolution()\n'
           'for _, c in {"coil", "bar"} do\n  i_, v_, f_ = mo_getcircuitproperties(c)\n'
           '  write(format("VAL %%.12g %%.12g %%.12g %%.12g %%.12g %%.12g\\n", re(i_), im(i_), re(v_), im(v_), re(f_), im(f_)))\nend\n'
           'mo_groupselectblock()\nfor _, k in {0, 2, 3} do\n  e_ = mo_blockintegral(k)\n  write(format("VAL %%.12g %%.12g\\n", re(e_), im(e_)))\nend\nwrite("DONE\\n")\n')
    vals = {}
    for who, solver in (("ours", ours), ("reference", ref_solver)):
        d = tmp_path / who
        d.mkdir()
        base = stage_case(case, str(d))
        r = subprocess.run([solver, base], cwd=d, capture_output=True, text=True)
        assert os.path.exists(base + ".ans"), r.stdout[-500:] + r.stderr[-500:]
        (d / "post.lua").write_text(lua % os.path.basename(base))
        r = subprocess.run([femmcli, "--lua-script=post.lua"], cwd=d, capture_output=True, text=True)
        assert "DONE" in r.stdout, r.stdout[-800:] + r.stderr[-800:]
        vals[who] = np.array([float(x) for ln in r.stdout.splitlines() if ln.startswith("VAL") for x in ln.split()[1:]])
    a, b = vals["ours"], vals["reference"]
    assert a.shape == b.shape and len(a) == 18 and np.all(np.isfinite(b))
    print(case, "reference post-processor on our .ans:", a, "\n   on the reference's:", b)
    assert np.all(np.abs(a - b) <= 1e-4 * np.maximum(np.abs(b), 1e-12 + 1e-6 * np.abs(b).max()))
