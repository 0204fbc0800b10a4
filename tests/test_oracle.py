"""The oracle restatement (oracle/oracle_core.c + oracle/fsolver_oracle.py) against the golden
vectors produced by the UNMODIFIED reference (oracle/_ref; see tests/golden/make_golden.py).
This is what pins the oracle: renumbering, element order, boundary tags, the assembled +
boundary-conditioned system, SSOR-PCG iteration counts and the solution."""
import tempfile

import numpy as np
import pytest

from oracle import fsolver_oracle as O
from oracle.age_matrix import age_matrix
from util import load_ref, rel_l2, stage_case


@pytest.mark.parametrize("name", ["tb", "tb30", "temp1"])
def test_oracle_matches_reference_golden(name):
    ref, meta = load_ref(name)
    o = O.solve_static2d(stage_case(name, tempfile.mkdtemp()))
    M = o["mesh"]
    # integer work: bit exact
    assert M["bandwidth"] == meta["bandwidth"]
    assert np.array_equal(M["xy"], ref["nodes"])
    assert np.array_equal(M["node_bc"], ref["node_bc"])
    assert np.array_equal(M["tri"], ref["elems"][:, :3])
    assert np.array_equal(M["lbl"], ref["elems"][:, 3])
    assert np.array_equal(M["e"], ref["elems"][:, 5:8])
    assert np.array_equal(M["pbc"], ref["pbc"])
    assert o["iters"] == meta["iters"]
    last = meta["solves"] - 1
    for step in sorted({0, last}):
        s = o["systems"][step]
        assert np.array_equal(s["rowptr"], ref[f"step{step}_rowptr"])
        assert np.array_equal(s["col"], ref[f"step{step}_col"])
        scale = np.abs(ref[f"step{step}_val"]).max()
        assert np.abs(s["val"] - ref[f"step{step}_val"]).max() <= 1e-12 * scale
        assert np.abs(s["rhs"] - ref[f"step{step}_rhs"]).max() <= 1e-12 * np.abs(ref[f"step{step}_rhs"]).max()
    assert rel_l2(o["V"], ref["V"]) <= 1e-12
    assert rel_l2(o["A"], ref["A"]) <= 1e-12
    refmu = ref["mu"][:, [0, 2]]
    assert np.abs(o["mu"] - refmu).max() <= 1e-9 * np.abs(refmu).max()


def test_oracle_sparse_primitives():
    """Put/Get/AddTo/SetValue/Periodicity semantics of spars.cpp on a hand-made 5x5 case."""
    import ctypes as C
    L = O.lib()
    m = L.orc_create(5, 0)
    for i in range(5):
        L.orc_put(m, 4.0, i, i)
    for i in range(4):
        L.orc_addto(m, -1.0, i + 1, i)       # swapped indices land in the upper triangle
    assert L.orc_get(m, 0, 1) == -1.0 and L.orc_get(m, 1, 0) == -1.0 and L.orc_get(m, 0, 3) == 0.0
    b = np.ctypeslib.as_array(L.orc_b(m), shape=(5,))
    b[:] = 1.0
    L.orc_setvalue(m, 0, 2.0)                 # b[1] -= A10*2 ; A10 = 0 ; b[0] = A00*2
    assert L.orc_get(m, 0, 1) == 0.0 and b[1] == 3.0 and b[0] == 8.0
    L.orc_periodicity(m, 1, 4, 0)             # rows 2 and 3 couple to 1 / 4; fill (2,4) and (1,3) appears
    assert L.orc_get(m, 2, 1) == -0.5 and L.orc_get(m, 2, 4) == -0.5
    assert L.orc_get(m, 3, 4) == -0.5 and L.orc_get(m, 3, 1) == -0.5
    assert b[1] == b[4] == 2.0
    L.orc_periodicity(m, 1, 4, 1)             # antiperiodic: opposite signs
    assert L.orc_get(m, 2, 1) == -L.orc_get(m, 2, 4)
    L.orc_destroy(m)


def test_age_matrix_symmetric_and_null_space():
    """The air-gap cell matrix is symmetric and annihilates constants for every shift pair
    (a constant potential has no flux) — a property check of the 55 closed forms."""
    rng = np.random.default_rng(0)
    for _ in range(20):
        ci, co, K = rng.random(), rng.random(), rng.uniform(0.2, 5.0)
        MG = age_matrix(ci, co, K)
        assert np.allclose(MG, MG.T)
        assert np.abs(MG @ np.ones(10)).max() < 1e-12
