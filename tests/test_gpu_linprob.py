"""xfemm_linprob_* (the CBigLinProb backend, spars.cpp:80-474) on synthetic matrices through the C ABI:
MultA / MultPC / PCGSolve against scipy, including the layouts the mesh cases do not stress —
rows far wider than a SELL slice (air-gap rings with the angle-independent pattern, circuit rows),
ragged sizes that are not multiples of 32, and a 1x1 system."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import xfemm_b200 as X

pytestmark = pytest.mark.gpu


def _matrix(n, wide_rows=(), wide_width=0, seed=0):
    """SPD, narrow band + optional very wide rows (symmetric)."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for i in range(n):
        for j in range(max(0, i - 3), i):
            if rng.random() < 0.8:
                v = -rng.random()
                rows += [i, j]
                cols += [j, i]
                vals += [v, v]
    for r in wide_rows:
        others = rng.choice(np.setdiff1d(np.arange(n), [r]), size=min(wide_width, n - 1), replace=False)
        for j in others:
            v = -0.01 * rng.random()
            rows += [r, j]
            cols += [j, r]
            vals += [v, v]
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    d = np.asarray(abs(A).sum(1)).ravel() + 1.0
    A = (A + sp.diags(d)).tocsr()
    A.sort_indices()
    return A


CASES = [
    ("band_1000", dict(n=1000)),
    ("ragged_77", dict(n=77)),
    ("one_wide_row", dict(n=3000, wide_rows=(1500,), wide_width=700)),
    ("wide_rows_multi_segment", dict(n=5000, wide_rows=(0, 17, 2500, 4999), wide_width=1000)),
    ("wide_cluster", dict(n=4000, wide_rows=tuple(range(100, 164)), wide_width=150)),
]


@pytest.mark.parametrize("name,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("precond", [X.PRECOND_JACOBI, X.PRECOND_MCSSOR, X.PRECOND_AMG])
def test_multa_multpc_pcg(name, kw, precond):
    A = _matrix(**kw)
    n = A.shape[0]
    rng = np.random.default_rng(1)
    L = X.LinProb(A.indptr, A.indices)
    L.set_values(A.data)
    x = rng.standard_normal(n)
    y = L.multa(x)
    assert np.abs(y - A @ x).max() <= 1e-12 * np.abs(A @ x).max()
    # the preconditioner must be symmetric (CG needs it): x.M^-1 y == y.M^-1 x
    z = rng.standard_normal(n)
    a, b = x @ L.multpc(z, precond), z @ L.multpc(x, precond)
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b), 1e-300)
    rhs = rng.standard_normal(n)
    L.set_rhs(rhs)
    info = L.pcg(1e-12, precond=precond)
    V = L.get_solution()
    ref = spla.spsolve(A.tocsc(), rhs)
    assert np.linalg.norm(V - ref) <= 1e-9 * np.linalg.norm(ref)
    assert info.iterations > 0 and info.relres < 1e-9
    # warm start from the solution: converged at once
    info2 = L.pcg(1e-8, warm=True, precond=precond)
    assert info2.iterations <= 2
    L.close()


def test_one_by_one_and_zero_rhs():
    L = X.LinProb(np.array([0, 1]), np.array([0]))
    L.set_values(np.array([4.0]))
    L.set_rhs(np.array([2.0]))
    L.pcg(1e-12, precond=X.PRECOND_JACOBI)
    assert abs(L.get_solution()[0] - 0.5) < 1e-15
    L.set_rhs(np.array([0.0]))
    info = L.pcg(1e-12, precond=X.PRECOND_JACOBI)
    # spars.cpp:255-257 returns on res_o == 0 BEFORE V is zeroed: the previous solution stays
    assert info.iterations == 0 and L.get_solution()[0] == 0.5
    L.close()


@pytest.mark.parametrize("precond", [X.PRECOND_JACOBI, X.PRECOND_MCSSOR, X.PRECOND_AMG])
def test_degenerate_systems(precond):
    """Diagonal matrix (no couplings at all), all rows fixed by SetValue, and an empty periodic list."""
    n = 100
    rowptr = np.arange(n + 1, dtype=np.int32)
    col = np.arange(n, dtype=np.int32)
    L = X.LinProb(rowptr, col)
    d = 1.0 + np.arange(n)
    L.set_values(d)
    b = np.cos(np.arange(n))
    L.set_rhs(b)
    L.pcg(1e-12, precond=precond)
    assert np.abs(L.get_solution() - b / d).max() < 1e-12
    L.close()
    # every row fixed: the solve returns the prescribed values
    A = _matrix(n=200)
    L = X.LinProb(A.indptr, A.indices)
    L.set_values(A.data)
    L.set_rhs(np.ones(200))
    vals = np.linspace(-1, 1, 200)
    L.dirichlet(np.arange(200, dtype=np.int32), vals)
    L.periodic(np.zeros((0, 3), dtype=np.int32))
    L.pcg(1e-12, precond=precond)
    assert np.abs(L.get_solution() - vals).max() < 1e-10
    L.close()


def test_pattern_build_on_the_device_equals_the_host_build():
    """xfemm_pattern_build_device (mesh graph by a radix sort of (row, column) keys on the GPU; used by FSolver) gives
    the same CSR as the host build: a fixture with air-gap blocks and periodic fill, a random triangle soup with
    repeated and degenerate pairs, isolated nodes, and the error for an index out of range."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from util import load_ref
    ref, meta = load_ref("tb")
    tri = ref["elems"][:, :3]
    for kw in (dict(pbc=ref["pbc"][:, :2]), dict(), dict(pbc=ref["pbc"][:40, :2], extra=np.array([[0, 4891], [17, 3000]]), with_fill=True)):
        h = X.pattern_build(meta["n_nodes"], tri, **kw)
        d = X.pattern_build(meta["n_nodes"], tri, device=True, **kw)
        assert all(np.array_equal(a, b) for a, b in zip(h, d))
    rng = np.random.default_rng(11)
    n = 5000
    soup = rng.integers(0, n - 100, size=(20000, 3)).astype(np.int32)  # the last 100 nodes stay isolated
    soup[::7, 1] = soup[::7, 0]  # degenerate elements (a node twice)
    h = X.pattern_build(n, soup)
    d = X.pattern_build(n, soup, device=True)
    assert np.array_equal(h[0], d[0]) and np.array_equal(h[1], d[1])
    assert np.all(np.diff(h[0])[-100:] == 1)
    soup[123, 2] = n + 5
    with pytest.raises(X.XfemmError):
        X.pattern_build(n, soup, device=True)
