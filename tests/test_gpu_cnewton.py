"""xfemm_clinprob_* with the Newton-AC side matrices (cspars.cpp:146-390, 983-1056) on synthetic systems through the
C ABI: matrix-vector products with the selector, and KludgeSolve against a dense solve of the real-linear system
M x + Mh x + Ms conj(x) + Ma x = b.  Edge cases: calls before xfemm_clinprob_enable_newton, zero right-hand side,
a one-row system, a size that is not a multiple of the slice width, a very wide (circuit) row."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

import xfemm_b200 as X

pytestmark = pytest.mark.gpu


class CProb:
    def __init__(self, A, n_nodes=None):
        self.A = A.tocsr()
        self.A.sort_indices()
        self.n = A.shape[0]
        self.h = C.c_void_p()
        rp = np.ascontiguousarray(self.A.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(self.A.indices, dtype=np.int32)
        rc = X.lib().xfemm_clinprob_create(self.n, self.n if n_nodes is None else n_nodes, rp, ci, C.byref(self.h))
        assert rc == 0, X.lib().xfemm_last_error()

    def call(self, name, *args):
        return getattr(X.lib(), "xfemm_clinprob_" + name)(self.h, *args)

    def ok(self, name, *args):
        rc = self.call(name, *args)
        assert rc == 0, (name, X.lib().xfemm_last_error())

    @staticmethod
    def ptr(a):
        return a.ctypes.data_as(C.c_void_p)

    def close(self):
        X.lib().xfemm_clinprob_destroy(self.h)


def _system(n, seed, wide_row=None, side_scale=0.003):
    """complex-symmetric M with a negative real diagonal (like the harmonic formulations) and small side matrices of
    the three symmetry classes on the same pattern"""
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    for i in range(n):
        for j in range(max(0, i - 3), i):
            if rng.random() < 0.8:
                rows.append(i), cols.append(j)
    if wide_row is not None:
        for j in rng.choice(np.setdiff1d(np.arange(n), [wide_row]), size=min(300, n - 1), replace=False):
            rows.append(max(wide_row, j)), cols.append(min(wide_row, j))
    L = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n)).tocsr()
    L.sum_duplicates()
    L.data[:] = 1.0
    L = sp.tril(L, -1).tocsr()
    m = L.nnz

    def off(scale):
        return sp.csr_matrix((scale * (rng.standard_normal(m) + 1j * rng.standard_normal(m)), L.indices, L.indptr), shape=(n, n))

    Lm = off(0.3)
    M = Lm + Lm.T
    M = M - sp.diags(np.asarray(abs(M).sum(1)).ravel() + 1.0 + 0.3j * rng.random(n))
    Lh, Ls, La = off(side_scale), off(side_scale), off(side_scale)
    H = Lh + Lh.conj().T + sp.diags(side_scale * rng.standard_normal(n))
    S = Ls + Ls.T + sp.diags(side_scale * (rng.standard_normal(n) + 1j * rng.standard_normal(n)))
    Am = La - La.conj().T + sp.diags(1j * side_scale * rng.standard_normal(n))
    pattern = (abs(M) + abs(H) + abs(S) + abs(Am)).tocsr()
    pattern.sort_indices()

    def on_pattern(G):
        G = G.tocsr()
        out = np.zeros(pattern.nnz, dtype=np.complex128)
        for i in range(n):
            s0, s1 = pattern.indptr[i], pattern.indptr[i + 1]
            cols_i = pattern.indices[s0:s1]
            g0, g1 = G.indptr[i], G.indptr[i + 1]
            out[s0:s1][np.searchsorted(cols_i, G.indices[g0:g1])] = G.data[g0:g1]
        return out

    return pattern, [M.tocsr(), H.tocsr(), S.tocsr(), Am.tocsr()], [on_pattern(G) for G in (M, H, S, Am)]


def _dense_newton_solve(mats, b):
    """x with M x + Mh x + Ms conj(x) + Ma x = b through the equivalent real 2n x 2n system"""
    M, H, S, Am = [g.toarray() for g in mats]
    Lx = M + H + Am
    n = len(b)
    K = np.block([[Lx.real + S.real, -Lx.imag + S.imag], [Lx.imag + S.imag, Lx.real - S.real]])
    xy = np.linalg.solve(K, np.concatenate([b.real, b.imag]))
    return xy[:n] + 1j * xy[n:]


@pytest.mark.parametrize("n,wide", [(1, None), (77, None), (1000, None), (2500, 1200)])
@pytest.mark.parametrize("precond", [X.PRECOND_JACOBI, X.PRECOND_AMG])
def test_products_and_kludge_solve(n, wide, precond):
    pattern, mats, vals = _system(n, seed=n, wide_row=wide)
    P = CProb(pattern)
    P.ok("set_values", P.ptr(vals[0]))
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = np.empty(n, dtype=np.complex128)
    # side matrices do not exist yet: the selector is refused, the solver too
    assert P.call("mult", 1, 0, P.ptr(x), P.ptr(y)) != 0
    info = X.SolveInfo()
    assert P.call("kludge_solve", 1e-10, 0, precond, 0, C.byref(info)) != 0
    P.ok("mult", 0, 0, P.ptr(x), P.ptr(y))
    assert np.abs(y - mats[0] @ x).max() <= 1e-13 * max(1.0, np.abs(y).max())
    P.ok("enable_newton")
    for k in (1, 2, 3):
        P.ok("set_side_values", k, P.ptr(vals[k]))
        back = np.empty_like(vals[k])
        P.ok("get_side_values", k, P.ptr(back))
        assert np.array_equal(back, vals[k])
    for k in range(4):
        for cj in (0, 1):
            P.ok("mult", k, cj, P.ptr(x), P.ptr(y))
            want = (mats[k].conj() if cj else mats[k]) @ x
            assert np.abs(y - want).max() <= 1e-13 * max(1.0, np.abs(want).max()), (k, cj)
    P.ok("mult_newton", P.ptr(x), P.ptr(y))
    want = mats[0] @ x + mats[1] @ x + mats[2] @ x.conj() + mats[3] @ x
    assert np.abs(y - want).max() <= 1e-13 * max(1.0, np.abs(want).max())
    # KludgeSolve: cold, then warm with another right-hand side
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    for warm, rhs in ((0, b), (1, 2.0 * b[::-1].copy())):
        P.ok("set_rhs", P.ptr(rhs))
        P.ok("kludge_solve", 1e-11, warm, precond, 0, C.byref(info))
        V = np.empty(n, dtype=np.complex128)
        P.ok("get_solution", P.ptr(V))
        exact = _dense_newton_solve(mats, rhs)
        err = np.linalg.norm(V - exact) / np.linalg.norm(exact)
        print(n, "precond", precond, "warm", warm, "inner COCG iterations", info.iterations, "relres", info.relres, "err", err)
        assert info.relres < 1e-10 and err <= 1e-8  # KludgeSolve stops at 10 x precision (or after ten outer iterations)
    # zero right-hand side: nothing to do, the cold solution is zero
    zero = np.zeros(n, dtype=np.complex128)
    P.ok("set_rhs", P.ptr(zero))
    P.ok("kludge_solve", 1e-10, 0, precond, 0, C.byref(info))
    P.ok("get_solution", P.ptr(V))
    assert info.iterations == 0 and not V.any()
    # Wipe clears the side matrices as well (cspars.cpp:556-590)
    P.ok("wipe")
    P.ok("get_side_values", 2, P.ptr(back))
    assert not back.any()
    P.close()
