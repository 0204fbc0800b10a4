"""CPU-side checks of the product's host layer (no GPU compute): the C-ABI library loads and exports
every symbol include/xfemm_b200.h declares, .fem/mesh parsing, Cuthill–McKee renumbering, the comb
sort and the sparsity pattern are bit-exact with the reference golden data, and the solver refuses
to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import tempfile

import numpy as np
import pytest

import xfemm_b200 as X
from oracle import fsolver_oracle as O
from oracle.age_matrix import age_matrix
from util import age_cells, load_ref, stage_case, upper_from_full

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "xfemm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(xfemm_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) > 50
    lib = C.CDLL(X.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    # the Python binding covers the same set
    assert declared == set(X.SIGNATURES), sorted(declared ^ set(X.SIGNATURES))


@pytest.mark.parametrize("name", ["tb", "tb30", "temp1"])
def test_load_renumber_pattern_bit_exact(name, tmp_path):
    ref, meta = load_ref(name)
    S = X.FSolver(stage_case(name, tmp_path))
    S.load_mesh(False)
    S.renumber(False)
    c = S.counts()
    assert (c["n_nodes"], c["n_elems"], c["n_pbc"], c["bandwidth"]) == (meta["n_nodes"], meta["n_elems"], meta["n_pbc"], meta["bandwidth"])
    xy, bc = S.nodes()
    assert np.array_equal(xy, ref["nodes"]) and np.array_equal(bc, ref["node_bc"])
    assert np.array_equal(S.elems(), ref["elems"])
    assert np.array_equal(S.pbc(), ref["pbc"])
    # pattern: every entry of the reference's upper-triangular lists is present; the only extras are
    # flagged periodic fill (the reference creates those only where a value is non-zero at run time)
    cells = age_cells(ref["age0_qn"]) if "age0_qn" in ref else None
    rowptr, col, fill = X.pattern_build(c["n_nodes"], S.elems()[:, :4], cells, S.pbc(), with_fill=True)
    n = c["n_nodes"]
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    up = col >= rows
    mine = set(zip(rows[up].tolist(), col[up].tolist()))
    mine_nofill = set(zip(rows[up & (fill == 0)].tolist(), col[up & (fill == 0)].tolist()))
    rrows = np.repeat(np.arange(n), np.diff(ref["step0_rowptr"]))
    refset = set(zip(rrows.tolist(), ref["step0_col"].tolist()))
    assert refset <= mine
    assert mine_nofill <= refset
    if name != "temp1":
        assert mine == refset
    # symmetric structure, sorted rows, diagonal present
    assert all(np.all(np.diff(col[rowptr[i]:rowptr[i + 1]]) > 0) for i in range(0, n, 97))
    import scipy.sparse as sp
    P = sp.csr_matrix((np.ones(len(col)), col, rowptr), shape=(n, n))
    assert (P != P.T).nnz == 0 and P.diagonal().min() == 1


def test_renumbering_primitives_match_oracle_on_random_graphs():
    rng = np.random.default_rng(1)
    for n in (1, 2, 17, 300):
        # random planar-ish graph: ring + chords, plus an isolated component to exercise the restart
        edges = [(i, (i + 1) % n) for i in range(n)] if n > 2 else ([(0, 1)] if n == 2 else [])
        edges += [tuple(rng.integers(0, n, 2)) for _ in range(n // 2)]
        edges = np.array([e for e in edges if e[0] != e[1]], dtype=np.int32).reshape(-1, 2)
        newnum, bw = X.renumber_cuthill(n, edges)
        onew = np.empty(n, dtype=np.int32)
        obw = C.c_int(0)
        O.lib().orc_cuthill(n, len(edges), edges.ctypes.data_as(C.c_void_p), onew.ctypes.data_as(C.c_void_p), C.byref(obw))
        assert np.array_equal(newnum, onew) and bw == obw.value
        assert sorted(newnum.tolist()) == list(range(n))
    for ne in (0, 1, 9, 10, 11, 12, 13, 14, 1000):
        tri = rng.integers(0, 500, (ne, 3)).astype(np.int32)
        order = X.sort_elements(tri) if ne else np.zeros(0, dtype=np.int32)
        oorder = np.empty(ne, dtype=np.int32)
        if ne:
            O.lib().orc_sort_elements(ne, tri.ctypes.data_as(C.c_void_p), oorder.ctypes.data_as(C.c_void_p))
        assert np.array_equal(order, oorder)


def test_age_cell_matrix_reduced_form_matches_55_closed_forms():
    lib = X.lib()
    rng = np.random.default_rng(2)
    for _ in range(50):
        ci, co, K = rng.random(), rng.random(), rng.uniform(0.1, 8.0)
        mine = X.age_cell_matrix(ci, co, K)
        ref = age_matrix(ci, co, K)
        assert np.abs(mine - ref).max() <= 1e-13 * max(1.0, np.abs(ref).max())


def test_bad_inputs_fail_loudly(tmp_path):
    with pytest.raises(X.XfemmError):
        X.FSolver(str(tmp_path / "does_not_exist"))
    S = X.FSolver(stage_case("tb", tmp_path))
    os.remove(str(tmp_path / "tb.ele"))
    with pytest.raises(X.XfemmError):
        S.load_mesh(False)
    assert X.fsolver_run(str(tmp_path / "nope")) == 1   # main.cpp: exit code 1 = problem loading .fem
    with pytest.raises(X.XfemmError):
        X.pattern_build(3, np.array([[0, 1, 7]], dtype=np.int32))


@pytest.mark.skipif(X.device_count() > 0, reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback(tmp_path):
    """Without a CUDA device the hot path must refuse to run rather than compute on the CPU."""
    with pytest.raises(X.XfemmError):
        X.cuda_init(0)
    S = X.FSolver(stage_case("tb", tmp_path))
    S.load_mesh(False)
    S.renumber(False)
    with pytest.raises(X.XfemmError):
        S.solve()
    rowptr = np.array([0, 1, 2], dtype=np.int32)
    col = np.array([0, 1], dtype=np.int32)
    with pytest.raises(X.XfemmError):
        X.LinProb(rowptr, col)
    assert X.fsolver_run(stage_case("tb30", tmp_path)) == 2   # main.cpp: exit code 2 = solve failed


def test_edge_cases_of_the_host_primitives():
    """Empty and degenerate inputs of the renumbering / pattern entry points."""
    # a single node without edges: numbered 0, bandwidth 1
    newnum, bw = X.renumber_cuthill(1, np.zeros((0, 2), dtype=np.int32))
    assert list(newnum) == [0] and bw == 1
    # two disconnected components (the restart branch of cuthill.cpp:269-301): a permutation results
    newnum, _ = X.renumber_cuthill(5, np.array([[0, 1], [3, 4]], dtype=np.int32))
    assert sorted(newnum) == [0, 1, 2, 3, 4]
    # no elements: diagonal-only pattern
    rp, col = X.pattern_build(4, np.zeros((0, 3), dtype=np.int32))
    assert list(rp) == [0, 1, 2, 3, 4] and list(col) == [0, 1, 2, 3]
    # one element, a periodic pair between a mesh node and an isolated node: the neighbours k of node 0 get
    # entries (k,3) and (3,k); the pair itself stays uncoupled (spars.cpp:421-474 skips k == i, k == j)
    rp, col, fill = X.pattern_build(4, np.array([[0, 1, 2]], dtype=np.int32), pbc=np.array([[0, 3]], dtype=np.int32), with_fill=True)
    rows = [set(col[rp[i]:rp[i + 1]]) for i in range(4)]
    assert rows[3] == {1, 2, 3} and rows[1] == {0, 1, 2, 3} and rows[2] == {0, 1, 2, 3} and rows[0] == {0, 1, 2}
    assert fill.sum() > 0 and len(fill) == len(col)
    # element sort of nothing / of one element
    assert len(X.sort_elements(np.zeros((0, 3), dtype=np.int32))) == 0
    assert list(X.sort_elements(np.array([[2, 1, 0]], dtype=np.int32))) == [0]
    # out-of-range indices are refused
    with pytest.raises(X.XfemmError):
        X.renumber_cuthill(3, np.array([[0, 5]], dtype=np.int32))
    with pytest.raises(X.XfemmError):
        X.pattern_build(3, np.array([[0, 1, 2]], dtype=np.int32), pbc=np.array([[0, 9]], dtype=np.int32))


@pytest.mark.parametrize("threads", ["1", "4"])
def test_layout_and_multigrid_setup_invariants(threads, tmp_path):
    """SELL layout with wide rows (plain, Z-curve, coloured), aggregation and Galerkin maps: tests/layout_check.cpp
    verifies on the CPU that every matrix entry is stored exactly once, the permutations are consistent and
    the coarse patterns / slot lists equal a brute-force P^T A P."""
    import subprocess
    exe = str(tmp_path / "layout_check")
    libdir = os.path.join(ROOT, "xfemm_b200")
    r = subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(ROOT, "tests", "layout_check.cpp"), "-o", exe, "-L" + libdir,
                        "-lxfemm_b200", "-Wl,-rpath," + libdir, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for n in ("60", "33"):
        r = subprocess.run([exe, n], capture_output=True, text=True, env=dict(os.environ, XFEMM_HOST_THREADS=threads))
        assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_magdir_expression_evaluator():
    """MagDirFctn strings (the reference runs them through Lua 4.0 in radians mode, static2d.cpp:513-583,
    liblua/lmathlib.cpp:24): operator precedence incl. right-associative ^ above unary minus, variables, functions;
    anything outside the expression sublanguage is rejected with a message, never mis-evaluated."""
    import ctypes, math
    L = X.lib()

    def ev(expr, x, y):
        out = ctypes.c_double(0)
        rc = L.xfemm_magdir_eval(expr.encode(), x, y, ctypes.byref(out))
        return rc, out.value

    rng = np.random.default_rng(5)
    cases = [
        ("theta", lambda x, y, th, R: th),
        ("theta + 20*cos(3*theta*PI/180) - R^2/4", lambda x, y, th, R: th + 20 * math.cos(3 * th * math.pi / 180) - R ** 2 / 4),
        ("-(x+1)^2*10 + deg(atan2(y+0.5, x+1.5)) + max(1, 2*y, abs(x))",
         lambda x, y, th, R: -(x + 1) ** 2 * 10 + math.degrees(math.atan2(y + 0.5, x + 1.5)) + max(1, 2 * y, abs(x))),
        ("-2^2", lambda *a: -4.0), ("2^3^2", lambda *a: 512.0), ("2^-1", lambda *a: 0.5), ("1-2-3", lambda *a: -4.0),
        ("8/4/2", lambda *a: 1.0), ("r*z - x*y", lambda *a: 0.0), ("mod(theta+360, 90) + floor(R) + ceil(.5) + 1e1",
                                                                 lambda x, y, th, R: math.fmod(th + 360, 90) + math.floor(R) + 1 + 10),
        ("sqrt(x*x+y*y) - R", lambda *a: 0.0), ("min(3, 2, 5) + exp(0) + log(1) + log10(100) + rad(180)/PI + re(2) + im(2)",
                                                lambda *a: 2 + 1 + 0 + 2 + 1 + 2 + 0),
    ]
    for expr, f in cases:
        for _ in range(5):
            x, y = rng.uniform(-3, 3, 2)
            th, R = math.degrees(math.atan2(y, x)), math.hypot(x, y)
            rc, v = ev(expr, x, y)
            assert rc == 0, (expr, L.xfemm_last_error())
            assert abs(v - f(x, y, th, R)) <= 1e-12 * max(1.0, abs(v)), (expr, v, f(x, y, th, R))
    for bad in ["", "theta +", "foo", "sin()", "atan2(1)", "x = 3", "theta; 2", "(1", "1 2", "cos(1", "a.b", "'s'", "sqrt(-1)", "1/0"]:
        rc, _ = ev(bad, 1.0, 2.0)
        assert rc != 0, bad
        assert L.xfemm_last_error()


@pytest.mark.parametrize("newton", [0, 1])
def test_complex_facade_host_logic_matches_the_reference_class(newton, tmp_path):
    """Host side of CBigComplexLinProb (no GPU involved): Put/Get with the selector k and its conjugation rules, SetValue,
    Periodicity / AntiPeriodicity on all four matrices.  tests/cfacade_driver.cpp built against this repo's facade and,
    by oracle/Makefile, against the unmodified reference class; every entry of M, Mh, Ms, Ma and the right-hand side
    after the boundary conditions must agree."""
    import subprocess
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "cfacade_ref")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/cfacade_ref not built (run make -C oracle where /root/reference is available)")
    exe = str(tmp_path / "cfacade_driver")
    host = os.path.join(ROOT, "xfemm_b200", "host")
    libdir = os.path.join(ROOT, "xfemm_b200")
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-I" + host, os.path.join(ROOT, "tests", "cfacade_driver.cpp"), "-o", exe,
                        "-L" + libdir, "-lxfemm_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outs = []
    for e in (exe, ref_exe):
        r = subprocess.run([e, "7", str(newton), "dump"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        outs.append(np.array([[float(x) for x in ln.split()] for ln in r.stdout.splitlines() if not ln.startswith("#")]))
    mine, theirs = outs
    n = 7 * 7 + 2
    assert mine.shape == theirs.shape == (4 * n * n + n, 2)
    scale = np.abs(theirs).max()
    assert np.abs(mine - theirs).max() <= 1e-15 * scale
    if newton:
        assert np.abs(theirs[n * n:4 * n * n]).max() > 0  # the side matrices are populated


def test_real_facade_host_logic_matches_the_reference_class(tmp_path):
    """Host side of CBigLinProb (no GPU involved): AddTo/Put/Get, SetValue (first/last value semantics) and
    Periodicity / AntiPeriodicity as hsolver/esolver use them.  tests/facade_driver.cpp built against this repo's
    facade and against the unmodified reference class: every entry and the right-hand side agree."""
    import subprocess
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "facade_ref")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/facade_ref not built (run make -C oracle where /root/reference is available)")
    exe = str(tmp_path / "facade_driver")
    host = os.path.join(ROOT, "xfemm_b200", "host")
    libdir = os.path.join(ROOT, "xfemm_b200")
    r = subprocess.run(["g++", "-std=c++14", "-O1", "-I" + host, os.path.join(ROOT, "tests", "facade_driver.cpp"), "-o", exe,
                        "-L" + libdir, "-lxfemm_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outs = []
    for e in (exe, ref_exe):
        r = subprocess.run([e, "9", "dump"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        outs.append(np.array([float(ln) for ln in r.stdout.splitlines() if not ln.startswith("#")]))
    mine, theirs = outs
    assert mine.shape == theirs.shape == (81 * 81 + 81,)
    assert np.abs(mine - theirs).max() <= 1e-15 * np.abs(theirs).max()


def test_magdir_evaluator_against_the_reference_lua_interpreter(tmp_path):
    """The same expressions evaluated by the reference's own Lua 4.0 (oracle/_ref/femmcli) with the variable block
    Static2D prepends (static2d.cpp:530-531): precedence of ^ and unary minus, radians-mode functions, PI."""
    import ctypes, math, subprocess
    femmcli = os.path.join(ROOT, "oracle", "_ref", "femmcli")
    if not os.path.exists(femmcli):
        pytest.skip("oracle/_ref/femmcli not built")
    exprs = ["theta", "theta + 20*cos(3*theta*PI/180) - R^2/4", "-(x+1)^2*10 + deg(atan2(y+0.5, x+1.5)) + max(1, 2*y, abs(x))",
             "-2^2", "2^3^2", "2^-1", "-x^2", "1-2-3", "8/4/2", "r*z - x*y", "mod(theta+360, 90) + floor(R) + ceil(.5) + 1e1",
             "min(3, 2, 5) + exp(0) + log(1) + log10(100) + rad(180)/PI", "sin(x)*cos(y) + tan(x/4) + asin(x/4) + acos(y/4) + atan(x)",
             "sqrt(R) + tanh(x) + cosh(y/3) + sinh(x/3)", "2*-3", "-x*-y", "(1+2)*3^2/4 - -1"]
    rng = np.random.default_rng(8)
    pts = rng.uniform(-3, 3, size=(3, 2))
    lines = []
    pts = [(float(x), float(y)) for x, y in pts]
    for x, y in pts:
        th, R = math.degrees(math.atan2(y, x)), math.hypot(x, y)
        for e in exprs:
            lines.append(f"x={x!r}\ny={y!r}\nr=x\nz=y\ntheta={th!r}\nR={R!r}\nwrite(format(\"VAL %.17g\\n\", {e}))")
    script = tmp_path / "expr.lua"
    script.write_text("\n".join(lines) + "\nwrite(\"DONE\\n\")\n")
    r = subprocess.run([femmcli, "--lua-script=expr.lua"], cwd=tmp_path, capture_output=True, text=True)
    assert "DONE" in r.stdout, r.stdout[-500:] + r.stderr[-500:]
    ref = [float(ln.split()[1]) for ln in r.stdout.splitlines() if ln.startswith("VAL")]
    assert len(ref) == len(pts) * len(exprs)
    L = X.lib()
    k = 0
    for x, y in pts:
        for e in exprs:
            out = ctypes.c_double(0)
            assert L.xfemm_magdir_eval(e.encode(), float(x), float(y), ctypes.byref(out)) == 0, (e, L.xfemm_last_error())
            assert abs(out.value - ref[k]) <= 1e-12 * max(1.0, abs(ref[k])), (e, out.value, ref[k])
            k += 1


def test_bh_curve_tables_against_the_reference_material_class(tmp_path):
    """host/bh_curve.cpp (row a20: GetSlopes, LaminatedBH, GetH, GetdHdB, Get_v, GetBHProps, IncrementalPermeability)
    against the UNMODIFIED reference class: tests/bh_driver.cpp is compiled against this repo's host mirror here and
    against cfemm/libfemm/CMaterialProp.cpp by oracle/Makefile (oracle/_ref/bh_ref).  Eight probe curves — static, stacking
    factor, one that needs smoothing passes, AC with hysteresis lag, AC with eddy currents in the laminations — every table
    entry and evaluations over 0..3 T, printed as bit patterns.  Bar: 1e-10 of the largest value of the same quantity in
    the same case (the two solve the spline system with different but equivalent eliminations: last-bit differences in
    the slopes, which the cancellation in d(nu)/dB^2 at high flux densities amplifies relative to that tiny value, and
    which the lamination model — a fixed-point iteration stopped at a residual of 1e-8 — carries to 2e-12)."""
    import struct
    import subprocess
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "bh_ref")
    if not os.path.exists(ref_exe):
        pytest.skip("oracle/_ref/bh_ref not built (run make -C oracle where /root/reference is available)")
    exe = str(tmp_path / "bh_ours")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "xfemm_b200", "host"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "bh_driver.cpp"), os.path.join(ROOT, "xfemm_b200", "host", "bh_curve.cpp"), "-o", exe], check=True)
    ours = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines()
    ref = subprocess.run([ref_exe], check=True, capture_output=True, text=True).stdout.splitlines()
    assert len(ours) == len(ref) and len(ref) > 250

    def decode(line):
        return [struct.unpack(">d", bytes.fromhex(t))[0] for t in line.split()]

    def is_data(line):
        t = line.split()
        return bool(t) and all(len(x) == 16 for x in t)

    # split into cases (a name line starts each); inside a case group the lines by their number of columns
    cases, cur = [], None
    for a, b in zip(ref, ours):
        if not is_data(a):
            assert a == b
            cur = {"name": a, "rows": []}
            cases.append(cur)
        else:
            cur["rows"].append((decode(a), decode(b)))
    assert len(cases) == 8
    worst = 0.0
    identical = total = 0
    for c in cases:
        by_width = {}
        for ra, rb in c["rows"]:
            assert len(ra) == len(rb), c["name"]
            by_width.setdefault(len(ra), []).append((ra, rb))
        for width, rows in by_width.items():
            A = np.array([r[0] for r in rows])
            B = np.array([r[1] for r in rows])
            scale = np.maximum(np.abs(A).max(axis=0), 1e-300)
            err = (np.abs(A - B) / scale).max()
            worst = max(worst, err)
            identical += int((A == B).sum())
            total += A.size
            c["worst"] = max(c.get("worst", 0.0), float(err))
            assert err <= 1e-10, (c["name"], width, err)
    print({c["name"]: c["worst"] for c in cases})
    print(f"bh tables: {identical} of {total} values bit-identical to the reference, worst deviation {worst:.2e} of the column scale")


def test_periodic_unit_levels_reproduce_list_order_semantics():
    """Row a8.  The device applies periodic pairs unit by unit (one unit per coupled row k, one for the two diagonals, one
    for the two right-hand-side entries), units grouped into dependency levels: level = 1 + the highest level of an
    earlier unit that shares an entry (csrc/linprob.cu, xfemm_linprob_periodic / k_periodic_units).  This is the same
    rule in numpy, run on random banded systems with chains of neighbouring pairs, the units of a level applied in a
    SHUFFLED order, against the pinned CPU restatement of the reference's sequential loop (oracle orc_periodicity =
    spars.cpp:366-474): the results have to be bit-identical, whatever the order inside a level."""
    import ctypes as C
    from oracle import fsolver_oracle as O
    L = O.lib()
    rng = np.random.default_rng(7)
    for trial in range(6):
        n, bw = 90, 4
        half = n // 2
        A = np.zeros((n, n))
        for i in range(n):
            for j in range(i, min(n, i + bw + 1)):
                if i == j or rng.random() < 0.8:
                    A[i, j] = A[j, i] = rng.normal() + (6.0 if i == j else 0.0)
        b = rng.normal(size=n)
        # a chain of neighbouring pairs (i, i + half): consecutive pairs share entries; a few anti-periodic, one node in two pairs
        pairs = [(i, i + half, int(rng.random() < 0.3)) for i in range(3, 3 + 18)]
        pairs.append((5, 80, 0))
        # ---- reference semantics: the oracle's sequential loop on upper-triangular storage
        m = L.orc_create(n, 0)
        for i in range(n):
            for j in range(i, n):
                if A[i, j] != 0.0:
                    L.orc_put(m, float(A[i, j]), i, j)
        bref = np.ctypeslib.as_array(L.orc_b(m), shape=(n,))
        bref[:] = b
        for i, j, t in pairs:
            L.orc_periodicity(m, i, j, t)
        Aref = np.array([[L.orc_get(m, min(i, j), max(i, j)) for j in range(n)] for i in range(n)])
        bref = bref.copy()
        L.orc_destroy(m)
        # ---- structural pattern incl. the fill a symbolic replay creates (what xfemm_pattern_build does)
        P = A != 0.0
        for i, j, _ in pairs:
            ks = [k for k in range(n) if (P[k, i] or P[k, j]) and k not in (i, j)]
            for k in ks:
                P[k, i] = P[i, k] = P[k, j] = P[j, k] = True
        # ---- units and their levels, pairs in list order
        units, last = [], {}

        def add(kind, ids, payload):
            lv = max((last[e] + 1 for e in ids if e in last), default=0)
            for e in ids:
                last[e] = lv
            units.append((lv, kind, payload))

        for i, j, t in pairs:
            if j < i:
                i, j = j, i
            s = -1.0 if t else 1.0
            for k in range(n):
                if k in (i, j) or not (P[k, i] or P[k, j]):
                    continue
                add("item", [("a", k, i), ("a", i, k), ("a", k, j), ("a", j, k)], (k, i, j, s))
            add("diag", [("a", i, i), ("a", j, j)], (i, j, s))
            add("rhs", [("b", i), ("b", j)], (i, j, s))
        nlev = max(u[0] for u in units) + 1
        assert nlev < len(pairs)  # far fewer levels than a pair-by-pair chain would need
        Aw, bw_ = A.copy(), b.copy()
        for lv in range(nlev):
            todo = [u for u in units if u[0] == lv]
            rng.shuffle(todo)
            for _, kind, p in todo:
                if kind == "item":
                    k, i, j, s = p
                    v1, v2 = Aw[k, i], Aw[k, j]
                    if v1 != 0.0 or v2 != 0.0:
                        c = (v1 + s * v2) / 2.0
                        Aw[k, i] = Aw[i, k] = c
                        Aw[k, j] = Aw[j, k] = s * c
                elif kind == "diag":
                    i, j, s = p
                    c = 0.5 * (Aw[i, i] + Aw[j, j])
                    Aw[i, i] = Aw[j, j] = c
                else:
                    i, j, s = p
                    c = 0.5 * (bw_[i] + s * bw_[j])
                    bw_[i], bw_[j] = c, s * c
        assert np.array_equal(Aw, Aref), trial
        assert np.array_equal(bw_, bref), trial
