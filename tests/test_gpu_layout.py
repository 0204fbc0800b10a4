"""The solver structure built on the device against the host builders that remain as its reference:

  XFEMM_LAYOUT_CHECK=1     linprob_build_layout builds the SELL layout both ways (csrc/layout_device.cu, csrc/layout.cpp)
                           and compares every table, slot for slot — on every level of the multigrid hierarchy: Z-curve
                           order and wide rows on the finest level, the 224-column slices of the small levels
  XFEMM_HIERARCHY_CHECK=1  amg_build does the same for the coarse patterns, the gather lists of the Galerkin products,
                           the centroids and the transfer maps (csrc/amg_setup.cu against the loops in csrc/amg.cu)

A mismatch fails the solve with an error that names the table.  The results of a device-built and a host-built
problem then agree to solver precision (same structure; the Krylov loop itself is not bit-reproducible)."""
import os

import numpy as np
import pytest

import xfemm_b200 as X
from util import stage_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _solve(base, options, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        S = X.FSolver(base)
        for k, v in options.items():
            S.set_option(k, v)
        S.load_mesh(False)
        S.renumber(False)
        S.solve()
        V, A = S.solution()
        st = S.stats()
        S.close()
        return A, st
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("options", [{"precond": 2}, {"precond": 2, "locality_order": 1, "age_union_pattern": 1}, {"precond": 0}])
def test_device_layout_equals_host_layout_small(tmp_path, options):
    options = dict(options, deterministic_assembly=1)  # same matrix bits in both runs
    base = stage_case("tb", str(tmp_path))
    A, st = _solve(base, options, {"XFEMM_LAYOUT_CHECK": "1", "XFEMM_HIERARCHY_CHECK": "1"})
    os.makedirs(tmp_path / "h")
    B, _ = _solve(stage_case("tb", str(tmp_path / "h")), options, {"XFEMM_HOST_LAYOUT": "1", "XFEMM_HOST_HIERARCHY": "1"})
    assert np.linalg.norm(A - B) <= 1e-7 * np.linalg.norm(B)


def test_device_layout_equals_host_layout_231k():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    work = os.environ.get("XFEMM_BENCH_DIR", "/tmp/xfemm_bench")
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fmesher")):
        pytest.skip("oracle/_ref/fmesher not built")
    base = bench.prepare_mesh(os.path.join(work, "ms0.005"), 0.005)
    A, st = _solve(base, {"precond": 2, "locality_order": 1, "age_union_pattern": 1}, {"XFEMM_LAYOUT_CHECK": "1", "XFEMM_HIERARCHY_CHECK": "1"})
    assert st["iters"][0] < 200


@pytest.mark.parametrize("case", ["tbh", "harm2d"])
def test_device_layout_equals_host_layout_complex(tmp_path, case):
    """The complex problem (and its real companion carrying the multigrid hierarchy) through the same builders."""
    base = stage_case(case, str(tmp_path))
    env = {"XFEMM_LAYOUT_CHECK": "1", "XFEMM_HIERARCHY_CHECK": "1"}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        S = X.FSolver(base)
        S.load_mesh(False)
        S.renumber(False)
        S.solve()  # a mismatch fails the solve with an error that names the table
        S.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
