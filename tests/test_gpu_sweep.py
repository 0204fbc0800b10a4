"""Rotor-angle sweeps on one mesh (BASELINE config 5a) and derived quantities through the
reference's own post-processor.

* re-targeting the air-gap element to a new rotor angle without re-meshing reproduces what
  fmesher + fsolver give for that angle (golden case tb30);
* with the angle-independent pattern the whole sweep runs on the resident device problem;
* the gap torque computed by the REFERENCE fpproc (oracle/_ref/femmcli, mo_gapintegral) on OUR .ans
  matches the reference's torque within 1e-4 relative and the analytic T = sin(theta) within the
  tolerance of the reference's own test (cfemm/femmcli/test/femmcli_TorqueBenchmark.lua:36-37).
"""
import math
import os
import shutil
import subprocess

import numpy as np
import pytest

import xfemm_b200 as X
from util import load_ref, rel_l2, stage_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FEMMCLI = os.path.join(ROOT, "oracle", "_ref", "femmcli")
REF_FSOLVER = os.path.join(ROOT, "oracle", "_ref", "fsolver")


def test_retarget_matches_remeshed_reference(tmp_path):
    ref30, _ = load_ref("tb30")
    S = X.FSolver(stage_case("tb", tmp_path))
    S.load_mesh(False)
    S.renumber(False)
    S.set_rotor_angle(30.0)
    S.solve()
    V, A = S.solution()
    assert rel_l2(A, ref30["A"]) <= 1e-6


def test_sweep_on_resident_problem(tmp_path):
    ref0, _ = load_ref("tb")
    ref30, _ = load_ref("tb30")
    S = X.FSolver(stage_case("tb", tmp_path))
    S.set_option("age_union_pattern", 1)
    S.load_mesh(False)
    S.renumber(False)
    S.solve()
    assert rel_l2(S.solution()[1], ref0["A"]) <= 1e-6
    for angle, ref in ((30.0, ref30), (0.0, ref0), (30.0, ref30)):
        S.set_rotor_angle(angle)
        S.resolve()
        S_A = S.solution()[1]
        assert rel_l2(S_A, ref["A"]) <= 1e-6, angle


def _gap_torque(workdir, base):
    lua = os.path.join(workdir, "tq.lua")
    with open(lua, "w") as f:
        f.write(f'open("{base}.fem")\nmi_loadsolution()\nprint("TORQUE " .. mo_gapintegral("AGE", 0))\n'
                'mo_groupselectblock()\nprint("ENERGY " .. mo_blockintegral(2))\n')
    r = subprocess.run([FEMMCLI, f"--lua-script={lua}"], cwd=workdir, capture_output=True, text=True)
    out = {}
    for ln in r.stdout.splitlines():
        if ln.startswith("TORQUE ") or ln.startswith("ENERGY "):
            k, v = ln.split()
            out[k] = float(v)
    assert "TORQUE" in out, r.stdout + r.stderr
    return out


@pytest.mark.skipif(not (os.path.exists(FEMMCLI) and os.path.exists(REF_FSOLVER)), reason="oracle/_ref not built")
def test_gap_torque_via_reference_postprocessor(tmp_path):
    tq_tolerance = 0.000042  # femmcli_TorqueBenchmark.lua:36
    for angle in (0.0, 30.0, 90.0):
        ours = tmp_path / f"ours{int(angle)}"
        ours.mkdir()
        base = stage_case("tb", str(ours))
        S = X.FSolver(base)
        S.load_mesh(False)
        S.renumber(False)
        S.set_rotor_angle(angle)
        S.solve()
        S.write()
        mine = _gap_torque(str(ours), "tb")
        assert abs(mine["TORQUE"] - math.sin(math.radians(angle))) <= tq_tolerance
        # the same integral evaluated on the device from the resident solution (no .ans, no nodal D2H)
        dev = S.gap_torque(0)
        assert abs(dev - mine["TORQUE"]) <= 1e-9 * max(abs(mine["TORQUE"]), 1e-3), (dev, mine["TORQUE"])
        if angle in (0.0, 30.0):
            # the reference solver on the mesh fmesher produced for this angle
            name = "tb" if angle == 0.0 else "tb30"
            theirs = tmp_path / f"ref{int(angle)}"
            theirs.mkdir()
            stage_case(name, str(theirs))
            subprocess.run([REF_FSOLVER, name], cwd=str(theirs), check=True, stdout=subprocess.DEVNULL)
            refq = _gap_torque(str(theirs), name)
            print("angle", angle, "torque ours", mine["TORQUE"], "reference", refq["TORQUE"], "energy", mine.get("ENERGY"), refq.get("ENERGY"))
            assert abs(mine["TORQUE"] - refq["TORQUE"]) <= 1e-4 * max(abs(refq["TORQUE"]), 1e-2)
            if "ENERGY" in refq and refq["ENERGY"] != 0:
                assert abs(mine["ENERGY"] - refq["ENERGY"]) <= 1e-4 * abs(refq["ENERGY"])


def test_torque_only_sweep_leaves_nodal_values_on_device(tmp_path):
    """A sweep that reads only the torque: no nodal values are copied back between positions."""
    S = X.FSolver(stage_case("tb", tmp_path))
    S.set_option("age_union_pattern", 1)
    S.load_mesh(False)
    S.renumber(False)
    S.solve()
    S.set_option("fetch_solution", 0)
    import ctypes as C
    h2d, d2h = C.c_longlong(0), C.c_longlong(0)
    X.lib().xfemm_transfer_bytes(C.byref(h2d), C.byref(d2h), 1)
    tq = []
    for angle in (0.0, 10.0, 30.0, 60.0, 90.0):
        S.set_rotor_angle(angle)
        S.resolve()
        tq.append(S.gap_torque(0))
    X.lib().xfemm_transfer_bytes(C.byref(h2d), C.byref(d2h), 0)
    assert d2h.value < 5 * 64 * 1024  # scalars, flags and the 2x96 flux densities only
    for angle, t in zip((0.0, 10.0, 30.0, 60.0, 90.0), tq):
        assert abs(t - math.sin(math.radians(angle))) <= 0.000042 + 2e-4 * abs(math.sin(math.radians(angle))), (angle, t)


def test_cached_static_part_gives_the_same_systems_and_answers(tmp_path):
    """Sweep mode (default with age_union_pattern on a linear problem): the boundary-conditioned, position-independent
    part of the system is restored from a device copy and only the air-gap blocks are assembled per position.  Same
    nodal solution and torque as re-assembling and re-conditioning everything (option cache_static_part = 0)."""
    out = {}
    for cache in (1, 0):
        d = tmp_path / str(cache)
        d.mkdir()
        S = X.FSolver(stage_case("tb", str(d)))
        S.set_option("age_union_pattern", 1)
        S.set_option("cache_static_part", cache)
        S.load_mesh(False)
        S.renumber(False)
        S.solve()
        res = []
        for deg in (7.0, 30.0, 123.0):
            S.set_rotor_angle(deg)
            S.resolve()
            res.append((S.solution()[1].copy(), S.gap_torque(0), sum(S.stats()["assemble_ms"])))
        out[cache] = res
        S.close()
    for (a1, t1, ms1), (a0, t0, ms0) in zip(out[1], out[0]):
        assert rel_l2(a1, a0) <= 1e-9
        assert abs(t1 - t0) <= 1e-9 * max(1.0, abs(t0))
    ref30, _ = load_ref("tb30")
    assert rel_l2(out[1][1][0], ref30["A"]) <= 1e-6
