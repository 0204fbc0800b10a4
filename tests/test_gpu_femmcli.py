"""The reference's Lua console with `mi_analyze` on the GPU (oracle/_ref/femmcli_b200: femmcli with the one source that
touches FSolver recompiled against oracle/adapters/fsolver_femmcli.h, everything else the reference's objects).

The sweep is the reference's only tight known-answer test (cfemm/femmcli/test/femmcli_TorqueBenchmark.lua:39-70:
torque of the benchmark machine = sin(angle) N*m at 0..90 degrees, absolute margin 4.2e-5, relative margin 0.006 %),
driven by our own script tests/golden/femmcli/torque_sweep.lua: mesher, Lua and post-processor are the reference's,
the solve in between is this repo's."""
import math
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _sweep(exe, d):
    shutil.copy(os.path.join(ROOT, "tests", "golden", "tb", "tb.fem"), d)
    shutil.copy(os.path.join(ROOT, "tests", "golden", "femmcli", "torque_sweep.lua"), d)
    # femmcli may crash in its teardown after the script has finished: judge the output, not the exit code
    r = subprocess.run([os.path.join(REF, exe), "--lua-script=torque_sweep.lua"], cwd=d, capture_output=True, text=True)
    out = r.stdout
    assert "DONE" in out, f"exit code {r.returncode}\n" + out[-1500:] + r.stderr[-1500:]
    return {int(ln.split()[1]): float(ln.split()[2]) for ln in out.splitlines() if ln.startswith("TORQUE")}


def test_reference_console_analyses_on_the_gpu(tmp_path):
    for exe in ("femmcli", "femmcli_b200"):
        if not os.path.exists(os.path.join(REF, exe)):
            pytest.skip(f"oracle/_ref/{exe} not built (run make -C oracle where /root/reference is available)")
    tq = {}
    for exe in ("femmcli", "femmcli_b200"):
        d = tmp_path / exe
        d.mkdir()
        tq[exe] = _sweep(exe, str(d))
    assert sorted(tq["femmcli_b200"]) == list(range(0, 91, 10))
    for deg in range(0, 91, 10):
        ours, theirs, exact = tq["femmcli_b200"][deg], tq["femmcli"][deg], math.sin(math.radians(deg))
        print(f"{deg:3d} deg  torque on the GPU {ours:.9f}  reference {theirs:.9f}  analytic {exact:.6f}")
        # margins of femmcli_TorqueBenchmark.lua:39-40
        assert abs(ours - exact) <= 4.2e-5
        if deg:
            assert abs(ours - exact) / exact * 100 <= 0.006
        assert abs(ours - theirs) <= 1e-8
