"""Block integrals from the resident solution: stored energy and circuit flux linkage reduced on the
device, against the REFERENCE post-processor (oracle/_ref/femmcli: mo_blockintegral(2), mo_getcircuitproperties)
reading the .ans of the same solve, and against the reference solver + post-processor end to end (bar 1e-4).

  tb     linear, two permanent magnets (second-quadrant energy, fpproc.cpp:3786-3800)
  sat2d  saturating 1117 steel (energy along the B-H curve, CMaterialProp.cpp:536-585), a 200-turn series coil
         with a bulk conductor (flux linkage = integral of A.J over the coil / I, fpproc.cpp:5166-5196)
"""
import os
import subprocess

import pytest

import xfemm_b200 as X
from util import stage_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FEMMCLI = os.path.join(ROOT, "oracle", "_ref", "femmcli")
REF_FSOLVER = os.path.join(ROOT, "oracle", "_ref", "fsolver")


def _postprocess(workdir, name, circuit=None):
    lua = os.path.join(workdir, "q.lua")
    with open(lua, "w") as f:
        f.write(f'open("{name}.fem")\nmi_loadsolution()\nmo_groupselectblock()\nprint("ENERGY " .. mo_blockintegral(2))\n')
        if circuit:
            f.write(f'i, v, fl = mo_getcircuitproperties("{circuit}")\nprint("CURRENT " .. i)\nprint("LINKAGE " .. fl)\n')
    r = subprocess.run([FEMMCLI, f"--lua-script={lua}"], cwd=workdir, capture_output=True, text=True)
    out = {}
    for ln in r.stdout.splitlines():
        p = ln.split()
        if len(p) == 2 and p[0] in ("ENERGY", "CURRENT", "LINKAGE"):
            out[p[0]] = float(p[1])
    assert "ENERGY" in out, r.stdout + r.stderr
    return out


@pytest.mark.skipif(not (os.path.exists(FEMMCLI) and os.path.exists(REF_FSOLVER)), reason="oracle/_ref not built")
@pytest.mark.parametrize("case,circuit", [("tb", None), ("sat2d", "coil")])
def test_energy_and_linkage_against_the_reference_postprocessor(case, circuit, tmp_path):
    ours = tmp_path / "ours"
    ours.mkdir()
    S = X.FSolver(stage_case(case, str(ours)))
    S.load_mesh(False)
    S.renumber(False)
    S.solve()
    S.write()
    mine = _postprocess(str(ours), case, circuit)
    dev_e = S.stored_energy()
    print(case, "stored energy: device", dev_e, "reference fpproc on our .ans", mine["ENERGY"])
    assert abs(dev_e - mine["ENERGY"]) <= 1e-9 * abs(mine["ENERGY"])
    if circuit:
        dev_l = S.flux_linkage(0)
        print(case, "flux linkage: device", dev_l, "reference fpproc on our .ans", mine["LINKAGE"], "current", mine["CURRENT"])
        assert abs(dev_l - mine["LINKAGE"]) <= 1e-9 * abs(mine["LINKAGE"])
    # the reference solver + post-processor end to end
    theirs = tmp_path / "ref"
    theirs.mkdir()
    stage_case(case, str(theirs))
    subprocess.run([REF_FSOLVER, case], cwd=str(theirs), check=True, stdout=subprocess.DEVNULL)
    ref = _postprocess(str(theirs), case, circuit)
    assert abs(dev_e - ref["ENERGY"]) <= 1e-4 * abs(ref["ENERGY"])
    if circuit:
        assert abs(S.flux_linkage(0) - ref["LINKAGE"]) <= 1e-4 * abs(ref["LINKAGE"])
    S.close()
