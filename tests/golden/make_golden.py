"""Regenerates the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref, built by `make -C oracle` from /root/reference) in this container:

  fmesher  → <case>.node/.ele/.edge/.pbc      (input fixtures: what fsolver reads)
  ref_harness dump → renumbered mesh, the assembled + boundary-conditioned system handed to
                     PCGSolve at every Newton step, the solution, iteration counts

Only small cases are committed.  Usage: python tests/golden/make_golden.py [case ...]
"""
import gzip
import json
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
SRC = "/root/reference"


def _write_fem(path, text):
    with open(path, "w", newline="") as f:
        f.write(text.replace("\r\n", "\n"))


def set_mesh_size(text, size):
    """Field 4 of every [NumBlockLabels] row is the mesh size constraint (CBlockLabel.cpp:130-134)."""
    lines = text.replace("\r\n", "\n").split("\n")
    out = []
    n = 0
    for ln in lines:
        if n > 0 and ln.strip():
            f = ln.split("\t")
            f[3] = repr(size)
            ln = "\t".join(f)
            n -= 1
        elif ln.strip().lower().startswith("[numblocklabels]"):
            n = int(ln.split("=")[1])
        out.append(ln)
    return "\n".join(out)


def set_key(text, key, value):
    out = []
    for ln in text.replace("\r\n", "\n").split("\n"):
        if ln.strip().lower().startswith(key.lower()):
            ln = f"{key} = {value}"
        out.append(ln)
    return "\n".join(out)


def case_tb(tmp):
    """config 1: test/TorqueBenchmark.fem as shipped (planar, linear, AGE + periodic BCs)."""
    _write_fem(os.path.join(tmp, "tb.fem"), open(os.path.join(SRC, "test", "TorqueBenchmark.fem"), newline="").read())
    return "tb"


def case_tbmag(tmp):
    """TorqueBenchmark with a functional magnetisation direction on both magnet labels (the string the reference
    evaluates per element with its Lua interpreter, static2d.cpp:513-583): operators, ^, variables and functions."""
    t = open(os.path.join(SRC, "test", "TorqueBenchmark.fem"), newline="").read().replace("\r\n", "\n")
    head, sep, tail = t.partition("[NumBlockLabels]")
    lines = tail.split("\n")
    fn = {0: "theta + 20*cos(3*theta*PI/180) - R^2/4", 2: "-(x+1)^2*10 + deg(atan2(y+0.5, x+1.5)) + max(1, 2*y, abs(x))"}
    for k, expr in fn.items():
        lines[1 + k] = lines[1 + k].rstrip() + '\t"' + expr + '"'
    _write_fem(os.path.join(tmp, "tbmag.fem"), head + sep + "\n".join(lines))
    return "tbmag"


def case_tb30(tmp):
    """TorqueBenchmark with the rotor turned by 30 degrees (AGE shift path)."""
    t = open(os.path.join(SRC, "test", "TorqueBenchmark.fem"), newline="").read().replace("\r\n", "\n")
    # <innerangle> of the AGE boundary block
    parts = t.split('<BdryName> = "AGE"')
    parts[1] = parts[1].replace("<innerangle> = 0", "<innerangle> = 30", 1)
    _write_fem(os.path.join(tmp, "tb30.fem"), '<BdryName> = "AGE"'.join(parts))
    return "tb30"


def case_temp1(tmp):
    """config 3 stand-in: cfemm/fsolver/test/Temp1.fem — nonlinear 1117 steel, Newton iterations."""
    _write_fem(os.path.join(tmp, "temp1.fem"), open(os.path.join(SRC, "cfemm", "fsolver", "test", "Temp1.fem"), newline="").read())
    return "temp1"


# ---------------------------------------------------------------------------------------------
# authored fixtures (the reference ships no axisymmetric, harmonic or strongly saturated case)
# ---------------------------------------------------------------------------------------------
STEEL_BH = [(0, 0), (0.7004, 238.7325), (1.351, 795.775), (1.624, 3183.1), (1.77, 7957.75), (2, 31831),
            (2.13, 79577.5), (2.25, 159155), (2.46, 318310)]  # 1117 steel of cfemm/fsolver/test/Temp1.fem


def fem_text(problemtype, frequency, precision, bdrys, blocks, circuits, points, segments, arcs, labels, units="centimeters",
             acsolver=0):
    L = ["[Format]      =  4.0", f"[Frequency]   =  {frequency}", f"[Precision]   =  {precision}", "[MinAngle]    =  30",
         "[DoSmartMesh] =  1", "[Depth]       =  1", f"[LengthUnits] =  {units}", f"[ProblemType] =  {problemtype}",
         "[Coordinates] =  cartesian", f"[ACSolver]    =  {acsolver}", "[PrevType]    =  0", '[PrevSoln]    =  ""',
         '[Comment]     =  "authored fixture"', "[PointProps]   = 0", f"[BdryProps]   = {len(bdrys)}"]
    for b in bdrys:
        d = {"BdryType": 0, "A_0": 0, "A_1": 0, "A_2": 0, "Phi": 0, "c0": 0, "c0i": 0, "c1": 0, "c1i": 0, "Mu_ssd": 0,
             "Sigma_ssd": 0, "innerangle": 0, "outerangle": 0}
        d.update(b)
        L += ["  <BeginBdry>", f'    <BdryName> = "{d.pop("name")}"'] + [f"    <{k}> = {v}" for k, v in d.items()] + ["  <EndBdry>"]
    L.append(f"[BlockProps]  = {len(blocks)}")
    for b in blocks:
        d = {"Mu_x": 1, "Mu_y": 1, "H_c": 0, "H_cAngle": 0, "J_re": 0, "J_im": 0, "Sigma": 0, "d_lam": 0, "Phi_h": 0,
             "Phi_hx": 0, "Phi_hy": 0, "LamType": 0, "LamFill": 1, "NStrands": 0, "WireD": 0}
        bh = b.get("bh", [])
        d.update({k: v for k, v in b.items() if k not in ("name", "bh")})
        L += ["  <BeginBlock>", f'    <BlockName> = "{b["name"]}"'] + [f"    <{k}> = {v}" for k, v in d.items()]
        L.append(f"    <BHPoints> = {len(bh)}")
        L += [f"      {x!r}\t{y!r}" for x, y in bh]
        L.append("  <EndBlock>")
    L.append(f"[CircuitProps]  = {len(circuits)}")
    for c in circuits:
        L += ["  <BeginCircuit>", f'    <CircuitName> = "{c["name"]}"', f"    <TotalAmps_re> = {c.get('amps', 0)}",
              f"    <TotalAmps_im> = {c.get('amps_im', 0)}", f"    <CircuitType> = {c.get('type', 1)}", "  <EndCircuit>"]
    L.append(f"[NumPoints] = {len(points)}")
    L += [f"{x!r}\t{y!r}\t0\t0" for x, y in points]
    L.append(f"[NumSegments] = {len(segments)}")
    L += [f"{a}\t{b}\t-1\t{bd}\t0\t0" for a, b, bd in segments]
    L.append(f"[NumArcSegments] = {len(arcs)}")
    L += [f"{a}\t{b}\t{ang}\t{seg}\t{bd}\t0\t0\t{seg}" for a, b, ang, seg, bd in arcs]
    L.append("[NumHoles] = 0")
    L.append(f"[NumBlockLabels] = {len(labels)}")
    L += [f"{x!r}\t{y!r}\t{blk}\t{size!r}\t{circ}\t{magdir}\t0\t{turns}\t{ext}" for x, y, blk, size, circ, magdir, turns, ext in labels]
    return "\n".join(L) + "\n"


def solenoid(problemtype, frequency, precision, mesh=0.25, amps=25, steel_extra=None, coil_sigma=0, magnet_sigma=0, acsolver=0,
             bar_amps=None):
    """Solenoid with a saturating 1117-steel plunger on the axis, a series-connected 200-turn coil and
    a magnet ring, inside a half disc of radius 10 cm with A=0 on the arc (cross-section in r>=0)."""
    steel = {"name": "1117 Steel", "Mu_x": 1777, "Mu_y": 1777, "bh": STEEL_BH}
    steel.update(steel_extra or {})
    blocks = [{"name": "Air"}, {"name": "Coil", "Sigma": coil_sigma}, steel,
              {"name": "Magnet", "Mu_x": 1.049, "Mu_y": 1.049, "H_c": 979000, "Sigma": magnet_sigma}]
    pts = [(0.0, -10.0), (0.0, 10.0), (1.0, -2.0), (2.0, -2.0), (2.0, 2.0), (1.0, 2.0), (0.0, -1.0), (0.8, -1.0), (0.8, 3.0),
           (0.0, 3.0), (3.0, -0.5), (4.0, -0.5), (4.0, 0.5), (3.0, 0.5)]
    segs = [(0, 6, 0), (6, 9, 0), (9, 1, 0), (2, 3, 0), (3, 4, 0), (4, 5, 0), (5, 2, 0), (6, 7, 0), (7, 8, 0), (8, 9, 0),
            (10, 11, 0), (11, 12, 0), (12, 13, 0), (13, 10, 0)]
    arcs = [(0, 1, 180, 5, 1)]
    labels = [(5.0, 5.0, 1, 4 * mesh, 0, 0, 1, 0), (1.5, 0.0, 2, mesh, 1, 0, 200, 0), (0.4, 1.0, 3, mesh, 0, 0, 1, 0),
              (3.5, 0.0, 4, mesh, 0, 90, 1, 0)]
    circuits = [{"name": "coil", "amps": amps, "type": 1}]
    if bar_amps is not None:
        # solid copper bar carrying a prescribed total current: parallel circuit over a conductor →
        # the voltage gradient becomes an extra unknown in the harmonic formulations (Case 2)
        blocks.append({"name": "Copper", "Sigma": 58})
        n0 = len(pts)
        pts += [(5.0, -1.0), (6.0, -1.0), (6.0, 1.0), (5.0, 1.0)]
        segs += [(n0, n0 + 1, 0), (n0 + 1, n0 + 2, 0), (n0 + 2, n0 + 3, 0), (n0 + 3, n0, 0)]
        labels.append((5.5, 0.0, 5, mesh, 2, 0, 1, 0))
        circuits.append({"name": "bar", "amps": bar_amps, "type": 0})
    return fem_text(problemtype, frequency, precision, [{"name": "A=0", "BdryType": 0}], blocks,
                    circuits, pts, segs, arcs, labels, acsolver=acsolver)


def case_axi1(tmp):
    """config 4 stand-in: StaticAxisymmetric, nodes on the r=0 axis, series circuit, magnet, B-H steel (Newton)."""
    _write_fem(os.path.join(tmp, "axi1.fem"), solenoid("axisymmetric", 0, "1e-08", amps=400))
    return "axi1"


def case_sat2d(tmp):
    """config 3 stand-in: the same cross-section solved as a planar problem — strongly saturated
    steel, several Newton steps, parallel circuit with a bulk conductor (Case 0 voltage gradient)."""
    _write_fem(os.path.join(tmp, "sat2d.fem"), solenoid("planar", 0, "1e-08", coil_sigma=58, amps=2000))
    return "sat2d"


def case_harm2d(tmp):
    """config 4 stand-in: Harmonic2D at 50 Hz — eddy currents in a conductive magnet and in the steel,
    B-H steel (successive approximation), wound series coil (Case 1) and a solid bar with prescribed
    total current (Case 2: circuit unknown appended to the system)."""
    _write_fem(os.path.join(tmp, "harm2d.fem"), solenoid("planar", 50, "1e-08", amps=400, magnet_sigma=0.667, bar_amps=500,
                                                          steel_extra={"Sigma": 5.8, "Phi_h": 10}))
    return "harm2d"


def case_harmaxi(tmp):
    """config 4 stand-in: HarmonicAxisymmetric at 50 Hz, same ingredients, nodes on the axis."""
    _write_fem(os.path.join(tmp, "harmaxi.fem"), solenoid("axisymmetric", 50, "1e-08", amps=100, magnet_sigma=0.667, bar_amps=200,
                                                           steel_extra={"Sigma": 5.8}))
    return "harmaxi"


def case_harm2dn(tmp):
    """harm2d solved with the Newton AC solver ([ACSolver]=1): side matrices Mh/Ms/Ma and KludgeSolve."""
    _write_fem(os.path.join(tmp, "harm2dn.fem"), set_key(solenoid("planar", 50, "1e-08", amps=400, magnet_sigma=0.667, bar_amps=500,
                                                                  steel_extra={"Sigma": 5.8, "Phi_h": 10}), "[ACSolver]", 1))
    return "harm2dn"


def case_harmaxin(tmp):
    """harmaxi solved with the Newton AC solver."""
    _write_fem(os.path.join(tmp, "harmaxin.fem"), set_key(solenoid("axisymmetric", 50, "1e-08", amps=100, magnet_sigma=0.667, bar_amps=200,
                                                                   steel_extra={"Sigma": 5.8}), "[ACSolver]", 1))
    return "harmaxin"


def case_tbh(tmp):
    """TorqueBenchmark geometry (air-gap element + periodic pairs) at 50 Hz with a conducting,
    current-carrying rotor block: exercises the harmonic sign of the AGE blocks and complex periodicity."""
    t = open(os.path.join(SRC, "test", "TorqueBenchmark.fem"), newline="").read().replace("\r\n", "\n")
    t = set_key(t, "[Frequency]", 50)
    t = set_key(t, "[Precision]", "1e-08")
    head, _, tail = t.partition("<BeginBlock>")           # first block = Air
    blk2_head, _, blk2_tail = tail.partition("<BeginBlock>")  # second block: give it sigma and J
    blk2_tail = blk2_tail.replace("<J_re> = 0", "<J_re> = 1", 1).replace("<Sigma> = 0", "<Sigma> = 0.667", 1)
    _write_fem(os.path.join(tmp, "tbh.fem"), head + "<BeginBlock>" + blk2_head + "<BeginBlock>" + blk2_tail)
    return "tbh"


def case_tbhn(tmp):
    """Newton AC solver with periodic pairs and the air-gap element: the tbh problem with the outer regions ("Air",
    "Ext": the blocks on either side of the periodic boundary circles) given the 1117-steel B-H curve, so that the
    side matrices Mh/Ms/Ma carry entries on periodic nodes (cspars.cpp:648-668, 732-752)."""
    case_tbh(tmp)
    t = open(os.path.join(tmp, "tbh.fem")).read()
    os.remove(os.path.join(tmp, "tbh.fem"))
    t = set_key(t, "[ACSolver]", 1)
    parts = t.split("<BeginBlock>")
    bh = "<BHPoints> = %d\n" % len(STEEL_BH) + "".join(f"      {x!r}\t{y!r}\n" for x, y in STEEL_BH)
    for i in (1, 2):  # Air, Ext
        parts[i] = parts[i].replace("<Mu_x> = 1", "<Mu_x> = 1777", 1).replace("<Mu_y> = 1", "<Mu_y> = 1777", 1)
        parts[i] = parts[i].replace("<BHPoints> = 0\n", bh, 1)
    _write_fem(os.path.join(tmp, "tbhn.fem"), "<BeginBlock>".join(parts))
    return "tbhn"


CASES = {"tb": case_tb, "tbmag": case_tbmag, "tb30": case_tb30, "temp1": case_temp1, "axi1": case_axi1, "sat2d": case_sat2d,
         "harm2d": case_harm2d, "harmaxi": case_harmaxi, "tbh": case_tbh,
         "harm2dn": case_harm2dn, "harmaxin": case_harmaxin, "tbhn": case_tbhn}


def run_case(name):
    tmp = os.path.join("/tmp", "xfemm_golden_" + name)
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    base = CASES[name](tmp)
    subprocess.run([os.path.join(REF, "fmesher"), base + ".fem"], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = os.path.join(HERE, name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    shutil.copy(os.path.join(tmp, base + ".fem"), os.path.join(out, name + ".fem"))
    for ext in (".node", ".ele", ".edge", ".pbc"):
        # mesh files are stored gzip-compressed; tests/util.py stages them into a temp dir
        with open(os.path.join(tmp, base + ext), "rb") as fi, gzip.open(os.path.join(out, name + ext + ".gz"), "wb", 9) as fo:
            fo.write(fi.read())
    dump = os.path.join(tmp, "dump")
    r = subprocess.run([os.path.join(REF, "ref_harness"), "dump", os.path.join(tmp, base), dump, "1000"], check=True,
                       stdout=subprocess.PIPE, text=True)
    meta = json.loads(r.stdout.strip().splitlines()[-1])
    keep = ["nodes", "node_bc", "elems", "pbc", "V", "A", "mu", "circ"]
    steps = meta["solves"]
    for s in sorted({0, steps - 1}):
        keep += [f"step{s}_{k}" for k in ("rowptr", "col", "val", "rhs", "vin", "vout")]
        keep += [f"step{s}_side{h}_{k}" for h in (1, 2, 3) for k in ("rowptr", "col", "val")]
    arrays = {}
    for k in keep:
        p = os.path.join(dump, k + ".npy")
        if os.path.exists(p):
            arrays[k] = np.load(p)
    a = 0
    while os.path.exists(os.path.join(dump, f"age{a}_meta.npy")):
        for k in ("meta", "qn", "qw"):
            arrays[f"age{a}_{k}"] = np.load(os.path.join(dump, f"age{a}_{k}.npy"))
        a += 1
    np.savez_compressed(os.path.join(out, "ref.npz"), **arrays)
    with open(os.path.join(out, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(name, {k: meta[k] for k in ("n_nodes", "n_elems", "n_pbc", "bandwidth", "solves", "iters")})


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
