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


CASES = {"tb": case_tb, "tb30": case_tb30, "temp1": case_temp1}


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
    r = subprocess.run([os.path.join(REF, "ref_harness"), "dump", os.path.join(tmp, base), dump], check=True,
                       stdout=subprocess.PIPE, text=True)
    meta = json.loads(r.stdout.strip().splitlines()[-1])
    keep = ["nodes", "node_bc", "elems", "pbc", "V", "A", "mu", "circ"]
    steps = meta["solves"]
    for s in sorted({0, steps - 1}):
        keep += [f"step{s}_{k}" for k in ("rowptr", "col", "val", "rhs", "vin", "vout")]
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
