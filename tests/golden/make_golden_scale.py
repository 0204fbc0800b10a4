"""Golden vectors at BASELINE scale (configs[1] and a 231 k-node stand-in), produced by FULL runs of the UNMODIFIED
reference (oracle/_ref, built from /root/reference by oracle/Makefile) in this container.

The meshes are far too big to commit (a 2 M-node mesh is ~450 MB of text); the reference's mesher is
deterministic, so the tests re-mesh the same .fem with the same oracle/_ref/fmesher binary and compare against what
is stored here:

  golden.json   counts, bandwidth, SSOR-PCG iteration count (replayed with the reference's own MultA/MultPC/Dot),
                sha256 of the renumbered node order, of the sorted element order and of the upper-triangular
                (rowptr, col) pattern handed to PCGSolve, gap torque and stored energy from the reference's fpproc
                reading the reference's .ans, and the phase timings of one full, unscaled reference run
                (ref_harness time <base> 0 0: every periodic pair, every iteration)
  sample.npz    every 64th entry of V (internal units) and A (Wb/m), plus |A|_2 of the whole vector

Usage: python tests/golden/make_golden_scale.py [tb_231k] [tb_2M]      (19 s / ~25 min of one core each)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, HERE)
from make_golden import set_mesh_size  # noqa: E402

CASES = {"tb_231k": 0.0053, "tb_2M": 0.0018}
STRIDE = 64


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def lua_quantities(workdir, base):
    lua = os.path.join(workdir, "q.lua")
    with open(lua, "w") as f:
        f.write(f'open("{base}.fem")\nmi_loadsolution()\nprint("TORQUE " .. mo_gapintegral("AGE", 0))\n'
                'mo_groupselectblock()\nprint("ENERGY " .. mo_blockintegral(2))\n')
    r = subprocess.run([os.path.join(REF, "femmcli"), f"--lua-script={lua}"], cwd=workdir, capture_output=True, text=True)
    out = {}
    for ln in r.stdout.splitlines():
        if ln.startswith(("TORQUE ", "ENERGY ")):
            k, v = ln.split()
            out[k.lower()] = float(v)
    return out


def run_case(name):
    size = CASES[name]
    tmp = os.path.join("/tmp", "xfemm_golden_" + name)
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    src = open(os.path.join(HERE, "tb", "tb.fem"), newline="").read()
    with open(os.path.join(tmp, "case.fem"), "w", newline="") as f:
        f.write(set_mesh_size(src, size))
    t0 = time.perf_counter()
    subprocess.run([os.path.join(REF, "fmesher"), "case.fem"], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    t_mesh = time.perf_counter() - t0
    # three runs of the unmodified reference, each on its own core:
    #  a) the fsolver binary itself (writes case.ans, wall clock)   b) white-box dump   c) white-box phase timers, full run
    rund = os.path.join(tmp, "run")
    os.makedirs(rund)
    for ext in (".fem", ".node", ".ele", ".edge", ".pbc"):
        shutil.copy(os.path.join(tmp, "case" + ext), rund)
    ta = time.perf_counter()
    pa = subprocess.Popen(["taskset", "-c", "1", os.path.join(REF, "fsolver"), "case"], cwd=rund, stdout=subprocess.DEVNULL)
    pb = subprocess.Popen(["taskset", "-c", "2", os.path.join(REF, "ref_harness"), "dump", os.path.join(tmp, "case"),
                           os.path.join(tmp, "dump"), "1"], stdout=subprocess.PIPE, text=True)
    pc = subprocess.Popen(["taskset", "-c", "3", os.path.join(REF, "ref_harness"), "time", os.path.join(tmp, "case"), "0", "0"],
                          stdout=subprocess.PIPE, text=True)
    pa.wait()
    wall_fsolver = time.perf_counter() - ta
    meta = json.loads(pb.communicate()[0].strip().splitlines()[-1])
    full = json.loads(pc.communicate()[0].strip().splitlines()[-1])
    q = lua_quantities(rund, "case")
    d = os.path.join(tmp, "dump")
    nodes, elems = np.load(os.path.join(d, "nodes.npy")), np.load(os.path.join(d, "elems.npy"))
    rowptr, col = np.load(os.path.join(d, "step0_rowptr.npy")), np.load(os.path.join(d, "step0_col.npy"))
    V, A = np.load(os.path.join(d, "V.npy")), np.load(os.path.join(d, "A.npy"))
    out = os.path.join(HERE, name)
    os.makedirs(out, exist_ok=True)
    gold = {
        "mesh_size": size, "n_nodes": meta["n_nodes"], "n_elems": meta["n_elems"], "n_pbc": meta["n_pbc"],
        "bandwidth": meta["bandwidth"], "precision": meta["precision"], "nnz_upper": int(len(col)),
        "reference_iterations": meta["iters"], "sha256_nodes": sha(nodes), "sha256_elems": sha(elems),
        "sha256_pattern_upper": sha(rowptr.astype(np.int64), col.astype(np.int32)),
        "torque_Nm": q.get("torque"), "energy_J": q.get("energy"), "stride": STRIDE,
        "norm_A": float(np.linalg.norm(A)), "norm_V": float(np.linalg.norm(V)),
        "full_reference_run": {
            "where": "build container (not the GPU box), one core each, three runs side by side",
            "fmesher_s": t_mesh, "fsolver_binary_wall_s": wall_fsolver,
            "phases_s": {k: full[k] for k in ("t_load_fem", "t_load_mesh", "t_cuthill", "t_formulation", "t_assembly_first",
                                               "t_pbc", "t_pcg")},
            "pcg_secs": full["pcg_secs"], "pbc_applied": full["pbc_applied"],
            "ms_per_iteration": 1e3 * sum(full["pcg_secs"]) / max(1, sum(meta["iters"])),
        },
    }
    with open(os.path.join(out, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    np.savez_compressed(os.path.join(out, "sample.npz"), V=V[::STRIDE].copy(), A=A[::STRIDE].copy())
    print(name, json.dumps({k: gold[k] for k in ("n_nodes", "bandwidth", "reference_iterations", "torque_Nm", "energy_J")}),
          "fsolver wall %.1f s" % wall_fsolver)


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        run_case(c)
