"""One mesh over several GPUs with the GLOBAL multigrid hierarchy (xfemm_dpcg_*, csrc/dist_amg.cu): peer stores + flags
over NVLink inside one persistent kernel per GPU.

* one rank: the partitioned kernel degenerates to the single-GPU recurrence — same iteration count, same solution;
* >= 2 GPUs (skipped on a one-GPU box): 2 / all ranks over real peer memory against the reference dump, iteration
  count equal to the one-rank count (the hierarchy is the same whatever the rank count).
Two ranks cannot share one GPU here: each rank's kernel is a cooperative launch that fills the device and waits for
its peers on the device.
"""
import ctypes as C
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import xfemm_b200 as X
from util import load_ref, rel_l2, stage_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(base, out, nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           "-m", "xfemm_b200.dist_run", base, "--out", out, "--backend", "nccl", "--solver", "peer"]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), XFEMM_DIST_TIMEOUT_MS="15000")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.parametrize("case", ["tb30", "temp1"])
def test_one_rank_equals_the_single_gpu_solver(case, tmp_path):
    from xfemm_b200.dist_amg import DistAmgPCG
    ref, meta = load_ref(case)
    S = X.FSolver(stage_case(case, tmp_path))
    S.set_option("assemble_only", 1)
    S.load_mesh(False)
    S.renumber(False)
    S.solve()
    L = X.lib().xfemm_fsolver_linprob(S.h)
    info1 = X.SolveInfo()
    X.check(X.lib().xfemm_linprob_pcg(L, meta["precision"], 0, X.PRECOND_AMG, 0, C.byref(info1)), "pcg")
    n = X.lib().xfemm_linprob_n(L)
    V1 = np.empty(n)
    X.check(X.lib().xfemm_linprob_get_solution(L, V1))
    D = DistAmgPCG(S, rank=0, world=1)
    info = D.solve(meta["precision"])
    V = D.solution()
    print(case, "single-GPU kernel", info1.iterations, "iterations; partitioned kernel, one rank:", info, D.stats())
    assert abs(info["iterations"] - info1.iterations) <= 1  # same recurrence, another summation order
    assert rel_l2(V, V1) <= 10 * meta["precision"]  # two runs of the same recurrence with different summation orders
    # warm start from the solution: converged at once
    info2 = D.solve(meta["precision"], warm=True)
    assert info2["iterations"] <= 2
    D.close()
    S.close()


def test_ranks_over_peer_memory(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    ref, _ = load_ref("tb30")
    base = stage_case("tb30", tmp_path)
    counts = {}
    for nproc in sorted({2, min(n, 8)}):
        out = str(tmp_path / f"x{nproc}.npz")
        info = _run(base, out, nproc)
        print(info)
        assert rel_l2(np.load(out)["V"], ref["V"]) <= 1e-6
        counts[nproc] = info["iterations"]
    one = _run(base, str(tmp_path / "x1.npz"), 1)["iterations"]
    assert all(abs(c - one) <= 1 for c in counts.values()), (one, counts)


def test_fsolver_command_line_on_two_gpus(tmp_path):
    """The C++ front end: two `fsolver_b200` processes (XFEMM_DIST_RANK / XFEMM_DIST_WORLD, one GPU each) solve ONE
    problem together — FSolver::Static2D routes its linear solve through xfemm_dpcg_*, the processes find each other
    through files — and both write the same .ans as a single-GPU run (nodal A within the solver precision)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cli = os.path.join(ROOT, "xfemm_b200", "fsolver_b200")
    job = f"xfemm_test_{os.getpid()}"
    procs, dirs = [], []
    for r in range(2):
        d = tmp_path / f"rank{r}"
        d.mkdir()
        stage_case("tb30", str(d))
        dirs.append(d)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(r), XFEMM_DIST_RANK=str(r), XFEMM_DIST_WORLD="2", XFEMM_DIST_JOB=job,
                   XFEMM_DIST_TIMEOUT_MS="15000")
        procs.append(subprocess.Popen([cli, "tb30"], cwd=str(d), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r in range(2):
        for f in (f"/dev/shm/{job}.{r}",):
            if os.path.exists(f):
                os.remove(f)
    assert all(p.returncode == 0 for p in procs), outs
    single = tmp_path / "single"
    single.mkdir()
    stage_case("tb30", str(single))
    r = subprocess.run([cli, "tb30"], cwd=str(single), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr

    def nodal_A(path):
        lines = open(path).read().split("[Solution]")[1].strip().splitlines()
        n = int(lines[0])
        return np.array([float(ln.split()[2]) for ln in lines[1:1 + n]])

    a1 = nodal_A(single / "tb30.ans")
    for d in dirs:
        assert rel_l2(nodal_A(d / "tb30.ans"), a1) <= 1e-8
