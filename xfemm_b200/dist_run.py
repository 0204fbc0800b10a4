"""One mesh over several GPUs: `torchrun --nproc-per-node N -m xfemm_b200.dist_run <basename>`.

Every rank loads `<basename>.fem` + mesh, renumbers, assembles and applies the boundary conditions
on its own GPU, then the ranks solve the one linear system together with the row-partitioned PCG
(xfemm_b200/dist_pcg.py).  Rank 0 prints one JSON line and (with --out) saves the solution.
Linear magnetostatic problems only (the Newton loop of a saturating problem would call this once
per pass; not wired yet).

--backend gloo --same-device lets two ranks share ONE GPU (functional check on a 1-GPU box; NCCL
refuses two ranks per device).
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import tempfile
import time

import numpy as np


def _fem_precision(fem):
    with open(fem, errors="replace") as f:
        for ln in f:
            if ln.strip().lower().startswith("[precision]"):
                return float(ln.split("=", 1)[1])
    return 1e-8


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("basename")
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--same-device", action="store_true")
    ap.add_argument("--precision", type=float, default=None)
    ap.add_argument("--repeats", type=int, default=1)
    ap.add_argument("--check-every", type=int, default=None)
    ap.add_argument("--precond", default="jacobi", choices=["amg", "jacobi"])
    ap.add_argument("--solver", default="nccl", choices=["nccl", "peer"],
                    help="nccl: round-1 path (kernels + torch.distributed collectives per iteration, Jacobi or block-local "
                         "multigrid); peer: ONE persistent kernel per GPU with the global multigrid hierarchy, halo and "
                         "scalars as peer stores over NVLink (xfemm_dpcg_*, csrc/dist_amg.cu)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--zero-rhs", action="store_true", help="test hook: solve with a zero right-hand side (exactly zero residual)")
    a = ap.parse_args(argv)

    import torch
    import torch.distributed as dist

    import xfemm_b200 as X
    from xfemm_b200.dist_pcg import DistPCG

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = 0 if a.same_device else int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("xfemm_b200.dist_run: no CUDA device (there is no CPU solve path)")
    torch.cuda.set_device(local)
    kw = {"device_id": torch.device("cuda", local)} if a.backend == "nccl" else {}
    dist.init_process_group(a.backend, **kw)
    try:
        # every rank works on its own copy of the input files (LoadMesh/Cuthill may delete them)
        tmp = tempfile.mkdtemp(prefix=f"xfemm_dist_r{rank}_")
        stem = os.path.basename(a.basename)
        for ext in (".fem", ".node", ".ele", ".edge", ".pbc"):
            if os.path.exists(a.basename + ext):
                shutil.copy(a.basename + ext, os.path.join(tmp, stem + ext))
        base = os.path.join(tmp, stem)
        t0 = time.perf_counter()
        S = X.FSolver(base)
        S.set_option("assemble_only", 1)
        S.set_option("precond", X.PRECOND_JACOBI)
        S.load_mesh()
        S.renumber()
        S.solve()  # assembly + boundary conditions only
        if a.zero_rhs:
            glin = X.lib().xfemm_fsolver_linprob(S.h)
            X.check(X.lib().xfemm_linprob_set_rhs(glin, np.zeros(X.lib().xfemm_linprob_n(glin), dtype=np.float64)), "set_rhs")
        t_setup = time.perf_counter() - t0
        if a.solver == "peer":
            from xfemm_b200.dist_amg import DistAmgPCG
            D = DistAmgPCG(S)
            prec = a.precision if a.precision is not None else _fem_precision(base + ".fem")
            infos = [D.solve(prec) for _ in range(a.repeats)]
            ms = torch.tensor([min(i["solve_ms"] for i in infos)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            V = D.solution()
            if rank == 0:
                line = {"mode": "row-partitioned multigrid pcg, peer stores", "n_gpus": world, "backend": a.backend, "n": int(D.n),
                        "iterations": infos[-1]["iterations"], "er": infos[-1]["er"], "solve_ms": float(ms.item()),
                        "ms_per_iteration": float(ms.item()) / max(1, infos[-1]["iterations"]), "setup_s": t_setup,
                        "plan": D.stats()}
                print(json.dumps(line), flush=True)
                if a.out:
                    np.savez(a.out, V=V, perm=S.permutation())
            D.close()
            S.close()
            shutil.rmtree(tmp, ignore_errors=True)
            return
        D = DistPCG(S, precond=X.PRECOND_AMG if a.precond == "amg" else X.PRECOND_JACOBI)
        D.refresh()
        prec = a.precision if a.precision is not None else _fem_precision(base + ".fem")
        infos = []
        for _ in range(a.repeats):
            dist.barrier()
            torch.cuda.synchronize()
            infos.append(D.solve(prec, check_every=a.check_every))
        ms = torch.tensor([min(i["solve_ms"] for i in infos)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        V = D.gather_solution()
        halo = [None] * world
        dist.all_gather_object(halo, (D.n_own, D.n_halo, D.n_send))
        if rank == 0:
            line = {"mode": "row-partitioned pcg", "precond": a.precond, "n_gpus": world, "backend": a.backend, "n": int(D.n),
                    "iterations": infos[-1]["iterations"], "er": infos[-1]["er"], "solve_ms": float(ms.item()),
                    "ms_per_iteration": float(ms.item()) / max(1, infos[-1]["iterations"]),
                    "setup_s": t_setup, "blocks": [{"n_own": h[0], "n_halo": h[1], "n_send": h[2]} for h in halo]}
            print(json.dumps(line), flush=True)
            if a.out:
                np.savez(a.out, V=V, perm=S.permutation())
        D.close()
        S.close()
        shutil.rmtree(tmp, ignore_errors=True)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
