"""One mesh over several B200s: plumbing around the C ABI `xfemm_dpcg_*` (csrc/dist_amg.cu).

The solver itself is the library's: ONE persistent kernel per GPU, halo entries and the scalars of an iteration
travel as peer stores + flags over NVLink.  What is left for the host is what `torch.distributed` is for — handing
the 64-byte CUDA IPC handles of the communication arenas around once, and the barriers between the phases of a solve.
There is no CPU path: without the CUDA library the calls raise.

    S = X.FSolver(base); S.set_option("assemble_only", 1); S.load_mesh(); S.renumber(); S.solve()   # every rank
    D = DistAmgPCG(S)                   # plan + arena + handle exchange
    info = D.solve(1e-10)               # every rank; info["solve_ms"] = the kernel, CUDA events
    V = D.solution()                    # complete solution (caller order) on every rank
"""
from __future__ import annotations

import ctypes as C

import numpy as np


class DistAmgPCG:
    def __init__(self, solver=None, linprob=None, rank: int | None = None, world: int | None = None):
        import xfemm_b200 as X

        self.X = X
        self.lib = X.lib()
        self.dist = None
        if world is None or rank is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                rank, world = dist.get_rank(), dist.get_world_size()
            else:
                rank, world = 0, 1
        elif world > 1:
            import torch.distributed as dist
            self.dist = dist
        self.rank, self.world = rank, world
        self.glin = linprob if linprob is not None else self.lib.xfemm_fsolver_linprob(solver.h)
        if not self.glin:
            raise RuntimeError("DistAmgPCG needs a resident real system (solve with assemble_only=1 first)")
        self.n = self.lib.xfemm_linprob_n(self.glin)
        h = C.c_void_p()
        X.check(self.lib.xfemm_dpcg_create(self.glin, rank, world, C.byref(h)), "xfemm_dpcg_create")
        self.h = h
        if world > 1:
            import torch
            mine = (C.c_ubyte * 64)()
            X.check(self.lib.xfemm_dpcg_ipc_handle(h, mine), "xfemm_dpcg_ipc_handle")
            dev = torch.device("cuda", torch.cuda.current_device()) if self.dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=dev)
            allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
            self.dist.all_gather(allh, t)
            blob = b"".join(bytes(x.cpu().numpy().tobytes()) for x in allh)
            buf = (C.c_ubyte * (64 * world)).from_buffer_copy(blob)
            X.check(self.lib.xfemm_dpcg_connect(h, buf), "xfemm_dpcg_connect")
            self.dist.barrier()

    def stats(self):
        lp, ab = C.c_int(0), C.c_longlong(0)
        le = (C.c_longlong * 12)()
        self.X.check(self.lib.xfemm_dpcg_stats(self.h, C.byref(lp), le, C.byref(ab)), "xfemm_dpcg_stats")
        names = ["halo", "restriction", "prolongation"]
        return {"partitioned_levels": lp.value, "arena_bytes": ab.value,
                "send_entries": {f"level{l}_{names[k]}": int(le[3 * l + k]) for l in range(lp.value) for k in range(3)}}

    def solve(self, precision: float, warm: bool = False, max_iter: int = 0):
        info = self.X.SolveInfo()
        if self.dist is not None:
            self.dist.barrier()  # the kernels of all ranks start together (they wait for each other on the device)
        self.X.check(self.lib.xfemm_dpcg_solve(self.h, precision, int(warm), max_iter, C.byref(info)), "xfemm_dpcg_solve")
        return {"iterations": info.iterations, "er": info.er, "solve_ms": info.solve_ms, "status": info.status}

    def solution(self) -> np.ndarray:
        """Complete solution vector V (solver numbering) on every rank."""
        self.X.check(self.lib.xfemm_dpcg_share_solution(self.h), "xfemm_dpcg_share_solution")
        self.X.check(self.lib.xfemm_dpcg_finish(self.h), "xfemm_dpcg_finish")  # waits for the peers' rows on the device
        V = np.empty(self.n, dtype=np.float64)
        self.X.check(self.lib.xfemm_linprob_get_solution(self.glin, V), "get_solution")
        return V

    def close(self):
        if getattr(self, "h", None):
            if self.dist is not None:
                try:
                    self.dist.barrier()  # nobody is still storing into this rank's arena
                except Exception:
                    pass
            self.lib.xfemm_dpcg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.xfemm_dpcg_destroy(self.h)
                self.h = None
        except Exception:
            pass
