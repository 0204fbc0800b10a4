"""Row-block partitioned Jacobi-PCG of ONE mesh over several B200s (BASELINE config 5b).

One process per GPU.  Every rank loads the same problem, assembles and boundary-conditions the
global system on its own GPU (assembly is <1 % of a solve), then owns a contiguous block of rows of
the Cuthill-McKee ordered matrix (cuthill.cpp:94-391 keeps the band narrow, so a row block only
talks to its two neighbours for a simply connected mesh; periodic pairs add a few far columns).
The kernels are the library's (csrc/dist.cu, C ABI `xfemm_dist_*`); this module is the plumbing:
`torch.distributed` moves the halo of the search direction (all_to_all over NCCL/NVLink) and
all-reduces the two scalars of a CG iteration.  Nothing is copied to the host inside the loop
except the convergence check every `check_every` iterations.

The iteration is the reference's PCG recurrence (spars.cpp:219-324); the preconditioner is Jacobi
or (option) the multigrid V-cycle of each rank's own diagonal block (additive Schwarz without
overlap, no communication inside the cycle); the stopping rule is the same
||r||_M / ||r0||_M < Precision.

The planning helpers (`split_rows`, `plan_halo`) are pure host code and covered by the CPU tests
with the gloo backend; the solve itself needs the CUDA library and fails loudly without it.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np


def split_rows(rowptr: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous row blocks with (nearly) equal entry counts: [(r0, r1)] * world."""
    n = len(rowptr) - 1
    if world > n:
        raise ValueError(f"cannot split {n} rows over {world} ranks")
    total = int(rowptr[n])
    cuts = [0]
    for k in range(1, world):
        r = int(np.searchsorted(rowptr, total * k / world, side="left"))
        r = min(max(r, cuts[-1] + 1), n - (world - k))
        cuts.append(r)
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def owner_of(cols: np.ndarray, blocks: list[tuple[int, int]]) -> np.ndarray:
    starts = np.array([b[0] for b in blocks], dtype=np.int64)
    return (np.searchsorted(starts, np.asarray(cols, dtype=np.int64), side="right") - 1).astype(np.int64)


def plan_halo(halo_cols_per_rank: list[np.ndarray], blocks: list[tuple[int, int]], rank: int):
    """Exchange plan of `rank`.

    halo_cols_per_rank[q] = sorted global columns rank q needs from others.
    Returns (send_rows, send_splits, recv_splits):
      send_rows   — rows of `rank` to ship, concatenated in peer order (what peer q asked for),
      send_splits — how many go to each peer, recv_splits — how many arrive from each peer;
    the halo of `rank` in sorted order is exactly the concatenation of what arrives, because the
    blocks are ascending with the rank.
    """
    world = len(blocks)
    r0, r1 = blocks[rank]
    send_rows, send_splits = [], []
    for q in range(world):
        want = np.asarray(halo_cols_per_rank[q], dtype=np.int64)
        mine = want[(want >= r0) & (want < r1)] if q != rank else want[:0]
        send_rows.append(mine)
        send_splits.append(len(mine))
    own_halo = np.asarray(halo_cols_per_rank[rank], dtype=np.int64)
    if len(own_halo) and np.any((own_halo >= r0) & (own_halo < r1)):
        raise ValueError("halo lists an owned column")
    if len(own_halo) > 1 and np.any(np.diff(own_halo) <= 0):
        raise ValueError("halo columns must be sorted and unique")
    own = owner_of(own_halo, blocks) if len(own_halo) else np.zeros(0, dtype=np.int64)
    recv_splits = [int(np.sum(own == q)) for q in range(world)]
    return (np.concatenate(send_rows).astype(np.int32) if send_rows else np.zeros(0, np.int32)), send_splits, recv_splits


def exchange(dist, sendbuf, recvbuf, send_splits, recv_splits, rank):
    """Halo exchange of one vector.  all_to_all on NCCL; pairwise send/recv elsewhere (gloo)."""
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recvbuf, sendbuf, recv_splits, send_splits)
        return
    if sendbuf.is_cuda:  # gloo has no device send/recv: stage through the host (functional path only)
        hs, hr = sendbuf.cpu(), recvbuf.cpu()
        exchange(dist, hs, hr, send_splits, recv_splits, rank)
        recvbuf.copy_(hr)
        return
    ops, so, ro = [], 0, 0
    for q in range(len(send_splits)):
        if q != rank and send_splits[q]:
            ops.append(dist.P2POp(dist.isend, sendbuf[so:so + send_splits[q]], q))
        if q != rank and recv_splits[q]:
            ops.append(dist.P2POp(dist.irecv, recvbuf[ro:ro + recv_splits[q]], q))
        so += send_splits[q]
        ro += recv_splits[q]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


class DistPCG:
    """Partitioned solve of the resident real system of an `xfemm_b200.FSolver` handle.

    Usage (every rank):  S = FSolver(base); S.set_option("assemble_only", 1); S.load_mesh();
    S.renumber(); S.solve();  D = DistPCG(S);  info = D.solve(precision);  V = D.gather_solution()
    """

    def __init__(self, solver, rank: int | None = None, world: int | None = None, precond: int | None = None):
        import torch
        import torch.distributed as dist
        import xfemm_b200 as X

        self.torch, self.dist, self.X = torch, dist, X
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        self.lib = X.lib()
        self.glin = self.lib.xfemm_fsolver_linprob(solver.h)
        if not self.glin:
            raise RuntimeError("DistPCG needs a resident real system (solve with assemble_only=1 first)")
        self.n = self.lib.xfemm_linprob_n(self.glin)
        rowptr = np.zeros(self.n + 1, dtype=np.int32)
        X.check(self.lib.xfemm_linprob_get_pattern(self.glin, rowptr.ctypes.data_as(C.c_void_p), None))
        self.blocks = split_rows(np.asarray(rowptr), self.world)
        self.r0, self.r1 = self.blocks[self.rank]
        h = C.c_void_p()
        X.check(self.lib.xfemm_dist_create(self.glin, self.r0, self.r1, C.byref(h)))
        self.h = h
        # preconditioner: Jacobi (default), or the multigrid V-cycle of each rank's own diagonal block.
        # Measured at 2 M DOF / 2 ranks: the block-local V-cycle needs 2 360 iterations where the global
        # one needs 91 (exact subdomain solves without overlap or a shared coarse level leave the interface
        # error to CG), so it is slower than Jacobi (6 650 cheap iterations) and stays an option.
        self.precond = X.PRECOND_JACOBI if precond is None else precond
        X.check(self.lib.xfemm_dist_set_precond(h, self.precond))
        n_own, n_halo = C.c_int(), C.c_int()
        X.check(self.lib.xfemm_dist_sizes(h, C.byref(n_own), C.byref(n_halo)))
        self.n_own, self.n_halo = n_own.value, n_halo.value
        halo = np.zeros(max(1, self.n_halo), dtype=np.int32)
        X.check(self.lib.xfemm_dist_halo_columns(h, halo))
        halo = halo[: self.n_halo]
        lists = [None] * self.world
        dist.all_gather_object(lists, halo)
        send_rows, self.send_splits, self.recv_splits = plan_halo(lists, self.blocks, self.rank)
        self.n_send = len(send_rows)
        sr = np.ascontiguousarray(send_rows, dtype=np.int32)
        X.check(self.lib.xfemm_dist_set_send_rows(h, self.n_send, sr.ctypes.data_as(C.c_void_p)))
        self.stream = torch.cuda.ExternalStream(self.lib.xfemm_dist_stream(h))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.sendbuf = torch.zeros(max(1, self.n_send), dtype=torch.float64, device=dev)
        self.recvbuf = torch.zeros(max(1, self.n_halo), dtype=torch.float64, device=dev)
        self.g = torch.zeros(4, dtype=torch.float64, device=dev)  # res, pAp, res_new, res_0
        self._graphs = {}

    def close(self):
        self._graphs = {}
        if getattr(self, "h", None):
            self.lib.xfemm_dist_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh(self):
        """Pull this rank's rows of the (re-)assembled global system."""
        self.X.check(self.lib.xfemm_dist_refresh(self.h))

    def _iteration(self):
        """One CG iteration: halo of p -> u = A p, p.u -> x, r update, r.D^-1 r -> new direction."""
        dist, lib, X, g = self.dist, self.lib, self.X, self.g
        gp = g.data_ptr()
        X.check(lib.xfemm_dist_pack(self.h, C.c_void_p(self.sendbuf.data_ptr())))
        exchange(dist, self.sendbuf[: self.n_send], self.recvbuf[: self.n_halo], self.send_splits, self.recv_splits, self.rank)
        X.check(lib.xfemm_dist_unpack(self.h, C.c_void_p(self.recvbuf.data_ptr())))
        X.check(lib.xfemm_dist_spmv_dot(self.h, C.c_void_p(gp + 8)))
        dist.all_reduce(g[1:2])
        X.check(lib.xfemm_dist_update(self.h, C.c_void_p(gp), C.c_void_p(gp + 16)))
        dist.all_reduce(g[2:3])
        X.check(lib.xfemm_dist_direction(self.h, C.c_void_p(gp)))
        g[0:1].copy_(g[2:3])

    def solve(self, precision: float = 1e-8, max_iter: int | None = None, check_every: int | None = None,
              use_graph: bool | None = None):
        """Runs the iteration until ||r||_M/||r0||_M <= precision, checked every `check_every`
        iterations (the only host synchronisation).  On NCCL a batch of `check_every` iterations is
        captured once into a CUDA graph (kernels + collectives) and replayed, so the loop is not
        bound by launch overhead."""
        torch, dist, lib, X, g = self.torch, self.dist, self.lib, self.X, self.g
        max_iter = max_iter or max(4 * self.n, 1000)
        if check_every is None:
            # Jacobi needs thousands of cheap iterations, the multigrid about a hundred expensive ones
            check_every = 5 if self.precond == X.PRECOND_AMG else 50
        if use_graph is None:
            use_graph = dist.get_backend() == "nccl" and (self.precond != X.PRECOND_AMG or bool(os.environ.get("XFEMM_DIST_GRAPH")))
        its = 0
        with torch.cuda.stream(self.stream):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            X.check(lib.xfemm_dist_init(self.h, C.c_void_p(g.data_ptr())))
            dist.all_reduce(g[0:1])
            g[3:4].copy_(g[0:1])
            res0 = float(g[3].item())
            er = 0.0 if res0 == 0.0 else 1.0
            graph = self._graphs.get(check_every) if use_graph else None
            while er > precision and its < max_iter:
                if use_graph and graph is None and its > 0:
                    # communicators are warm after the first eager batch: capture the next one
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, stream=self.stream, capture_error_mode="thread_local"):
                        for _ in range(check_every):
                            self._iteration()
                    self._graphs[check_every] = graph
                if graph is not None:
                    graph.replay()
                    its += check_every
                else:
                    for _ in range(check_every):
                        self._iteration()
                    its += check_every
                res = float(g[0].item())
                if not math.isfinite(res) or res < 0.0:
                    # the single-GPU kernel reports XFEMM_ERR_DIVERGED for NaN / breakdown; so does this path
                    raise self.X.XfemmError(f"partitioned PCG diverged after {its} iterations (residual {res})")
                er = math.sqrt(res / res0)
            t1.record()
            t1.synchronize()
        return {"iterations": its, "er": er, "solve_ms": t0.elapsed_time(t1), "graph": graph is not None,
                "n_own": self.n_own, "n_halo": self.n_halo, "n_send": self.n_send}

    def gather_solution(self) -> np.ndarray:
        """Full solution vector V (solver order) on every rank."""
        own = np.zeros(self.n_own, dtype=np.float64)
        self.X.check(self.lib.xfemm_dist_get_solution(self.h, own))
        parts = [None] * self.world
        self.dist.all_gather_object(parts, own)
        return np.concatenate(parts)
