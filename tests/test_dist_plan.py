"""Host logic of the row-partitioned PCG (xfemm_b200/dist_pcg.py) on CPU: world_size-2 gloo run.

The kernels are replaced by numpy in this test — it pins the PARTITION, the HALO PLAN and the
exchange ordering (what the CUDA path relies on), not the arithmetic.  The CUDA path itself is
covered by tests/test_gpu_dist.py.
"""
import os
import socket

import numpy as np
import pytest

from xfemm_b200.dist_pcg import exchange, owner_of, plan_halo, split_rows


def _banded_spd(n, seed=3, far=5):
    """CSR of a narrow-band SPD matrix + a few far couplings (the shape periodic pairs produce)."""
    rng = np.random.default_rng(seed)
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - 3), i):
            if rng.random() < 0.7:
                A[i, j] = A[j, i] = -rng.random()
    for _ in range(far):
        i, j = rng.integers(0, n, 2)
        if i != j:
            A[i, j] = A[j, i] = -0.5
    A += np.diag(np.abs(A).sum(1) + 1.0)
    rowptr = [0]
    col, val = [], []
    for i in range(n):
        nz = np.nonzero(A[i])[0]
        col += list(nz)
        val += list(A[i, nz])
        rowptr.append(len(col))
    return A, np.array(rowptr), np.array(col), np.array(val)


def _halo_cols(rowptr, col, r0, r1):
    c = col[rowptr[r0]:rowptr[r1]]
    return np.unique(c[(c < r0) | (c >= r1)])


def test_split_rows_balances_entries():
    _, rowptr, _, _ = _banded_spd(200)
    for w in (1, 2, 3, 8):
        blocks = split_rows(rowptr, w)
        assert blocks[0][0] == 0 and blocks[-1][1] == 200
        assert all(blocks[k][1] == blocks[k + 1][0] for k in range(w - 1))
        assert all(b[1] > b[0] for b in blocks)
        nnz = [rowptr[b[1]] - rowptr[b[0]] for b in blocks]
        assert max(nnz) - min(nnz) <= 2 * 8 + 8  # one row of slack either side
    with pytest.raises(ValueError):
        split_rows(np.array([0, 1, 2]), 3)
    # ragged: as many ranks as rows still gives every rank one row
    assert split_rows(np.array([0, 5, 5, 9]), 3) == [(0, 1), (1, 2), (2, 3)]


def test_plan_halo_is_consistent_between_ranks():
    _, rowptr, col, _ = _banded_spd(120, far=9)
    for w in (2, 3, 4):
        blocks = split_rows(rowptr, w)
        halos = [_halo_cols(rowptr, col, *b) for b in blocks]
        plans = [plan_halo(halos, blocks, r) for r in range(w)]
        for r in range(w):
            send_rows, ss, rs = plans[r]
            assert ss[r] == 0 and rs[r] == 0
            assert sum(rs) == len(halos[r])
            # what r receives from q is what q sends to r, in the same (ascending) order
            o = 0
            for q in range(w):
                got = halos[r][o:o + rs[q]]
                o += rs[q]
                srq, ssq, _ = plans[q]
                off = sum(ssq[:r])
                assert np.array_equal(got, srq[off:off + ssq[r]])
                assert np.all(owner_of(got, blocks) == q)
    with pytest.raises(ValueError):
        plan_halo([np.array([5, 3]), np.array([], dtype=int)], [(0, 2), (2, 10)], 0)
    with pytest.raises(ValueError):
        plan_halo([np.array([1]), np.array([], dtype=int)], [(0, 2), (2, 10)], 0)


def _worker(rank, world, port, n, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, rowptr, col, val = _banded_spd(n)
        b = np.sin(np.arange(n) * 0.37) + 0.1
        blocks = split_rows(rowptr, world)
        r0, r1 = blocks[rank]
        halo = _halo_cols(rowptr, col, r0, r1)
        lists = [None] * world
        dist.all_gather_object(lists, halo)
        send_rows, ss, rs = plan_halo(lists, blocks, rank)
        n_own, n_halo = r1 - r0, len(halo)
        # local square block: owned rows, columns = owned + halo (sorted)
        cols = np.concatenate([np.arange(r0, r1), halo])
        Aloc = A[r0:r1][:, cols]
        dinv = 1.0 / np.diag(A)[r0:r1]
        x = np.zeros(n_own)
        r = b[r0:r1].copy()
        p = np.zeros(n_own + n_halo)
        p[:n_own] = dinv * r
        g = torch.zeros(3, dtype=torch.float64)
        g[0] = float(r @ (dinv * r))
        dist.all_reduce(g[0:1])
        res0 = float(g[0])
        snd = torch.zeros(max(1, len(send_rows)), dtype=torch.float64)
        rcv = torch.zeros(max(1, n_halo), dtype=torch.float64)
        its = 0
        while its < 10 * n:
            snd[:len(send_rows)] = torch.from_numpy(p[send_rows - r0])
            exchange(dist, snd[:len(send_rows)], rcv[:n_halo], ss, rs, rank)
            p[n_own:] = rcv[:n_halo].numpy()
            u = Aloc @ p
            g[1] = float(p[:n_own] @ u)
            dist.all_reduce(g[1:2])
            alpha = float(g[0] / g[1])
            x += alpha * p[:n_own]
            r -= alpha * u
            g[2] = float(r @ (dinv * r))
            dist.all_reduce(g[2:3])
            beta = float(g[2] / g[0])
            p[:n_own] = dinv * r + beta * p[:n_own]
            g[0] = g[2]
            its += 1
            if np.sqrt(float(g[0]) / res0) < 1e-12:
                break
        parts = [None] * world
        dist.all_gather_object(parts, x)
        if rank == 0:
            np.save(out, np.concatenate(parts))
    finally:
        dist.destroy_process_group()


def test_partitioned_pcg_gloo_world2(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = 150
    out = str(tmp_path / "x.npy")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    A, _, _, _ = _banded_spd(n)
    b = np.sin(np.arange(n) * 0.37) + 0.1
    x = np.load(out)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-10
