"""Test helpers: staging of golden fixtures and comparisons against the reference dumps."""
import gzip
import json
import os
import shutil

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def stage_case(name, tmpdir):
    """Copies <name>.fem and the (gzip-compressed) fmesher output into tmpdir; returns the basename path."""
    src = os.path.join(GOLDEN, name)
    shutil.copy(os.path.join(src, name + ".fem"), os.path.join(tmpdir, name + ".fem"))
    for ext in (".node", ".ele", ".edge", ".pbc"):
        with gzip.open(os.path.join(src, name + ext + ".gz"), "rb") as fi, open(os.path.join(tmpdir, name + ext), "wb") as fo:
            fo.write(fi.read())
    return os.path.join(str(tmpdir), name)


def load_ref(name):
    ref = dict(np.load(os.path.join(GOLDEN, name, "ref.npz")))
    meta = json.load(open(os.path.join(GOLDEN, name, "meta.json")))
    return ref, meta


def age_cells(qn):
    """10 ring nodes per air-gap cell from the quadNode table (static2d.cpp:277-330)."""
    n = len(qn) - 1
    cells = []
    for k in range(n):
        qm = qn[n - 1 if k - 1 < 0 else k - 1]
        q0, q1 = qn[k], qn[k + 1]
        q2 = qn[1 if k + 2 > n else k + 2]
        cells.append([qm[0], q0[0], q0[1], q1[1], q2[1], qm[2], q0[2], q0[3], q1[3], q2[3]])
    return np.array(cells, dtype=np.int32)


def upper_from_full(rowptr, col, vals=None):
    """Upper triangle (incl. diagonal) of a full CSR matrix as (rowptr, col[, vals]) — the storage
    order of the reference's per-row linked lists (spars.h:25-35)."""
    n = len(rowptr) - 1
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    m = col >= rows
    urp = np.zeros(n + 1, dtype=np.int64)
    np.add.at(urp, rows[m] + 1, 1)
    urp = np.cumsum(urp)
    if vals is None:
        return urp, col[m]
    return urp, col[m], vals[m]


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
