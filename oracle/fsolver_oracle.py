"""fsolver_oracle — CPU restatement of the reference's planar magnetostatic hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg; never by the product (xfemm_b200/).  Parity status: PINNED against the unmodified reference
(tests/test_oracle.py compares with the golden vectors in tests/golden/ that oracle/_ref produced).

Structure: numpy for the file parsing and table building, plain C (oracle_core.c → liboracle.so,
loaded with ctypes) for the loops the reference runs serially: Cuthill–McKee, the element comb sort,
Put/Get/AddTo on sorted upper-triangular rows, SetValue, Periodicity, the element loop of Static2D,
SSOR and the PCG recurrence.  Every piece cites the reference file:line it follows.

Scope of the restatement: Frequency == 0, planar, [PrevSoln] empty, no Lua magnetisation functions
(the configurations of BASELINE.json configs 1-3).  The harmonic and axisymmetric formulations are
checked against oracle/_ref directly.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
PI = 3.141592653589793238462643383
MUO = 1.2566370614359173e-6
UNITS = [2.54, 0.1, 1.0, 100.0, 0.00254, 1.0e-04]               # static2d.cpp:67
LENGTH_CONV_METERS = [0.0254, 0.001, 0.01, 1.0, 2.54e-5, 1.0e-6]  # femmconstants.cpp

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "oracle_core.c")):
            subprocess.run(["make", "-s", "-C", HERE, "liboracle"], check=True)
        L = C.CDLL(LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_b.restype = C.POINTER(C.c_double)
        L.orc_V.restype = C.POINTER(C.c_double)
        L.orc_get.restype = C.c_double
        L.orc_nnz.restype = C.c_long
        for f, a in {
            "orc_create": [C.c_int, C.c_int], "orc_destroy": [C.c_void_p], "orc_b": [C.c_void_p], "orc_V": [C.c_void_p],
            "orc_set_precision": [C.c_void_p, C.c_double], "orc_put": [C.c_void_p, C.c_double, C.c_int, C.c_int],
            "orc_get": [C.c_void_p, C.c_int, C.c_int], "orc_addto": [C.c_void_p, C.c_double, C.c_int, C.c_int],
            "orc_wipe": [C.c_void_p], "orc_setvalue": [C.c_void_p, C.c_int, C.c_double],
            "orc_periodicity": [C.c_void_p, C.c_int, C.c_int, C.c_int], "orc_pcgsolve": [C.c_void_p, C.c_int, C.c_int],
            "orc_nnz": [C.c_void_p], "orc_export": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
            "orc_sort_elements": [C.c_int, C.c_void_p, C.c_void_p],
            "orc_cuthill": [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)],
            "orc_static2d_elements": [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9,
            "orc_multa": [C.c_void_p, C.c_void_p, C.c_void_p], "orc_multpc": [C.c_void_p, C.c_void_p, C.c_void_p],
        }.items():
            getattr(L, f).argtypes = a
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------------------------
# .fem parsing (feasolver.cpp:181-563 and the C*Prop::fromStream parsers) — only the keys the
# static planar path reads
# --------------------------------------------------------------------------------------------
def parse_fem(path):
    lines = [ln.strip() for ln in open(path, newline="").read().replace("\r\n", "\n").split("\n")]
    P = {"precision": 1e-8, "frequency": 0.0, "lengthunits": 0, "coordinates": 0, "problemtype": 0, "depth": -1.0,
         "points": [], "bdrys": [], "blocks": [], "circuits": [], "labels": []}
    units = {"inches": 0, "millimeters": 1, "centimeters": 2, "meters": 3, "mils": 4, "microns": 5}
    i = 0

    def kv(ln):
        k, _, v = ln.partition("=")
        return k.strip().lower(), v.strip()

    def block(i, end):
        d = {}
        i += 1  # <Begin…>
        while lines[i].lower() != end:
            if lines[i]:
                k, v = kv(lines[i])
                if k == "<bhpoints>":
                    n = int(v)
                    pts = []
                    while len(pts) < 2 * n:
                        i += 1
                        pts += [float(x) for x in lines[i].split()]
                    d["bh"] = np.array(pts).reshape(n, 2)
                    d[k] = n
                else:
                    d[k] = v
            i += 1
        return d, i

    while i < len(lines):
        ln = lines[i]
        if not ln:
            i += 1
            continue
        k, v = kv(ln)
        if k in ("[precision]", "[frequency]", "[depth]"):
            P[k.strip("[]")] = float(v)
        elif k == "[lengthunits]":
            P["lengthunits"] = units.get(v.lower(), 0)
        elif k == "[coordinates]":
            P["coordinates"] = 1 if v.lower() == "polar" else 0
        elif k == "[problemtype]":
            P["problemtype"] = 1 if v.lower() == "axisymmetric" else 0
        elif k in ("[pointprops]", "[bdryprops]", "[blockprops]", "[circuitprops]"):
            end, key = {"[pointprops]": ("<endpoint>", "points"), "[bdryprops]": ("<endbdry>", "bdrys"),
                        "[blockprops]": ("<endblock>", "blocks"), "[circuitprops]": ("<endcircuit>", "circuits")}[k]
            for _ in range(int(v)):
                i += 1
                while not lines[i]:
                    i += 1
                d, i = block(i, end)
                P[key].append(d)
        elif k == "[numblocklabels]":
            for _ in range(int(v)):
                i += 1
                f = lines[i].split()
                ext = int(f[8]) if len(f) > 8 else 0
                P["labels"].append({"blocktype": int(f[2]) - 1, "incircuit": int(f[4]) - 1, "magdir": float(f[5]),
                                    "turns": int(float(f[7])), "isdefault": bool(ext & 2), "isexternal": bool(ext & 1),
                                    "magdirfctn": lines[i].split('"')[1] if '"' in lines[i] else ""})
        elif k in ("[numpoints]", "[numsegments]", "[numarcsegments]", "[numholes]"):
            i += int(v)
        i += 1
    return P


def fnum(d, key, default=0.0):
    return float(d.get(key, default))


# --------------------------------------------------------------------------------------------
# B-H slopes (CMaterialProp.cpp:127-348), omega = 0 branch
# --------------------------------------------------------------------------------------------
def get_slopes(B, H, lamtype, lamfill):
    B, H = B.copy(), H.copy()
    n = len(B)
    mu_x = B[1] / (MUO * abs(H[1]))
    ok, processed = False, False
    slope = None
    while not ok:
        M = np.zeros((n, n))
        b = np.zeros(n)
        l1 = B[1] - B[0]
        M[0, 0], M[0, 1], b[0] = 4. / l1, 2. / l1, 6. * (H[1] - H[0]) / (l1 * l1)
        l1 = B[n - 1] - B[n - 2]
        M[n - 1, n - 1], M[n - 1, n - 2], b[n - 1] = 4. / l1, 2. / l1, 6. * (H[n - 1] - H[n - 2]) / (l1 * l1)
        for i in range(1, n - 1):
            l1, l2 = B[i] - B[i - 1], B[i + 1] - B[i]
            M[i, i - 1], M[i, i], M[i, i + 1] = 2. / l1, 4. * (l1 + l2) / (l1 * l2), 2. / l2
            b[i] = 6. * (H[i] - H[i - 1]) / (l1 * l1) + 6. * (H[i + 1] - H[i]) / (l2 * l2)
        slope = np.linalg.solve(M, b)
        ok = True
        for i in range(1, n):
            d0, d1, u0, u1, L = slope[i - 1], slope[i], H[i - 1], H[i], B[i] - B[i - 1]
            c0 = d0
            c1 = -(2. * (2. * d0 * L + d1 * L + 3. * u0 - 3. * u1)) / (L * L)
            c2 = (3. * (d0 * L + d1 * L + 2. * u0 - 2. * u1)) / (L * L * L)
            X0 = X1 = -1.
            disc = c1 * c1 - 4. * c0 * c2
            if c2 == 0:
                if c1 != 0:
                    X0 = -c0 / c1
            elif disc > 0:
                disc = math.sqrt(disc)
                X0, X1 = -(c1 + disc) / (2. * c2), (-c1 + disc) / (2. * c2)
            if (0. <= X0 <= L) or (0. <= X1 <= L):
                ok = False
        if not ok:
            Bn, Hn = B.copy(), H.copy()
            for i in range(1, n - 1):
                Bn[i] = (B[i - 1] + B[i] + B[i + 1]) / 3.
                Hn[i] = (H[i - 1] + H[i] + H[i + 1]) / 3.
            B, H = Bn, Hn
        if ok and not processed:
            if lamtype == 0 and lamfill != 1:
                for i in range(1, n):
                    mu = lamfill * B[i] / H[i] + (1. - lamfill) * MUO
                    B[i] = abs(mu * H[i])
                    H[i] = B[i] / mu
                ok = False
            processed = True
    return B, H, slope, mu_x


# --------------------------------------------------------------------------------------------
# mesh files (fsolver.cpp:350-718)
# --------------------------------------------------------------------------------------------
def load_mesh(base, P):
    nd = np.loadtxt(base + ".node", skiprows=1, ndmin=2)
    xy = nd[:, 1:3] * (100 * LENGTH_CONV_METERS[P["lengthunits"]])
    mk = nd[:, 3].astype(int)
    node_bc = np.where(mk > 1, mk - 2, -1).astype(np.int32)
    el = np.loadtxt(base + ".ele", skiprows=1, ndmin=2).astype(np.int64)
    tri = el[:, 1:4].astype(np.int32)
    lbl = (el[:, 4] - 1).astype(np.int32)
    default = -1
    for i, L in enumerate(P["labels"]):
        if L["isdefault"]:
            default = i
    lbl[lbl < 0] = default
    assert (lbl >= 0).all(), "MISSINGMATPROPS"
    ed = np.loadtxt(base + ".edge", skiprows=1, ndmin=2).astype(np.int64)
    edges = ed[:, 1:3].astype(np.int32)
    emark = ed[:, 3]
    with open(base + ".pbc") as f:
        toks = f.read().split("\n")
    npbc = int(toks[0].split()[0])
    pbc = np.array([[int(x) for x in toks[1 + k].split()[1:4]] for k in range(npbc)], dtype=np.int32).reshape(-1, 3)
    pos = 1 + npbc
    ages = []
    nage = int(toks[pos].split()[0]) if pos < len(toks) and toks[pos].strip() else 0
    pos += 1
    for _ in range(nage):
        name = toks[pos]
        f = toks[pos + 1].split()
        g = {"name": name, "fmt": int(f[0]), "inner": float(f[1]), "outer": float(f[2]), "ri": float(f[3]),
             "ro": float(f[4]), "arclen": float(f[5]), "n": int(f[8]), "ishift": float(f[9]), "oshift": float(f[10])}
        q = np.array([toks[pos + 2 + k].split() for k in range(g["n"] + 1)], dtype=float)
        g["qn"] = q[:, 0::2].astype(np.int32)
        g["qw"] = q[:, 1::2]
        ages.append(g)
        pos += 2 + g["n"] + 1
    # boundary markers on element sides (fsolver.cpp:626-697)
    e3 = -np.ones((len(tri), 3), dtype=np.int32)
    members = [[] for _ in range(len(xy))]
    for i, t in enumerate(tri):
        for v in t:
            members[v].append(i)
    for (n0, n1), m in zip(edges, emark):
        if m < 0:
            j = -(m + 2)
            for q in members[n0]:
                t = tri[q]
                for s in range(3):
                    a, b = t[s], t[(s + 1) % 3]
                    if (a == n0 and b == n1) or (a == n1 and b == n0):
                        e3[q, s] = j
    return {"xy": xy, "node_bc": node_bc, "tri": tri, "lbl": lbl, "e": e3, "edges": edges, "pbc": pbc, "ages": ages}


def renumber(M):
    """cuthill.cpp:94-391 (+ fsolver.cpp:1341-1353 SortNodes): returns the solver-order mesh."""
    n = len(M["xy"])
    newnum = np.empty(n, dtype=np.int32)
    bw = C.c_int(0)
    edges = np.ascontiguousarray(M["edges"], dtype=np.int32)
    lib().orc_cuthill(n, len(edges), _p(edges), _p(newnum), C.byref(bw))
    out = dict(M)
    out["bandwidth"] = bw.value
    out["newnum"] = newnum
    xy = np.empty_like(M["xy"])
    xy[newnum] = M["xy"]
    bc = np.empty_like(M["node_bc"])
    bc[newnum] = M["node_bc"]
    tri = newnum[M["tri"]].astype(np.int32)
    pbc = M["pbc"].copy()
    if len(pbc):
        pbc[:, 0] = newnum[pbc[:, 0]]
        pbc[:, 1] = newnum[pbc[:, 1]]
    ages = []
    for g in M["ages"]:
        g = dict(g)
        g["qn"] = newnum[g["qn"]].astype(np.int32)
        ages.append(g)
    order = np.empty(len(tri), dtype=np.int32)
    t3 = np.ascontiguousarray(tri)
    lib().orc_sort_elements(len(tri), _p(t3), _p(order))
    out.update(xy=xy, node_bc=bc, tri=tri[order], lbl=M["lbl"][order], e=M["e"][order], pbc=pbc, ages=ages)
    return out


from oracle.age_matrix import age_matrix  # static2d.cpp:220-274, all 55 closed forms


def solve_static2d(base, max_newton=50):
    """Static2D (static2d.cpp:53-1033) on <base>.fem + mesh files; returns dict with the renumbered
    mesh, per-step systems (upper CSR), iteration counts, V and A."""
    P = parse_fem(base + ".fem")
    assert P["frequency"] == 0 and P["problemtype"] == 0
    c = PI * 4.e-05
    M = renumber(load_mesh(base, P))
    n, ne = len(M["xy"]), len(M["tri"])
    L = lib()
    # ---- materials / labels
    blocks = P["blocks"]
    bh_tables, bh_ofs = [], {}
    tot = 0
    mat = []
    for bi, bp in enumerate(blocks):
        d = {"mu_x": fnum(bp, "<mu_x>", 1), "mu_y": fnum(bp, "<mu_y>", 1), "H_c": fnum(bp, "<h_c>"), "J": fnum(bp, "<j_re>"),
             "sigma": fnum(bp, "<sigma>"), "lamtype": int(fnum(bp, "<lamtype>")), "lamfill": fnum(bp, "<lamfill>", 1),
             "bhpoints": int(bp.get("<bhpoints>", 0))}
        if d["bhpoints"] > 0:
            B, H, S, mux = get_slopes(bp["bh"][:, 0], bp["bh"][:, 1], d["lamtype"], d["lamfill"])
            d["mu_x"] = d["mu_y"] = mux
            bh_ofs[bi] = tot
            tot += len(B)
            bh_tables.append((B, H, S))
        mat.append(d)
    Bd = np.concatenate([t[0] for t in bh_tables]) if bh_tables else np.zeros(1)
    Hd = np.concatenate([t[1] for t in bh_tables]) if bh_tables else np.zeros(1)
    Sd = np.concatenate([t[2] for t in bh_tables]) if bh_tables else np.zeros(1)
    # circuits (static2d.cpp:84-167)
    circ = [{"type": int(fnum(cp, "<circuittype>")), "amps": fnum(cp, "<totalamps_re>"), "dvolts": fnum(cp, "<voltgradient_re>"),
             "case": 0, "J": 0.0, "dV": 0.0} for cp in P["circuits"]]
    labels = [dict(l) for l in P["labels"]]
    # serial circuits → one parallel circuit per label (fsolver.cpp:278-317)
    ncirc0 = len(circ)
    for l in labels:
        ic = l["incircuit"]
        if 0 <= ic < ncirc0 and circ[ic]["type"] == 1:
            nc = dict(circ[ic])
            nc["amps"] *= l["turns"]
            circ.append(nc)
            l["incircuit"] = len(circ) - 1
    for cp in circ:
        if cp["type"] == 1:
            cp["type"] = 0
    for l in labels:
        bt = mat[l["blocktype"]]
        l["wound"] = abs(l["turns"]) > 1 or bt["lamtype"] > 2
    xy, tri, lblarr = M["xy"], M["tri"], M["lbl"]
    if circ:
        I1, I2, I3 = np.zeros(len(circ)), np.zeros(len(circ)), np.zeros(len(circ))
        for i in range(ne):
            l = labels[lblarr[i]]
            ic = l["incircuit"]
            if ic == -1:
                continue
            t = tri[i]
            p0, p1 = xy[t[1], 1] - xy[t[2], 1], xy[t[2], 1] - xy[t[0], 1]
            q0, q1 = xy[t[2], 0] - xy[t[1], 0], xy[t[0], 0] - xy[t[2], 0]
            a = (p0 * q1 - p1 * q0) / 2.
            bt = mat[l["blocktype"]]
            cd = 0 if l["wound"] else bt["sigma"]
            I1[ic] += a
            I2[ic] += a * cd
            I3[ic] += bt["J"] * a * 100.
        for i, cp in enumerate(circ):
            if cp["type"] == 0:
                if I2[i] == 0:
                    cp["case"] = 1
                    cp["J"] = 0. if I1[i] == 0 else 0.01 * (cp["amps"] - I3[i]) / I1[i]
                else:
                    cp["case"] = 0
                    cp["dV"] = -0.01 * (cp["amps"] - I3[i]) / I2[i]
            else:
                cp["case"] = 0
                cp["dV"] = cp["dvolts"]
    lab = np.zeros((len(labels), 8))
    bh_off = np.zeros(len(labels), dtype=np.int32)
    nonlinear_label = np.zeros(len(labels), dtype=bool)
    for li, l in enumerate(labels):
        bt = mat[l["blocktype"]]
        t = bt["lamfill"]
        mu1 = mu2 = -1.0
        if bt["lamtype"] == 0:
            mu1, mu2 = bt["mu_x"] * t + (1. - t), bt["mu_y"] * t + (1. - t)
        if bt["lamtype"] == 1:
            mu = bt["mu_x"]
            mu1, mu2 = mu * t + (1. - t), mu / (t + mu * (1. - t))
        if bt["lamtype"] == 2:
            mu = bt["mu_y"]
            mu2, mu1 = mu * t + (1. - t), mu / (t + mu * (1. - t))
        if bt["lamtype"] > 2:
            mu1 = mu2 = 1.0
        tc = 0.0
        if l["incircuit"] >= 0:
            cp = circ[l["incircuit"]]
            if cp["case"] == 1:
                tc = cp["J"]
            if cp["case"] == 0:
                tc = -cp["dV"] * bt["sigma"]
        lab[li] = [mu1, mu2, bt["J"] + tc, bt["H_c"], l["magdir"], t, bt["lamtype"], bt["bhpoints"]]
        bh_off[li] = bh_ofs.get(l["blocktype"], 0)
        nonlinear_label[li] = bt["bhpoints"] != 0
    bdry = np.zeros((max(1, len(P["bdrys"])), 4))
    for bi, bp in enumerate(P["bdrys"]):
        bdry[bi] = [fnum(bp, "<bdrytype>"), fnum(bp, "<c0>"), fnum(bp, "<c1>"), 0]
    el8 = np.zeros((ne, 8), dtype=np.int32)
    el8[:, 0:3] = tri
    el8[:, 3] = lblarr
    el8[:, 4] = [labels[l]["blocktype"] for l in lblarr]
    el8[:, 5:8] = M["e"]
    el8 = np.ascontiguousarray(el8)
    xyc = np.ascontiguousarray(xy)
    mu12 = -np.ones((ne, 2))
    linear = not nonlinear_label[lblarr].any()

    mat_h = L.orc_create(n, M["bandwidth"])
    L.orc_set_precision(mat_h, P["precision"])
    b = np.ctypeslib.as_array(L.orc_b(mat_h), shape=(n,))
    V = np.ctypeslib.as_array(L.orc_V(mat_h), shape=(n,))
    relax = 1.0
    res = lastres = 0.0
    it = 0
    systems, iters = [], []
    while True:
        if it > 0:
            L.orc_wipe(mat_h)
        # air-gap elements, static2d.cpp:192-350
        for g in M["ages"]:
            N = g["n"]
            dt = (PI / 180.) * (g["arclen"] / N)
            K = 2. * (g["ro"] - g["ri"]) / (dt * (g["ro"] + g["ri"]))
            ci, co = g["ishift"], g["oshift"]
            if ci > co:
                ci, co = ci - co, 0
            else:
                ci, co = 1 - co + ci, 1
            MG = age_matrix(ci, co, K)
            qn, qw = g["qn"], g["qw"]
            for k in range(N):
                km, k2 = (N - 1 if k - 1 < 0 else k - 1), (1 if k + 2 > N else k + 2)
                nn = [qn[km, 0], qn[k, 0], qn[k, 1], qn[k + 1, 1], qn[k2, 1], qn[km, 2], qn[k, 2], qn[k, 3], qn[k + 1, 3], qn[k2, 3]]
                ww = [qw[km, 0], qw[k, 0], qw[k, 1], qw[k + 1, 1], qw[k2, 1], qw[km, 2], qw[k, 2], qw[k, 3], qw[k + 1, 3], qw[k2, 3]]
                if k == 0 and g["fmt"] == 1:
                    ww[0], ww[5] = -ww[0], -ww[5]
                if k + 1 == N and g["fmt"] == 1:
                    ww[4], ww[9] = -ww[4], -ww[9]
                for ii in range(10):
                    for jj in range(ii, 10):
                        L.orc_addto(mat_h, MG[ii, jj] * ww[ii] * ww[jj], int(nn[ii]), int(nn[jj]))
        L.orc_static2d_elements(mat_h, it, ne, _p(el8), _p(xyc), _p(lab), _p(bh_off), _p(Bd), _p(Hd), _p(Sd), _p(bdry), _p(mu12))
        # point currents and fixed values, static2d.cpp:818-926
        for i in range(n):
            if M["node_bc"][i] >= 0:
                b[i] += 0.01 * fnum(P["points"][M["node_bc"][i]], "<i_re>")
        for i in range(n):
            if M["node_bc"][i] >= 0:
                pp = P["points"][M["node_bc"][i]]
                if fnum(pp, "<i_re>") == 0 and fnum(pp, "<i_im>") == 0:
                    L.orc_setvalue(mat_h, i, fnum(pp, "<a_re>") / c)
        u = UNITS[P["lengthunits"]]
        for i in range(ne):
            for j in range(3):
                s = el8[i, 5 + j]
                if s >= 0 and bdry[s, 0] == 0:
                    bp = P["bdrys"][s]
                    for node in (el8[i, j], el8[i, (j + 1) % 3]):
                        x, y = xy[node]
                        if P["coordinates"] == 0:
                            a = fnum(bp, "<a_0>") + (x / u) * fnum(bp, "<a_1>") + (y / u) * fnum(bp, "<a_2>")
                        else:
                            r = math.sqrt(x * x + y * y)
                            th = 0 if (x == 0 and y == 0) else math.atan2(y, x) / 0.01745329251994329576923690768
                            a = fnum(bp, "<a_0>") + (r / u) * fnum(bp, "<a_1>") + th * fnum(bp, "<a_2>")
                        a *= math.cos(fnum(bp, "<phi>") * 0.01745329251994329576923690768)
                        L.orc_setvalue(mat_h, int(node), a / c)
        for (pi, pj, pt) in M["pbc"]:
            L.orc_periodicity(mat_h, int(pi), int(pj), int(pt))
        nnz = L.orc_nnz(mat_h)
        rp, col, val = np.empty(n + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        L.orc_export(mat_h, _p(rp), _p(col), _p(val))
        systems.append({"rowptr": rp, "col": col, "val": val, "rhs": b.copy(), "vin": V.copy()})
        V_old = V.copy()
        k = L.orc_pcgsolve(mat_h, it, 0)
        assert k >= 0, "singular"
        iters.append(k)
        systems[-1]["vout"] = V.copy()
        if not linear:
            x, y = float(((V - V_old) ** 2).sum()), float((V * V).sum())
            if y == 0:
                linear = True
            else:
                lastres, res = res, math.sqrt(x / y)
            if it > 5:
                if res > lastres and relax > 0.125:
                    relax /= 2.
                else:
                    relax += 0.1 * (1. - relax)
                V[:] = relax * V + (1.0 - relax) * V_old
        if res < 100. * P["precision"] and it > 0:
            linear = True
        it += 1
        if linear or it >= max_newton:
            break
    out = {"mesh": M, "systems": systems, "iters": iters, "V": V.copy(), "A": V * c, "mu": mu12.copy(),
           "rowptr": systems[-1]["rowptr"], "col": systems[-1]["col"], "precision": P["precision"]}
    L.orc_destroy(mat_h)
    return out
