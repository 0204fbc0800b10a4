"""xfemm_b200 — B200-native fsolver hot path (element assembly + sparse symmetric solves).

This Python module is only a thin ctypes binding of the C ABI declared in include/xfemm_b200.h
(the product is the shared library xfemm_b200/libxfemm_b200.so built from csrc/ and host/).
It mirrors the way the reference is driven (`fsolver <basename>` / mexfsolver: LoadProblemFile,
LoadMesh, Cuthill, Static2D…, Write…) and is used by tests/ and bench.py.

There is no CPU fallback: if the library is missing, or no CUDA device is present when a solve
is requested, the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XFEMM_B200_LIB", os.path.join(_HERE, "libxfemm_b200.so"))  # override: experiments only

PRECOND_JACOBI = 0
PRECOND_MCSSOR = 1
PRECOND_AMG = 2


class XfemmError(RuntimeError):
    pass


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("precond", C.c_int), ("colors", C.c_int), ("status", C.c_int),
                ("er", C.c_double), ("relres", C.c_double), ("solve_ms", C.c_double), ("setup_ms", C.c_double),
                ("bytes_per_iter", C.c_double)]


class MagLabel(C.Structure):
    _fields_ = [("mu1", C.c_double), ("mu2", C.c_double), ("Jsrc", C.c_double), ("H_c", C.c_double),
                ("magdir", C.c_double), ("lamfill", C.c_double), ("lamtype", C.c_int), ("bh_offset", C.c_int),
                ("bh_points", C.c_int), ("is_external", C.c_int), ("mu1_im", C.c_double), ("mu2_im", C.c_double),
                ("sigma", C.c_double), ("Jsrc_im", C.c_double), ("dVs", C.c_double), ("circuit", C.c_int), ("pad_", C.c_int)]


_P = C.c_void_p
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")

# name -> (restype, argtypes); every symbol include/xfemm_b200.h declares
SIGNATURES = {
    "xfemm_cuda_init": (C.c_int, [C.c_int]),
    "xfemm_device_count": (C.c_int, []),
    "xfemm_last_error": (C.c_char_p, []),
    "xfemm_version": (C.c_char_p, []),
    "xfemm_transfer_bytes": (None, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]),
    "xfemm_launch_count": (C.c_longlong, []),
    "xfemm_fsolver_nnz": (C.c_longlong, [_P]),
    "xfemm_linprob_create": (C.c_int, [C.c_int, _ip, _ip, C.POINTER(_P)]),
    "xfemm_linprob_destroy": (C.c_int, [_P]),
    "xfemm_linprob_n": (C.c_int, [_P]),
    "xfemm_linprob_nnz": (C.c_longlong, [_P]),
    "xfemm_linprob_wipe": (C.c_int, [_P]),
    "xfemm_linprob_snapshot": (C.c_int, [_P]),
    "xfemm_linprob_restore": (C.c_int, [_P]),
    "xfemm_linprob_set_values": (C.c_int, [_P, _dp]),
    "xfemm_linprob_get_values": (C.c_int, [_P, _dp]),
    "xfemm_linprob_set_rhs": (C.c_int, [_P, _dp]),
    "xfemm_linprob_restore_solution": (C.c_int, [_P]),
    "xfemm_host_register": (C.c_int, [_P, C.c_size_t]),
    "xfemm_host_unregister": (C.c_int, [_P]),
    "xfemm_linprob_get_rhs": (C.c_int, [_P, _dp]),
    "xfemm_linprob_set_solution": (C.c_int, [_P, _dp]),
    "xfemm_linprob_get_solution": (C.c_int, [_P, _dp]),
    "xfemm_linprob_dirichlet": (C.c_int, [_P, C.c_int, _ip, _dp, _dp, C.c_int]),
    "xfemm_linprob_periodic": (C.c_int, [_P, C.c_int, _ip, _ip, _ip]),
    "xfemm_linprob_pcg": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(SolveInfo)]),
    "xfemm_linprob_multa": (C.c_int, [_P, _dp, _dp]),
    "xfemm_linprob_multpc": (C.c_int, [_P, C.c_int, _dp, _dp]),
    "xfemm_linprob_bench_spmv": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "xfemm_linprob_amg_levels": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "xfemm_linprob_set_coords": (C.c_int, [_P, _P]),
    "xfemm_linprob_prebuild": (C.c_int, [_P, C.c_int]),
    "xfemm_linprob_history_push": (C.c_int, [_P]),
    "xfemm_linprob_history_guess": (C.c_int, [_P, C.c_int]),
    "xfemm_linprob_history_clear": (C.c_int, [_P]),
    "xfemm_linprob_save_solution": (C.c_int, [_P]),
    "xfemm_linprob_newton_norms": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "xfemm_linprob_relax": (C.c_int, [_P, C.c_double]),
    "xfemm_pattern_build": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, C.c_int, _P, _P,
                                       C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
                                       C.POINTER(C.POINTER(C.c_ubyte))]),
    "xfemm_pattern_build_device": (C.c_int, [C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P, C.c_int, _P, _P,
                                       C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
                                       C.POINTER(C.POINTER(C.c_ubyte))]),
    "xfemm_linprob_set_fill_mask": (C.c_int, [_P, _P]),
    "xfemm_linprob_get_exists_mask": (C.c_int, [_P, _P]),
    "xfemm_free": (None, [_P]),
    "xfemm_renumber_cuthill": (C.c_int, [C.c_int, C.c_int, _ip, _ip, C.POINTER(C.c_int)]),
    "xfemm_renumber_sort_elements": (C.c_int, [C.c_int, _ip, _ip]),
    "xfemm_age_cell_matrix": (C.c_int, [C.c_double, C.c_double, C.c_double, _dp]),
    "xfemm_mag_create": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "xfemm_mag_destroy": (C.c_int, [_P]),
    "xfemm_mag_set_labels": (C.c_int, [_P, C.c_int, _P]),
    "xfemm_mag_set_age": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _P]),
    "xfemm_mag_assemble_static": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "xfemm_mag_assemble_static_parts": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "xfemm_mag_get_mu": (C.c_int, [_P, _dp]),
    "xfemm_mag_solution_to_A": (C.c_int, [_P, _dp]),
    "xfemm_mag_gap_torque": (C.c_int, [_P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), _P]),
    "xfemm_fsolver_gap_torque": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "xfemm_fsolver_amg_levels": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "xfemm_clinprob_amg_levels": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "xfemm_fsolver_stored_energy": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "xfemm_fsolver_flux_linkage": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "xfemm_mag_set_deterministic": (C.c_int, [_P, C.c_int]),
    "xfemm_mag_label_integrals": (C.c_int, [_P, C.c_int, _P, _dp, _dp]),
    "xfemm_clinprob_create": (C.c_int, [C.c_int, C.c_int, _ip, _ip, C.POINTER(_P)]),
    "xfemm_clinprob_destroy": (C.c_int, [_P]),
    "xfemm_clinprob_n": (C.c_int, [_P]),
    "xfemm_clinprob_nnz": (C.c_longlong, [_P]),
    "xfemm_clinprob_wipe": (C.c_int, [_P]),
    "xfemm_clinprob_set_values": (C.c_int, [_P, _P]),
    "xfemm_clinprob_get_values": (C.c_int, [_P, _P]),
    "xfemm_clinprob_set_rhs": (C.c_int, [_P, _P]),
    "xfemm_clinprob_get_rhs": (C.c_int, [_P, _P]),
    "xfemm_clinprob_set_solution": (C.c_int, [_P, _P]),
    "xfemm_clinprob_get_solution": (C.c_int, [_P, _P]),
    "xfemm_clinprob_dirichlet": (C.c_int, [_P, C.c_int, _ip, _P, _P, C.c_int]),
    "xfemm_clinprob_periodic": (C.c_int, [_P, C.c_int, _ip, _ip, _ip]),
    "xfemm_clinprob_set_fill_mask": (C.c_int, [_P, _P]),
    "xfemm_clinprob_get_exists_mask": (C.c_int, [_P, _P]),
    "xfemm_clinprob_cocg": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(SolveInfo)]),
    "xfemm_clinprob_prebuild": (C.c_int, [_P, C.c_int]),
    "xfemm_clinprob_save_solution": (C.c_int, [_P]),
    "xfemm_clinprob_newton_norms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "xfemm_clinprob_relax": (C.c_int, [_P, C.c_double]),
    "xfemm_clinprob_enable_newton": (C.c_int, [_P]),
    "xfemm_clinprob_set_side_values": (C.c_int, [_P, C.c_int, _P]),
    "xfemm_clinprob_get_side_values": (C.c_int, [_P, C.c_int, _P]),
    "xfemm_clinprob_mult_newton": (C.c_int, [_P, _P, _P]),
    "xfemm_clinprob_mult": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "xfemm_clinprob_kludge_solve": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(SolveInfo)]),
    "xfemm_cmag_create": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "xfemm_cmag_destroy": (C.c_int, [_P]),
    "xfemm_cmag_set_labels": (C.c_int, [_P, C.c_int, _P]),
    "xfemm_cmag_assemble": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "xfemm_cmag_set_newton": (C.c_int, [_P, C.c_int]),
    "xfemm_cmag_fix_circuit_diagonals": (C.c_int, [_P]),
    "xfemm_cmag_get_mu": (C.c_int, [_P, _P]),
    "xfemm_cmag_solution_to_A": (C.c_int, [_P, _P]),
    "xfemm_linprob_get_pattern": (C.c_int, [_P, _P, _P]),
    "xfemm_dist_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "xfemm_dist_destroy": (C.c_int, [_P]),
    "xfemm_dist_sizes": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "xfemm_dist_set_precond": (C.c_int, [_P, C.c_int]),
    "xfemm_dist_halo_columns": (C.c_int, [_P, _ip]),
    "xfemm_dist_set_send_rows": (C.c_int, [_P, C.c_int, _P]),
    "xfemm_dist_refresh": (C.c_int, [_P]),
    "xfemm_dist_init": (C.c_int, [_P, _P]),
    "xfemm_dist_pack": (C.c_int, [_P, _P]),
    "xfemm_dist_unpack": (C.c_int, [_P, _P]),
    "xfemm_dist_spmv_dot": (C.c_int, [_P, _P]),
    "xfemm_dist_update": (C.c_int, [_P, _P, _P]),
    "xfemm_dist_direction": (C.c_int, [_P, _P]),
    "xfemm_dist_get_solution": (C.c_int, [_P, _dp]),
    "xfemm_dist_stream": (C.c_ulonglong, [_P]),
    "xfemm_dpcg_create": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P)]),
    "xfemm_dpcg_destroy": (C.c_int, [_P]),
    "xfemm_dpcg_ipc_handle": (C.c_int, [_P, _P]),
    "xfemm_dpcg_connect": (C.c_int, [_P, _P]),
    "xfemm_dpcg_connect_files": (C.c_int, [_P, C.c_char_p, C.c_char_p, C.c_int]),
    "xfemm_dpcg_stats": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "xfemm_dpcg_solve": (C.c_int, [_P, C.c_double, C.c_int, C.c_int, C.POINTER(SolveInfo)]),
    "xfemm_dpcg_share_solution": (C.c_int, [_P]),
    "xfemm_dpcg_finish": (C.c_int, [_P]),
    "xfemm_fsolver_linprob": (_P, [_P]),
    "xfemm_fsolver_run": (C.c_int, [C.c_char_p, C.c_int]),
    "xfemm_fsolver_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "xfemm_fsolver_load_mesh": (C.c_int, [_P, C.c_int]),
    "xfemm_fsolver_renumber": (C.c_int, [_P, C.c_int]),
    "xfemm_fsolver_solve": (C.c_int, [_P]),
    "xfemm_fsolver_resolve": (C.c_int, [_P]),
    "xfemm_fsolver_set_rotor_angle": (C.c_int, [_P, C.c_int, C.c_double, C.c_double]),
    "xfemm_fsolver_write": (C.c_int, [_P]),
    "xfemm_fsolver_close": (C.c_int, [_P]),
    "xfemm_fsolver_counts": (C.c_int, [_P] + [C.POINTER(C.c_int)] * 7),
    "xfemm_fsolver_get_nodes": (C.c_int, [_P, _dp, _ip]),
    "xfemm_fsolver_get_elems": (C.c_int, [_P, _ip]),
    "xfemm_fsolver_get_pbc": (C.c_int, [_P, _ip]),
    "xfemm_fsolver_get_permutation": (C.c_int, [_P, _ip]),
    "xfemm_fsolver_get_pattern": (C.c_int, [_P, C.POINTER(C.c_longlong), _P, _P]),
    "xfemm_fsolver_get_system": (C.c_int, [_P, _P, _P]),
    "xfemm_magdir_eval": (C.c_int, [C.c_char_p, C.c_double, C.c_double, C.POINTER(C.c_double)]),
    "xfemm_fsolver_get_side_matrix": (C.c_int, [_P, C.c_int, _P]),
    "xfemm_fsolver_get_solution": (C.c_int, [_P, _P, _P]),
    "xfemm_fsolver_get_mu": (C.c_int, [_P, _P]),
    "xfemm_fsolver_get_stats": (C.c_int, [_P, C.POINTER(C.c_int), _ip, _dp, _dp, C.POINTER(C.c_double)]),
    "xfemm_fsolver_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
}

_lib = None


def lib():
    """Loads the in-tree shared library (built by `python -m xfemm_b200.build` / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XfemmError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build) first; "
                             "there is no Python/CPU fallback for the solver")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc, what=""):
    if rc != 0:
        msg = lib().xfemm_last_error().decode(errors="replace")
        raise XfemmError(f"{what} failed (code {rc}): {msg}")


def device_count():
    return lib().xfemm_device_count()


def cuda_init(device=0):
    _check(lib().xfemm_cuda_init(device), "xfemm_cuda_init")


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check(rc, what="xfemm"):
    _check(rc, what)


def age_cell_matrix(ci, co, K):
    """10x10 air-gap-element cell matrix (static2d.cpp:200-274)."""
    MG = np.zeros(100, dtype=np.float64)
    _check(lib().xfemm_age_cell_matrix(ci, co, K, MG), "age_cell_matrix")
    return MG.reshape(10, 10)


def renumber_cuthill(n_nodes, edges):
    """Cuthill–McKee exactly as the reference (cuthill.cpp:94-350): returns (newnum, bandwidth)."""
    edges = _i32(edges).reshape(-1, 2)
    newnum = np.empty(n_nodes, dtype=np.int32)
    bw = C.c_int(0)
    _check(lib().xfemm_renumber_cuthill(n_nodes, len(edges), edges.reshape(-1), newnum, C.byref(bw)), "cuthill")
    return newnum, bw.value


def sort_elements(tri):
    """The reference's incomplete comb sort (cuthill.cpp:40-85): returns the element order."""
    tri = _i32(tri).reshape(-1, 3)
    order = np.empty(len(tri), dtype=np.int32)
    _check(lib().xfemm_renumber_sort_elements(len(tri), tri.reshape(-1), order), "sort_elements")
    return order


def pattern_build(n_rows, tri, age_nodes=None, pbc=None, extra=None, with_fill=False, device=False):
    """Full CSR pattern (rowptr, col[, fill]) of mesh graph + AGE blocks + periodic fill.
    device=True: mesh-graph pass on the GPU (xfemm_pattern_build_device, what FSolver uses)."""
    fp = C.POINTER(C.c_ubyte)()
    tri = _i32(tri)
    stride = tri.shape[1]
    age = _i32(age_nodes).reshape(-1, 10) if age_nodes is not None and len(age_nodes) else None
    pi = pj = None
    npbc = 0
    if pbc is not None and len(pbc):
        pbc = _i32(pbc)
        pi, pj = _i32(pbc[:, 0]), _i32(pbc[:, 1])
        npbc = len(pi)
    ei = ej = None
    nex = 0
    if extra is not None and len(extra):
        extra = _i32(extra)
        ei, ej = _i32(extra[:, 0]), _i32(extra[:, 1])
        nex = len(ei)
    rp = C.POINTER(C.c_int)()
    cp = C.POINTER(C.c_int)()
    build = lib().xfemm_pattern_build_device if device else lib().xfemm_pattern_build
    _check(build(n_rows, len(tri), _ptr(tri), stride, 0 if age is None else len(age), _ptr(age),
                                     nex, _ptr(ei), _ptr(ej), npbc, _ptr(pi), _ptr(pj), C.byref(rp), C.byref(cp),
                                     C.byref(fp) if with_fill else None),
           "pattern_build")
    rowptr = np.ctypeslib.as_array(rp, shape=(n_rows + 1,)).copy()
    col = np.ctypeslib.as_array(cp, shape=(int(rowptr[-1]),)).copy()
    lib().xfemm_free(rp)
    lib().xfemm_free(cp)
    if with_fill:
        fill = np.ctypeslib.as_array(fp, shape=(int(rowptr[-1]),)).copy()
        lib().xfemm_free(fp)
        return rowptr, col, fill
    return rowptr, col


class LinProb:
    """Device-side real symmetric system — the backend of CBigLinProb (spars.h:38-83)."""

    def __init__(self, rowptr, col):
        self.rowptr, self.col = _i32(rowptr), _i32(col)
        self.n = len(self.rowptr) - 1
        self.h = C.c_void_p()
        _check(lib().xfemm_linprob_create(self.n, self.rowptr, self.col, C.byref(self.h)), "linprob_create")

    def close(self):
        if self.h:
            lib().xfemm_linprob_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nnz(self):
        return lib().xfemm_linprob_nnz(self.h)

    def wipe(self):
        _check(lib().xfemm_linprob_wipe(self.h), "wipe")

    def set_values(self, v):
        _check(lib().xfemm_linprob_set_values(self.h, _f64(v)), "set_values")

    def get_values(self):
        v = np.empty(self.nnz, dtype=np.float64)
        _check(lib().xfemm_linprob_get_values(self.h, v), "get_values")
        return v

    def set_rhs(self, b):
        _check(lib().xfemm_linprob_set_rhs(self.h, _f64(b)), "set_rhs")

    def get_rhs(self):
        b = np.empty(self.n, dtype=np.float64)
        _check(lib().xfemm_linprob_get_rhs(self.h, b), "get_rhs")
        return b

    def set_solution(self, V):
        _check(lib().xfemm_linprob_set_solution(self.h, _f64(V)), "set_solution")

    def get_solution(self):
        V = np.empty(self.n, dtype=np.float64)
        _check(lib().xfemm_linprob_get_solution(self.h, V), "get_solution")
        return V

    def dirichlet(self, nodes, x_first, x_last=None, bdw=0):
        nodes = _i32(nodes)
        xf = _f64(x_first)
        xl = xf if x_last is None else _f64(x_last)
        _check(lib().xfemm_linprob_dirichlet(self.h, len(nodes), nodes, xf, xl, bdw), "dirichlet")

    def periodic(self, pairs):
        pairs = _i32(pairs).reshape(-1, 3)
        _check(lib().xfemm_linprob_periodic(self.h, len(pairs), _i32(pairs[:, 0]), _i32(pairs[:, 1]), _i32(pairs[:, 2])),
               "periodic")

    def pcg(self, precision, warm=False, precond=PRECOND_MCSSOR, max_iter=0, check=True):
        info = SolveInfo()
        rc = lib().xfemm_linprob_pcg(self.h, precision, int(bool(warm)), precond, max_iter, C.byref(info))
        if check:
            _check(rc, "pcg")
        return info

    def multa(self, x):
        y = np.empty(self.n, dtype=np.float64)
        _check(lib().xfemm_linprob_multa(self.h, _f64(x), y), "multa")
        return y

    def multpc(self, x, precond=PRECOND_MCSSOR):
        y = np.empty(self.n, dtype=np.float64)
        _check(lib().xfemm_linprob_multpc(self.h, precond, _f64(x), y), "multpc")
        return y

    def bench_spmv(self, reps=20):
        ms = C.c_double(0)
        _check(lib().xfemm_linprob_bench_spmv(self.h, reps, C.byref(ms)), "bench_spmv")
        return ms.value


class FSolver:
    """File-level mirror of the reference's FSolver use (main.cpp:39-84): open → load_mesh →
    renumber → solve → write, with array getters for the parity tests."""

    def __init__(self, basename, verbose=False):
        self.h = C.c_void_p()
        self.basename = basename
        _check(lib().xfemm_fsolver_open(basename.encode(), int(verbose), C.byref(self.h)), "LoadProblemFile")

    def close(self):
        if self.h:
            lib().xfemm_fsolver_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        _check(lib().xfemm_fsolver_set_option(self.h, key.encode(), float(value)), "set_option")

    def load_mesh(self, delete_files=False):
        _check(lib().xfemm_fsolver_load_mesh(self.h, int(delete_files)), "LoadMesh")

    def renumber(self, delete_files=False):
        _check(lib().xfemm_fsolver_renumber(self.h, int(delete_files)), "Cuthill")

    def solve(self):
        _check(lib().xfemm_fsolver_solve(self.h), "solve")

    def resolve(self):
        """Re-runs the Newton loop (assembly + BCs + solve) on the resident device problem."""
        _check(lib().xfemm_fsolver_resolve(self.h), "resolve")

    def set_rotor_angle(self, inner_deg, outer_deg=0.0, age_index=0):
        """Re-targets an air-gap element to a new rotor angle without re-meshing."""
        _check(lib().xfemm_fsolver_set_rotor_angle(self.h, age_index, float(inner_deg), float(outer_deg)), "set_rotor_angle")

    def gap_torque(self, age_index=0):
        """Torque [N m] through an air-gap element from the resident solution (no .ans file, no nodal D2H)."""
        t = C.c_double(0)
        _check(lib().xfemm_fsolver_gap_torque(self.h, age_index, C.byref(t)), "gap_torque")
        return t.value

    def stored_energy(self):
        """Stored energy [J] of all blocks from the resident solution (= mo_groupselectblock(); mo_blockintegral(2))."""
        e = C.c_double(0)
        _check(lib().xfemm_fsolver_stored_energy(self.h, C.byref(e)), "stored_energy")
        return e.value

    def flux_linkage(self, circuit=0):
        """Flux linkage [Wb] of circuit `circuit` of the .fem file from the resident solution."""
        w = C.c_double(0)
        _check(lib().xfemm_fsolver_flux_linkage(self.h, circuit, C.byref(w)), "flux_linkage")
        return w.value

    def write(self):
        _check(lib().xfemm_fsolver_write(self.h), "write")

    def counts(self):
        v = [C.c_int(0) for _ in range(7)]
        _check(lib().xfemm_fsolver_counts(self.h, *[C.byref(x) for x in v]), "counts")
        keys = ["n_nodes", "n_elems", "n_pbc", "n_age", "bandwidth", "n_unknowns", "is_complex"]
        return {k: x.value for k, x in zip(keys, v)}

    def nodes(self):
        n = self.counts()["n_nodes"]
        xy = np.empty((n, 2), dtype=np.float64)
        bc = np.empty(n, dtype=np.int32)
        _check(lib().xfemm_fsolver_get_nodes(self.h, xy.reshape(-1), bc), "get_nodes")
        return xy, bc

    def elems(self):
        n = self.counts()["n_elems"]
        e = np.empty((n, 8), dtype=np.int32)
        _check(lib().xfemm_fsolver_get_elems(self.h, e.reshape(-1)), "get_elems")
        return e

    def pbc(self):
        n = self.counts()["n_pbc"]
        p = np.empty((n, 3), dtype=np.int32)
        if n:
            _check(lib().xfemm_fsolver_get_pbc(self.h, p.reshape(-1)), "get_pbc")
        return p

    def permutation(self):
        n = self.counts()["n_nodes"]
        p = np.empty(n, dtype=np.int32)
        _check(lib().xfemm_fsolver_get_permutation(self.h, p), "get_permutation")
        return p

    def system(self):
        """(rowptr, col, vals, rhs) of the last assembled + BC'd system (needs option keep_system=1)."""
        nnz = C.c_longlong(0)
        _check(lib().xfemm_fsolver_get_pattern(self.h, C.byref(nnz), None, None), "get_pattern")
        n = self.counts()["n_unknowns"]
        rowptr = np.empty(n + 1, dtype=np.int32)
        col = np.empty(nnz.value, dtype=np.int32)
        _check(lib().xfemm_fsolver_get_pattern(self.h, C.byref(nnz), _ptr(rowptr), _ptr(col)), "get_pattern")
        cplx = self.counts()["is_complex"]
        vals = np.empty(nnz.value, dtype=np.complex128 if cplx else np.float64)
        rhs = np.empty(n, dtype=np.complex128 if cplx else np.float64)
        _check(lib().xfemm_fsolver_get_system(self.h, _ptr(vals), _ptr(rhs)), "get_system")
        return rowptr, col, vals, rhs

    def side_matrix(self, k):
        """Newton AC: values of Mh (k=1), Ms (k=2) or Ma (k=3) of the last system, on the slots of system()."""
        nnz = C.c_longlong(0)
        _check(lib().xfemm_fsolver_get_pattern(self.h, C.byref(nnz), None, None), "get_pattern")
        vals = np.empty(nnz.value, dtype=np.complex128)
        _check(lib().xfemm_fsolver_get_side_matrix(self.h, k, _ptr(vals)), "get_side_matrix")
        return vals

    def solution(self):
        c = self.counts()
        dt = np.complex128 if c["is_complex"] else np.float64
        V = np.empty(c["n_unknowns"], dtype=dt)
        A = np.empty(c["n_unknowns"], dtype=dt)
        _check(lib().xfemm_fsolver_get_solution(self.h, _ptr(V), _ptr(A)), "get_solution")
        return V, A

    def mu(self):
        c = self.counts()
        m = np.empty((c["n_elems"], 4 if c["is_complex"] else 2), dtype=np.float64)
        _check(lib().xfemm_fsolver_get_mu(self.h, _ptr(m)), "get_mu")
        return m

    def stats(self):
        steps = C.c_int(0)
        iters = np.zeros(64, dtype=np.int32)
        sms = np.zeros(64, dtype=np.float64)
        ams = np.zeros(64, dtype=np.float64)
        rr = C.c_double(0)
        _check(lib().xfemm_fsolver_get_stats(self.h, C.byref(steps), iters, sms, ams, C.byref(rr)), "get_stats")
        k = steps.value
        return {"newton_steps": k, "iters": iters[:k].tolist(), "solve_ms": sms[:k].tolist(),
                "assemble_ms": ams[:k].tolist(), "relres": rr.value}


def fsolver_run(basename, verbose=False):
    """`fsolver <basename>`: 0 ok, 1 the .fem could not be loaded, 2 the solve failed."""
    return lib().xfemm_fsolver_run(basename.encode(), int(verbose))
