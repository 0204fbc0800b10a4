"""Build driver: compiles the CUDA kernels (sm_100a), the host C++ mirror of the reference
interface and links them into ONE in-tree shared library, xfemm_b200/libxfemm_b200.so, plus the
`fsolver` command line.  nvcc cross-compiles without a GPU.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libxfemm_b200.so")
CLI = os.path.join(HERE, "fsolver_b200")

CU = ["csrc/linprob.cu", "csrc/mag_assemble.cu", "csrc/clinprob.cu", "csrc/mag_harmonic.cu", "csrc/dist.cu", "csrc/dist_amg.cu", "csrc/amg.cu", "csrc/cnewton.cu", "csrc/pattern_device.cu", "csrc/layout_device.cu", "csrc/amg_setup.cu"]
CPP = ["csrc/layout.cpp", "host/pattern.cpp", "host/fsolver_io.cpp", "host/fsolver_solve.cpp", "host/fsolver_harmonic.cpp",
       "host/bh_curve.cpp", "host/age_matrix.cpp", "host/spars.cpp", "host/cspars.cpp",
       "host/capi_fsolver.cpp", "host/magdir_expr.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("XFEMM_NVCC_EXTRA", "").split()
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wno-unused-result"]


def _newer(src, deps, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in [src] + deps)


def _headers():
    hs = []
    for d in ("csrc", "host"):
        for f in os.listdir(os.path.join(HERE, d)):
            if f.endswith((".h", ".cuh")):
                hs.append(os.path.join(HERE, d, f))
    hs.append(os.path.join(ROOT, "include", "xfemm_b200.h"))
    return hs


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    hdrs = _headers()
    jobs, objs = [], []
    for s in CU:
        src = os.path.join(HERE, s)
        obj = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(obj)
        if force or _newer(src, hdrs, obj):
            jobs.append(([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj], obj + ".log"))
    for s in CPP:
        src = os.path.join(HERE, s)
        obj = os.path.join(BUILD, os.path.basename(s) + ".o")
        objs.append(obj)
        if force or _newer(src, hdrs, obj):
            jobs.append((["g++"] + CXX_FLAGS + ["-c", src, "-o", obj], obj + ".log"))
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(lambda j: _run(*j), jobs))
    if verbose:
        for o in outs:
            print(o)
    if jobs or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
             "-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt",
             # -Bsymbolic: references inside the library bind to its own definitions.  A host program may define
             # classes of the same names (the reference's FSolver / CBigLinProb live next to this library in
             # oracle/_ref/femmcli_b200); without it the dynamic linker would route our internal calls to them.
             "-Xlinker", "-Bsymbolic"], os.path.join(BUILD, "link.log"))
    main_src = os.path.join(HERE, "host", "main.cpp")
    if jobs or _newer(main_src, hdrs, CLI):
        _run(["g++"] + CXX_FLAGS + [main_src, "-o", CLI, "-L" + HERE, "-lxfemm_b200",
              "-Wl,-rpath,$ORIGIN"], os.path.join(BUILD, "cli.log"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
