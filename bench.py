#!/usr/bin/env python
"""bench.py — fsolver hot path on B200: refined TorqueBenchmark (~2 M DOF), planar magnetostatics.

One "step" = one pass of the hot path over one problem: element assembly (+ air-gap element,
Dirichlet and periodic boundary conditions) and the FP64 PCG solve to the .fem's Precision,
i.e. everything FSolver::Static2D does (cfemm/fsolver/static2d.cpp:53-1033).

  value      seconds per problem with the mesh, pattern and tables already resident in HBM
             (xfemm_fsolver_resolve), device-timed with CUDA events by the library and bracketed
             by barrier + synchronize here; max over ranks; whole job = N ranks x K problems.
  e2e        the same metric through the reference-facing call with HOST buffers
             (xfemm_fsolver_solve: host set-up of tables + pattern, H2D of mesh/pattern, kernels,
             D2H of V and A inside the timed region).
  roofline   the persistent PCG kernel: algorithmic bytes/iteration x iterations
             / CUDA-event time of the kernel, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference
             the UNMODIFIED reference (oracle/_ref/ref_harness, built from /root/reference) on
             the host cores, single thread (the reference has no threading), on a bounded sample:
             full assembly + Dirichlet, `--ref-pbc` periodic pairs and `--ref-iters` SSOR-PCG
             iterations, scaled to the full pair count / the reference's iteration count.

Inputs are synthetic in the sense of the task: the shipped TorqueBenchmark.fem geometry meshed by
the reference's own fmesher (oracle/_ref/fmesher) at mesh size 0.0018 cm → 1 991 745 nodes.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")
GOLD_FEM = os.path.join(ROOT, "tests", "golden", "tb", "tb.fem")

PROFILES = os.path.join(ROOT, "profiles")


def reference_record(mesh_size):
    """Full, unscaled run of the unmodified reference on this workload, recorded by tests/golden/make_golden_scale.py
    (build container) — iteration count, phase times — or None.  The SSOR-PCG iteration count is a property of the
    assembled system, not of the machine: it is what the sampled reference arm scales its iteration sample to."""
    gdir = os.path.join(ROOT, "tests", "golden")
    for name in sorted(os.listdir(gdir)):
        f = os.path.join(gdir, name, "golden.json")
        if os.path.exists(f):
            g = json.load(open(f))
            if abs(g.get("mesh_size", -1) - mesh_size) < 1e-12:
                return g
    return None


def workload_config(workload, mesh_size, n_nodes, n_elems, n_pbc):
    """The `config` object: identical for the GPU arm and the reference arm (everything that is specific to one
    implementation lives under `implementation`)."""
    return {"workload": workload, "mesh_size_cm": mesh_size, "n_nodes": n_nodes, "n_elems": n_elems, "n_pbc": n_pbc,
            "precision": 1e-10,
            "l2": ("inputs larger than L2" if 140.0 * n_nodes > 126e6 else "inputs SMALLER than L2 (not the headline size)")
                  + ": matrix + vectors of one problem are about %.0f MB (140 B per node) against 126 MB" % (140.0 * n_nodes / 1e6)}


def set_mesh_size(text, size):
    lines = text.replace("\r\n", "\n").split("\n")
    out, n = [], 0
    for ln in lines:
        if n > 0 and ln.strip():
            f = ln.split("\t")
            f[3] = repr(size)
            ln = "\t".join(f)
            n -= 1
        elif ln.strip().lower().startswith("[numblocklabels]"):
            n = int(ln.split("=")[1])
        out.append(ln)
    return "\n".join(out)


def prepare_mesh(workdir, mesh_size, rank=0):
    """Writes <workdir>/case.fem and meshes it with the reference's fmesher (cached per mesh size)."""
    os.makedirs(workdir, exist_ok=True)
    base = os.path.join(workdir, "case")
    done = base + ".meshed"
    if os.path.exists(done) and all(os.path.exists(base + e) for e in (".node", ".ele", ".edge", ".pbc")):
        return base
    text = open(GOLD_FEM, newline="").read()
    if mesh_size > 0:
        text = set_mesh_size(text, mesh_size)
    with open(base + ".fem", "w", newline="") as f:
        f.write(text)
    fm = os.path.join(REF, "fmesher")
    if not os.path.exists(fm):
        raise RuntimeError("oracle/_ref/fmesher is missing (run __graft_entry__.build() where /root/reference exists); "
                           "the benchmark meshes its input with the reference's mesher")
    subprocess.run([fm, "case.fem"], cwd=workdir, check=True, stdout=subprocess.DEVNULL)
    open(done, "w").write("ok")
    return base


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 2 ms (the timed region of
    the multigrid solver is tens of milliseconds per step), nvidia-smi as a fallback."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.active = True   # samples are kept only while set (start the thread early: NVML initialisation takes ~0.1 s)
        self.max_mhz = None
        self.source = None

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES if it is a plain index list
        idx = self.gpu
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[self.gpu])
            except Exception:
                pass
        h = N.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        self.source = "nvml"
        while not self.stop_flag:
            if not self.active:
                time.sleep(0.001)
                continue
            self.samples.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
            try:
                mask = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for nm, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "source": self.source}


def run_reference(base, ref_iters, ref_pbc, total_ref_iters):
    """Times the unmodified reference (oracle/_ref/ref_harness: the reference's own objects, one thread — it has no
    threading) on one whole problem.  ref_iters / ref_pbc > 0: a bounded sample — full assembly + Dirichlet, `ref_pbc`
    periodic pairs, `ref_iters` SSOR-PCG iterations — scaled to the full pair count and to the reference's
    iteration count for this system; both <= 0: the full run, nothing scaled."""
    h = os.path.join(REF, "ref_harness")
    if not os.path.exists(h):
        raise RuntimeError("oracle/_ref/ref_harness is missing")
    t0 = time.time()
    r = subprocess.run(["taskset", "-c", "0", h, "time", base, str(ref_iters), str(ref_pbc)], stdout=subprocess.PIPE,
                       text=True, check=True)
    wall = time.time() - t0
    m = json.loads(r.stdout.strip().splitlines()[-1])
    n_pbc = m["n_pbc"]
    pbc_applied = max(1, m["pbc_applied"]) if n_pbc else 0
    t_pbc_full = m["t_pbc"] * (n_pbc / pbc_applied) if n_pbc else 0.0
    it_done = sum(m["iters"]) or 1
    full = ref_iters <= 0 and ref_pbc <= 0
    if full:
        # the harness does not replay the recurrence in a full run: the count comes from the record of the same system
        it_done = total_ref_iters or it_done
    ms_per_iter = 1e3 * sum(m["pcg_secs"]) / it_done
    iters_full = total_ref_iters if total_ref_iters else it_done
    t_pcg_full = sum(m["pcg_secs"]) if full else ms_per_iter * 1e-3 * iters_full
    t_asm = m["t_assembly_first"]
    sample = (f"FULL run, nothing scaled: assembly+Dirichlet {t_asm:.2f}s, {n_pbc} periodic pairs {t_pbc_full:.1f}s, "
              f"SSOR-PCG {t_pcg_full:.1f}s ({iters_full} iterations); harness wall {wall:.0f}s" if full else
              f"full assembly+Dirichlet ({t_asm:.2f}s), {pbc_applied}/{n_pbc} periodic pairs (spread evenly over the list), "
              f"{it_done} of {iters_full} SSOR-PCG iterations ({ms_per_iter:.1f} ms/iter); harness wall {wall:.0f}s")
    return {"s_per_problem": t_asm + t_pbc_full + t_pcg_full, "t_assembly": t_asm, "t_pbc_scaled": t_pbc_full,
            "t_pcg_scaled": t_pcg_full, "ms_per_iter": ms_per_iter, "iters_assumed": iters_full, "full": full,
            "sample": sample, "t_load_mesh": m["t_load_mesh"], "t_cuthill": m["t_cuthill"], "n_nodes": m["n_nodes"],
            "n_elems": m["n_elems"], "n_pbc": n_pbc, "harness_wall_s": wall}


def reference_calibration(info, mesh_size):
    """The sampled arm next to the full runs on record: tests/golden (build container) and, when one was made on a GPU
    box, profiles/r02_reference_full_gpubox_<mesh>.json.  sample_over_full compares like with like (same host) only
    for the latter."""
    cal = {}
    g = reference_record(mesh_size)
    if g:
        f = g["full_reference_run"]
        cal["build_container_full_run"] = {"s_per_problem": f["phases_s"]["t_assembly_first"] + f["phases_s"]["t_pbc"] + f["phases_s"]["t_pcg"],
                                           "iterations": g["reference_iterations"], "ms_per_iteration": f["ms_per_iteration"],
                                           "periodic_pairs_s": f["phases_s"]["t_pbc"], "source": "tests/golden (make_golden_scale.py)"}
    p = os.path.join(PROFILES, f"r02_reference_full_gpubox_ms{mesh_size}.json")
    if os.path.exists(p):
        try:
            f = json.loads(open(p).read().strip().splitlines()[-1])
            cal["gpu_box_full_run"] = {"s_per_problem": f["value"], "sample": f["cpu_baseline"]["sample"], "source": os.path.relpath(p, ROOT)}
            if not info["full"]:
                r = info["s_per_problem"] / f["value"]
                cal["sample_over_gpu_box_full_run"] = r
                cal["within_5pct"] = bool(abs(r - 1.0) <= 0.05)
        except Exception as ex:
            cal["gpu_box_full_run"] = f"unreadable: {ex}"
    return cal


def pcg_bytes_per_iteration(X, S, precond, nnz, n, scatter_ok=True):
    """Algorithmic bytes of ONE preconditioned CG iteration (FP64 values, int32 indices).

    Krylov part: SpMV 12 nnz + 20 n, dot 16 n, fused x/r update 48 n, direction update 24 n.
    Jacobi adds 24 n + 16 n (z = D^-1 r, z.r); multicolour SSOR 12 nnz + 40 n + 16 n;
    the multigrid V-cycle, per level l: pre-smooth 24 n_l, residual 12 nnz_l + 28 n_l, restriction 12 n_l + 24 n_{l+1}
    (by scatter from the residual pass: 24 n_{l+1} only), prolongation 28 n_l, post-smooth 12 nnz_l + 36 n_l (z.r fused),
    dense coarsest solve 8 n_c^2.
    """
    import ctypes as C
    krylov = 12.0 * nnz + 108.0 * n
    if precond == X.PRECOND_JACOBI:
        return krylov + 40.0 * n, None
    if precond == X.PRECOND_MCSSOR:
        return krylov + 12.0 * nnz + 56.0 * n, None
    L = X.lib().xfemm_fsolver_linprob(S.h)
    nr = (C.c_int * 16)()
    nz = (C.c_longlong * 16)()
    nl = X.lib().xfemm_linprob_amg_levels(L, 16, nr, nz)
    if nl <= 0:
        return krylov, None
    levels = [{"rows": int(nr[l]), "nnz": int(nz[l])} for l in range(nl)]
    # restriction by scatter (the default of the one-GPU real solver with a dense coarsest solve; off in deterministic
    # runs): the residual pass adds into the coarse iterate instead of storing the residual, and the restriction pass
    # with its reads of the residual and of the member lists (12 n_l) is gone
    scatter = scatter_ok and nr[nl - 1] <= 128 and not (os.environ.get("XFEMM_DETERMINISTIC") or os.environ.get("XFEMM_AMG_ORDERED_RESTRICT"))
    per_row = 116.0 if scatter else 128.0
    cyc = 0.0
    for l in range(nl - 1):
        cyc += 24.0 * nz[l] + per_row * nr[l] + 24.0 * nr[l + 1]
    cyc += 8.0 * nr[nl - 1] ** 2
    return krylov + cyc, levels


def kernel_source_sha16():
    """Hash of the CUDA sources: an ncu capture is only quoted for the kernels it was taken from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "xfemm_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh")):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_traffic(precond, iterations):
    """DRAM bytes (read + write) of one launch of the persistent kernel: the per-iteration figure of the committed
    ncu --set full capture of this command (profiles/r02_ncu_traffic.json: dram bytes of the captured launch / its
    iteration count) times the iterations of the timed launches.  (None, why) when there is no capture of THESE
    sources: the file records the hash of xfemm_b200/csrc it was taken from."""
    try:
        f = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
        d = f[precond]
        if f.get("source_sha16") != kernel_source_sha16():
            return None, "the committed ncu capture (profiles/r02_ncu_traffic.json) is of an earlier version of the kernels: not quoted"
        return d["dram_bytes_per_launch"] / d["iterations"] * iterations, "profiles/r02_ncu_traffic.json"
    except Exception as ex:
        return None, f"no capture: {ex}"


def measured_hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def run_partitioned(args, base, workload, rank, local_rank, world):
    """Config 5b: all N GPUs solve ONE linear system.  Row-partitioned multigrid-preconditioned CG with the SAME
    hierarchy as on one GPU: one persistent kernel per GPU, halo entries and the two scalars of an iteration as peer
    stores + flags over NVLink (xfemm_dpcg_*, csrc/dist_amg.cu).  `--dist-solver nccl` runs the round-1 path instead
    (Jacobi, torch.distributed collectives per iteration) for comparison.  scaling = strong."""
    import torch
    import torch.distributed as dist
    import xfemm_b200 as X
    import ctypes as C

    if world == 1 and not dist.is_initialized():
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29517", rank=0, world_size=1,
                                device_id=torch.device(f"cuda:{local_rank}"))
    t0 = time.time()
    S = X.FSolver(base)
    S.set_option("assemble_only", 1)
    S.set_option("precond", X.PRECOND_AMG)
    S.set_option("locality_order", 1)  # Z-curve row order: a rank's rows are a compact patch of the mesh, its halo a few thousand rows
    S.load_mesh(False)
    S.renumber(False)
    S.solve()
    t_setup = time.time() - t0
    c = S.counts()
    nnz = X.lib().xfemm_fsolver_nnz(S.h)
    n = c["n_unknowns"]
    if args.dist_solver == "nccl":
        from xfemm_b200.dist_pcg import DistPCG
        D = DistPCG(S, precond=X.PRECOND_JACOBI)
        D.refresh()
        solve = lambda: D.solve(1e-10)
        bytes_iter, levels, label = 12.0 * nnz + 148.0 * n, None, "jacobi, NCCL collectives per iteration (round-1 path)"
    else:
        from xfemm_b200.dist_amg import DistAmgPCG
        t0 = time.time()
        D = DistAmgPCG(S)
        t_plan = time.time() - t0
        solve = lambda: D.solve(1e-10)
        bytes_iter, levels = pcg_bytes_per_iteration(X, S, X.PRECOND_AMG, nnz, n, scatter_ok=False)  # (ordered restriction pass)
        label = "multigrid V-cycle, global hierarchy"
    launches0 = int(X.lib().xfemm_launch_count())
    infos = []
    sampler = ClockSampler(local_rank)
    for i in range(args.warmup + args.steps):
        dist.barrier()
        torch.cuda.synchronize()
        if i == args.warmup:
            launches0 = int(X.lib().xfemm_launch_count())
            if rank == 0:
                sampler.start()
        infos.append(solve())
    sampler.stop_flag = True
    infos = infos[args.warmup:]
    ms = torch.tensor([sum(i["solve_ms"] for i in infos) / len(infos)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    V = D.solution() if args.dist_solver != "nccl" else D.gather_solution()
    plan = D.stats() if args.dist_solver != "nccl" else None
    plans = [None] * world
    dist.all_gather_object(plans, plan)
    if rank == 0:
        its = infos[-1]["iterations"]
        v = float(ms.item()) * 1e-3
        peak, peak_src = measured_hbm_peak()
        achieved = bytes_iter * its / v / 1e9
        rec = reference_record(args.mesh_size)
        line = {
            "metric": "fsolver_static2d_s_per_problem", "value": v, "unit": "s/problem", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(workload + " — ONE mesh row-partitioned over the GPUs (config 5b)", args.mesh_size,
                                      c["n_nodes"], c["n_elems"], c["n_pbc"]),
            "implementation": {"nnz_full": nnz, "preconditioner": label,
                               "parallelism": f"{world} ranks, contiguous ranges of the Z-curve ordered rows of the big multigrid levels; small "
                                              "levels cycled by every rank; halo + scalars as peer stores and flags over NVLink inside one "
                                              "persistent kernel per GPU" if args.dist_solver != "nccl" else
                                              f"{world} ranks, row blocks of the Cuthill-McKee ordered matrix, all_to_all + 2 all_reduce per iteration (NCCL)",
                               "step": "the linear solve of one problem (assembly and multigrid value refresh are replicated on every rank, untimed here: <2 ms)",
                               "plan_per_rank": plans, "multigrid_levels": levels},
            "pcg": {"iterations": its, "reference_iterations": sum(rec["reference_iterations"]) if rec else None,
                    "ms": v * 1e3, "ms_per_iter": v * 1e3 / max(1, its), "er": infos[-1]["er"]},
            "roofline": {"bound": "hbm", "kernel": "k_dpcg (persistent partitioned PCG, all ranks)" if args.dist_solver != "nccl" else "kd_* + NCCL",
                         "achieved": achieved, "peak": peak * world, "peak_source": peak_src + f" x {world} GPUs", "unit": "GB/s",
                         "frac": achieved / (peak * world), "traffic": None, "bytes_per_iter": bytes_iter},
            "e2e": None,
            "host_setup_s": {"load_renumber_assemble_hierarchy": t_setup},
            "gpu_launches": int(X.lib().xfemm_launch_count()) - launches0,
            "clocks": sampler.summary(),
            "solution_norm": float((V * V).sum() ** 0.5)}
        emit(line)
    D.close()
    S.close()
    dist.destroy_process_group()
    return 0


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 3 and 4: the other three formulations at ~2 M unknowns (authored inputs: the reference ships
# no saturating / time-harmonic / axisymmetric fixture; the .fem files are the tracked test fixtures,
# tests/golden/<case>/<case>.fem, refined with the reference's mesher)
# ---------------------------------------------------------------------------------------------------------------
CONFIG_CASES = {
    # name: (fixture, mesh size → ~2 M nodes, what it is)
    "3": ("sat2d", 0.00925, "config 3: saturating 1117 steel (planar magnetostatic, Newton iterations with per-element reluctivity / Jacobian re-assembly)"),
    "4a": ("tbh", 0.0018, "config 4: Harmonic2D, TorqueBenchmark geometry at 50 Hz with a conducting, current-carrying rotor block (complex-symmetric COCG)"),
    "4b": ("axi1", 0.00925, "config 4: StaticAxisymmetric solenoid actuator, B-H steel plunger (Newton iterations)"),
}


def prepare_case_mesh(workdir, fixture, mesh_size):
    os.makedirs(workdir, exist_ok=True)
    base = os.path.join(workdir, "case")
    if os.path.exists(base + ".meshed") and all(os.path.exists(base + e) for e in (".node", ".ele", ".edge", ".pbc")):
        return base
    text = open(os.path.join(ROOT, "tests", "golden", fixture, fixture + ".fem"), newline="").read()
    with open(base + ".fem", "w", newline="") as f:
        f.write(set_mesh_size(text, mesh_size))
    subprocess.run([os.path.join(REF, "fmesher"), "case.fem"], cwd=workdir, check=True, stdout=subprocess.DEVNULL)
    open(base + ".meshed", "w").write("ok")
    return base


def cocg_bytes_per_iteration(levels, nnz, n):
    """Algorithmic bytes of one multigrid-preconditioned COCG iteration (complex FP64 = 16 B, int32 indices; SURVEY §8d:
    complex SpMV 20 nnz + 36 n).  Krylov part: SpMV 20 nnz + 36 n, p.Ap 32 n, x/r update 96 n, direction 48 n, ||r|| and
    z.r 48 n; preconditioner: the real V-cycle of the companion matrix applied to two right-hand sides at once —
    matrix bytes once (12 nnz_l per pass), vector bytes twice."""
    krylov = 20.0 * nnz + 260.0 * n
    cyc = 0.0
    for l in range(len(levels) - 1):
        cyc += 24.0 * levels[l]["nnz"] + 2 * 128.0 * levels[l]["rows"] + 2 * 24.0 * levels[l + 1]["rows"]
    if levels:
        cyc += 2 * 8.0 * levels[-1]["rows"] ** 2
    return krylov + cyc + 64.0 * n  # + residual split / recombination of the two components


def run_config(args, key):
    """One JSON line for configs 3 / 4a / 4b on ONE GPU: value = s/problem for the whole solve (every Newton or
    successive-approximation pass: assembly + boundary conditions + Krylov solve) through the reference-facing call on
    the resident mesh, roofline of the Krylov kernel and of the assembly kernel, cpu_baseline from the unmodified
    reference on a bounded sample."""
    import ctypes as C
    import torch
    import xfemm_b200 as X

    fixture, mesh_size, what = CONFIG_CASES[key]
    if args.mesh_size != 0.0018:
        mesh_size = args.mesh_size
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    X.cuda_init(local_rank)
    workdir = os.path.join(args.workdir, f"cfg{key}_ms{mesh_size}")
    base = prepare_case_mesh(workdir, fixture, mesh_size)
    harmonic = key == "4a"
    precond = {"jacobi": X.PRECOND_JACOBI, "mcssor": X.PRECOND_MCSSOR, "amg": X.PRECOND_AMG}[args.precond]

    h2d, d2h = C.c_longlong(0), C.c_longlong(0)
    X.lib().xfemm_transfer_bytes.restype = None

    def one_e2e():
        S = X.FSolver(base)
        S.set_option("precond", precond)
        S.load_mesh(False)
        S.renumber(False)
        X.lib().xfemm_transfer_bytes(C.byref(h2d), C.byref(d2h), 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        S.solve()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        X.lib().xfemm_transfer_bytes(C.byref(h2d), C.byref(d2h), 0)
        return S, dt

    S, _ = one_e2e()  # warm-up of allocator, kernels, hierarchy
    S.close()
    e2e_t = []
    for _ in range(max(1, min(args.steps, 2))):
        S, dt = one_e2e()
        e2e_t.append(dt)
        if _ + 1 < max(1, min(args.steps, 2)):
            S.close()
    c = S.counts()
    st0 = S.stats()
    # resident steps: the whole non-linear loop again on the device problem, from a zero field (static formulations;
    # the time-harmonic front end has no resident re-solve: its steps are full solves and value = e2e)
    sampler = ClockSampler(local_rank)
    launches0 = int(X.lib().xfemm_launch_count())
    step_t, stats = [], []
    if not harmonic:
        S.set_option("warm_resolve", 0)
        for i in range(args.warmup):
            S.resolve()
        sampler.start()
        torch.cuda.synchronize()
        launches0 = int(X.lib().xfemm_launch_count())
        for i in range(args.steps):
            t0 = time.perf_counter()
            S.resolve()
            torch.cuda.synchronize()
            step_t.append(time.perf_counter() - t0)
            stats.append(S.stats())
    else:
        sampler.start()
        for i in range(args.steps):
            S.close()
            launches0 = int(X.lib().xfemm_launch_count()) if i == 0 else launches0
            S, dt = one_e2e()
            step_t.append(dt)
            stats.append(S.stats())
    sampler.stop_flag = True
    launches = int(X.lib().xfemm_launch_count()) - launches0
    value = sum(step_t) / len(step_t)
    st = stats[-1]
    its, kms, ams = sum(st["iters"]), sum(st["solve_ms"]), sum(st["assemble_ms"])
    nnz = X.lib().xfemm_fsolver_nnz(S.h)
    n = c["n_unknowns"]
    nr = (C.c_int * 16)()
    nz = (C.c_longlong * 16)()
    nl = X.lib().xfemm_fsolver_amg_levels(S.h, 16, nr, nz)
    levels = [{"rows": int(nr[l]), "nnz": int(nz[l])} for l in range(nl)]
    if harmonic:
        bytes_iter = cocg_bytes_per_iteration(levels, nnz, n) if precond == X.PRECOND_AMG else 20.0 * nnz + 260.0 * n + 64.0 * n
        kernel = "k_cocg<%s> (persistent complex-symmetric CG: complex SpMV + fused updates + preconditioner)" % args.precond
        asm_bytes = 312.0 + 9 * 16.0  # complex slots
    else:
        krylov = 12.0 * nnz + 108.0 * n
        cyc = sum(24.0 * levels[l]["nnz"] + 128.0 * levels[l]["rows"] + 24.0 * levels[l + 1]["rows"] for l in range(nl - 1))
        cyc += 8.0 * levels[-1]["rows"] ** 2 if levels else 0.0
        bytes_iter = krylov + (cyc if precond == X.PRECOND_AMG else 40.0 * n)
        kernel = "k_pcg<%s> (persistent PCG, all Newton passes of a solve)" % args.precond
        asm_bytes = 312.0
    peak, peak_src = measured_hbm_peak()
    achieved = bytes_iter * its / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    passes = st["newton_steps"]
    asm_gbs = c["n_elems"] * asm_bytes * passes / (ams * 1e-3) / 1e9 if ams > 0 else None
    metric = "fsolver_harmonic2d_s_per_problem" if harmonic else ("fsolver_staticaxi_s_per_problem" if key == "4b" else "fsolver_static2d_s_per_problem")
    line = {
        "metric": metric, "value": value, "unit": "s/problem", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "2xf64" if harmonic else "f64", "data": "synthetic",
        "config": {"workload": what + f"; tests/golden/{fixture}/{fixture}.fem meshed by the reference's fmesher at mesh size {mesh_size}",
                   "mesh_size": mesh_size, "n_nodes": c["n_nodes"], "n_elems": c["n_elems"], "n_unknowns": n, "n_pbc": c["n_pbc"],
                   "l2": "inputs larger than L2: matrix + vectors %.0f MB against 126 MB" % (((20.0 if harmonic else 12.0) * nnz + (128.0 if harmonic else 56.0) * n) / 1e6)},
        "implementation": {"nnz_full": nnz, "preconditioner": args.precond, "multigrid_levels": levels,
                           "step": ("one full solve from host arrays (the time-harmonic front end has no resident re-solve)" if harmonic else
                                    "the whole Newton loop on the resident mesh from a zero field: per pass assembly with B-H lookup and Jacobian terms, "
                                    "Dirichlet + periodic conditions, multigrid value refresh, PCG")},
        "passes": {"count": passes, "krylov_iterations": st["iters"], "krylov_ms": st["solve_ms"], "assembly_ms": st["assemble_ms"]},
        "pcg": {"iterations": its, "ms": kms, "ms_per_iter": kms / max(1, its), "relres": st["relres"]},
        "assembly": {"ms_per_pass": ams / max(1, passes), "elements_per_s": c["n_elems"] * passes / (ams * 1e-3) if ams > 0 else None,
                     "bytes_per_element": asm_bytes, "achieved_gbs": asm_gbs, "frac_of_peak": (asm_gbs / peak) if asm_gbs else None,
                     "kernel": "kh_assemble" if harmonic else ("k_assemble_staticaxi" if key == "4b" else "k_assemble_static2d")},
        "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "bytes_per_iter": bytes_iter},
        "e2e": {"value": sum(e2e_t) / len(e2e_t), "unit": "s/problem", "h2d_bytes_per_step": int(h2d.value), "d2h_bytes_per_step": int(d2h.value)},
        "gpu_launches": launches, "clocks": sampler.summary(),
    }
    if harmonic and not args.no_cpu_baseline:
        line["cpu_baseline"] = {"value": None, "unit": "s/problem", "cores": 1, "kind": "reference",
                                "sample": "not sampled: the reference's complex solver has no iteration cap to sample through the harness, "
                                          "and a full run at this size is hours (Temp1 at 15 306 nodes and 50 Hz: 2.3 s, SURVEY §6)"}
    elif not args.no_cpu_baseline:
        try:
            # the unmodified reference: first pass only (assembly, boundary conditions, a sample of its SSOR iterations)
            r = subprocess.run(["taskset", "-c", "0", os.path.join(REF, "ref_harness"), "time", base, str(args.ref_iters), str(args.ref_pbc)],
                               stdout=subprocess.PIPE, text=True, check=True, timeout=1500)
            m = json.loads(r.stdout.strip().splitlines()[-1])
            it_done = sum(m["iters"]) or 1
            ms_it = 1e3 * sum(m["pcg_secs"]) / it_done if not harmonic else None
            line["cpu_baseline"] = {
                "value": None, "unit": "s/problem", "cores": 1, "kind": "reference",
                "first_pass": {"assembly_and_conditions_s": m["t_assembly_first"], "ms_per_ssor_pcg_iteration": ms_it,
                               "iterations_sampled": it_done, "krylov_s_sampled": sum(m["pcg_secs"])},
                "sample": "the reference's FIRST pass only: full assembly + boundary conditions, %d SSOR-PCG iterations (%s); a full "
                          "reference run of this configuration (%d passes here, ~20x our iteration counts per pass at config 2) is "
                          "hours and was not made: no s/problem figure is extrapolated" % (it_done, "%.1f ms each" % ms_it if ms_it else "complex solver: whole first solve", passes)}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "s/problem", "cores": 1, "kind": "reference", "sample": f"unavailable: {ex}"}
    emit(line)
    S.close()
    return 0


_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    out = _OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mesh-size", type=float, default=0.0018, help="block-label mesh size in cm (0.0018 → ~2M DOF)")
    ap.add_argument("--precond", default=os.environ.get("XFEMM_PRECOND", "amg"), choices=["jacobi", "mcssor", "amg"])
    ap.add_argument("--ref-iters", type=int, default=60)
    ap.add_argument("--ref-pbc", type=int, default=60)
    ap.add_argument("--ref-full", action="store_true",
                    help="reference arm: ONE full, unscaled run of the reference (all periodic pairs, all iterations; "
                         "minutes at 2 M nodes) instead of the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the runs with the other preconditioners")
    ap.add_argument("--config", default="2", choices=["2", "3", "4a", "4b"],
                    help="BASELINE.json workload: 2 = refined TorqueBenchmark, linear (default, the headline); 3 = saturating steel, "
                         "Newton; 4a = Harmonic2D 50 Hz; 4b = StaticAxisymmetric.  Config 5: default mode with --gpus N (sweep) and --partitioned")
    ap.add_argument("--partitioned", action="store_true",
                    help="config 5b: ONE mesh row-partitioned over the N GPUs (strong scaling) instead of the sweep")
    ap.add_argument("--dist-solver", default="peer", choices=["peer", "nccl"], help="--partitioned: peer-memory multigrid kernel (default) or the round-1 NCCL/Jacobi path")
    ap.add_argument("--workdir", default=os.environ.get("XFEMM_BENCH_DIR", "/tmp/xfemm_bench"))
    args = ap.parse_args()
    # stdout carries one JSON line and nothing else: whatever a library prints there (NCCL's version banner, solver
    # chatter) is sent to stderr from here on
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config != "2" and args.impl == "ours":
        return run_config(args, args.config) if rank == 0 else 0
    workload = f"TorqueBenchmark planar magnetostatic, fmesher mesh size {args.mesh_size} cm, linear, Static2D FP64 PCG"
    workdir = os.path.join(args.workdir, f"ms{args.mesh_size}")
    rec = reference_record(args.mesh_size)
    total_ref_iters = sum(rec["reference_iterations"]) if rec else 0

    if args.impl == "reference":
        if rank != 0:
            return 0
        base = prepare_mesh(workdir, args.mesh_size)
        vals = []
        info = None
        if args.ref_full:
            info = run_reference(base, 0, 0, total_ref_iters)
            vals.append(info["s_per_problem"])
        else:
            for _ in range(max(1, min(args.steps, 2))):
                info = run_reference(base, args.ref_iters, args.ref_pbc, total_ref_iters)
                vals.append(info["s_per_problem"])
        v = sum(vals) / len(vals)
        line = {"impl": "reference", "metric": "fsolver_static2d_s_per_problem", "value": v, "unit": "s/problem",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(workload, args.mesh_size, info["n_nodes"], info["n_elems"], info["n_pbc"]),
                "cpu_baseline": {"value": v, "unit": "s/problem", "cores": 1, "kind": "reference", "sample": info["sample"]},
                "e2e": {"value": v, "unit": "s/problem", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "reference_phases_s": {"assembly": info["t_assembly"], "periodic_pairs": info["t_pbc_scaled"], "pcg": info["t_pcg_scaled"],
                                       "ms_per_iteration": info["ms_per_iter"], "iterations": info["iters_assumed"],
                                       "load_mesh": info["t_load_mesh"], "cuthill": info["t_cuthill"]},
                "calibration": reference_calibration(info, args.mesh_size),
                "timed_region_s": info["harness_wall_s"] if args.ref_full else None}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    import xfemm_b200 as X

    if world > 1:
        # stdout carries the ONE JSON line: NCCL's own messages (its version banner at the first communicator) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    X.cuda_init(local_rank)

    # ---- set-up (untimed): mesh by the reference's fmesher, LoadProblemFile/LoadMesh/Cuthill on the host
    if rank == 0:
        prepare_mesh(workdir, args.mesh_size)
    if world > 1:
        dist.barrier()
    base = os.path.join(workdir, "case")
    precond = {"jacobi": X.PRECOND_JACOBI, "mcssor": X.PRECOND_MCSSOR, "amg": X.PRECOND_AMG}[args.precond]
    if args.partitioned:
        return run_partitioned(args, base, workload, rank, local_rank, world)
    t0 = time.time()
    S = X.FSolver(base)
    S.set_option("precond", precond)
    S.set_option("age_union_pattern", 0 if os.environ.get("XFEMM_BENCH_NO_UNION") else 1)  # pattern independent of the rotor angle → sweeps stay resident
    S.load_mesh(False)
    t_load = time.time() - t0
    t0 = time.time()
    S.renumber(False)
    t_cm = time.time() - t0
    c = S.counts()

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- e2e: host buffers → result on host, through the FSolver-level call, the way the reference's
    # command line drives it (a plain problem handle: no sweep options, the problem is solved once)
    e2e_times = []
    n_e2e = max(1, min(args.steps, 2))
    X.lib().xfemm_transfer_bytes.restype = None
    import ctypes as C
    h2d, d2h = C.c_longlong(0), C.c_longlong(0)
    E = X.FSolver(base)
    E.set_option("precond", precond)
    E.load_mesh(False)
    E.renumber(False)
    for i in range(1 + n_e2e):
        X.lib().xfemm_transfer_bytes(C.byref(h2d), C.byref(d2h), 1)
        sync()
        t0 = time.perf_counter()
        E.solve()
        sync()
        if i > 0:
            e2e_times.append(time.perf_counter() - t0)
        X.lib().xfemm_transfer_bytes(C.byref(h2d), C.byref(d2h), 0)
    E.close()
    e2e = sum(e2e_times) / len(e2e_times)
    S.solve()  # sets the sweep handle up (angle-independent pattern, locality order)
    if world > 1:
        t = torch.tensor([e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())

    # ---- device-resident steps: every step is one rotor position of the 360-position sweep
    # (rank r takes the contiguous block of positions r*360/N, r*360/N + 1, … — independent problems, no
    # communication; consecutive positions per rank keep the extrapolated initial guesses sharp)
    def position(i):
        return float((rank * (360 // world) + i) % 360)

    # headline steps: the whole per-problem path every time (assembly of all elements, Dirichlet, periodic pairs, solve);
    # the library's sweep mode that keeps the position-independent part on the device is timed separately below
    S.set_option("cache_static_part", 0)
    sampler = ClockSampler(local_rank)
    sampler.active = False
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        S.set_rotor_angle(position(i))
        S.resolve()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    sampler.active = True
    t0 = time.perf_counter()
    ev0.record()
    launches0 = int(X.lib().xfemm_launch_count())
    solve_ms, asm_ms, iters = [], [], []
    for i in range(args.steps):
        S.set_rotor_angle(position(args.warmup + i))
        S.resolve()
        st = S.stats()
        solve_ms.append(sum(st["solve_ms"]))
        asm_ms.append(sum(st["assemble_ms"]))
        iters.append(sum(st["iters"]))
    ev1.record()
    launches = int(X.lib().xfemm_launch_count()) - launches0
    sync()
    elapsed = time.perf_counter() - t0
    st = S.stats()
    nnz = X.lib().xfemm_fsolver_nnz(S.h)
    bytes_iter, levels = pcg_bytes_per_iteration(X, S, precond, nnz, c["n_unknowns"])
    # a sweep that only reads the torque: nodal values stay on the device
    torque_only = None
    if rank == 0 and world == 1 and c["n_age"] > 0:
        S.set_option("fetch_solution", 0)
        S.set_option("cache_static_part", 1)  # library default for sweeps: only the air-gap blocks are re-assembled
        S.set_rotor_angle(position(args.warmup + args.steps))  # the sweep goes on; this position takes the device copy
        S.resolve()                                            # of the boundary-conditioned static part (untimed)
        tq = []
        sync()
        t1 = time.perf_counter()
        for i in range(args.steps):
            S.set_rotor_angle(position(args.warmup + args.steps + 1 + i))
            S.resolve()
            tq.append(S.gap_torque(0))
        sync()
        torque_only = {"s_per_position": (time.perf_counter() - t1) / args.steps, "torque_Nm": tq,
                       "note": "nodal values stay on the device (torque reduced there) and the position-independent, "
                               "boundary-conditioned part of the system is restored from a device copy: only the air-gap "
                               "blocks are assembled per position",
                       "positions_deg": [position(args.warmup + args.steps + 1 + i) for i in range(args.steps)]}
        # the other device-side reductions a sweep reads instead of nodal values (outside the timed loop): stored energy
        # of all blocks; flux linkage per circuit where the problem has circuits (TorqueBenchmark has none)
        try:
            torque_only["stored_energy_J_last_position"] = S.stored_energy()
            torque_only["flux_linkage_Wb_last_position"] = S.flux_linkage(0) if c.get("n_circuits", 0) > 0 else None
        except Exception as ex:  # noqa: BLE001 — a missing reduction must not cost the bench line
            torque_only["stored_energy_J_last_position"] = f"unavailable: {ex}"
        S.set_option("fetch_solution", 1)
        S.set_option("cache_static_part", 0)
        # the same problem solved from a zero initial guess (what a stand-alone solve of one position costs)
        S.set_option("warm_resolve", 0)
        S.resolve()
        S.resolve()
        cs = S.stats()
        torque_only["cold_start"] = {"pcg_ms": sum(cs["solve_ms"]), "iterations": sum(cs["iters"]), "relres": cs["relres"]}
        S.set_option("warm_resolve", 1)
    # (an NVML query takes tens of milliseconds while the GPU is busy: the sampler keeps running through the
    # torque-only steps — the same kernels under the same load — so that the 0.2 s of headline steps are not
    # represented by a single sample)
    sampler.stop_flag = True
    # the deterministic (gather-based) assembly next to the default one (FP64 atomics), same resident problem
    det_asm = None
    if rank == 0 and world == 1:
        S.set_option("deterministic_assembly", 1)
        S.resolve()  # builds the contributor lists (once per mesh)
        tms = []
        for i in range(3):
            S.resolve()
            tms.append(sum(S.stats()["assemble_ms"]))
        S.set_option("deterministic_assembly", 0)
        det_asm = {"ms": min(tms), "atomic_ms": min(asm_ms), "elements_per_s": c["n_elems"] / (1e-3 * min(tms)),
                   "note": "assembly + boundary conditions of one step; deterministic = element contributions gathered per CSR slot in a fixed order"}
    # the other GPU-parallel preconditioners on the same resident problem, for the record
    variant = None
    if rank == 0 and world == 1 and not args.no_variants:
        variant = []
        names = {X.PRECOND_JACOBI: "jacobi", X.PRECOND_MCSSOR: "mcssor", X.PRECOND_AMG: "amg"}
        for other in names:
            if other == precond:
                continue
            S.set_option("precond", other)
            S.set_option("warm_resolve", 0)  # the comparison is between full solves from a zero guess
            S.resolve()
            S.resolve()
            vs = S.stats()
            vb, _ = pcg_bytes_per_iteration(X, S, other, X.lib().xfemm_fsolver_nnz(S.h), c["n_unknowns"])
            vms, vit = sum(vs["solve_ms"]), sum(vs["iters"])
            variant.append({"preconditioner": names[other], "iterations": vit, "pcg_ms": vms, "relres": vs["relres"],
                            "ms_per_iter": vms / max(1, vit), "algorithmic_gbs": vb * vit / (vms * 1e-3) / 1e9 if vms > 0 else None})
        S.set_option("precond", precond)
        S.set_option("warm_resolve", 1)
    if world > 1:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    # ---- config 5b next to the sweep: the SAME mesh solved ONCE by all N GPUs together (row-partitioned multigrid PCG
    # over peer memory, csrc/dist_amg.cu).  An extra leg after the timed region; a watchdog keeps it from ever costing
    # the headline line.  At this mesh size the solve is latency-bound and does not speed up with N — the strong-
    # scaling records at 20 M nodes are under profiles/ (bench.py --partitioned --mesh-size 0.00057).
    single_mesh = None
    leg_done = threading.Event()
    held = {"line": None}

    def watchdog():
        if not leg_done.wait(timeout=float(os.environ.get("XFEMM_BENCH_LEG_TIMEOUT", "240"))):
            if rank == 0 and held["line"] is not None:
                held["line"]["single_mesh_all_gpus"] = {"error": "the extra leg did not finish in time; skipped"}
                emit(held["line"])
            os._exit(0)

    def partitioned_leg():
        from xfemm_b200.dist_amg import DistAmgPCG
        P = X.FSolver(base)
        P.set_option("assemble_only", 1)
        P.set_option("precond", X.PRECOND_AMG)
        P.set_option("locality_order", 1)
        P.load_mesh(False)
        P.renumber(False)
        P.solve()
        D = DistAmgPCG(P)
        infos = [D.solve(1e-10) for _ in range(3)]
        msl = torch.tensor([min(i["solve_ms"] for i in infos)], dtype=torch.float64, device="cuda")
        dist.all_reduce(msl, op=dist.ReduceOp.MAX)
        Vn = float(np.linalg.norm(D.solution()))
        out = {"what": "ONE problem of this mesh solved by all ranks together (strong scaling; extra leg, not in `value`)",
               "n_gpus": world, "iterations": infos[-1]["iterations"], "solve_ms": float(msl.item()),
               "ms_per_iter": float(msl.item()) / max(1, infos[-1]["iterations"]), "solution_norm": Vn, "plan_rank0": D.stats(),
               "data_path": "peer stores + arrival counters over NVLink inside one persistent kernel per GPU; NCCL only for the host barriers",
               "records_at_20M_nodes": "profiles/r02_bench_partitioned_20M_{1,4,8}gpu.json"}
        D.close()
        P.close()
        return out

    problems = args.steps * world
    if world > 1 and not os.environ.get("XFEMM_BENCH_NO_SINGLE_MESH"):
        import numpy as np
        S.close()  # frees the sweep handle's device memory first
        if rank != 0:
            threading.Thread(target=watchdog, daemon=True).start()
            try:
                partitioned_leg()
            except Exception:
                pass
            leg_done.set()
            dist.destroy_process_group()
            return 0
    elif rank != 0:
        dist.destroy_process_group()
        return 0

    value = elapsed / problems
    n = c["n_unknowns"]
    pcg_ms = sum(solve_ms) / len(solve_ms)
    its = sum(iters) / len(iters)
    achieved = bytes_iter * its / (pcg_ms * 1e-3) / 1e9 if pcg_ms > 0 else 0.0
    peak, peak_src = measured_hbm_peak()
    traffic, traffic_src = ncu_traffic(args.precond, its) if args.mesh_size == 0.0018 else (None, "captured at mesh size 0.0018 only")
    line = {
        "metric": "fsolver_static2d_s_per_problem", "value": value, "unit": "s/problem", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": value * 1e3 * world, "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(workload, args.mesh_size, c["n_nodes"], c["n_elems"], c["n_pbc"]),
        "implementation": {"nnz_full": nnz, "preconditioner": args.precond,
                   "l2_note": "matrix+vectors (%.0f MB) exceed the 126 MB L2" % ((12.0 * nnz + 56.0 * n) / 1e6),
                   "parallelism": f"rotor-angle sweep: {world} independent positions at a time, one per GPU (each rank a contiguous block of the 360 positions), no collective",
                   "step": "one rotor position of the sweep: AGE re-target + assembly of all elements + Dirichlet + periodic pairs "
                           "+ PCG on the resident mesh (the library's cache of the position-independent part is switched "
                           "off here and reported under sweep_torque_only); "
                           "initial guess = quadratic extrapolation of the three previous positions (same stopping rule; "
                           "sweep_torque_only.cold_start gives the solve from a zero guess)"},
        "pcg": {"iterations": its, "reference_iterations": total_ref_iters or None, "ms": pcg_ms,
                "ms_per_iter": pcg_ms / its if its else None, "relres": st["relres"], "multigrid_levels": levels},
        "assembly": {"ms": sum(asm_ms) / len(asm_ms), "elements_per_s": c["n_elems"] / (1e-3 * sum(asm_ms) / len(asm_ms))},
        "roofline": {"bound": "hbm", "kernel": "k_pcg<%s> (persistent PCG: SpMV + fused vector updates + preconditioner)" % args.precond,
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "bytes_per_iter": bytes_iter,
                     "achieved_note": "algorithmic bytes per launch (bytes_per_iter x iterations) / CUDA-event launch duration; "
                                      "traffic = DRAM bytes of one launch from the committed ncu capture (profiles/)"},
        "e2e": {"value": e2e, "unit": "s/problem", "h2d_bytes_per_step": int(h2d.value), "d2h_bytes_per_step": int(d2h.value)},
        "variant": variant,
        "assembly_deterministic": det_asm,
        "sweep_torque_only": torque_only,
        "host_setup_s": {"load_mesh": t_load, "cuthill": t_cm},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline and world == 1:  # the CPU baseline is timed at N = 1 only
        try:
            info = run_reference(base, args.ref_iters, args.ref_pbc, total_ref_iters)
            line["cpu_baseline"] = {"value": info["s_per_problem"], "unit": "s/problem", "cores": 1, "kind": "reference",
                                    "sample": info["sample"], "calibration": reference_calibration(info, args.mesh_size)}
        except Exception as ex:  # the line must still print
            line["cpu_baseline"] = {"value": None, "unit": "s/problem", "cores": 1, "kind": "reference",
                                    "sample": f"unavailable: {ex}"}
    if world > 1 and not os.environ.get("XFEMM_BENCH_NO_SINGLE_MESH"):
        held["line"] = line
        threading.Thread(target=watchdog, daemon=True).start()
        try:
            line["single_mesh_all_gpus"] = partitioned_leg()
        except Exception as ex:
            line["single_mesh_all_gpus"] = {"error": str(ex)[:300]}
        leg_done.set()
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
