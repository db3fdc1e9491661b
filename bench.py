#!/usr/bin/env python
"""Benchmark of the partial_schur hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (``config.workload``): BASELINE config 2 -- 2-D 5-point Laplacian, N = 4096
(n = 16 777 216 rows, 83 869 696 entries, float64), K = 10, max_dim = 40 (p = 15),
sort = largest real part, tol = 1e-8, v0 = ``np.random.seed(0); randn(n)`` normalised.

A *step* is one Krylov-Schur restart cycle of that solve at full size: host Schur +
reorder of H (40 x 40), the in-place truncation V[:, :15] = V Q, and the Arnoldi
expansion from column 15 to 40 (25 SpMV + 25 CGS2/DGKS orthogonalisations).  The full
solve needs on the order of 10^4 such cycles at this size, far
beyond a benchmark run and far beyond what the CPU reference can finish, so the
headline is the cycle throughput in Arnoldi matvecs per second:

  value  device-timed (CUDA events on the solver's stream), A and V resident in HBM
  e2e    the public ``partial_schur(A_host, ...)`` call with a bounded number of restarts:
         CSR + v0 uploaded from pinned host memory and Q, T read back inside the timing
  roofline       the kernel class with the largest share of the step, algorithmic bytes
                  / CUDA-event time, against MEASURED_PEAKS.json
  cpu_baseline   the UNMODIFIED reference (baseline/_ref, installed by
                 tools/install_reference.py) on the host cores, same operator, same size:
                 Arnoldi steps 15..39 of its first expansion (the column counts of one
                 restart cycle), timed at the operator
  converged      BASELINE.json's own metric, time-to-k-converged, on an instance both arms
                 can finish: mark(400) (n = 80 200, nonsymmetric), K = 20, max_dim = 60
  parity         seeded solves against the reference's records (tests/golden) run before
                 the timing, at whatever rank count the benchmark runs on

``--impl reference`` times the reference alone on the same operator at the same size: its own
``partial_schur`` is run on a time-stamping operator; the timed region is whole restart cycles
(boundaries = the first operator application of each expansion).  A cycle costs ~25 s on the
box's host cores, so the timed region is ``--ref-cycles`` cycles (default 2) whatever K is, and
``ms_per_step`` is that region divided by K: a bounded sample of the workload, as measured.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
_REFERENCE_ARM = "reference" in sys.argv[1:] and "--impl" in sys.argv[1:]
if _REFERENCE_ARM or os.environ.get("AB200_BENCH_CPU_THREADS"):
    # torchrun exports OMP_NUM_THREADS=1 to its children: the CPU arm would silently run its
    # BLAS on one core.  The thread count of the CPU arm is set explicitly, before NumPy loads.
    _n = os.environ.get("AB200_BENCH_CPU_THREADS") or str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

for _p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

NEV, MAX_DIM, TOL = 10, 40, 1e-8
P = min(NEV + 5, MAX_DIM - 1)
GRID_FULL = 4096          # config 2
METRIC = "partial_schur Arnoldi matvecs/s (restart cycles, config 2)"
UNIT = "matvec/s"
CONV = dict(m=400, nev=20, max_dim=60)      # the converged leg: mark(400)
FAST_SCHUR = True     # GPU arm: dgees instead of zgees while H is real (--exact-schur turns it off)


# ----------------------------------------------------------------------------- helpers
def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region, time-stamped; `stop(t0, t1)`
    keeps the samples taken inside [t0, t1] (perf_counter).

    Source: NVML (nvidia_ml_py -- the library behind nvidia-smi; same fields as the recipe's
    `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` line), polled every
    5 ms from a thread.  A looping `nvidia-smi -lms` child is the fallback; it needs a few hundred
    milliseconds before its first line while the timed region lasts 75 ms at 8 GPUs, so the
    sampler is started before the warm-up cycles either way."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.samples = []      # (t, sm_mhz, max_mhz, {reasons})
        self.source = None
        self._stop = False

    # -- NVML
    def _nvml_loop(self, nv, handle):
        bits = [(nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"),
                (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap")]
        mx = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(handle))
                self.samples.append((time.perf_counter(), sm, mx, {n for b, n in bits if mask & b}))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:      # CUDA ordinal -> physical index
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if idx < len(ids) and ids[idx].isdigit():
                    idx = int(ids[idx])
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 50"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            rs = {n for n, v in zip(self.NAMES, f[2:6]) if v.lower().startswith("active")}
            self.samples.append((time.perf_counter(), sm, mx, rs))

    def stop(self, t0=None, t1=None):
        self._stop = True
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML, no nvidia-smi"]}
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        else:
            self.thread.join(timeout=1.0)
        inside = [x for x in self.samples if t0 is None or (t0 <= x[0] <= t1)]
        window = "timed region"
        if not inside and t0 is not None:
            # a region shorter than the polling period: the samples under the same load right
            # around it (the warm-up cycles before, the un-instrumented cycles after)
            inside = [x for x in self.samples if t0 - 0.25 <= x[0] <= t1 + 0.25]
            window = "timed region +- 0.25 s (same load: warm-up cycles before, re-run after)"
        reasons = set()
        for x in inside:
            reasons |= x[3]
        return {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None,
                "sm_max_mhz": float(max(x[2] for x in inside)) if inside else None,
                "samples": len(inside), "window": window, "source": self.source,
                "reasons": sorted(reasons)}


def set_host_threads():
    """All host cores for the CPU arm, whatever the launcher exported; returns the count
    OpenBLAS actually uses."""
    want = int(os.environ.get("AB200_BENCH_CPU_THREADS") or os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        return int(max((d.get("num_threads", 1) for d in threadpool_info()), default=1))
    except Exception:
        return want


def pinned_csr(dev_lib, A):
    """Copy the CSR arrays into pinned host memory (done once, outside any timing) so
    the e2e upload is a pinned -> device copy as the contract asks."""
    import ctypes as C

    import scipy.sparse as sp
    keep = []

    def pin(a):
        ptr = C.c_void_p()
        rc = dev_lib.ab200_host_alloc(C.byref(ptr), int(a.nbytes))
        if rc != 0:
            return a
        buf = (C.c_char * a.nbytes).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=a.dtype, count=a.size)
        out[:] = a
        keep.append(ptr)
        return out

    M = sp.csr_matrix((pin(A.data), pin(A.indices), pin(A.indptr)), shape=A.shape, copy=False)
    return M, keep


# ----------------------------------------------------------------------------- CPU arm
def load_cpu_arm():
    """The CPU implementation to time: the unmodified reference installed under baseline/_ref
    (kind "reference"), else the oracle's restatement of it (kind "port")."""
    import importlib.util
    init = os.path.join(ROOT, "baseline", "_ref", "arnoldi", "__init__.py")
    if os.path.exists(init):
        spec = importlib.util.spec_from_file_location(
            "arnoldi_reference", init, submodule_search_locations=[os.path.dirname(init)])
        mod = importlib.util.module_from_spec(spec)
        sys.modules["arnoldi_reference"] = mod
        spec.loader.exec_module(mod)
        import arnoldi_reference.utils as ru
        return dict(kind="reference", partial_schur=mod.partial_schur,
                    sort=ru.arg_largest_real,
                    what="unmodified cournape/arnoldi-py partial_schur (baseline/_ref)")
    import oracle
    return dict(kind="port", partial_schur=oracle.partial_schur, sort=oracle.arg_largest_real,
                what="oracle/krylov.py (restatement of the reference: baseline/_ref is absent)")


class _Enough(Exception):
    pass


class StampedOperator:
    """Duck-typed operator (shape, dtype, @ -- all the reference uses, decomposition.py:44,58)
    that records the wall clock at every application and stops the solve after `limit`."""

    def __init__(self, A, limit):
        self.A = A
        self.shape = A.shape
        self.dtype = A.dtype
        self.limit = limit
        self.stamps = []

    def __matmul__(self, x):
        self.stamps.append(time.perf_counter())
        if len(self.stamps) > self.limit:
            raise _Enough()
        return self.A @ x


def cpu_partial_cycles(arm, A, cycles):
    """Run the CPU arm's partial_schur on config 2 and time `cycles` restart cycles
    (cycles = 0: only steps 15..39 of the first expansion).  Returns (matvecs, seconds, note)."""
    first = MAX_DIM
    per = MAX_DIM - P
    limit = first + cycles * per if cycles > 0 else first
    op = StampedOperator(A, limit)
    np.random.seed(0)
    t0 = time.perf_counter()
    try:
        arm["partial_schur"](op, NEV, max_dim=MAX_DIM, stopping_criterion=TOL,
                             sort_function=arm["sort"], max_restarts=cycles + 2)
        raise RuntimeError("the CPU solve ended before the sample was complete")
    except _Enough:
        pass
    total = time.perf_counter() - t0
    st = op.stamps
    if cycles > 0:
        # cycle i = [first application of expansion i, first application of expansion i + 1):
        # 25 operator applications + orthogonalisations, then Schur / reorder / restart GEMM
        seconds = st[first + cycles * per] - st[first]
        matvecs = cycles * per
        note = (f"{cycles} whole restart cycles ({matvecs} matvecs) of {arm['what']} on config 2 at "
                f"full size, after its first expansion ({first} matvecs) as warm-up; cycle "
                "boundaries are the first operator application of each expansion")
    else:
        # steps j = P .. MAX_DIM-1 of the first expansion have the column counts of a cycle; the
        # sample runs from the operator application of step P to the first application after
        # the first restart, so it also holds one Schur + reorder + restart GEMM: one cycle's work
        seconds = st[first] - st[P]
        matvecs = per
        note = (f"Arnoldi steps {P}..{first - 1} of the first expansion plus the first restart "
                f"(Schur, reorder, V[:, :p] = V Q) = the work of one restart cycle, {matvecs} "
                f"matvecs, of {arm['what']} on config 2 at full size")
    return matvecs, seconds, note, total


def cpu_converged(arm):
    """Time-to-k-converged of the CPU arm on the converged-leg instance (scripts/utils.py:161-174:
    perf_counter around the call)."""
    from arnoldi_b200.matrices import mark
    A = mark(CONV["m"])
    np.random.seed(0)
    t0 = time.perf_counter()
    Q, T, hist = arm["partial_schur"](A, CONV["nev"], max_dim=CONV["max_dim"],
                                      stopping_criterion=TOL, sort_function=arm["sort"],
                                      max_restarts=5000)
    dt = time.perf_counter() - t0
    return dict(time_to_k_converged_s=dt, restarts=int(hist.restarts[0]),
                diagT=np.array(np.diag(T)))


def conv_workload():
    m = CONV["m"]
    return (f"mark({m}) n={m * (m + 1) // 2} (BASELINE config 3 family), K={CONV['nev']}, "
            f"max_dim={CONV['max_dim']}, LR, tol={TOL}, seed 0, to convergence")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from arnoldi_b200.matrices import lap2d
    threads = set_host_threads()
    arm = load_cpu_arm()
    steps = max(1, args.steps)
    cycles = max(1, min(steps, args.ref_cycles))
    A = lap2d(args.grid)
    matvecs, seconds, note, total = cpu_partial_cycles(arm, A, cycles)
    value = matvecs / seconds
    conv = None
    if not args.no_converged:
        c = cpu_converged(arm)
        conv = {"workload": conv_workload(), "time_to_k_converged_s": c["time_to_k_converged_s"],
                "restarts": c["restarts"], "ritz_real_top3": [float(x) for x in c["diagT"].real[:3]]}
    base = {"value": value, "unit": UNIT, "cores": threads, "kind": arm["kind"],
            "sample": note + f"; {seconds:.1f} s timed of {total:.1f} s run; csr_matvec is "
                             "single-threaded, OpenBLAS uses all threads",
            "os_cpu_count": os.cpu_count()}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * seconds / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic",
        "config": workload_config(args.grid),
        "timed_cycles": cycles, "timed_matvecs": matvecs, "timed_region_s": seconds,
        "step_note": (f"timed region = {cycles} restart cycles; ms_per_step = timed region / "
                      f"{steps} (a bounded sample: one reference cycle costs "
                      f"{seconds / cycles:.1f} s)"),
        "cpu_baseline": base,
        "converged": conv,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def storage_note(st):
    """complex128 arithmetic as in the reference; while A, v0 and every Schur basis applied are
    real the imaginary parts are exactly zero and the device holds the basis as float64."""
    return ("c128 (basis stored f64: provably real on this operator, results identical)"
            if st.get("real_storage") else "c128")


def workload_config(grid=GRID_FULL):
    if grid != GRID_FULL:
        return {"workload": f"lap2d({grid}) REDUCED (not config 2)"}
    return {"workload": f"lap2d({GRID_FULL}) n={GRID_FULL**2} nnz={5*GRID_FULL**2-4*GRID_FULL} "
                        f"float64 CSR, K={NEV}, max_dim={MAX_DIM}, p={P}, LR, tol={TOL}, seed 0",
            "step": f"one Krylov-Schur restart cycle: Schur+reorder (host, {MAX_DIM}x{MAX_DIM}), "
                    f"truncation V[:, :{P}] = V Q, expansion {P}->{MAX_DIM} "
                    f"({MAX_DIM-P} SpMV + CGS2/DGKS)",
            "l2": "inputs larger than L2 (V = 11.0 GB, A = 1.07 GB)",
            "host_schur_b200_arm": ("fast_real_schur=True: while H is real, dsyevd + a column "
                                    "permutation when H is symmetric to 1e-12 (this operator), "
                                    "else dgees + 2x2 block rotations; the reference arm runs its "
                                    "own zgees + ztrexc"),
            "ritz_parity_note": "this operator has double eigenvalues: its Ritz-value parity is "
                                "pinned at N <= 64 and by eigenvalue membership + residual beyond; "
                                "the `parity` block is on operators with simple spectra"}


# ----------------------------------------------------------------------------- parity
def parity_block(comm=None, device=0):
    """Seeded solves against the reference's records (tests/golden), through the public API at
    the rank count of this run.  Returns a JSON-able dict (`ok` per case, `all_ok`)."""
    import scipy.sparse as sp

    from arnoldi_b200 import partial_schur
    from arnoldi_b200.matrices import mark
    from arnoldi_b200.utils import arg_largest_real
    g1 = np.load(os.path.join(ROOT, "tests", "golden", "solves.npz"))
    g2 = np.load(os.path.join(ROOT, "tests", "golden", "solves_r2.npz"))

    def rect(N):
        return sp.csr_matrix((g2[f"rect{N}_data"], g2[f"rect{N}_indices"], g2[f"rect{N}_indptr"]),
                             shape=tuple(int(x) for x in g2[f"rect{N}_shape"]))

    cases = [("mark50_s0", g1, mark(50), 5, 20), ("mark100_s0", g1, mark(100), 20, 60),
             ("rect64_s0", g2, rect(64), 10, 40)]
    out = {}
    for tag, g, A, nev, md in cases:
        np.random.seed(0)
        Q, T, hist = partial_schur(A, nev, max_dim=md, stopping_criterion=TOL, max_restarts=2000,
                                   sort_function=arg_largest_real, device=device, comm=comm,
                                   fast_real_schur=FAST_SCHUR)
        if comm is not None and comm.world > 1:
            pieces = comm.all_gather_bytes(np.ascontiguousarray(Q).tobytes())
            Q = np.concatenate([np.frombuffer(b, np.complex128).reshape(-1, nev) for b in pieces])
        lam, ref = np.diag(T), g[f"{tag}_diagT"]
        rel = np.abs(lam - ref) / np.abs(ref)
        res = np.linalg.norm(A @ Q - Q @ T, axis=0)
        R, Rref = int(hist.restarts[0]), int(g[f"{tag}_hist_restarts"][0])
        ok = bool(R == Rref and np.sum(rel > 1e-10) <= max(1, nev // 10) and rel.max() < 1e-8
                  and res.max() < 1e-7)
        out[tag] = {"R": R, "R_ref": Rref, "max_rel": float(rel.max()),
                    "n_above_1e-10": int(np.sum(rel > 1e-10)), "max_residual": float(res.max()),
                    "ok": ok}
    out["all_ok"] = all(v["ok"] for v in out.values())
    out["rule"] = ("identical restart counts; Ritz values 1e-10 relative except the last-converged "
                   "pair(s) (<= max(1, K//10) of them, < 1e-8: determined only to their residual, "
                   "SURVEY.md section 8c); Schur residual < 1e-7")
    return out


def converged_leg(cpu, comm=None, device=0):
    """BASELINE.json's metric on an instance both arms finish: our partial_schur from host CSR,
    wall clock around the call, in the default (reference-faithful) and the real-arithmetic mode."""
    from arnoldi_b200 import partial_schur
    from arnoldi_b200.matrices import mark
    from arnoldi_b200.utils import arg_largest_real
    A = mark(CONV["m"])
    out = {"workload": conv_workload()}
    for mode in ("lossless", "pairs"):
        best = None
        for rep in range(2):            # rep 0 warms the CUDA context / allocator
            np.random.seed(0)
            stats = {}
            if comm is not None:
                comm.barrier()
            t0 = time.perf_counter()
            Q, T, hist = partial_schur(A, CONV["nev"], max_dim=CONV["max_dim"],
                                       stopping_criterion=TOL, sort_function=arg_largest_real,
                                       max_restarts=5000, stats=stats, device=device, comm=comm,
                                       real_arith=mode, fast_real_schur=FAST_SCHUR)
            dt = time.perf_counter() - t0
            if comm is not None:
                dt = comm.max_float(dt)
            best = dt if best is None or rep > 0 else best
        ent = {"time_to_k_converged_s": best, "restarts": int(hist.restarts[0]),
               "matvecs": int(stats["true_matvecs"]), "real_storage_at_end": int(stats["real_storage"])}
        if cpu is not None:
            ref = cpu["diagT"]
            lam = np.diag(T)
            a, b = np.sort_complex(lam), np.sort_complex(ref)
            rel = np.abs(a - b) / np.abs(b)
            ent["max_rel_ritz_diff_vs_cpu"] = float(rel.max())
            ent["ritz_above_1e-10"] = int(np.sum(rel > 1e-10))
        out["b200_" + mode] = ent
    if cpu is not None:
        out["cpu"] = {"time_to_k_converged_s": cpu["time_to_k_converged_s"],
                      "restarts": cpu["restarts"]}
        out["speedup_lossless"] = cpu["time_to_k_converged_s"] / out["b200_lossless"]["time_to_k_converged_s"]
        out["speedup_pairs"] = cpu["time_to_k_converged_s"] / out["b200_pairs"]["time_to_k_converged_s"]
    out["note"] = ("lossless = the reference's iteration (restart counts identical up to summation "
                   "order); pairs = real arithmetic with conjugate pairs kept whole, a different "
                   "truncation, so its restart count is stated, not equal")
    return out


# ----------------------------------------------------------------------------- GPU arm
def warm_for_sampler(cycle, comm, seconds=1.0, cap=400):
    """Extra warm-up cycles, the same number on every rank, so that the nvidia-smi sampler
    (started before the warm-up) is delivering lines under load when the timed region starts."""
    t0 = time.perf_counter()
    cycle()
    dt = time.perf_counter() - t0
    if comm is not None:
        dt = comm.max_float(dt)
    n = int(min(cap, max(0, np.ceil(seconds / max(dt, 1e-4)))))
    for _ in range(n):
        cycle()
    return n + 1


def event_note(steps):
    idx = list(range(0, steps, EVENT_EVERY))
    return (f"CUDA events around every launch of cycles {idx} of the {steps}-cycle timed region "
            f"(every {EVENT_EVERY}th cycle: an event record between two kernels costs the device "
            "~3 us, 0.36 ms per cycle -- 4% of a cycle at 8 GPUs; tools/event_cost.py); "
            "per-class launches / bytes are those of the instrumented cycles")


EVENT_EVERY = 4      # cycles 0, 4, 8, ... of the timed region carry the per-kernel CUDA events


def kernel_table(st, ms, peak, frac=1.0):
    """Per-class figures from the launches that carried CUDA events: `frac` of the timed
    region's cycles (every cycle launches the same kernels, so launches and bytes scale by it);
    `ms` = duration of those instrumented cycles."""
    classes = {}
    for key in ("spmv", "ortho_pass1", "ortho_fused", "ortho_pass2", "restart"):
        if st[key + "_launches"] and st[key + "_ms"] > 0:
            launches = int(round(st[key + "_launches"] * frac))
            nbytes = st[key + "_bytes"] * frac
            classes[key] = dict(ms=st[key + "_ms"], bytes=nbytes, launches=launches,
                                gbs=nbytes / st[key + "_ms"] / 1e6)
    top = max(classes, key=lambda k: classes[k]["ms"])
    kernels = {k: {"launches": v["launches"], "avg_ms": v["ms"] / v["launches"],
                   "achieved_gbs": v["gbs"], "frac_of_measured": v["gbs"] / peak,
                   "frac_of_8tbs": v["gbs"] / 8000.0,
                   "share_of_step": v["ms"] / ms} for k, v in classes.items()}
    return classes, top, kernels


def run_b200(args):
    from arnoldi_b200 import _lib, partial_schur
    from arnoldi_b200.matrices import lap2d
    from arnoldi_b200.rotate import rotate
    from arnoldi_b200.solver import DeviceSolver
    from arnoldi_b200.utils import arg_largest_real, rand_normalized_vector

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        return run_b200_multi(args, rank, world, local)

    lib = _lib.load()
    if lib.ab200_device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    parity = None if args.no_parity else parity_block(device=local)
    grid = args.grid
    A = lap2d(grid)
    n = A.shape[0]
    np.random.seed(0)
    v0 = rand_normalized_vector(n, np.complex128)
    H = np.zeros((MAX_DIM + 1, MAX_DIM), np.complex128)

    dev = DeviceSolver(n, MAX_DIM, device=local)
    if args.complex_storage:
        dev.set_option("real_mode", 0)
    for kv in args.option:
        k, v = kv.split("=")
        dev.set_option(k, int(v))
    dev.set_timing(True)
    dev.set_csr(A.indptr, A.indices, A.data)
    dev.set_columns(0, v0)
    host_ms = []

    def grow(start):
        cols, n_iter, brk = dev.expand(start, MAX_DIM, TOL)
        assert n_iter == MAX_DIM and not brk
        for j in range(start, n_iter):
            H[: j + 2, j] = cols[: j + 2, j]

    def cycle():
        m = MAX_DIM
        t0 = time.perf_counter()
        T2, Q = rotate(H[:m, :m], arg_largest_real, fast_real=FAST_SCHUR and not args.complex_storage)
        spike = H[m, :m] @ Q[:, :P]
        host_ms.append(1e3 * (time.perf_counter() - t0))
        dev.restart(Q, m, P)
        H[:P, :P] = T2[:P, :P]
        H[P, :P] = spike
        H[P, P:] = 0
        grow(P)

    grow(0)
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(3, args.warmup)):
        cycle()
    extra = warm_for_sampler(cycle, None)
    dev.synchronize()
    dev.reset_stats()
    host_ms.clear()
    cyc_s = []
    dev.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        dev.set_timing(i % EVENT_EVERY == 0)
        tc = time.perf_counter()
        cycle()                      # ends with a stream synchronise: the host clock brackets it
        cyc_s.append(time.perf_counter() - tc)
    ms = dev.timer_stop()
    t1 = time.perf_counter()
    wall = t1 - t0
    st = dev.stats()
    matvecs = args.steps * (MAX_DIM - P)
    assert st["arnoldi_steps"] == matvecs, (st["arnoldi_steps"], matvecs)
    value = matvecs / (ms * 1e-3)
    timed_cycles = len(range(0, args.steps, EVENT_EVERY))
    ms_instr = 1e3 * sum(cyc_s[0::EVENT_EVERY])
    # the same cycles without any per-kernel CUDA events, reported beside the timed region
    dev.set_timing(False)
    dev.timer_start()
    for _ in range(args.steps):
        cycle()
    ms_plain = dev.timer_stop()
    clk = clocks.stop(t0, t1)
    clk["warmup_extra_cycles"] = extra

    # ---- roofline of the dominant kernel class (CUDA events inside the timed region)
    peak, peak_src = measured_peak()
    classes, top, kernels = kernel_table(st, ms_instr, peak, timed_cycles / args.steps)
    roofline = {"bound": "hbm", "kernel": top, "achieved": classes[top]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": classes[top]["gbs"] / peak, "peak_source": peak_src,
                "traffic": traffic_from_profile(top, grid, st, classes[top]["bytes"] / classes[top]["launches"]),
                "algorithmic_bytes_per_launch": classes[top]["bytes"] / classes[top]["launches"],
                "kernel_events": event_note(args.steps),
                "kernels": kernels,
                "kernel_sum_ms_per_step": sum(v["ms"] for v in classes.values()) / timed_cycles,
                "ms_per_instrumented_step": ms_instr / timed_cycles,
                "host_rotate_ms_per_step": float(np.mean(host_ms[:args.steps])),
                "ms_per_step_without_kernel_events": ms_plain / args.steps,
                "dgks_second_round_fraction": st["second_rounds"] / max(1, st["arnoldi_steps"])}
    dev.close()

    # ---- e2e through the public API: host CSR -> (Q, T, history) on the host
    e2e = None
    if not args.no_e2e:
        Ap, keep = pinned_csr(lib, A)
        restarts = args.e2e_restarts
        out = {}
        h2d = A.data.nbytes + A.indices.nbytes + A.indptr.nbytes + 16 * n
        d2h = 16 * n * NEV + 16 * (MAX_DIM + 1) * MAX_DIM
        times = []
        for rep in range(args.e2e_reps + 1):
            np.random.seed(0)
            stats = {}
            t0 = time.perf_counter()
            Q, T, hist = partial_schur(Ap, NEV, max_dim=MAX_DIM, stopping_criterion=TOL,
                                       sort_function=arg_largest_real, max_restarts=restarts,
                                       raise_on_no_convergence=False, stats=stats, device=local,
                                       real_storage=not args.complex_storage,
                                       fast_real_schur=FAST_SCHUR)
            dt = time.perf_counter() - t0
            if rep > 0:
                times.append(dt)
            out = stats
        mv = out["true_matvecs"]
        e2e = {"value": mv / float(np.mean(times)), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "call": f"partial_schur(A_host, {NEV}, max_dim={MAX_DIM}, max_restarts={restarts}) "
                       f"= {mv} matvecs per call, {float(np.mean(times)):.3f} s per call incl. "
                       "v0 = randn(n) on the host, CSR upload from pinned memory, Q/T download; "
                       "bytes are per CALL (one upload of A and v0, one download of Q and T)",
               "reps": len(times),
               "host_phases_s": {k: round(v, 4) for k, v in out.get("host_phases_s", {}).items()}}
        for ptr in keep:
            lib.ab200_host_free(ptr)

    cpu = None
    cpu_conv = None
    if not args.no_cpu:
        threads = set_host_threads()
        arm = load_cpu_arm()
        matv, secs, note, total = cpu_partial_cycles(arm, A, 0)
        cpu = {"value": matv / secs, "unit": UNIT, "cores": threads, "kind": arm["kind"],
               "sample": note + f"; {secs:.1f} s timed of {total:.1f} s run; csr_matvec is "
                                "single-threaded, OpenBLAS uses all threads",
               "os_cpu_count": os.cpu_count()}
        if not args.no_converged:
            cpu_conv = cpu_converged(arm)
    del A
    converged = None if args.no_converged else converged_leg(cpu_conv, device=local)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": storage_note(st), "data": "synthetic",
        "config": workload_config(grid),
        "wall_ms_per_step": 1e3 * wall / args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "converged": converged,
        "parity": parity,
        "gpu_launches": int(st["kernel_launches"]), "clocks": clk,
    }
    print(json.dumps(line))


def traffic_from_profile(kernel, grid, st, bytes_per_launch=None):
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/traffic.json: measured DRAM bytes / algorithmic bytes of one launch, per
    storage mode) scaled to the average launch of this run; None when no capture covers it."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        mode = "real" if st.get("real_storage") else "complex"
        ent = t.get(f"{kernel}@lap2d({grid})@{mode}") or t.get(f"{kernel}@lap2d({grid})")
        if ent is None or bytes_per_launch is None:
            return None
        return float(ent["ratio"]) * float(bytes_per_launch)
    except Exception:
        return None


def run_b200_multi(args, rank, world, local):
    """Strong scaling on one box: the SAME config-2 solve, rows of A and V block-row
    sharded over `world` GPUs (one process each); halo of v read from peer HBM over
    NVLink inside the SpMV, inner products combined inside the reducing kernels over peer
    memory.  torch.distributed (NCCL) only bootstraps and takes the max of the per-rank
    timings."""
    import torch
    import torch.distributed as dist

    from arnoldi_b200 import partial_schur
    from arnoldi_b200.distributed import RowPartition, TorchComm, build_halo_plan, slice_rows
    from arnoldi_b200.matrices import lap2d
    from arnoldi_b200.rotate import rotate
    from arnoldi_b200.solver import DeviceSolver
    from arnoldi_b200.utils import arg_largest_real, rand_normalized_vector

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = TorchComm()
    parity = None if args.no_parity else parity_block(comm=comm, device=local)
    grid = args.grid
    A = lap2d(grid)
    n = A.shape[0]
    part = RowPartition(n, world)
    r0, r1 = part.rows(rank)
    plan = build_halo_plan(slice_rows(A, r0, r1))
    np.random.seed(0)
    v0 = rand_normalized_vector(n, np.complex128)
    H = np.zeros((MAX_DIM + 1, MAX_DIM), np.complex128)

    dev = DeviceSolver(n, MAX_DIM, device=local, row0=r0, nrows_local=r1 - r0)
    if args.complex_storage:
        dev.set_option("real_mode", 0)
    for kv in args.option:
        k, v = kv.split("=")
        dev.set_option(k, int(v))
    dev.set_timing(True)
    dev.connect(comm, part)
    dev.set_halo(plan.ghost_cols)
    dev.set_csr(plan.indptr, plan.indices, plan.data)
    dev.set_columns(0, v0[r0:r1])
    host_ms = []

    def grow(start):
        cols, n_iter, brk = dev.expand(start, MAX_DIM, TOL)
        assert n_iter == MAX_DIM and not brk
        for j in range(start, n_iter):
            H[: j + 2, j] = cols[: j + 2, j]

    def cycle():
        m = MAX_DIM
        t0 = time.perf_counter()
        T2, Q = rotate(H[:m, :m], arg_largest_real, fast_real=FAST_SCHUR and not args.complex_storage)
        spike = H[m, :m] @ Q[:, :P]
        host_ms.append(1e3 * (time.perf_counter() - t0))
        dev.restart(Q, m, P)
        H[:P, :P] = T2[:P, :P]
        H[P, :P] = spike
        H[P, P:] = 0
        grow(P)

    grow(0)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(max(3, args.warmup)):
        cycle()
    extra = warm_for_sampler(cycle, comm)
    dev.synchronize()
    dev.reset_stats()
    host_ms.clear()
    cyc_s = []
    comm.barrier()
    dev.timer_start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        dev.set_timing(i % EVENT_EVERY == 0)
        tc = time.perf_counter()
        cycle()
        cyc_s.append(time.perf_counter() - tc)
    ms = dev.timer_stop()
    t1 = time.perf_counter()
    wall = t1 - t0
    comm.barrier()
    ms_max = comm.max_float(ms)
    wall_max = comm.max_float(wall)
    st = dev.stats()
    matvecs = args.steps * (MAX_DIM - P)
    value = matvecs / (ms_max * 1e-3)
    timed_cycles = len(range(0, args.steps, EVENT_EVERY))
    ms_instr = 1e3 * sum(cyc_s[0::EVENT_EVERY])
    # the same cycles without any per-kernel CUDA events (see run_b200)
    dev.set_timing(False)
    comm.barrier()
    dev.timer_start()
    for _ in range(args.steps):
        cycle()
    ms_plain = comm.max_float(dev.timer_stop())
    clk = clocks.stop(t0, t1) if rank == 0 else None
    if clk is not None:
        clk["warmup_extra_cycles"] = extra
    peak, peak_src = measured_peak()
    classes, top, kernels = kernel_table(st, ms_instr, peak, timed_cycles / args.steps)
    roofline = {"bound": "hbm", "kernel": top, "achieved": classes[top]["gbs"], "peak": peak,
                "unit": "GB/s", "frac": classes[top]["gbs"] / peak, "peak_source": peak_src,
                "traffic": None, "kernel_events": event_note(args.steps),
                "kernels": kernels, "note": "rank 0, per-GPU bytes / per-GPU time",
                "kernel_sum_ms_per_step": sum(v["ms"] for v in classes.values()) / timed_cycles,
                "ms_per_instrumented_step": ms_instr / timed_cycles,
                "host_rotate_ms_per_step": float(np.mean(host_ms[:args.steps])),
                "ms_per_step_without_kernel_events": ms_plain / args.steps,
                "halo_entries_rank0": int(len(plan.ghost_cols))}
    launches = int(st["kernel_launches"])
    dev.disconnect()
    comm.barrier()
    dev.close()

    e2e = None
    if not args.no_e2e:
        restarts = args.e2e_restarts
        h2d = plan.data.nbytes + plan.indices.nbytes + plan.indptr.nbytes + 16 * (r1 - r0)
        d2h = 16 * (r1 - r0) * NEV + 16 * (MAX_DIM + 1) * MAX_DIM
        times = []
        mv = 0
        for rep in range(args.e2e_reps + 1):
            np.random.seed(0)
            stats = {}
            comm.barrier()
            t0 = time.perf_counter()
            partial_schur(A, NEV, max_dim=MAX_DIM, stopping_criterion=TOL,
                          sort_function=arg_largest_real, max_restarts=restarts,
                          raise_on_no_convergence=False, stats=stats, device=local, comm=comm,
                          real_storage=not args.complex_storage, fast_real_schur=FAST_SCHUR)
            dt = comm.max_float(time.perf_counter() - t0)
            if rep > 0:
                times.append(dt)
            mv = stats["true_matvecs"]
            phases = stats.get("host_phases_s", {})
        e2e = {"value": mv / float(np.mean(times)), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(d2h) * world,
               "call": f"partial_schur(A_host, {NEV}, max_dim={MAX_DIM}, max_restarts={restarts}, "
                       f"comm=...) = {mv} matvecs per call, {float(np.mean(times)):.3f} s per call "
                       "(max over ranks) incl. row slicing + halo plan on the host, uploads, "
                       "local Q download; bytes are per CALL, summed over ranks",
               "reps": len(times),
               "host_phases_s_rank0": {k: round(v, 4) for k, v in phases.items()}}
    del A
    converged = None if args.no_converged else converged_leg(None, comm=comm, device=local)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": storage_note(st), "data": "synthetic",
            "config": workload_config(grid), "parallelism": f"block-row x{world}",
            "wall_ms_per_step": 1e3 * wall_max / args.steps,
            "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "converged": converged,
            "parity": parity,
            "gpu_launches": launches, "clocks": clk,
        }
        print(json.dumps(line))
    comm.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=GRID_FULL, help="lap2d grid side (4096 = config 2)")
    ap.add_argument("--e2e-restarts", type=int, default=60)
    ap.add_argument("--e2e-reps", type=int, default=1)
    ap.add_argument("--ref-cycles", type=int, default=2,
                    help="restart cycles the reference arm times (one costs ~25 s of CPU)")
    ap.add_argument("--complex-storage", action="store_true",
                    help="keep the basis as complex128 even while it is provably real")
    ap.add_argument("--option", action="append", default=[], metavar="KEY=INT",
                    help="ab200_set_option for the value leg (A/B runs)")
    ap.add_argument("--exact-schur", action="store_true",
                    help="GPU arm: factor H with zgees exactly as the reference (default: dgees "
                         "while H is real, fast_real_schur=True)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-converged", action="store_true")
    args = ap.parse_args()
    global FAST_SCHUR
    FAST_SCHUR = not args.exact_schur
    # stdout carries exactly ONE line (the JSON record): anything libraries print while the
    # benchmark runs (e.g. NCCL's version banner) is diverted to stderr at the fd level
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import contextlib
    import io
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    for ln in lines[:-1]:
        print(ln, file=sys.stderr)
    if lines:
        print(lines[-1], flush=True)


if __name__ == "__main__":
    main()
