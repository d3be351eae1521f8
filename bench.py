"""bench.py -- GP fit + log-marginal-likelihood throughput on B200 (BASELINE.json metric), and the other
BASELINE configurations on request.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c4|c5|all]

Default (what the driver runs): config C3 -- one "step" = GaussianProcess.fit_gp(x, y) followed by
log_marginal(kernel, {}, 1.0), Matern nu=2.5, n=65536, d=8, float64, synthetic data.
`value` is whole-job algorithmic TFLOP/s, F = n^3/3 + 2 d n^2 + 4 n^2 per step, with x and y resident in HBM;
`e2e` is the same metric through the public API with pinned HOST tensors in and host results out.  The
dominant kernel (trailing SYRK update of the blocked Cholesky) is timed live with CUDA events by the library's
instrumentation and reported against the cuBLAS DGEMM rate measured in the same run (and the nominal 40
TFLOP/s); the HBM-bound stages (Gram, triangular solves) are reported in GB/s against the measured copy rate.
With N > 1 (torchrun) the same workload runs through DistributedGP and the line carries a `parity` object.

--config c1|c2|c4|c5 prints one line per configuration with seconds, TFLOP/s or GB/s and roofline fractions
against both the nominal and the measured denominators, next to a bounded CPU figure.

CPU arm (`cpu_baseline`, and --impl reference): stpy is pure Python and cannot travel to the GPU box in source
form (licence: no redistribution), so the timed CPU code is the oracle's port of the reference's AS-WRITTEN
algorithm (oracle/stpy_oracle.py: dense Sigma^T Sigma, second Gram, pivoted-QR lstsq with 1 and n right-hand
sides, slogdet + solve; cross-checked against the unmodified reference's timings in this container,
profiles/cpu_port_vs_reference_r02.txt), on n in {2048, 4096, 8192} with a fitted c n^3 extrapolation to the
full size, next to the fair single-Cholesky evaluation (estimator.py:32-40).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "gp_fit_plus_lml_fp64_tflops"
UNIT = "TFLOP/s"
N_FULL, D_FULL = 65536, 8
CPU_STEP_N = 4096   # per-step sample of the --impl reference arm (2.2 s per step on 24 cores)
# log marginal likelihood of the C3 workload (seed 0) as printed by the single-GPU path (BENCH_r01 / SCALE_r01,
# N = 1: 173870.2726076183; N = 2: 173870.27260761836); the N > 1 lines are checked against it
C3_LML_REFERENCE = 173870.2726076183


def flops_fit_lml(n, d):
    return n ** 3 / 3.0 + 2.0 * d * n ** 2 + 4.0 * n ** 2


def make_data(n, d, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, dtype=torch.float64, generator=g) * 2 - 1
    y = torch.sin(3 * x.sum(dim=1, keepdim=True)) + 0.1 * torch.randn(n, 1, dtype=torch.float64, generator=g)
    return x, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [t.strip() for t in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def captured_traffic():
    """DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture (None if absent)."""
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "ncu_full_summaries_r01.json")))["syrk_final_k1024_banded"]
        row = cap["launches"][0]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        total = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            val, u = row[key].split()
            total += float(val) * unit[u]
        return total
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------------ CPU arms
def _c3_kernel():
    from oracle import stpy_oracle as O
    return lambda a, b: O.matern_kernel(a, b, gamma=1.0, nu=2.5)


def cpu_as_written_seconds(n, d, kern=None):
    """One run of the reference's as-written fit_gp + _log_marginal_squared (oracle port) on the host cores."""
    from oracle import stpy_oracle as O
    x, y = O.make_data(n, d, seed=0)
    kern = kern or _c3_kernel()
    t0 = time.perf_counter()
    O.fit_gp_as_written(kern, x, y, 0.1)
    O.lml_as_written(kern, x, y, 0.1, 1.0)
    return time.perf_counter() - t0


def cpu_fair_seconds(n, d, kern=None):
    """The fair CPU figure: ONE Gram + ONE Cholesky + cholesky_solve (Estimator.log_marginal, estimator.py:32-40),
    which yields alpha and the evidence -- what the GPU path computes."""
    from oracle import stpy_oracle as O
    x, y = O.make_data(n, d, seed=0)
    kern = kern or _c3_kernel()
    t0 = time.perf_counter()
    O.lml_cholesky(kern, x, y, 0.1, 1.0)
    return time.perf_counter() - t0


def fit_cubic(ns, secs):
    """Least-squares c in t = c n^3 (the as-written path is n^3-dominated from n ~ 2000)."""
    num = sum(t * n ** 3 for n, t in zip(ns, secs))
    den = sum(float(n) ** 6 for n in ns)
    return num / den


def cpu_baseline_c3(n_full, d, sizes=(2048, 4096, 8192)):
    """bounded CPU sample of the C3 workload + fitted c n^3 extrapolation + fair figure."""
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_as_written_seconds(1024, d)  # warm the thread pool / allocator
    secs = [cpu_as_written_seconds(n, d) for n in sizes]
    fair = [cpu_fair_seconds(n, d) for n in sizes]
    c, cf = fit_cubic(sizes, secs), fit_cubic(sizes, fair)
    big = sizes[-1]
    return {"value": flops_fit_lml(big, d) / secs[-1] / 1e12, "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": "stpy's as-written fit_gp + _log_marginal_squared (oracle port of gauss_procc.py:136-177, 336-378, "
                      "631-638), Matern-5/2, d=%d, one run each at n=%s of %d; value = F(n,d)/seconds at n=%d"
                      % (d, list(sizes), n_full, big),
            "seconds": dict(zip(map(str, sizes), secs)),
            "fitted_c_n3": c, "extrapolated_seconds_full_n": c * float(n_full) ** 3,
            "extrapolated_tflops_full_n": flops_fit_lml(n_full, d) / (c * float(n_full) ** 3) / 1e12,
            "fair_single_cholesky": {"what": "Estimator.log_marginal (estimator.py:32-40): one Gram, one Cholesky, "
                                             "cholesky_solve", "seconds": dict(zip(map(str, sizes), fair)),
                                     "fitted_c_n3": cf, "extrapolated_seconds_full_n": cf * float(n_full) ** 3,
                                     "tflops_at_largest_sample": flops_fit_lml(big, d) / fair[-1] / 1e12}}


def run_reference(args):
    """Reference arm: stpy's own algorithm (CPU port) on a bounded sample of the workload; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    torch.set_num_threads(os.cpu_count() or 1)
    n, d = CPU_STEP_N, D_FULL
    for _ in range(min(args.warmup, 2)):
        cpu_as_written_seconds(n, d)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_as_written_seconds(n, d)
    sec = (time.perf_counter() - t0) / max(1, args.steps)
    val = flops_fit_lml(n, d) / sec / 1e12
    base = cpu_baseline_c3(N_FULL, d)
    base.update({"value": val, "sample": "each step: as-written fit_gp + _log_marginal_squared (oracle port), Matern-5/2, "
                                          "n=%d of %d, d=%d; TFLOP/s counts the same algorithmic F(n,d); plus one run each "
                                          "at n=2048/4096/8192 for the fitted c n^3 extrapolation" % (n, N_FULL, d)})
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 Matern-5/2 GP fit_gp + log_marginal, fp64, d=8 (CPU sample n=%d)" % n,
                       "n": n, "d": d},
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    from stpy_b200 import _lib as L
    from stpy_b200.kernels import KernelFunction
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    L.load()
    if world > 1:
        from stpy_b200 import distributed as D
        dgemm = measure_dgemm().get("cublas_dgemm_tflops_burst")
        ref = C3_LML_REFERENCE if (args.n, args.d) == (N_FULL, D_FULL) else None
        return D.bench_main(args, METRIC, UNIT, flops_fit_lml, make_data, ClockSampler, measured_peaks,
                            lml_reference=ref, dgemm_tflops=dgemm)

    n, d = args.n, args.d
    x, y = make_data(n, d, seed=0)
    x_dev, y_dev = x.cuda(), y.cuda()
    kernel = KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, kappa=1.0, d=d)
    gp = GaussianProcess(kernel=kernel, s=0.1)
    gp.outer_block = args.outer
    F = flops_fit_lml(n, d)

    def step(xx, yy):
        gp.fit_gp(xx, yy)
        return gp.log_marginal(kernel, {}, 1.0)

    for _ in range(args.warmup):
        lml = step(x_dev, y_dev)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    L.call("stpyb_profile", 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        lml = step(x_dev, y_dev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    prof = (ctypes.c_double * 21)()
    launches = ctypes.c_longlong(0)
    L.call("stpyb_profile_read", prof, ctypes.byref(launches))
    L.call("stpyb_profile", 0)
    clocks = sampler.stop()
    value = F / (ms * 1e-3) / 1e12

    # the same kernel with the look-ahead off (nothing shares the SMs with it): one extra, untimed-for-`value` step
    prof_iso = (ctypes.c_double * 21)()
    old = ctypes.c_longlong(0)
    L.call("stpyb_set_lookahead_min_n", -1, ctypes.byref(old))
    L.call("stpyb_profile", 1)
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    L.call("stpyb_profile_read", prof_iso, ctypes.byref(ctypes.c_longlong(0)))
    L.call("stpyb_profile", 0)
    L.call("stpyb_set_lookahead_min_n", old.value, None)
    syrk_iso = (prof_iso[10] / (prof_iso[9] * 1e-3) / 1e12) if prof_iso[9] > 0 else None

    # end to end: pinned host inputs, host outputs (A, LML) -- wall clock around the public API
    xh, yh = x.pin_memory(), y.pin_memory()
    step(xh, yh)
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = step(xh, yh)
        _ = float(out)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e = {"value": F / e2e_s / 1e12, "unit": UNIT, "seconds_per_step": e2e_s,
           "h2d_bytes_per_step": (n * d + n) * 8, "d2h_bytes_per_step": n * 8 + 4 + 24}

    cats = ["potrf_diag", "panel_trsm", "panel_update", "trailing_syrk", "gram", "other", "solves"]
    breakdown = {c: {"ms_per_step": prof[3 * i] / args.steps, "tflops": (prof[3 * i + 1] / (prof[3 * i] * 1e-3) / 1e12)
                     if prof[3 * i] > 0 else None, "launches_per_step": prof[3 * i + 2] / args.steps}
                 for i, c in enumerate(cats)}
    syrk_ms, syrk_fl, syrk_n = prof[9], prof[10], prof[11]
    # HBM-bound stages against the measured copy rate (MEASURED_PEAKS.json) and the nominal 8 TB/s:
    # Gram: 4 n^2 bytes written (lower tiles only); a triangular solve reads half of L once: 4 n^2 bytes
    hbm_meas = measured_peaks().get("hbm_gbs") or 6553.6
    gram_ms = prof[12] / args.steps
    solve_ms, solve_fl, solve_n = prof[18] / args.steps, prof[19] / args.steps, prof[20] / args.steps
    hbm_stages = {
        "gram": {"ms_per_step": gram_ms, "algorithmic_bytes": 4.0 * n * n,
                 "GBps": (4.0 * n * n / (gram_ms * 1e-3) / 1e9) if gram_ms > 0 else None},
        "triangular_solves": {"ms_per_step": solve_ms, "launches_per_step": solve_n, "algorithmic_bytes": 4.0 * solve_fl,
                              "GBps": (4.0 * solve_fl / (solve_ms * 1e-3) / 1e9) if solve_ms > 0 else None}}
    for st in hbm_stages.values():
        if st["GBps"]:
            st["frac_of_measured_hbm"] = st["GBps"] / hbm_meas
            st["frac_of_nominal_8TBps"] = st["GBps"] / 8000.0

    comparator = {}
    dgemm_tf = None
    if not args.no_comparator:
        comparator = run_comparators(n, gp)
        dgemm_tf = comparator.get("cublas_dgemm_tflops_burst")
    peaks = measured_peaks()
    peak = dgemm_tf if dgemm_tf else 40.0
    achieved = (syrk_fl / (syrk_ms * 1e-3) / 1e12) if syrk_ms > 0 else None
    roofline = {"bound": "tensor", "kernel": "gemm_nt_kernel<128x64 tile, BK=32 x 2 stages, EpiAccum> (trailing SYRK of POTRF)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if achieved else None,
                "frac_of_nominal_40": (achieved / 40.0) if achieved else None,
                "achieved_without_overlap": syrk_iso,
                "frac_without_overlap": (syrk_iso / peak) if syrk_iso else None,
                "overlap_note": ("`achieved` is live in the timed region, where the next panel's factorisation runs on a "
                                 "side stream and shares the SMs with this kernel (look-ahead); `achieved_without_overlap` "
                                 "is the same kernel on the same launches in one extra step with the look-ahead off"),
                "traffic": captured_traffic(),
                "traffic_note": ("bytes, dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the ncu --set full capture "
                                 "in profiles/ncu_full_summaries_r01.json[syrk_final_k1024_banded]: trailing SYRK of order 31744, "
                                 "K=1024 (1.036e12 flop, 30.1 ms); algorithmic bytes of that launch 8.35e9 (C tiles read + "
                                 "written, panel read once); tensor pipe 92.7 % active, DRAM at 5 % of peak -> tensor-bound"),
                "peak_source": ("cuBLAS DGEMM 8192^3 fp64 measured in this run (MEASURED_PEAKS.json has no fp64 entry; "
                                "nominal B200 fp64 tensor peak 40 TFLOP/s)" if dgemm_tf else "nominal 40 TFLOP/s (fallback)"),
                "launches_per_step": syrk_n / args.steps, "ms_per_launch": syrk_ms / syrk_n if syrk_n else None,
                "algorithmic_flops_per_launch": syrk_fl / syrk_n if syrk_n else None,
                "share_of_step": syrk_ms / args.steps / ms,
                "hbm_gbs_measured": peaks.get("hbm_gbs")}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_baseline_c3(n, d, sizes=(2048, 4096, 8192) if n >= 16384 else (1024, 2048))

    # size-independent parity at full size, outside the timed region: the same checks the N > 1 lines carry
    from stpy_b200.distributed import parity_checks

    class _Single:  # the single-GPU model seen through the interface parity_checks drives
        world, rank, group = 1, 0, None
        A = property(lambda self: gp._A_dev.view(-1, 1))

        def mean_std(self, xt):
            return gp.mean_std(xt)
    gp.fit_gp(x_dev, y_dev)
    parity = parity_checks(_Single(), kernel, x_dev, y_dev, 0.1, float(lml),
                           C3_LML_REFERENCE if (n, d) == (N_FULL, D_FULL) else None)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "seconds_per_step": ms * 1e-3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3: Matern nu=2.5 GP, fit_gp + log_marginal, n=%d, d=%d, fp64, 1 GPU" % (n, d),
                       "n": n, "d": d, "kernel": "matern nu=2.5 gamma=1 kappa=1", "noise_s": 0.1,
                       "outer_block": args.outer, "flops_per_step": F,
                       "l2": "working set (%.1f GB factor) far larger than the 126 MB L2; no flush needed" %
                             (n * n * 8 / 1e9)},
            "lml": float(lml), "e2e": e2e, "gpu_launches": int(launches.value), "clocks": clocks,
            "roofline": roofline, "breakdown": breakdown, "hbm_bound_stages": hbm_stages, "parity": parity,
            "cpu_baseline": cpu, "comparator": comparator}
    print(json.dumps(line))
    if not parity["ok"]:
        print("PARITY FAILURE: %s" % json.dumps(parity), file=sys.stderr)
        return 3
    return 0


# ------------------------------------------------------------------------------------ other BASELINE configs
def _events(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best, out = float("inf"), None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best, out


def _wall(fn, reps=3):
    torch.cuda.synchronize()
    best, out = float("inf"), None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def _tensor_roofline(tflops, dgemm):
    return {"bound": "tensor", "achieved": tflops, "peak": dgemm or 40.0, "unit": "TFLOP/s",
            "frac": tflops / (dgemm or 40.0), "frac_of_measured_dgemm": (tflops / dgemm) if dgemm else None,
            "frac_of_nominal_40": tflops / 40.0, "traffic": None,
            "peak_source": "cuBLAS DGEMM 8192^3 measured in this run" if dgemm else "nominal 40 TFLOP/s"}


def config_c1(args, dgemm):
    """C1: SE GP n=1024 d=2, fit + mean_std(256) + LML (the reference's CPU-runnable case): latency-bound."""
    from oracle import stpy_oracle as O
    from stpy_b200 import _lib as L
    from stpy_b200.kernels import KernelFunction as KF
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    n, d, nt = 1024, 2, 256
    x, y = O.make_data(n, d, seed=0)
    xt, _ = O.make_data(nt, d, seed=1)
    xd, yd, xtd = x.cuda(), y.cuda(), xt.cuda()
    k = KF(kernel_name="squared_exponential", gamma=0.5, kappa=1., d=d)
    gp = GaussianProcess(kernel=k, s=0.1)

    def step(xx, yy, xq):
        gp.fit_gp(xx, yy)
        mu, sd = gp.mean_std(xq)
        return mu, sd, gp.log_marginal(k, {}, 1.0)
    L.call("stpyb_profile", 0)
    t_dev, (mu, sd, lml) = _events(lambda: step(xd, yd, xtd), reps=10, warm=3)
    t_fit, _ = _events(lambda: gp.fit_gp(xd, yd), reps=10)
    t_ms, _ = _events(lambda: gp.mean_std(xtd), reps=10)
    t_lml, _ = _events(lambda: gp.log_marginal(k, {}, 1.0), reps=10)
    xh, yh, xth = x.pin_memory(), y.pin_memory(), xt.pin_memory()
    t_e2e, _ = _wall(lambda: step(xh, yh, xth), reps=10)
    F = flops_fit_lml(n, d) + float(n) * n * nt
    kern = lambda a, b: O.se_kernel(a, b, gamma=0.5)
    torch.set_num_threads(os.cpu_count() or 1)
    O.fit_gp_as_written(kern, x, y, 0.1)
    t0 = time.perf_counter()
    K_, A_ = O.fit_gp_as_written(kern, x, y, 0.1)
    O.mean_std_as_written(kern, x, y, 0.1, xt, K=K_)
    ref_lml = O.lml_as_written(kern, x, y, 0.1)
    t_cpu = time.perf_counter() - t0
    ref = O.gp_cholesky(kern, x, y, 0.1, xt)
    tf = F / t_dev / 1e12
    return {"metric": "c1_fit_meanstd_lml_seconds", "value": t_dev, "unit": "s", "higher_is_better": False,
            "config": {"workload": "C1: SE GP n=1024 d=2 fp64: fit_gp + mean_std(256) + log_marginal, 1 GPU", "n": n, "d": d,
                       "nt": nt, "flops_per_step": F},
            "seconds": {"fit": t_fit, "mean_std_256": t_ms, "log_marginal_after_fit": t_lml, "all_three": t_dev},
            "tflops": tf, "e2e": {"value": t_e2e, "unit": "s", "h2d_bytes_per_step": (n * d + n + nt * d) * 8,
                                  "d2h_bytes_per_step": (n + 2 * nt) * 8 + 24},
            "roofline": dict(_tensor_roofline(tf, dgemm), bound="latency",
                             note="8 dependent 128-column block steps: bound by kernel-launch and single-SM latencies, "
                                  "not by either roofline; the fractions are given for completeness"),
            "parity": {"mean_rel_inf": float((mu.cpu() - ref["mean"]).abs().max() / ref["mean"].abs().max()),
                       "var_rel_inf": float((sd.cpu() ** 2 - ref["std"] ** 2).abs().max() / (ref["std"] ** 2).abs().max()),
                       "lml_abs": abs(float(lml) - float(ref_lml))},
            "cpu_baseline": {"value": t_cpu, "unit": "s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "the full C1 workload once: as-written fit_gp + mean_std(256) + _log_marginal_squared"}}


def config_c2(args, dgemm):
    """C2: ARD-SE GP n=16384 d=10: fit + LML value + gradient w.r.t. the 10 lengthscales."""
    from oracle import stpy_oracle as O
    from stpy_b200.kernels import KernelFunction as KF
    from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
    n, d = 16384, 10
    x, y = O.make_data(n, d, seed=0)
    xd, yd = x.cuda(), y.cuda()
    ard0 = torch.linspace(0.8, 1.6, d, dtype=torch.float64)
    k = KF(kernel_name="ard", ard_gamma=ard0.clone(), d=d)
    gp = GaussianProcess(kernel=k, s=0.1)

    def step(xx, yy):
        gp.fit_gp(xx, yy)
        a = ard0.clone().requires_grad_(True)
        v = gp.log_marginal(k, {'0': {'ard_gamma': a}}, 1.0)
        v.backward()
        return float(v), a.grad
    t_dev, (v, g) = _events(lambda: step(xd, yd), reps=3, warm=2)
    xh, yh = x.pin_memory(), y.pin_memory()
    t_e2e, _ = _wall(lambda: step(xh, yh), reps=3)
    F = float(n) ** 3 + 2.0 * d * n * n  # n^3/3 factor + 2 n^3/3 inverse + Gram / derivative passes
    tf = F / t_dev / 1e12
    # CPU: as-written value + autograd backward (the reference's optimiser evaluation) on bounded sizes
    torch.set_num_threads(os.cpu_count() or 1)
    sizes, secs = (1024, 2048, 4096), []
    for m_ in sizes:
        xs, ys = O.make_data(m_, d, seed=0)
        t0 = time.perf_counter()
        O.fit_gp_as_written(lambda a, b: O.ard_kernel(a, b, ard0), xs, ys, 0.1)
        O.lml_grad_ard(xs, ys, 0.1, ard0)
        secs.append(time.perf_counter() - t0)
    c = fit_cubic(sizes, secs)
    return {"metric": "c2_fit_lml_grad_fp64_tflops", "value": tf, "unit": UNIT, "higher_is_better": True,
            "config": {"workload": "C2: ARD-SE GP n=16384 d=10 fp64: fit_gp + log_marginal + gradient w.r.t. 10 "
                                   "lengthscales, 1 GPU", "n": n, "d": d, "flops_per_step": F},
            "seconds": {"fit_plus_value_plus_gradient": t_dev}, "lml": v,
            "e2e": {"value": F / t_e2e / 1e12, "unit": UNIT, "seconds_per_step": t_e2e,
                    "h2d_bytes_per_step": (n * d + n) * 8, "d2h_bytes_per_step": n * 8 + (d + 3 + 18) * 8},
            "roofline": _tensor_roofline(tf, dgemm),
            "cpu_baseline": {"value": F * 0 + (float(sizes[-1]) ** 3 + 2.0 * d * sizes[-1] ** 2) / secs[-1] / 1e12, "unit": UNIT,
                             "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "as-written fit_gp + _log_marginal_squared + autograd backward (oracle port) at n=%s, "
                                       "d=10, one run each; value at n=%d" % (list(sizes), sizes[-1]),
                             "seconds": dict(zip(map(str, sizes), secs)), "fitted_c_n3": c,
                             "extrapolated_seconds_full_n": c * float(n) ** 3}}


def config_c4(args, dgemm):
    """C4: RFF m=8192, n=1e6, d=16: embedding + Bayesian linear regression posterior (m x m Cholesky) + mean_std(256)."""
    import numpy as np
    from oracle import stpy_oracle as O
    from stpy_b200.embeddings.embedding import RFFEmbedding
    from stpy_b200.continuous_processes.kernelized_features import KernelizedFeatures
    n, d, m, nt = 10 ** 6, 16, 8192, 256
    x, y = O.make_data(n, d, seed=0)
    xt, _ = O.make_data(nt, d, seed=1)
    np.random.seed(0)
    emb = RFFEmbedding(gamma=1.0, m=m, d=d)
    kf = KernelizedFeatures(embedding=emb, m=m, s=0.1, lam=1.0, d=d)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    kf.distributed = world > 1
    xd, yd, xtd = x.cuda(), y.cuda(), xt.cuda()

    def step(xx, yy, xq):
        kf.fit_gp(xx, yy)
        return kf.mean_std(xq)
    t_dev, (mu, sd) = _events(lambda: step(xd, yd, xtd), reps=2, warm=1)
    xh, yh, xth = x.pin_memory(), y.pin_memory(), xt.pin_memory()
    t_e2e, _ = _wall(lambda: step(xh, yh, xth), reps=2)
    F = float(n) * m * m + 2.0 * n * m * d + 2.0 * n * m + float(m) ** 3 / 3.0
    tf = F / t_dev / 1e12
    # CPU: the as-written primal path at (n=1e5, m=2048), scaled linearly in n and quadratically in m -- on rank 0 at
    # N = 1 only (under torchrun every rank would time it on the same host cores)
    cpu = None
    if world == 1:
        torch.set_num_threads(os.cpu_count() or 1)
        nc, mc = 10 ** 5, 2048
        np.random.seed(0)
        W = torch.from_numpy(np.random.normal(size=(mc, d)))
        t0 = time.perf_counter()
        Phi = O.rff_embed(x[:nc], W)
        O.blr_as_written(Phi, y[:nc], 0.1, 1.0, O.rff_embed(xt, W))
        t_cpu = time.perf_counter() - t0
        Fc = float(nc) * mc * mc + 2.0 * nc * mc * d + 2.0 * nc * mc + float(mc) ** 3 / 3.0
        cpu = {"value": Fc / t_cpu / 1e12, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "as-written embed + precompute (pinverse) + theta_mean + mean_std(256) (oracle port) at "
                         "n=1e5, m=2048, d=16, one run; the full size needs 65 GB for Phi alone",
               "seconds": t_cpu, "extrapolated_seconds_full_size": t_cpu * (n / nc) * (m / mc) ** 2}
    return {"metric": "c4_rff_blr_fp64_tflops", "value": tf, "unit": UNIT, "higher_is_better": True, "n_gpus": world,
            "config": {"workload": "C4: RFF m=8192, n=1e6, d=16 fp64: embedding + normal equations (Phi never stored) + "
                                   "m x m Cholesky + mean_std(256)", "n": n, "d": d, "m": m, "flops_per_step": F},
            "seconds": {"fit_plus_mean_std": t_dev},
            "e2e": {"value": F / t_e2e / 1e12, "unit": UNIT, "seconds_per_step": t_e2e,
                    "h2d_bytes_per_step": (n * d + n + nt * d) * 8, "d2h_bytes_per_step": 2 * nt * 8},
            "roofline": _tensor_roofline(tf / world, dgemm),
            "parity": {"pred_finite": bool(torch.isfinite(mu).all() and torch.isfinite(sd).all())},
            "cpu_baseline": cpu}


def config_c5(args, dgemm):
    """C5: 64 SE / Matern kernels on one dataset (n=8192, d=4): log marginal likelihood per kernel."""
    import numpy as np
    from oracle import stpy_oracle as O
    from stpy_b200.kernels import KernelFunction as KF
    from stpy_b200.sweep import lml_sweep, lml_sweep_distributed
    n, d = 8192, 4
    x, y = O.make_data(n, d, seed=0)
    gam = np.logspace(-1, 0.5, 32)
    ks = [KF(kernel_name="squared_exponential", gamma=float(g), d=d) for g in gam] + \
         [KF(kernel_name="matern", gamma=float(g), nu=2.5, d=d) for g in gam]
    xd, yd = x.cuda(), y.cuda()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sweep = lml_sweep_distributed if world > 1 else lml_sweep
    t_dev, vals = _events(lambda: sweep(ks, xd, yd, s=0.1), reps=3, warm=1)
    xh, yh = x.pin_memory(), y.pin_memory()
    t_e2e, _ = _wall(lambda: sweep(ks, xh, yh, s=0.1), reps=2)
    F = 64.0 * (float(n) ** 3 / 3.0 + 2.0 * d * n * n + 2.0 * n * n)
    tf = F / t_dev / 1e12
    cpu, parity = None, None
    if world == 1:  # CPU figure and the check against it on rank 0 at N = 1 only
        torch.set_num_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        refs = [float(O.lml_as_written(lambda a, b, g=g: O.se_kernel(a, b, gamma=float(g)), x, y, 0.1)) for g in gam[[3, 20]]]
        t_cpu = (time.perf_counter() - t0) / 2
        parity = {"lml_abs_vs_cpu_port": max(abs(float(vals[3]) - refs[0]), abs(float(vals[20]) - refs[1]))}
        cpu = {"value": (float(n) ** 3 / 3.0 + 2.0 * d * n * n + 2.0 * n * n) / t_cpu / 1e12, "unit": UNIT,
               "cores": torch.get_num_threads(), "kind": "port",
               "sample": "_log_marginal_squared as written (Gram + slogdet + solve) for 2 of the 64 kernels at "
                         "the full n=8192; seconds per kernel", "seconds": t_cpu,
               "extrapolated_seconds_64_kernels": 64 * t_cpu}
    return {"metric": "c5_sweep_64_kernels_fp64_tflops", "value": tf, "unit": UNIT, "higher_is_better": True,
            "n_gpus": world,
            "config": {"workload": "C5: 32 SE + 32 Matern-5/2 kernels, shared n=8192 d=4 dataset, LML per kernel", "n": n,
                       "d": d, "kernels": 64, "flops_per_step": F},
            "seconds": {"sweep": t_dev, "per_kernel": t_dev / 64},
            "e2e": {"value": F / t_e2e / 1e12, "unit": UNIT, "seconds_per_step": t_e2e,
                    "h2d_bytes_per_step": (n * d + n) * 8, "d2h_bytes_per_step": 64 * 8 * 3 + 64 * 4},
            "roofline": _tensor_roofline(tf / world, dgemm),
            "parity": parity, "cpu_baseline": cpu}


def run_config(name, args):
    """One JSON line for config c1|c2|c4|c5 (c3 is run_b200 / distributed.bench_main)."""
    from stpy_b200 import _lib as L
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    L.load()
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    if world > 1 and name in ("c1", "c2"):
        raise SystemExit("config %s is a single-GPU configuration" % name)
    dgemm = measure_dgemm().get("cublas_dgemm_tflops_burst")
    line = {"c1": config_c1, "c2": config_c2, "c4": config_c4, "c5": config_c5}[name](args, dgemm)
    line.setdefault("n_gpus", 1)
    line.update({"dtype": "f64", "data": "synthetic", "vs_baseline": None, "scaling": "strong"})
    if rank == 0:
        print(json.dumps(line), flush=True)
    torch.cuda.empty_cache()
    return 0


def measure_dgemm():
    """cuBLAS DGEMM 8192^3 through torch (comparator / roofline denominator; never on the product path)."""
    out = {}
    try:
        m = 8192
        a = torch.randn(m, m, dtype=torch.float64, device="cuda")
        b = torch.randn(m, m, dtype=torch.float64, device="cuda")
        for _ in range(2):
            torch.matmul(a, b)
        best = float("inf")
        for _ in range(5):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            torch.matmul(a, b)
            s1.record()
            torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        out["cublas_dgemm_tflops_burst"] = 2.0 * m ** 3 / (best * 1e-3) / 1e12
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(20):
            torch.matmul(a, b)
        s1.record()
        torch.cuda.synchronize()
        out["cublas_dgemm_tflops_sustained"] = 20 * 2.0 * m ** 3 / (s0.elapsed_time(s1) * 1e-3) / 1e12
        del a, b
    except Exception as e:  # pragma: no cover
        out["cublas_error"] = repr(e)
    torch.cuda.empty_cache()
    return out


def run_comparators(n, gp):
    """cuBLAS DGEMM and cuSOLVER POTRF (through torch), timed only -- never on the product path."""
    out = measure_dgemm()
    try:
        from stpy_b200.kernels import KernelFunction
        from stpy_b200.continuous_processes.gauss_procc import GaussianProcess
        nc = min(n, 32768)  # cuSOLVER through torch needs ~1 minute at 65536; the rate is what is compared
        xs, ys = make_data(nc, 8, seed=0)
        gpc = GaussianProcess(kernel=KernelFunction(kernel_name="matern", gamma=1.0, nu=2.5, d=8), s=0.1)
        gpc.fit_gp(xs.cuda(), ys.cuda())
        K = gpc.K  # full symmetric Gram + s^2 I on the device
        n = nc
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        Lc = torch.linalg.cholesky(K)
        s1.record()
        torch.cuda.synchronize()
        sec = s0.elapsed_time(s1) * 1e-3
        out["cusolver_potrf_seconds"] = sec
        out["cusolver_potrf_n"] = n
        out["cusolver_potrf_tflops"] = n ** 3 / 3.0 / sec / 1e12
        del K, Lc, gpc
    except Exception as e:  # pragma: no cover
        out["cusolver_error"] = repr(e)
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5", "all"],
                    help="BASELINE.json configuration (default c3: the one the metric is quoted on)")
    ap.add_argument("--problem-n", dest="n", type=int, default=N_FULL)
    ap.add_argument("--problem-d", dest="d", type=int, default=D_FULL)
    ap.add_argument("--outer", type=int, default=0,
                    help="panel width / K depth of the trailing update (default: 1024 on one or two GPUs, 512 on four or eight)")
    ap.add_argument("--depth", type=int, default=-1,
                    help="multi-GPU: steps the panel chain may lead the bulk updates (default: the number of GPUs)")
    ap.add_argument("--no-comparator", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.outer <= 0:
        # one GPU: 1024 is the measured optimum (profiles/outer_block_probe_r01.txt); distributed: 512 keeps the
        # latency-bound backward sweep short (one 512-block solve per hop)
        # (2 GPUs: 1024 measured 1478 ms against 1502 ms at 512; 4 and 8 GPUs: 512)
        args.outer = 1024 if int(os.environ.get("WORLD_SIZE", "1")) <= 2 else 512
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "all":
        world = int(os.environ.get("WORLD_SIZE", "1"))
        for name in (("c1", "c2") if world == 1 else ()) + ("c4", "c5"):
            run_config(name, args)
        return run_b200(args)
    if args.config != "c3":
        return run_config(args.config, args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
