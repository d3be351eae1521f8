"""bench.py -- GP fit + log-marginal-likelihood throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = GaussianProcess.fit_gp(x, y) followed by log_marginal(kernel, {}, 1.0) on the
C3 workload of BASELINE.json: Matern nu=2.5, n=65536, d=8, float64, synthetic data
(SURVEY.md section 8d).  `value` is whole-job algorithmic TFLOP/s, F = n^3/3 + 2 d n^2 + 4 n^2
per step, with x and y resident in HBM; `e2e` is the same metric through the public API
with pinned HOST tensors in and host results out.  The dominant kernel (trailing SYRK
update of the blocked Cholesky) is timed live with CUDA events by the library's
instrumentation and reported against the measured cuBLAS DGEMM rate of the same run.

--impl reference times the reference algorithm's CPU port (oracle/stpy_oracle.py: the
as-written fit_gp + _log_marginal_squared of stpy, which cannot travel to the GPU box in
source form) on the host cores, on a bounded sample of the same workload.
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
CPU_SAMPLE_N = 3072
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
def cpu_fit_lml_seconds(n, d, repeats=1):
    """Reference algorithm (as written in stpy) on the host cores: best-of-`repeats` seconds."""
    from oracle import stpy_oracle as O
    x, y = O.make_data(n, d, seed=0)
    kern = lambda a, b: O.matern_kernel(a, b, gamma=1.0, nu=2.5)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        O.fit_gp_as_written(kern, x, y, 0.1)
        O.lml_as_written(kern, x, y, 0.1, 1.0)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args):
    """Reference arm: stpy's own algorithm (CPU port) on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    torch.set_num_threads(os.cpu_count() or 1)
    n, d = CPU_SAMPLE_N, D_FULL
    for _ in range(args.warmup):
        cpu_fit_lml_seconds(n, d)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_fit_lml_seconds(n, d)
    sec = (time.perf_counter() - t0) / max(1, args.steps)
    val = flops_fit_lml(n, d) / sec / 1e12
    sample = ("fit_gp + _log_marginal_squared as written in stpy (dense Sigma^T Sigma, 2 Grams, gelsy lstsq with 1 and n "
              "right-hand sides, slogdet + solve), Matern-5/2, n=%d of 65536, d=8; TFLOP/s counts the same "
              "algorithmic F(n,d)" % n)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 Matern-5/2 GP fit_gp + log_marginal, fp64, d=8 (CPU sample n=%d)" % n,
                       "n": n, "d": d},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
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
        torch.set_num_threads(os.cpu_count() or 1)
        sec = cpu_fit_lml_seconds(CPU_SAMPLE_N, d)
        cpu = {"value": flops_fit_lml(CPU_SAMPLE_N, d) / sec / 1e12, "unit": UNIT, "cores": torch.get_num_threads(),
               "kind": "port", "seconds": sec,
               "sample": "stpy's as-written fit_gp + _log_marginal_squared (oracle port), Matern-5/2, n=%d of %d, d=%d, "
                         "one run" % (CPU_SAMPLE_N, n, d)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "seconds_per_step": ms * 1e-3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3: Matern nu=2.5 GP, fit_gp + log_marginal, n=%d, d=%d, fp64, 1 GPU" % (n, d),
                       "n": n, "d": d, "kernel": "matern nu=2.5 gamma=1 kappa=1", "noise_s": 0.1,
                       "outer_block": args.outer, "flops_per_step": F,
                       "l2": "working set (%.1f GB factor) far larger than the 126 MB L2; no flush needed" %
                             (n * n * 8 / 1e9)},
            "lml": float(lml), "e2e": e2e, "gpu_launches": int(launches.value), "clocks": clocks,
            "roofline": roofline, "breakdown": breakdown, "cpu_baseline": cpu, "comparator": comparator}
    print(json.dumps(line))
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
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
