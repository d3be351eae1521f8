"""Host-side pieces of bench.py that run without a GPU: the CPU arm (oracle port timed on bounded samples with a
fitted c n^3 extrapolation and the fair single-Cholesky figure) and the --impl reference JSON line."""
import json
import os
import subprocess
import sys

from conftest import ROOT

sys.path.insert(0, ROOT)


def test_cpu_baseline_record_has_the_contract_keys():
    import bench
    rec = bench.cpu_baseline_c3(4096, 4, sizes=(128, 256, 384))
    assert rec["kind"] == "port" and rec["cores"] >= 1 and rec["unit"] == "TFLOP/s" and rec["value"] > 0
    assert set(rec["seconds"]) == {"128", "256", "384"}
    assert rec["extrapolated_seconds_full_n"] > rec["seconds"]["384"]
    fair = rec["fair_single_cholesky"]
    assert fair["extrapolated_seconds_full_n"] > 0 and set(fair["seconds"]) == {"128", "256", "384"}
    assert abs(bench.fit_cubic([10, 20], [2e-3 * 10 ** 3, 2e-3 * 20 ** 3]) - 2e-3) < 1e-15


def test_reference_arm_prints_one_json_line(monkeypatch):
    env = dict(os.environ, PYTHONPATH=ROOT)
    code = ("import bench, sys; bench.CPU_STEP_N = 200; "
            "bench.cpu_baseline_c3 = lambda n, d, sizes=(64, 96): bench.__dict__['_orig'](n, d, sizes=(64, 96)); "
            "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; sys.exit(bench.main())")
    code = "import bench; bench._orig = bench.cpu_baseline_c3; " + code
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gp_fit_plus_lml_fp64_tflops" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 exit without work under torchrun
    r2 = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT, timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == ""
