#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --n 8192 --steps 3 --warmup 2 --no-cpu-baseline --no-comparator > gpurun_out/bench_8k.log 2>&1
timeout 600 python bench.py --n 16384 --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_16k.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_64k.log 2>&1
tail -n 4 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ("bench_8k","bench_16k","bench_64k"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, "value %.2f TF  ms %.1f  e2e %.2f  syrk %.2f TF share %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["share_of_step"]))
        print("   ", {k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["breakdown"].items()})
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-800:])
PY
