#!/bin/bash
mkdir -p gpurun_out
python tools/diag_phases.py > gpurun_out/diag_phases.txt 2>&1; cat gpurun_out/diag_phases.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python tools/run_configs.py --only c1,c5 > gpurun_out/configs15.log 2>&1
cut -c1-400 gpurun_out/configs15.log
timeout 600 python bench.py --problem-n 8192 --steps 5 --warmup 3 --no-cpu-baseline --no-comparator > gpurun_out/bench_8k.log 2>&1
timeout 900 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_64k.log 2>&1
python - <<'PY'
import json
for f in ("bench_8k","bench_64k"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, "value %.2f TF  ms %.1f  syrk %.2f TF share %.2f"%(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"]))
        print("   ", {k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["breakdown"].items()})
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-800:])
PY
