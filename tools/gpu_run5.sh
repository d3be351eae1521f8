#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python tools/gemm_probe.py > gpurun_out/gemm_probe.txt 2>&1
cat gpurun_out/gemm_probe.txt
for o in 256 512; do
timeout 600 python bench.py --problem-n 32768 --outer $o --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_32k_o$o.log 2>&1
done
timeout 900 python bench.py --outer 512 --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_64k_o512.log 2>&1
python - <<'PY'
import json
for f in ("bench_32k_o256","bench_32k_o512","bench_64k_o512"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, "value %.2f TF  ms %.1f  syrk %.2f TF share %.2f"%(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"]))
        print("   ", {k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["breakdown"].items()})
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-800:])
PY
