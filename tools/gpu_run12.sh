#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_N1.log 2>&1
echo "exit $?" >> gpurun_out/bench_N1.log
timeout 600 python tools/run_configs.py --only c4,c5 > gpurun_out/configs45.log 2>&1
cut -c1-500 gpurun_out/configs45.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_N2.log 2>&1
echo "exit $?" >> gpurun_out/bench_N2.log
python - <<'PY'
import json
for f in ("bench_N1","bench_N2"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "value %.2f TF  ms %.1f  e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("breakdown_rank0_ms"))
        if "breakdown" in d: print("   ", {k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["breakdown"].items()}, d["roofline"]["achieved"], d["roofline"]["peak"], d["comparator"], d["cpu_baseline"])
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-800:])
PY
