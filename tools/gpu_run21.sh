#!/bin/bash
mkdir -p gpurun_out
for N in 8 4 2; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_N$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_N$N.log
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-comparator > gpurun_out/bench_N1.log 2>&1
python - <<'PY'
import json
for f in ("bench_N1","bench_N2","bench_N4","bench_N8"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().strip().splitlines() if l.startswith("{")][-1])
        b=dict(d.get("breakdown_rank0_ms") or {}); st=b.pop("step_ms",[])
        print(f, "value %.2f TF  ms %.1f  e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), b, d["config"].get("backward_sweep_transport"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-2500:])
PY
