#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/dist_multi_check.py > gpurun_out/dist_multi2.log 2>&1; echo "exit $?" >> gpurun_out/dist_multi2.log; tail -n 4 gpurun_out/dist_multi2.log
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_N2.log 2>&1
echo "exit $?" >> gpurun_out/bench_N2.log
timeout 300 python -m pytest tests -m gpu -q -k "edge or optimize" -p no:cacheprovider 2>&1 | tail -n 3
python - <<'PY'
import json
for f in ("bench_N2",):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().strip().splitlines() if l.startswith("{")][-1])
        b=d.get("breakdown_rank0_ms"); st=b.pop("step_ms",[])
        print(f, "value %.2f TF  ms %.1f  e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), b, "first", st[:4], "last", st[-6:])
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-1500:])
PY
