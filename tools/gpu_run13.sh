#!/bin/bash
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518"
timeout 300 $TR tools/dist_multi_check.py > gpurun_out/dist_multi8.log 2>&1; echo "exit $?" >> gpurun_out/dist_multi8.log; tail -n 5 gpurun_out/dist_multi8.log
for o in 512 256; do
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --outer $o > gpurun_out/bench_N8_o$o.log 2>&1
echo "exit $?" >> gpurun_out/bench_N8_o$o.log
done
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 600 $TR4 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_N4.log 2>&1
python - <<'PY'
import json
for f in ("bench_N8_o512","bench_N8_o256","bench_N4"):
    try:
        d=json.loads([l for l in open("gpurun_out/%s.log"%f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "value %.2f TF  ms %.1f  e2e %.2f"%(d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("breakdown_rank0_ms"))
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-1500:])
PY
