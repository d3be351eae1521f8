#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_probe.py > gpurun_out/gemm_probe.txt 2>&1
for o in 512 1024; do
timeout 600 python bench.py --problem-n 32768 --outer $o --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_32k_o$o.log 2>&1
done
timeout 600 python bench.py --problem-n 32768 --outer 256 --steps 2 --warmup 1 --no-cpu-baseline --no-comparator > gpurun_out/bench_32k_o256.log 2>&1
CMD="python bench.py --problem-n 8192 --steps 1 --warmup 1 --no-comparator --no-cpu-baseline"
$CMD > gpurun_out/plain8k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:potrf_diag_kernel -s 3 -c 1 -o gpurun_out/prof_diag2 $CMD > gpurun_out/ncu_full_diag2.log 2>&1
cat gpurun_out/gemm_probe.txt
python - <<'PY'
import json
for f in ("bench_32k_o256","bench_32k_o512","bench_32k_o1024"):
    try:
        d=json.loads(open("gpurun_out/%s.log"%f).read().strip().splitlines()[-1])
        print(f, "value %.2f TF  ms %.1f  syrk %.2f TF share %.2f"%(d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["share_of_step"]))
        print("   ", {k:(round(v["ms_per_step"],2), v["launches_per_step"]) for k,v in d["breakdown"].items()})
    except Exception as e:
        print(f, "ERR", e, open("gpurun_out/%s.log"%f).read()[-800:])
PY
