#!/bin/bash
mkdir -p gpurun_out
./tools/microbench > gpurun_out/microbench.txt 2>&1
timeout 300 python tools/dist_single_check.py > gpurun_out/dist_single.log 2>&1
echo "exit $?" >> gpurun_out/dist_single.log
CMD="python bench.py --problem-n 16384 --steps 1 --warmup 1 --no-comparator --no-cpu-baseline"
$CMD > gpurun_out/plain16k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches16k.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain16k_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 60 -c 4 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
$CMD > gpurun_out/plain16k_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:potrf_diag_kernel -s 3 -c 1 -o gpurun_out/prof_diag $CMD > gpurun_out/ncu_full_diag.log 2>&1
cat gpurun_out/microbench.txt | head -60
tail -n 3 gpurun_out/dist_single.log
