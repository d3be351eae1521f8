#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 16 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
CMD="python bench.py --problem-n 32768 --steps 1 --warmup 1 --no-comparator --no-cpu-baseline"
$CMD > gpurun_out/plain32k.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiAccum -s 31 -c 1 -o gpurun_out/prof_syrk_k32_r01 $CMD > gpurun_out/ncu_full_syrk.log 2>&1
tail -n 2 gpurun_out/ncu_full_syrk.log | cut -c1-200
