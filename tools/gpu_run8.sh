#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-comparator --no-cpu-baseline"
$CMD > gpurun_out/plain64k.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches64k.csv $CMD > gpurun_out/ncu_launch64k.log 2>&1
$CMD > gpurun_out/plain64k_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:EpiAccum -s 8 -c 2 -o gpurun_out/prof_syrk_r01 $CMD > gpurun_out/ncu_full_syrk.log 2>&1
$CMD > gpurun_out/plain64k_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:EpiGram -c 1 -o gpurun_out/prof_gram_r01 $CMD > gpurun_out/ncu_full_gram.log 2>&1
tail -c 600 gpurun_out/plain64k.log
ls -la gpurun_out/
