#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
python tools/diag_phases.py > gpurun_out/diag_phases.txt 2>&1; cat gpurun_out/diag_phases.txt
for pct in 0 50 100 150; do echo "STAGGER_PCT=$pct"; STPYB_STAGGER_PCT=$pct python tools/gemm_probe.py 2>&1 | grep -E "K=256 accumulate|K=512 accumulate|SYRK-like"; done > gpurun_out/stagger_sweep.txt 2>&1
cat gpurun_out/stagger_sweep.txt
CMD="python bench.py --problem-n 32768 --steps 1 --warmup 1 --no-comparator --no-cpu-baseline"
$CMD > gpurun_out/plain32k.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiAccum -s 30 -c 2 -o gpurun_out/prof_syrk_r01 $CMD > gpurun_out/ncu_full_syrk.log 2>&1
$CMD > gpurun_out/plain32k_b.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:EpiGram -c 1 -o gpurun_out/prof_gram_r01 $CMD > gpurun_out/ncu_full_gram.log 2>&1
tail -n 3 gpurun_out/ncu_full_syrk.log gpurun_out/ncu_full_gram.log
