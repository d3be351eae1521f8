#!/bin/bash
# first GPU pass: tests, smoke, small and full bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --problem-n 8192 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_8k.log 2>&1
echo "exit $?" >> gpurun_out/bench_8k.log
timeout 600 python bench.py --problem-n 32768 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_32k.log 2>&1
echo "exit $?" >> gpurun_out/bench_32k.log
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_64k.log 2>&1
echo "exit $?" >> gpurun_out/bench_64k.log
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_8k.log gpurun_out/bench_32k.log gpurun_out/bench_64k.log
