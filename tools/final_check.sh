#!/bin/bash
# what the driver runs at round end, on one GPU: gpu tests, smoke, both bench arms
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/final_smoke.log
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/final_bench_ref.log 2>&1; echo "ref exit $?"
( time python bench.py ) > gpurun_out/final_bench.log 2>&1; echo "bench exit $?"
tail -n 3 gpurun_out/final_pytest.log; tail -n 2 gpurun_out/final_smoke.log
grep '^{' gpurun_out/final_bench_ref.log | cut -c1-300; grep real gpurun_out/final_bench_ref.log
grep '^{' gpurun_out/final_bench.log | cut -c1-1600; grep real gpurun_out/final_bench.log
