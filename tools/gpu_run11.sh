#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
for N in 8 4 2; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
if [ $N = 8 ]; then timeout 300 $TR tools/dist_multi_check.py > gpurun_out/dist_multi8.log 2>&1; echo "exit $?" >> gpurun_out/dist_multi8.log; tail -n 4 gpurun_out/dist_multi8.log; fi
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_N$N.log 2>&1
echo "exit $?" >> gpurun_out/bench_N$N.log
tail -n 2 gpurun_out/bench_N$N.log | cut -c1-420
done
