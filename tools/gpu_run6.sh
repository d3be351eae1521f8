#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/dist_multi_check.py > gpurun_out/dist_multi.log 2>&1
echo "exit $?" >> gpurun_out/dist_multi.log
tail -n 12 gpurun_out/dist_multi.log
timeout 600 $TR bench.py --gpus 2 --problem-n 32768 --steps 2 --warmup 1 > gpurun_out/bench2_32k.log 2>&1
echo "exit $?" >> gpurun_out/bench2_32k.log
tail -c 1500 gpurun_out/bench2_32k.log
timeout 900 $TR bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench2_64k.log 2>&1
echo "exit $?" >> gpurun_out/bench2_64k.log
tail -c 1500 gpurun_out/bench2_64k.log
