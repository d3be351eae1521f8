#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/run_configs.py --out gpurun_out/configs_r01.json > gpurun_out/configs.log 2>&1
echo "exit $?" >> gpurun_out/configs.log
cat gpurun_out/configs.log | cut -c1-900
