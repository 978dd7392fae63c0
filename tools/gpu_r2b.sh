#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_baseline_config_gpu.py -m gpu -q -s > gpurun_out/r02b_baseline_tests.log 2>&1
grep -E "worst loss|gradient-norm|AssertionError|passed|failed" gpurun_out/r02b_baseline_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -x -k conv2 > gpurun_out/r02b_conv2_tests.log 2>&1
tail -15 gpurun_out/r02b_conv2_tests.log | cut -c1-300
timeout 600 python tools/conv_bench2.py > gpurun_out/r02b_conv_bench2.txt 2>&1
cat gpurun_out/r02b_conv_bench2.txt | cut -c1-260
