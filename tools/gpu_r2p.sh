#!/bin/bash
# row N3 (DINO ViT extractor): parity on the GPU, then its bench line
set -u
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_dino_gpu.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -25 | cut -c1-220 | tee gpurun_out/r02_dino_tests.log
timeout 120 python bench.py --workload dino > gpurun_out/r02_bench_dino.json 2> gpurun_out/r02_bench_dino.err; tail -c 600 gpurun_out/r02_bench_dino.err; cut -c1-1500 gpurun_out/r02_bench_dino.json
