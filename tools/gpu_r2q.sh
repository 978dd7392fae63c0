#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 90 ncu --set full --clock-control none -k regex:knn_candidates_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_knn_candidates -f python tools/ncu_knn.py 2>&1 | tail -2
