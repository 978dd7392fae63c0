#!/bin/bash
mkdir -p gpurun_out
timeout 40 ncu --set full --clock-control none -k regex:paste_masks_kernel -s 1 -c 1 -o gpurun_out/r02_ncu_paste_masks -f python tools/ncu_paste.py 2>&1 | tail -2
