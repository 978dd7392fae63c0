#!/bin/bash
set -u
timeout 200 python -m pytest tests/test_conv_gpu.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-200
