#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -m gpu -q > gpurun_out/r02i_knn_tests.log 2>&1; tail -8 gpurun_out/r02i_knn_tests.log | cut -c1-300
timeout 600 python tools/knn_quick.py 2>&1 | tail -6
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err || tail -c 800 gpurun_out/bench_i.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_i.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches")))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_WGRAD2_MIN_GF=20
run U2B_WGRAD2_MIN_GF=35
run U2B_WGRAD2_MIN_GF=35 U2B_MULTI_STREAM=0
