#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_model_gpu.py::test_training_gradients_match_oracle" "tests/test_kmeans_gpu.py::test_run_kmeans_save_load_and_decode_json" \
   tests/test_ops_gpu.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/r02l_tests.log; grep "relative L2\|worst loss\|passed\|failed" gpurun_out/r02l_tests.log | cut -c1-200
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err || tail -c 800 gpurun_out/bench_l.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_l.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches")))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_ROI_CHW=0
run U2B_ROI_CHW=1
run U2B_ROI_CHW=0
