#!/bin/bash
# PDL between libu2b200 kernels: correctness (BN / conv / model tests run with it on) and step time on / off
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_conv_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_baseline_config_gpu.py -m gpu -q -x 2>&1 | tail -6 | cut -c1-200
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err || tail -c 800 gpurun_out/bench_m.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_m.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s | loss %.4f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches"), l.get("final_loss", 0)))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_PDL=0
run U2B_PDL=1
run U2B_PDL=0
run U2B_PDL=1
