#!/bin/bash
# conv2 staging depth: parity, then the step with the short-K pipeline on / off
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err || tail -c 800 gpurun_out/bench_n.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_n.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | loss %.4f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("final_loss", 0)))
    for g in l["roofline"]["groups"]:
        sh = g["shape_N_H_W_Cin_Cout_k_stride"]
        if sh[5] == 1 and g["kind"] != "wgrad":
            print("     %-6s %-34s x%-2d %7.1f us each" % (g["kind"], sh, g["launches"], g["ms"] / g["launches"] * 1e3))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_CONV2_STAGING=0
run U2B_CONV2_STAGING=1
run U2B_CONV2_STAGING=0
run U2B_CONV2_STAGING=1
