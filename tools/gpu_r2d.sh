#!/bin/bash
set -u
mkdir -p gpurun_out
U2B_RUN_DRAFT_TESTS=1 timeout 300 python -m pytest tests/test_fused_losses_gpu.py -m gpu -q -k relabel 2>&1 | grep -E "differ|passed|failed" | cut -c1-400
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "conv2 or wgrad2 or autograd" > gpurun_out/r02d_conv_tests.log 2>&1; tail -12 gpurun_out/r02d_conv_tests.log | cut -c1-300
timeout 600 python tools/conv_bench2.py > gpurun_out/r02d_conv_bench2.txt 2>&1; cat gpurun_out/r02d_conv_bench2.txt | cut -c1-250
timeout 600 python tools/conv_bench_bwd.py > gpurun_out/r02d_conv_bench_bwd.txt 2>&1; cat gpurun_out/r02d_conv_bench_bwd.txt | cut -c1-250
timeout 600 python -m pytest tests/test_baseline_config_gpu.py tests/test_model_gpu.py -m gpu -q -s > gpurun_out/r02d_model_tests.log 2>&1
grep -E "worst loss|gradient-norm|vs float64|AssertionError|passed|failed" gpurun_out/r02d_model_tests.log | cut -c1-300
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err || tail -c 800 gpurun_out/bench_d.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_d.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s | loss %.4f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches"), l.get("final_loss", 0)))
except Exception as e:
    print("bench failed", e)
PY
}
F="U2B_UPSAMPLE_KERNEL=1 U2B_FUSED_DET_LOSSES=1"
run $F U2B_CONV_POLICY=all U2B_WGRAD2=0
run $F U2B_CONV_POLICY=all U2B_WGRAD2=1
run $F U2B_CONV_POLICY=large3x3 U2B_WGRAD2=1
