#!/bin/bash
set -u
mkdir -p gpurun_out
U2B_RUN_DRAFT_TESTS=1 timeout 300 python -m pytest tests/test_fused_losses_gpu.py -m gpu -q -k relabel > gpurun_out/r02c_relabel.log 2>&1; tail -3 gpurun_out/r02c_relabel.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_conv_gpu.py -m gpu -q -x -k "not cluster and not conv2_forward" > gpurun_out/r02c_model_tests.log 2>&1; tail -5 gpurun_out/r02c_model_tests.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err || tail -c 800 gpurun_out/bench_c.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_c.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s | loss %.4f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches"), l.get("final_loss", 0)))
except Exception as e:
    print("bench failed", e)
PY
}
F="U2B_UPSAMPLE_KERNEL=1 U2B_FUSED_DET_LOSSES=1"
run $F U2B_CONV_POLICY=large3x3 U2B_CONV2=0 U2B_MULTI_STREAM=0
run $F U2B_CONV_POLICY=large3x3 U2B_CONV2=0 U2B_MULTI_STREAM=1
run $F U2B_CONV_POLICY=large3x3 U2B_CONV2=1 U2B_MULTI_STREAM=1
run $F U2B_CONV_POLICY=all U2B_CONV2=1 U2B_MULTI_STREAM=1
U2B_UPSAMPLE_KERNEL=1 U2B_FUSED_DET_LOSSES=1 timeout 300 python tools/timeline_static.py gpurun_out/r02c_timeline_static.txt 90 2>&1 | tail -3
