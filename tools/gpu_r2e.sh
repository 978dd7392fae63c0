#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_losses_gpu.py -m gpu -q > gpurun_out/r02e_fused_tests.log 2>&1; tail -6 gpurun_out/r02e_fused_tests.log | cut -c1-400
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "maxpool or fpn_lateral" 2>&1 | tail -4 | cut -c1-300
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "conv2 or wgrad2 or autograd" > gpurun_out/r02e_conv_tests.log 2>&1; tail -5 gpurun_out/r02e_conv_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_baseline_config_gpu.py tests/test_model_gpu.py -m gpu -q -s > gpurun_out/r02e_model_tests.log 2>&1
grep -E "worst loss|gradient-norm|vs float64|AssertionError|passed|failed" gpurun_out/r02e_model_tests.log | cut -c1-300
timeout 600 python tools/conv_bench_bwd.py fpn_output2 semseg_p2 mask_fcn res4_conv2 res5_conv2 res4_conv3 fc1 lateral2 > gpurun_out/r02e_conv_bench_bwd.txt 2>&1; cat gpurun_out/r02e_conv_bench_bwd.txt | cut -c1-250
timeout 300 python tools/conv_bench2.py semseg mask res2_conv3 > gpurun_out/r02e_conv_bench2.txt 2>&1; cat gpurun_out/r02e_conv_bench2.txt | cut -c1-250
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err || tail -c 800 gpurun_out/bench_e.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_e.json").read().strip().splitlines()[-1])
    r = l["roofline"]
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s | loss %.4f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches"), l.get("final_loss", 0)))
    print("   in-step roofline: %s: %.0f TF/s frac %.3f share %.3f | all tcgen05: %s" % (r["kernel"][:90], r["achieved"], r["frac"], r.get("share_of_step_time", 0), r.get("all_tcgen05_launches")))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_WGRAD2=0 U2B_FUSED_MASK_LOSS=0 U2B_FUSED_FPN_SUM=0 U2B_MAXPOOL_KERNEL=0
run U2B_WGRAD2=0 U2B_FUSED_MASK_LOSS=1 U2B_FUSED_FPN_SUM=0 U2B_MAXPOOL_KERNEL=0
run U2B_WGRAD2=0 U2B_FUSED_MASK_LOSS=1 U2B_FUSED_FPN_SUM=1 U2B_MAXPOOL_KERNEL=1
run U2B_WGRAD2=1 U2B_FUSED_MASK_LOSS=1 U2B_FUSED_FPN_SUM=1 U2B_MAXPOOL_KERNEL=1
for shp in res2_conv3 fpn_output2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv2_kernel -s 2 -c 1 -o gpurun_out/r02e_ncu_conv2_$shp -f python tools/ncu_conv2.py $shp 2>&1 | tail -1
done
U2B_WGRAD2=0 timeout 300 python tools/timeline_static.py gpurun_out/r02e_timeline_static.txt 100 2>&1 | tail -3
