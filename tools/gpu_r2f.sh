#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_fused_losses_gpu.py -m gpu -q > gpurun_out/r02f_ops_tests.log 2>&1; tail -6 gpurun_out/r02f_ops_tests.log | cut -c1-300
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -q -k "conv2 or wgrad2 or autograd or deconv" > gpurun_out/r02f_conv_tests.log 2>&1; tail -5 gpurun_out/r02f_conv_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_baseline_config_gpu.py tests/test_model_gpu.py -m gpu -q -s > gpurun_out/r02f_model_tests.log 2>&1
grep -E "worst loss|gradient-norm|AssertionError|Error|passed|failed" gpurun_out/r02f_model_tests.log | cut -c1-300
timeout 300 python tools/conv_bench2.py fpn_output2 semseg mask_fcn res2_conv3 res4_conv1 lateral2 > gpurun_out/r02f_conv_bench2.txt 2>&1; cat gpurun_out/r02f_conv_bench2.txt | cut -c1-250
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err || tail -c 800 gpurun_out/bench_f.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_f.json").read().strip().splitlines()[-1])
    r = l["roofline"]
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s | loss %.4f" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches"), l.get("final_loss", 0)))
    a = r.get("all_tcgen05_launches", {})
    print("   in-step roofline: %s: %.0f TF/s frac %.3f | all tcgen05: %s launches %.2f ms %.0f TF/s, %.2f of step flop" % (r["kernel"][:100], r["achieved"], r["frac"], a.get("launches_per_step"), a.get("ms_per_step", 0), a.get("achieved", 0), a.get("share_of_step_flop", 0)))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_WGRAD2=0
run U2B_WGRAD2=1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench_full.json 2> gpurun_out/r02f_bench_full.err; tail -c 300 gpurun_out/r02f_bench_full.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02f_bench_full.json").read().strip().splitlines()[-1])
    print("FULL: value %.2f | %.2f ms | e2e %.2f | kmeans %.3e (%.3f ms, e2e %.3e, cpu %s) | infer %s" % (
        l["value"], l["ms_per_step"], l["e2e"]["value"], l["kmeans"]["value"], l["kmeans"]["ms_per_step"], l["kmeans"]["e2e"]["value"],
        l["kmeans"].get("cpu_baseline", {}).get("value"), {k: (v if not isinstance(v, dict) else '...') for k, v in l.get("infer", {}).items() if k in ("value", "ms_per_step")}))
    print("   infer rooflines:", {k: round(v["frac"], 3) for k, v in l.get("infer", {}).get("rooflines", {}).items()})
    print("   cpu_baseline:", l.get("cpu_baseline"))
except Exception as e:
    print("full bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference.json 2> gpurun_out/r02f_bench_reference.err; cut -c1-400 gpurun_out/r02f_bench_reference.json
U2B_WGRAD2=0 timeout 300 python tools/timeline_static.py gpurun_out/r02f_timeline_static.txt 100 2>&1 | tail -3
