#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn_gpu.py -m gpu -q -x > gpurun_out/r02h_knn_tests.log 2>&1; tail -12 gpurun_out/r02h_knn_tests.log | cut -c1-300
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "fused_bn" 2>&1 | tail -3 | cut -c1-300
timeout 600 python bench.py --workload knn > gpurun_out/r02h_bench_knn.json 2> gpurun_out/r02h_bench_knn.err; tail -c 400 gpurun_out/r02h_bench_knn.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02h_bench_knn.json").read().strip().splitlines()[-1])
    print("knn: %.3e queries/s, %.1f ms total, candidate pass %.1f TF/s (frac %.3f), uncertified %s" % (l["value"], l["ms_per_step"], l["roofline"]["achieved"], l["roofline"]["frac"], l["config"]["uncertified_rows_recomputed"]))
except Exception as e:
    print("knn bench failed", e)
PY
run() { echo "== $*"; env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err || tail -c 800 gpurun_out/bench_h.err
  python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/bench_h.json").read().strip().splitlines()[-1])
    r = l["roofline"]
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches")))
    a = r.get("all_tcgen05_launches", {})
    print("   in-step roofline [%s]: %s: %.0f TF/s frac %.3f | all tcgen05: %s launches %.2f ms %.0f TF/s" % (r.get("method", "")[:60], r["kernel"][:90], r["achieved"], r["frac"], a.get("launches_per_step"), a.get("ms_per_step", 0), a.get("achieved", 0)))
    for g in r.get("groups", [])[:14]:
        print("     %-6s %-34s x%d %8.3f ms %7.1f TF/s" % (g["kind"], g["shape_N_H_W_Cin_Cout_k_stride"], g["launches"], g["ms"], g["tflops"]))
except Exception as e:
    print("bench failed", e)
PY
}
run U2B_BN_MASK_FROM_X=0
run U2B_BN_MASK_FROM_X=1
run U2B_BN_MASK_FROM_X=1 U2B_WGRAD2_MIN_GF=100
