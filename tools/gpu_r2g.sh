#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kmeans_gpu.py -m gpu -q > gpurun_out/r02g_kmeans_tests.log 2>&1; tail -6 gpurun_out/r02g_kmeans_tests.log | cut -c1-300
U2B_BENCH_SKIP_CPU=1 timeout 600 python bench.py --workload kmeans --steps 20 --warmup 3 > gpurun_out/r02g_bench_kmeans.json 2> gpurun_out/r02g_bench_kmeans.err; tail -c 300 gpurun_out/r02g_bench_kmeans.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02g_bench_kmeans.json").read().strip().splitlines()[-1])
print("kmeans: %.3e emb/s, %.3f ms/iter, assign %.3f ms (frac %.3f), e2e %.3e" % (l["value"], l["ms_per_step"], l["roofline"]["ms_per_launch"], l["roofline"]["frac"], l["e2e"]["value"]))
PY
U2B_WGRAD2=1 U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r02g_bench_wgrad2.json 2> gpurun_out/r02g_bench_wgrad2.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02g_bench_wgrad2.json").read().strip().splitlines()[-1])
print("train wgrad2=1: %.2f img/s %.2f ms" % (l["value"], l["ms_per_step"]))
for g in l["roofline"]["groups"]:
    print("  %-6s %-34s x%d %8.3f ms %7.1f TF/s" % (g["kind"], g["shape_N_H_W_Cin_Cout_k_stride"], g["launches"], g["ms"], g["tflops"]))
PY
