#!/bin/bash
# last check of the round: the driver's default bench command, then as much of the GPU suite as the remaining budget allows
set -u
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r02_last_bench_n1.json 2> gpurun_out/r02_last_bench_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_last_bench_n1.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02_last_bench_n1.json").read().strip().splitlines()[-1])
    r = l["roofline"]
    print("value %.2f | %.2f ms/step | e2e %.2f | roofline %.1f frac %.3f traffic %s | kmeans %.3e | infer %.1f" % (
        l["value"], l["ms_per_step"], l["e2e"]["value"], r["achieved"], r["frac"], r.get("traffic"), l["kmeans"]["value"], l["infer"]["value"]))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 330 python -m pytest tests -m gpu -q -x --deselect tests/test_kmeans_gpu.py::test_assign_full_baseline_size_sampled_rows_vs_oracle 2>&1 | tail -4 | cut -c1-200
