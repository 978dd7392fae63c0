#!/bin/bash
# round-2 GPU call: tests touched since the last full run, default bench line, timeline of the static step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_model_gpu.py::test_training_gradients_match_oracle" "tests/test_kmeans_gpu.py::test_run_kmeans_save_load_and_decode_json" \
   tests/test_ops_gpu.py -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/r02k_tests.log; tail -16 gpurun_out/r02k_tests.log | cut -c1-220
timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 600 gpurun_out/r02k_bench.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02k_bench.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches")))
    print("roofline", l["roofline"]["kernel"][:90], l["roofline"]["achieved"], l["roofline"]["frac"])
    km = l.get("kmeans", {}); print("kmeans", km.get("value"), km.get("ms_per_step"), km.get("e2e"), km.get("cpu_baseline"))
    inf = l.get("infer", {}); print("infer", inf.get("value"), inf.get("ms_per_step"), {k: (round(v["frac"], 3), round(v["ms_per_launch"]*1e3,1)) for k, v in inf.get("rooflines", {}).items()})
    print("cpu_baseline", l.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python tools/timeline_static.py > gpurun_out/r02k_timeline_static.txt 2>&1; head -45 gpurun_out/r02k_timeline_static.txt | cut -c1-170
timeout 600 python tools/aten_ops_static.py gpurun_out/r02k_aten_ops.txt 2>&1 | tail -52 | cut -c1-230
