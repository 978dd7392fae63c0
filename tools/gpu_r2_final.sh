#!/bin/bash
# round-2 closing run on one B200 (tight GPU budget): parity of what changed last, staging A/B, smoke(), the driver's bench
# line, reference arm, launch list of the step and one full ncu capture of the in-step dominant kernel
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_conv_gpu.py tests/test_model_gpu.py tests/test_baseline_config_gpu.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-200 | tee gpurun_out/r02_final_tests.log
quick() { env "$@" U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_INFER=1 timeout 200 python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_q.err | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(l['ms_per_step'],4))" ; }
a=$(quick U2B_CONV2_STAGING=0); b=$(quick U2B_CONV2_STAGING=1); a2=$(quick U2B_CONV2_STAGING=0); b2=$(quick U2B_CONV2_STAGING=1)
echo "staging off: $a $a2 ms/step   on: $b $b2 ms/step" | tee gpurun_out/r02_final_staging_ab.txt
best=$(python -c "print(1 if min(float('$b' or 99), float('$b2' or 99)) <= min(float('$a' or 99), float('$a2' or 99)) else 0)")
export U2B_CONV2_STAGING=$best; echo "using U2B_CONV2_STAGING=$best" | tee -a gpurun_out/r02_final_staging_ab.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; tail -c 300 gpurun_out/r02_final_bench_n1.err
python - <<'PY'
import json
try:
    l = json.loads(open("gpurun_out/r02_final_bench_n1.json").read().strip().splitlines()[-1])
    print("value %.2f | %.2f ms/step | e2e %.2f | launches %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches")))
    r = l["roofline"]; print("roofline", r["kernel"][:80], round(r["achieved"], 1), round(r["frac"], 3), "all tcgen05:", r.get("all_tcgen05_launches"))
    km = l.get("kmeans", {}); print("kmeans", km.get("value"), km.get("ms_per_step"), (km.get("roofline") or {}).get("frac"), (km.get("e2e") or {}).get("value"), (km.get("cpu_baseline") or {}).get("value"))
    inf = l.get("infer", {}); print("infer", inf.get("value"), inf.get("ms_per_step"), (inf.get("e2e") or {}).get("value"), {k: (round(v["frac"], 3), round(v["ms_per_launch"] * 1e3, 1)) for k, v in inf.get("rooflines", {}).items()})
    print("cpu_baseline", l.get("cpu_baseline")); print("clocks", l.get("clocks"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_final_launches_static.csv python tools/profile_static.py 2>&1 | tail -1 | cut -c1-120
python tools/launch_summary.py gpurun_out/r02_final_launches_static.csv 1 80 > gpurun_out/r02_final_launches_static_summary.txt
python tools/launch_phases.py gpurun_out/r02_final_launches_static.csv > gpurun_out/r02_final_launches_static_phases.txt
head -4 gpurun_out/r02_final_launches_static_summary.txt; grep "^==" gpurun_out/r02_final_launches_static_phases.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad2_kernel -s 2 -c 1 -o gpurun_out/r02_final_ncu_wgrad2_mask_head -f python tools/ncu_wgrad2.py 2>&1 | tail -1
ncu -i gpurun_out/r02_final_ncu_wgrad2_mask_head.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_raw_summary.py > gpurun_out/r02_final_ncu_wgrad2_mask_head_summary.txt; cat gpurun_out/r02_final_ncu_wgrad2_mask_head_summary.txt | cut -c1-250
timeout 200 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err; cut -c1-300 gpurun_out/r02_final_bench_reference.json
