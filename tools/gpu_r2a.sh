#!/bin/bash
# round-2 call: draft re-test, BASELINE-config parity tests, bench with/without the validated drafts, timeline + launch list
set -u
mkdir -p gpurun_out
U2B_RUN_DRAFT_TESTS=1 timeout 300 python -m pytest tests/test_fused_losses_gpu.py -m gpu -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_baseline_config_gpu.py -m gpu -q -s 2>&1 | grep -v Warning | tail -40
for flags in "U2B_NOFLAG=1" "U2B_UPSAMPLE_KERNEL=1 U2B_FUSED_DET_LOSSES=1"; do
  echo "== $flags"
  env $flags U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 timeout 200 python bench.py --steps 20 --warmup 3 \
      > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err || tail -c 600 gpurun_out/bench_a.err
  python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_a.json").read().strip().splitlines()[-1])
print("value %.2f | %.2f ms/step | e2e %.2f | launches %s" % (l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches")))
PY
done
U2B_UPSAMPLE_KERNEL=1 U2B_FUSED_DET_LOSSES=1 timeout 300 python tools/timeline_static.py gpurun_out/r02_timeline_static.txt 80 2>&1 | tail -14
U2B_UPSAMPLE_KERNEL=1 U2B_FUSED_DET_LOSSES=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r02_launches_static.csv python tools/profile_static.py 2>&1 | tail -1 | cut -c1-160
python tools/launch_summary.py gpurun_out/r02_launches_static.csv 1 90 > gpurun_out/r02_launches_static_summary.txt
python tools/launch_phases.py gpurun_out/r02_launches_static.csv > gpurun_out/r02_launches_static_phases.txt
head -3 gpurun_out/r02_launches_static_summary.txt; grep "^==" gpurun_out/r02_launches_static_phases.txt
