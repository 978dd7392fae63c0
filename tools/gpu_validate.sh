#!/bin/bash
# One-call GPU validation used during development (run through gpurun from the repo root):
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh quick'      tests of the detector path + short N=1 bench
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh full'      whole -m gpu suite + full bench line + reference arm
#   gpurun --timeout 900 -- 'bash tools/gpu_validate.sh profile'    ncu launch list of one graph replay + phase summary
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_validate.sh multi 2'   torchrun tests + N-GPU bench
#   gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh drafts'    round-2 draft kernels: parity tests, then the bench per flag
# Results land in gpurun_out/ (merged back by gpurun); every step is under its own `timeout`.
set -u
mode=${1:-quick}
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
km = l.get("kmeans", {})
print("value %.2f %s | %.2f ms/step | e2e %.2f | launches %s | final_loss %s | kmeans %.3e" % (
    l["value"], l["unit"], l["ms_per_step"], l["e2e"]["value"], l.get("gpu_launches"), l.get("final_loss"), km.get("value", 0)))
PY
}
case "$mode" in
  quick)
    timeout 500 python -m pytest tests/test_optimizer_gpu.py tests/test_ops_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5
    U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
    tail -c 400 gpurun_out/bench_quick.err; line gpurun_out/bench_quick.json ;;
  full)
    timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
    timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
    tail -c 400 gpurun_out/bench_n1.err; line gpurun_out/bench_n1.json
    timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
    cut -c1-200 gpurun_out/bench_ref.json ;;
  profile)
    timeout 700 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_static.csv python tools/profile_static.py 2>&1 | tail -1 | cut -c1-120
    python tools/launch_summary.py gpurun_out/launches_static.csv 1 70 > gpurun_out/launches_static_summary.txt
    python tools/launch_phases.py gpurun_out/launches_static.csv > gpurun_out/launches_static_phases.txt
    head -3 gpurun_out/launches_static_summary.txt; grep "^==" gpurun_out/launches_static_phases.txt ;;
  multi)
    n=${2:-2}
    timeout 400 bash tools/run_multigpu_tests.sh "$n" 2>&1 | tail -4
    U2B_BENCH_SKIP_CPU=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
        --master-port 29511 bench.py --gpus "$n" --steps 20 --warmup 3 > "gpurun_out/bench_n$n.json" 2> "gpurun_out/bench_n$n.err"
    echo "bench rc=$?"; tail -c 300 "gpurun_out/bench_n$n.err"; line "gpurun_out/bench_n$n.json" ;;
  drafts)
    # round-2 drafts (never run on hardware in round 1): their parity tests first, then the step with everything on
    U2B_RUN_DRAFT_TESTS=1 timeout 600 python -m pytest tests/test_fused_losses_gpu.py -m gpu -q 2>&1 | tail -15
    for flags in "U2B_UPSAMPLE_KERNEL=1" "U2B_ROI_CHW=1" "U2B_FUSED_DET_LOSSES=1" "U2B_TC_WGRAD=1" \
                 "U2B_UPSAMPLE_KERNEL=1 U2B_ROI_CHW=1 U2B_FUSED_DET_LOSSES=1 U2B_TC_WGRAD=1"; do
      echo "== $flags"
      env $flags U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_KMEANS=1 timeout 200 python bench.py --steps 20 --warmup 3 \
          > gpurun_out/bench_drafts.json 2> gpurun_out/bench_drafts.err || tail -c 600 gpurun_out/bench_drafts.err
      line gpurun_out/bench_drafts.json
    done ;;
  *) echo "usage: $0 quick|full|profile|multi [N]|drafts"; exit 2 ;;
esac
