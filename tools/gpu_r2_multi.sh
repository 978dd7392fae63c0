#!/bin/bash
# usage: bash tools/gpu_r2_multi.sh N   (gpurun --gpus N)
set -u
n=${1:-2}
mkdir -p gpurun_out
timeout 300 bash tools/run_multigpu_tests.sh "$n" > gpurun_out/r02_multi_tests_n$n.log 2>&1; tail -6 gpurun_out/r02_multi_tests_n$n.log | cut -c1-300
runb() {   # label, env...
  label=$1; shift
  env "$@" U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_INFER=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus "$n" --steps 20 --warmup 3 > "gpurun_out/r02_bench_n${n}_$label.json" 2> "gpurun_out/r02_bench_n${n}_$label.err" || tail -c 400 "gpurun_out/r02_bench_n${n}_$label.err"
  python - "gpurun_out/r02_bench_n${n}_$label.json" "$label" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s value %.2f img/s | %.2f ms/step | e2e %.2f" % (sys.argv[2], l["value"], l["ms_per_step"], l["e2e"]["value"]))
except Exception as e:
    print(sys.argv[2], "parse failed", e)
PY
}
runb xchg2 U2B_SYNCBN_XCHG2=1
runb xchg1 U2B_SYNCBN_XCHG2=0
runb xchg2_again U2B_SYNCBN_XCHG2=1
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 tools/timeline_static.py "gpurun_out/r02_timeline_static_n$n.txt" 70 2>&1 | tail -14 | cut -c1-170
