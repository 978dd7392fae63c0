#!/bin/bash
# usage: bash tools/gpu_r2_multi.sh N   (gpurun --gpus N)
set -u
n=${1:-2}
mkdir -p gpurun_out
timeout 420 bash tools/run_multigpu_tests.sh "$n" > gpurun_out/r02_multi_tests_n$n.log 2>&1; tail -6 gpurun_out/r02_multi_tests_n$n.log | cut -c1-300
for ov in 1 0; do
  skipkm=0; [ "$ov" = 0 ] && skipkm=1
  U2B_OVERLAP_ALLREDUCE=$ov U2B_BENCH_SKIP_KMEANS=$skipkm U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_INFER=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus "$n" --steps 20 --warmup 3 > "gpurun_out/r02_bench_n${n}_ov$ov.json" 2> "gpurun_out/r02_bench_n${n}_ov$ov.err"
  echo "bench n=$n overlap=$ov rc=$?"; tail -c 300 "gpurun_out/r02_bench_n${n}_ov$ov.err"
  python - "gpurun_out/r02_bench_n${n}_ov$ov.json" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    km = l.get("kmeans", {})
    print("value %.2f img/s | %.2f ms/step | e2e %.2f | kmeans %.3e (%.3f ms)" % (l["value"], l["ms_per_step"], l["e2e"]["value"], km.get("value", 0), km.get("ms_per_step", 0)))
except Exception as e:
    print("parse failed", e)
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29517 tools/timeline_static.py "gpurun_out/r02_timeline_static_n$n.txt" 70 2>&1 | tail -14 | cut -c1-170
