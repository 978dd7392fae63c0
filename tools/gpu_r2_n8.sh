#!/bin/bash
# lean multi-GPU sanity at N GPUs: torchrun tests, then the bench with the multi-CTA SyncBN exchange and with round 1's
set -u
n=${1:-8}
mkdir -p gpurun_out
timeout 240 bash tools/run_multigpu_tests.sh "$n" > gpurun_out/r02_multi_tests_n$n.log 2>&1; grep -h "passed\|failed\|rror" gpurun_out/r02_multi_tests_n$n.log | sort | uniq -c | head -8
for x in 1 0; do
  U2B_SYNCBN_XCHG2=$x U2B_BENCH_SKIP_KMEANS=1 U2B_BENCH_SKIP_CPU=1 U2B_BENCH_SKIP_INFER=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus "$n" --steps 20 --warmup 3 > "gpurun_out/r02_bench_n${n}_xchg2_$x.json" 2> "gpurun_out/r02_bench_n${n}_xchg2_$x.err" || tail -c 600 "gpurun_out/r02_bench_n${n}_xchg2_$x.err"
  python - "gpurun_out/r02_bench_n${n}_xchg2_$x.json" "xchg2=$x" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-10s n=%d value %.2f img/s | %.3f ms/step | e2e %.2f | loss %.4f" % (sys.argv[2], l["n_gpus"], l["value"], l["ms_per_step"], l["e2e"]["value"], l.get("final_loss", 0)))
except Exception as e:
    print(sys.argv[2], "parse failed", e)
PY
done
