#!/bin/bash
# multi-GPU tests: one pytest process per GPU under torchrun (usage: tools/run_multigpu_tests.sh [NGPUS])
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/pytest_then_exit.py tests/test_syncbn_multigpu.py -q -x -p no:cacheprovider
