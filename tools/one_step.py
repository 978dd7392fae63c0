"""Profiling driver: a few training steps of u2seg_R50_800 (2x1024^2 synthetic) for `ncu` launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer
from u2seg_b200.bench_train import _to_device
torch.backends.cudnn.benchmark = False
tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16)
dev = torch.device("cuda")
pool = [_to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=i), dev) for i in range(2)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(n):
    tr.run_step(pool[i % 2])
torch.cuda.synchronize()
print("done", n)
