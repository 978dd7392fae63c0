"""Developer tool: one replay of the whole-step CUDA graph between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer
from u2seg_b200.bench_train import _to_device

torch.backends.cudnn.benchmark = False
torch.manual_seed(0)
tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16, static_graph=True, g_max=20)
dev = torch.device("cuda", 0)
pool = [_to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=1234 + i), dev) for i in range(2)]
for i in range(2):
    tr.run_step(pool[i % 2])
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.run_step(pool[0])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", {k: float(v) for k, v in tr._static_out[0].items()})
