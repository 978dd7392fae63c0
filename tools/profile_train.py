"""Developer tool: where does the training step spend its time (torch profiler + phase timers)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer
from u2seg_b200.bench_train import _to_device

torch.backends.cudnn.benchmark = False
import torch.distributed as dist
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if WORLD > 1:
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
cfg = get_u2seg_cfg(800)
tr = Trainer(cfg, amp_dtype=torch.bfloat16, graph_backbone=('--graph' in sys.argv))
dev = torch.device("cuda", torch.cuda.current_device())
pool = [_to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=i), dev) for i in range(2)]
for i in range(4):
    tr.run_step(pool[i % 2])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(4):
    tr.run_step(pool[i % 2])
torch.cuda.synchronize()
print("ms/step (benchmark=False):", (time.perf_counter() - t0) / 4 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        tr.run_step(pool[i % 2])
    torch.cuda.synchronize()
if WORLD == 1 or dist.get_rank() == 0:
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))
