"""Developer tool: attribute the copy / cast kernels of the static training step to the Python lines that issue them
(torch profiler with stacks, eager execution of Trainer._static_step)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer
from u2seg_b200.bench_train import _to_device

torch.backends.cudnn.benchmark = False
torch.manual_seed(0)
tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16, static_graph=True, g_max=20)
dev = torch.device("cuda", 0)
batch = _to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=1234), dev)
tr._lr_t.fill_(0.001)
tr._load_static_inputs(batch)
for _ in range(2):
    tr._static_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    tr._static_step()
    torch.cuda.synchronize()
want = sys.argv[1:] or ["aten::copy_", "aten::add_", "aten::add", "aten::mul", "aten::upsample_bilinear2d", "aten::fill_"]
agg = collections.defaultdict(lambda: [0, 0.0, None])
for ev in prof.events():
    if ev.name not in want or ev.device_time_total <= 0:
        continue
    here = [f for f in ev.stack if "/u2seg_b200/" in f or "tools/" in f]
    site = here[0].split("/u2seg_b200/")[-1] if here else ("<autograd> " + (ev.stack[0] if ev.stack else "?"))[:90]
    key = (ev.name, site, str(ev.input_shapes)[:70])
    agg[key][0] += 1
    agg[key][1] += ev.device_time_total
tot = collections.Counter()
for (name, site, shp), (c, t, _) in agg.items():
    tot[name] += t
print({k: round(v / 1e3, 3) for k, v in tot.items()}, "ms per step by op")
for (name, site, shp), (c, t, _) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print("%8.1f us %4d  %-24s %-60s %s" % (t, c, name, site, shp))
