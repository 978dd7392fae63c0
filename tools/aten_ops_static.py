"""Developer tool: which ATen (library) ops remain in the static training step, with shapes and the u2seg_b200 source line
that issued them. Runs ONE eager (not captured) static step under torch.profiler with shapes + stacks.
usage: python tools/aten_ops_static.py [out.txt]"""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from u2seg_b200.bench_train import _to_device
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/aten_ops_static.txt"
torch.manual_seed(0)
tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16, static_graph=True, g_max=20)
dev = torch.device("cuda", 0)
batch = _to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=1234), dev)
tr.run_step(batch)                      # builds buffers, captures the graph (not used below)
torch.cuda.synchronize()
tr._load_static_inputs(batch)
for _ in range(2):
    tr._static_step()                   # eager warm-up
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    tr._static_step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
WANT = ("aten::copy_", "aten::add", "aten::add_", "aten::cat", "aten::mul", "aten::sum", "aten::index", "aten::index_select",
        "aten::gather", "aten::topk", "aten::sort", "aten::where", "aten::fill_", "aten::zero_", "aten::clone", "aten::_foreach_copy_",
        "aten::convolution_backward", "aten::cudnn_convolution", "aten::mm", "aten::addmm", "aten::bmm", "aten::threshold_backward",
        "aten::relu", "aten::sigmoid", "aten::div", "aten::sub", "aten::masked_fill_", "aten::scatter_", "aten::cumsum")
for e in prof.events():
    if e.name not in WANT or e.device_time_total <= 0:
        continue
    where = "?"
    for fr in (e.stack or []):
        if "u2seg_b200/" in fr and "/_lib.py" not in fr:
            where = fr.split("u2seg_b200/")[-1]
            break
    shapes = str([s for s in (e.input_shapes or []) if s])[:70]
    key = (e.name, shapes, where[:70])
    agg[key][0] += 1
    agg[key][1] += e.self_device_time_total
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
tot = sum(v for _, (c, v) in rows)
lines = ["ATen ops with device time in one eager static step: %.3f ms total (self device time)" % (tot / 1e3)]
for (name, shapes, where), (c, v) in rows[:70]:
    lines.append("%8.1f us %4d x  %-28s %-70s %s" % (v, c, name, shapes, where))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:50]))
