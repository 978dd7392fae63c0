"""Developer tool: which Python lines of the static training step issue large copy / cast / layout-change kernels.
A TorchDispatchMode records aten copy-like ops on tensors >= 1M elements with the innermost repo frame."""
import collections, os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.utils._python_dispatch import TorchDispatchMode
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer
from u2seg_b200.bench_train import _to_device

MIN = int(os.environ.get("MIN_NUMEL", 1 << 20))
WATCH = ("copy_", "_to_copy", "clone", "contiguous", "add_", "add", "mul", "sum", "fill_", "zero_", "zeros", "zeros_like",
         "cat", "stack", "index", "index_select", "where", "mul_")
agg = collections.defaultdict(lambda: [0, 0])


class Tracer(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        name = func.__name__.split(".")[0]
        if name in WATCH:
            ts = [a for a in list(args) + [out] if isinstance(a, torch.Tensor)]
            big = max((t.numel() for t in ts), default=0)
            if big >= MIN:
                frames = [f for f in traceback.extract_stack() if "/u2seg_b200/" in f.filename and "engine.py" not in f.filename]
                site = "%s:%d %s" % (frames[-1].filename.split("/u2seg_b200/")[-1], frames[-1].lineno, frames[-1].name) if frames else "<autograd engine>"
                desc = " ".join("%s%s%s" % (str(t.dtype).replace("torch.", ""), list(t.shape),
                                            "" if t.is_contiguous() else ("cl" if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) else "strided"))
                                for t in ts[:3])
                k = (name, site, desc)
                agg[k][0] += 1
                agg[k][1] += sum(t.numel() * t.element_size() for t in ts[:3])
        return out


torch.backends.cudnn.benchmark = False
torch.manual_seed(0)
tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16, static_graph=True, g_max=20)
dev = torch.device("cuda", 0)
batch = _to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=1234), dev)
tr._lr_t.fill_(0.001)
tr._load_static_inputs(batch)
tr._static_step()
torch.cuda.synchronize()
with Tracer():
    tr._static_step()
torch.cuda.synchronize()
tot = sum(v[1] for v in agg.values())
print("total bytes touched by watched ops on tensors >= %d elements: %.1f MB" % (MIN, tot / 1e6))
for (name, site, desc), (c, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:80]:
    print("%8.1f MB %3d  %-10s %-55s %s" % (b / 1e6, c, name, site, desc))
