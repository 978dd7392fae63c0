"""Developer tool: REAL timeline of one replay of the whole-step CUDA graph (torch.profiler / CUPTI activity records,
no serialisation, warm caches) - complements the ncu launch list, whose per-launch times are cold-cache and serialised.
Prints: step span, sum of kernel durations, idle time between kernels on the critical stream, and the top kernels by
summed duration.  usage: python tools/timeline_static.py [out.txt] [top_n]"""
import collections, os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from u2seg_b200.bench_train import _to_device
from u2seg_b200.config import get_u2seg_cfg
from u2seg_b200.data_synth import synthetic_batch
from u2seg_b200.engine import Trainer

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline_static.txt"
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
torch.manual_seed(0)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:        # torchrun: the data-parallel step (SyncBN exchanges + overlapped gradient all-reduce); rank 0 reports
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16, device=dev, static_graph=True, g_max=20)
if world > 1:
    tr.broadcast_parameters(0)
pool = [_to_device(synthetic_batch(2, 1024, 1024, 800, 28, seed=1234 + 97 * rank + i), dev) for i in range(2)]
for i in range(4):
    tr.run_step(pool[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        tr.run_step(pool[i % 2])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs if "memcpy" not in e.name.lower() or True))
# split into replays by the largest gaps: take the last replay
if not ks:
    print("no CUDA events recorded"); sys.exit(1)
gaps = sorted(((ks[i + 1][0] - max(k[1] for k in ks[:i + 1]), i) for i in range(len(ks) - 1)), reverse=True)
cuts = sorted(i for _, i in gaps[:2])
last = ks[cuts[-1] + 1:] if cuts else ks
t0, t1 = last[0][0], max(k[1] for k in last)
busy_union, cur_end, idle = 0.0, last[0][0], 0.0
for s, e, _ in last:
    if s > cur_end:
        idle += s - cur_end
        cur_end = s
    if e > cur_end:
        busy_union += e - cur_end
        cur_end = e
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n in last:
    n = re.sub(r"(void |at::native::|\(anonymous namespace\)::|<unnamed>::|at::)", "", n)
    n = re.sub(r"std::array<char \*.*", "", n)[:110]
    agg[n][0] += 1
    agg[n][1] += e - s
tot = sum(v for _, v in agg.values())
lines = ["one graph replay: %d kernels/copies, span %.3f ms, GPU busy (union over streams) %.3f ms, idle gaps %.3f ms, "
         "sum of kernel durations %.3f ms" % (len(last), (t1 - t0) / 1e3, busy_union / 1e3, idle / 1e3, tot / 1e3)]
small = sum(1 for s, e, _ in last if e - s < 5.0)
lines.append("kernels shorter than 5 us: %d (%.3f ms)" % (small, sum(e - s for s, e, _ in last if e - s < 5.0) / 1e3))
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    lines.append("%9.1f us %5d x %7.1f us  %s" % (v, c, v / c, n))
if rank == 0:
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))
if world > 1:
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)       # see tools/pytest_then_exit.py: interpreter shutdown with captured NCCL graphs can block
