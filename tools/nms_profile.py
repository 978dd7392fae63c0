"""Developer tool: where the single-CTA NMS scan spends its cycles (thread 0's view), RPN-sized input."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200 import _lib
from u2seg_b200.layers import batched_nms

g = torch.Generator().manual_seed(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8768
c = torch.rand(n, 2, generator=g) * 1024
wh = torch.exp(torch.rand(n, 2, generator=g) * 3 + 2.5)
b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1024).cuda()
s = torch.randn(n, generator=g).cuda()
lv = torch.randint(0, 5, (n,), generator=g).cuda()
L = _lib.lib()
for mk in (None, 2000):
    for _ in range(3):
        k = batched_nms(b, s, lv, 0.7, max_keep=mk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        batched_nms(b, s, lv, 0.7, max_keep=mk)
    e1.record()
    torch.cuda.synchronize()
    _lib.check(L.u2b_debug_nms_profile(1, None))
    reps = 10
    for _ in range(reps):
        batched_nms(b, s, lv, 0.7, max_keep=mk)
    out = (ctypes.c_uint64 * 6)()
    _lib.check(L.u2b_debug_nms_profile(0, out))
    cyc = [v / reps for v in out]
    print("n=%d max_keep=%s kept=%d  batched_nms %.1f us/call; scan cycles/call: wait_tile %.0f chain %.0f bar1 %.0f keep+or %.0f bar2 %.0f  (sum %.0f = %.1f us at 1.9 GHz)"
          % (n, mk, k.numel(), e0.elapsed_time(e1) / 20 * 1e3, cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], sum(cyc), sum(cyc) / 1.9e3))
