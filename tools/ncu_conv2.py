"""Developer tool for `ncu --set full -k regex:conv2_kernel`: a few launches of conv2 on one hot-path shape.
usage: python tools/ncu_conv2.py <shape name substring of tools/conv_bench2.SHAPES> [stats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.conv_bench2 import SHAPES
from u2seg_b200.modeling.conv_tc import conv2_nhwc
name = sys.argv[1]
stats = len(sys.argv) > 2
sh = next(s for s in SHAPES if name in s[0])
_, N, Cin, H, W, Cout, k, st = sh
x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
w = torch.randn(Cout, k, k, Cin, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    conv2_nhwc(x, w, st, k // 2, want_stats=stats)
torch.cuda.synchronize()
print("done", sh)
