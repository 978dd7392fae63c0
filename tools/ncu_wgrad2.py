"""Developer tool for `ncu --set full -k regex:conv_wgrad2_kernel`: a few launches of the 2-CTA tcgen05 weight gradient on one
shape. usage: python tools/ncu_wgrad2.py N H W Cin Cout k   (default: the mask head's 3x3, 256 ROIs x 14x14, 256->256)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.modeling.conv_tc import conv_wgrad2
a = [int(v) for v in sys.argv[1:7]] if len(sys.argv) >= 7 else [256, 14, 14, 256, 256, 3]
N, H, W, Cin, Cout, k = a
x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
gy = torch.randn(N, Cout, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_()
    conv_wgrad2(x, gy, k, k, 1, k // 2, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("done", a)
