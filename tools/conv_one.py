"""Developer tool: a few launches of the bench's roofline kernel (conv_tc 3x3 256->256 on 2x256x256) for
`ncu --set full -k regex:conv_tc_kernel -c 1`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.modeling.conv_tc import conv2d_nhwc
x = torch.randn(2, 256, 256, 256, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
w = (torch.randn(256, 3, 3, 256, device="cuda") * 0.02).bfloat16()
for _ in range(3):
    conv2d_nhwc(x, w, 1, 1)
torch.cuda.synchronize()
