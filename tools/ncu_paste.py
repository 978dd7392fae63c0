"""Developer tool for `ncu --set full -k regex:paste_masks_kernel`: the inference shape of the bench (100 masks of 28x28 pasted
into 800x1333). usage: python tools/ncu_paste.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.bench_infer import _size_law_boxes
from u2seg_b200.layers import paste_masks_in_image
g = torch.Generator().manual_seed(0)
masks = torch.rand(100, 28, 28, generator=g).cuda()
boxes = _size_law_boxes(100, 1333, 800, g).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(2):
    flush.zero_()
    paste_masks_in_image(masks, boxes, (800, 1333), 0.5)
torch.cuda.synchronize()
print("done")
