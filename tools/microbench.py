"""Developer tool / profiling driver: one invocation of each hand-written kernel at its BASELINE shape,
timed with CUDA events (prints algorithmic GB/s or TFLOP/s). Run under ncu for profiles/."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.layers import ROIPooler, paste_masks_in_image, crop_and_resize_masks, Matcher, batched_nms
from u2seg_b200.modeling.conv_tc import conv2d_nhwc

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
g = torch.Generator(device="cuda").manual_seed(0)
# ROIAlign, training shapes: pyramid of 2x1024^2, K=1024, 7x7 (box) and K=256, 14x14 (mask), bf16
feats = [torch.randn(2, 256, 1024 // s, 1024 // s, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last) for s in (4, 8, 16, 32)]
def boxes(n):
    c = torch.rand(n, 2, device="cuda", generator=g) * 1024
    wh = torch.exp(torch.rand(n, 2, device="cuda", generator=g) * 2.8 + 3.4)
    return torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1024)
b512 = [boxes(512), boxes(512)]
pool7 = ROIPooler(7, (0.25, 0.125, 0.0625, 0.03125))
pool14 = ROIPooler(14, (0.25, 0.125, 0.0625, 0.03125))
t = timeit(lambda: pool7(feats, b512))
out_bytes = 1024 * 256 * 49 * 2
res["roi_align_fwd_7x7_K1024"] = {"ms": t, "alg_MB": out_bytes / 1e6, "GBps_out_only": out_bytes / t / 1e6}
b128 = [boxes(128), boxes(128)]
t = timeit(lambda: pool14(feats, b128))
res["roi_align_fwd_14x14_K256"] = {"ms": t, "alg_MB": 256 * 256 * 196 * 2 / 1e6, "GBps_out_only": 256 * 256 * 196 * 2 / t / 1e6}
fr = [f.clone().requires_grad_(True) for f in feats]
o = pool7(fr, b512); go = torch.randn_like(o)
def bwd():
    o.backward(go, retain_graph=True)
t = timeit(bwd)
res["roi_align_bwd_7x7_K1024(+zero/cast)"] = {"ms": t}
# paste_masks: config 5, N=100, 800x1333
pm = torch.rand(100, 28, 28, device="cuda", generator=g)
pb = torch.cat([torch.rand(100, 2, device="cuda", generator=g) * 600, torch.rand(100, 2, device="cuda", generator=g) * 600 + 620], 1)
t = timeit(lambda: paste_masks_in_image(pm, pb, (800, 1333), 0.5))
res["paste_masks_N100_800x1333"] = {"ms": t, "alg_MB": 100 * 800 * 1333 / 1e6, "GBps": 100 * 800 * 1333 / t / 1e6}
# crop_and_resize: 128 fg rois on 20 gt masks of 1024^2
gm = torch.rand(20, 1024, 1024, device="cuda", generator=g) > 0.5
idx = torch.randint(0, 20, (128,), device="cuda", generator=g)
t = timeit(lambda: crop_and_resize_masks(gm, b128[0], 28, gt_index=idx))
res["crop_resize_masks_M128_28x28"] = {"ms": t}
# IoU + matcher: 20 gt x 261,888 anchors
an = boxes(261888); gt = boxes(20)
m = Matcher([0.3, 0.7], [0, -1, 1], True)
t = timeit(lambda: m.match_boxes(gt, an))
res["iou_match_G20_A261888"] = {"ms": t, "alg_MB": 261888 * (16 + 8 + 4 + 1) / 1e6, "GBps": 261888 * 29 / t / 1e6}
# NMS 10,000 boxes, 5 levels
nb = boxes(10000); ns = torch.randn(10000, device="cuda", generator=g); nl = torch.randint(0, 5, (10000,), device="cuda", generator=g)
t = timeit(lambda: batched_nms(nb, ns, nl, 0.65, max_keep=4000))
res["batched_nms_10000_keep4000"] = {"ms": t}
# tcgen05 conv: FPN output2 / RPN p2 shape
x = torch.randn(2, 256, 256, 256, device="cuda", generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
w = torch.randn(256, 3, 3, 256, device="cuda", generator=g).bfloat16()
t = timeit(lambda: conv2d_nhwc(x, w, 1, 1))
fl = 2.0 * 2 * 256 * 256 * 256 * 256 * 9
res["conv_tc_3x3_256_256x256x2"] = {"ms": t, "TFLOPs": fl / t / 1e9}
print(json.dumps(res, indent=1))
