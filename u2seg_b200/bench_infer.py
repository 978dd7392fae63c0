"""bench.py workload `infer`: u2seg_R50_300 panoptic inference on synthetic 800x1333 images, batch 1
(BASELINE.json configs[4]): backbone + RPN + 3 cascade stages (ROIAlign) + mask head + paste_masks + merge."""
import json
import time

import torch


def run_infer(args, ClockSampler, load_peaks, dist_info, emit=True):
    from . import _lib
    from .config import get_u2seg_cfg
    from .modeling import build_model
    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = build_model(get_u2seg_cfg(300)).to(memory_format=torch.channels_last)
    with torch.no_grad():   # random-init weights: give the class scores a usable spread so detections exist
        for k in range(3):
            model.roi_heads.box_predictor[k].cls_score.weight.mul_(60.0)
    g = torch.Generator().manual_seed(7)
    imgs = [torch.randint(0, 256, (3, 800, 1333), generator=g, dtype=torch.uint8).pin_memory() for _ in range(4)]
    dimgs = [i.to(dev) for i in imgs]
    model.train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):   # calibrate BN running statistics
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.momentum = 1.0
        model.backbone(model.preprocess_image([{"image": dimgs[0]}]).tensor)
    model.eval()

    def step(img):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model([{"image": img, "height": 800, "width": 1333}])[0]

    for i in range(max(3, args.warmup)):
        out = step(dimgs[i % 4])
    ndet = len(out["instances"])
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    l0 = _lib.launch_count
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(dimgs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.launch_count - l0
    t0 = time.perf_counter()
    n = max(3, min(args.steps, 10))
    for i in range(n):
        out = step(imgs[i % 4].to(dev, non_blocking=True))
        pan = out["panoptic_seg"][0].cpu()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    line = {"metric": "u2seg_R50_300_inference_images_per_sec", "value": 1e3 / ms, "unit": "images/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "u2seg_R50_300.yaml panoptic inference, synthetic 800x1333 image, batch 1, "
                                   "%d detections pasted to full-resolution masks" % ndet},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": 1.0 / dt, "unit": "images/s", "h2d_bytes_per_step": 3 * 800 * 1333,
                    "d2h_bytes_per_step": 800 * 1333 * 4,
                    "what": "model([{image (pinned uint8 host), height, width}]) incl. H2D and D2H of the panoptic map"}}
    peaks = load_peaks()
    line["rooflines"] = infer_kernel_rooflines(peaks, dev)
    line["roofline"] = line["rooflines"]["paste_masks_kernel"]
    del model
    torch.cuda.empty_cache()
    if not emit:
        return line
    print(json.dumps(line))


def _size_law_boxes(n, W, H, g, lo=32.0, hi=512.0):
    """centres uniform over the image, sides log-uniform lo..hi px, clipped to the image."""
    import math
    cx, cy = torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H
    w = torch.exp(torch.rand(n, generator=g) * (math.log(hi) - math.log(lo)) + math.log(lo))
    h = torch.exp(torch.rand(n, generator=g) * (math.log(hi) - math.log(lo)) + math.log(lo))
    b = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    b[:, 0::2] = b[:, 0::2].clamp(0, W)
    b[:, 1::2] = b[:, 1::2].clamp(0, H)
    return b


def infer_kernel_rooflines(peaks, dev):
    """The two HBM-bound kernels BASELINE.json configs[4] names, on the SURVEY 8(d) fixed-detections micro-benchmark:
    paste_masks_in_image (100 masks of 28x28 -> 100 x 800 x 1333 bool; algorithmic bytes = N*H*W written + N*784*4 read)
    and the multi-level ROIAlign (1000 boxes, 7x7, p2..p5 of an 800x1344 image, 256 channels bf16; algorithmic bytes =
    output + rois + the feature-map area under the boxes, capped per level by the level itself). CUDA events on the launch
    stream; a 256 MB buffer is written between launches (L2 flush). Peak = measured HBM copy bandwidth."""
    from .layers import ROIPooler, paste_masks_in_image
    from .structures import Boxes
    g = torch.Generator().manual_seed(3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / n

    out = {}
    Hh, Ww, N = 800, 1333, 100
    masks = torch.rand(N, 28, 28, generator=g).to(dev)
    boxes = _size_law_boxes(N, Ww, Hh, g).to(dev)
    ms = timeit(lambda: paste_masks_in_image(masks, boxes, (Hh, Ww), 0.5))
    nbytes = N * Hh * Ww + N * 784 * 4
    out["paste_masks_kernel"] = {"bound": "hbm", "kernel": "paste_masks_kernel (100 x 28x28 -> 100 x 800 x 1333 bool)",
                                 "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm"], "peak_source": peaks["src"] + " HBM copy",
                                 "traffic": None, "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms}
    K, C, P = 1000, 256, 7
    strides = (4, 8, 16, 32)
    feats = [torch.randn(1, C, 800 // s, 1344 // s, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
             for s in strides]
    rb = _size_law_boxes(K, 1333, 800, g).to(dev)
    # the kernel alone, through the C ABI (the ROIPooler call around it adds box conversion + level assignment launches and
    # their host gaps, which is what round 1's 6 % mostly measured)
    import ctypes
    from . import _lib
    from .layers import assign_boxes_to_levels_rois, convert_boxes_to_pooler_format
    L = _lib.lib()
    rois5 = convert_boxes_to_pooler_format([rb]).contiguous()
    levels = assign_boxes_to_levels_rois(rois5, 2, 5, 224, 4)
    nhwc = [f.permute(0, 2, 3, 1) for f in feats]
    assert all(t.is_contiguous() for t in nhwc)
    ptrs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in nhwc])
    hs = (ctypes.c_int32 * 4)(*[t.shape[1] for t in nhwc])
    ws = (ctypes.c_int32 * 4)(*[t.shape[2] for t in nhwc])
    sc = (ctypes.c_float * 4)(*[1.0 / s for s in strides])
    pooled = torch.empty((K, P, P, C), dtype=torch.bfloat16, device=dev)

    def launch():
        _lib.check(L.u2b_roi_align_fwd(2, 4, ptrs, hs, ws, sc, C, _lib.ptr(rois5), _lib.ptr(levels), K, P,
                                       ctypes.c_void_p(pooled.data_ptr()), _lib.stream_ptr()), "u2b_roi_align_fwd")

    ms = timeit(launch)
    pooler = ROIPooler(P, tuple(1.0 / s for s in strides), 0, "ROIAlignV2")
    with torch.no_grad():
        ref = pooler(feats, [Boxes(rb)])
    assert torch.equal(ref.permute(0, 2, 3, 1).contiguous(), pooled), "direct launch differs from the ROIPooler result"
    # feature bytes under the boxes at their assigned level (poolers.py:23-59), capped by the level's size
    area = ((rb[:, 2] - rb[:, 0]) * (rb[:, 3] - rb[:, 1])).clamp(min=1e-6)
    lvl = torch.floor(4 + torch.log2(torch.sqrt(area) / 224 + 1e-8)).clamp(2, 5).long() - 2
    touched = 0.0
    for li, s in enumerate(strides):
        m = lvl == li
        px = float((((rb[m, 2] - rb[m, 0]) / s + 2) * ((rb[m, 3] - rb[m, 1]) / s + 2)).sum())
        touched += min(px, (800 // s) * (1344 // s)) * C * 2
    nbytes = K * C * P * P * 2 + K * 20 + touched
    out["roi_align_fwd_kernel"] = {"bound": "hbm", "kernel": "roi_align_fwd_kernel (1000 boxes, 7x7, 4 levels, 256 ch bf16; one launch through u2b_roi_align_fwd)",
                                   "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                   "frac": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm"], "peak_source": peaks["src"] + " HBM copy",
                                   "traffic": None, "algorithmic_bytes_per_launch": nbytes, "ms_per_launch": ms}
    return out
