"""bench.py workload `infer`: u2seg_R50_300 panoptic inference on synthetic 800x1333 images, batch 1
(BASELINE.json configs[4]): backbone + RPN + 3 cascade stages (ROIAlign) + mask head + paste_masks + merge."""
import json
import time

import torch


def run_infer(args, ClockSampler, load_peaks, dist_info):
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
    print(json.dumps(line))
