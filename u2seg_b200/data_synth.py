"""Synthetic COCO-panoptic-shaped inputs (SURVEY §8d recipe): uint8 RGB image, G instances with
boxes over all FPN levels, elliptical bit masks, block-structured semantic labels with 5% ignore."""
import math

import torch

from .structures import BitMasks, Boxes, Instances


def synthetic_batch(n_images, H, W, num_classes, sem_classes, seed, G=20, min_size=32, max_size=512, device="cpu",
                    pin=False):
    g = torch.Generator().manual_seed(seed)
    ys = torch.arange(H, dtype=torch.float32)[:, None] + 0.5
    xs = torch.arange(W, dtype=torch.float32)[None, :] + 0.5
    out = []
    for _ in range(n_images):
        image = torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)
        cx, cy = torch.rand(G, generator=g) * W, torch.rand(G, generator=g) * H
        lo, hi = math.log(min(min_size, W / 4)), math.log(min(max_size, W / 2))
        bw = torch.exp(torch.rand(G, generator=g) * (hi - lo) + lo)
        bh = torch.exp(torch.rand(G, generator=g) * (hi - lo) + lo)
        x0, y0 = (cx - bw / 2).clamp(0, W - 8), (cy - bh / 2).clamp(0, H - 8)
        x1 = torch.maximum((cx + bw / 2).clamp(0, W), x0 + 8)
        y1 = torch.maximum((cy + bh / 2).clamp(0, H), y0 + 8)
        b = torch.stack([x0, y0, x1, y1], dim=1)
        classes = torch.randint(0, num_classes, (G,), generator=g)
        ecx, ecy = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
        rx, ry = (b[:, 2] - b[:, 0]) / 2, (b[:, 3] - b[:, 1]) / 2
        masks = (((xs[None] - ecx[:, None, None]) / rx[:, None, None]) ** 2
                 + ((ys[None] - ecy[:, None, None]) / ry[:, None, None]) ** 2) <= 1.0
        blk = 64 if H >= 256 else 16
        coarse = torch.randint(0, sem_classes, ((H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
        sem = coarse.repeat_interleave(blk, 0).repeat_interleave(blk, 1)[:H, :W].clone()
        sem[torch.rand(H, W, generator=g) < 0.05] = 255
        sem = sem.long()
        if pin:
            image, b, classes, masks, sem = [t.pin_memory() for t in (image, b, classes, masks, sem)]
        inst = Instances((H, W))
        inst.gt_boxes = Boxes(b.to(device))
        inst.gt_classes = classes.to(device)
        inst.gt_masks = BitMasks(masks.to(device))
        out.append({"image": image.to(device), "instances": inst, "sem_seg": sem.to(device), "height": H, "width": W})
    return out
