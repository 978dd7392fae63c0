"""ORACLE (test infrastructure, never shipped): functional CPU restatement of the u2seg_R50_{300,800}
Panoptic-FPN step (fp32, plain torch + torchvision CPU ops — the same library calls the reference
makes), operating on a state_dict with the reference's parameter names.

Every function cites the reference file:line it follows (paths under /root/reference/detectron2).
Pinned against the unmodified reference by oracle/make_golden_detector.py (seeded synthetic
inputs, shared state_dict): tests/test_detector_oracle.py compares the ten training losses and the
inference outputs with the committed fixtures, and — when /root/reference is present — live.

RNG: subsample_labels consumes torch.randperm on the CPU generator in the same order as the
reference, so seeded runs reproduce the reference's sampling exactly.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F
from torchvision.ops import nms as tv_nms
from torchvision.ops import roi_align as tv_roi_align


@dataclass
class DetCfg:
    """Values of configs/COCO-PanopticSegmentation/u2seg_R50_800.yaml + its _BASE_ chain + config/defaults.py."""
    num_classes: int = 800
    sem_classes: int = 28
    pixel_mean: Tuple[float, ...] = (123.675, 116.280, 103.530)
    pixel_std: Tuple[float, ...] = (58.395, 57.120, 57.375)
    size_divisibility: int = 32
    anchor_sizes: Tuple[int, ...] = (32, 64, 128, 256, 512)
    anchor_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    rpn_iou_thresholds: Tuple[float, float] = (0.3, 0.7)
    rpn_batch: int = 256
    rpn_pos_fraction: float = 0.5
    rpn_pre_topk_train: int = 2000
    rpn_pre_topk_test: int = 1000
    rpn_post_topk_train: int = 4000
    rpn_post_topk_test: int = 1000
    rpn_nms: float = 0.65
    roi_batch: int = 512
    roi_pos_fraction: float = 0.25
    cascade_ious: Tuple[float, ...] = (0.5, 0.6, 0.7)
    cascade_weights: Tuple[Tuple[float, ...], ...] = ((10.0, 10.0, 5.0, 5.0), (20.0, 20.0, 10.0, 10.0),
                                                      (30.0, 30.0, 15.0, 15.0))
    box_pool: int = 7
    mask_pool: int = 14
    score_thresh: float = 0.05
    nms_test: float = 0.5
    dets_per_image: int = 100
    sem_ignore: int = 255
    sem_loss_weight: float = 0.5
    sem_common_stride: int = 4
    combine_overlap: float = 0.5
    combine_stuff_area: int = 4096
    combine_inst_thresh: float = 0.5
    bn_eps: float = 1e-5
    scale_clamp: float = field(default_factory=lambda: math.log(1000.0 / 16))


# ----------------------------------------------------------------------------------------------
# structures / box ops
# ----------------------------------------------------------------------------------------------
def pairwise_iou(b1, b2):
    """structures/boxes.py:336-358 (pairwise_intersection :310-333)."""
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh.clamp_(min=0)
    inter = wh.prod(dim=2)
    iou = torch.where(inter > 0, inter / (area1[:, None] + area2 - inter),
                      torch.zeros(1, dtype=inter.dtype, device=inter.device))
    return iou


def clip_boxes(b, image_size):
    """structures/boxes.py:183-197 Boxes.clip (in place)."""
    h, w = image_size
    x1 = b[:, 0].clamp(min=0, max=w)
    y1 = b[:, 1].clamp(min=0, max=h)
    x2 = b[:, 2].clamp(min=0, max=w)
    y2 = b[:, 3].clamp(min=0, max=h)
    return torch.stack((x1, y1, x2, y2), dim=-1)


def nonempty(b, threshold=0.0):
    """structures/boxes.py:199-213."""
    return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)


def matcher(iou, thresholds, labels, allow_low_quality):
    """modeling/matcher.py:62-127. thresholds are the inner ones; -inf/+inf are added as in :50-55."""
    thr = [-float("inf")] + list(thresholds) + [float("inf")]
    if iou.numel() == 0:
        return (iou.new_full((iou.size(1),), 0, dtype=torch.int64),
                iou.new_full((iou.size(1),), labels[0], dtype=torch.int8))
    matched_vals, matches = iou.max(dim=0)
    match_labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
    for l, low, high in zip(labels, thr[:-1], thr[1:]):
        match_labels[(matched_vals >= low) & (matched_vals < high)] = l
    if allow_low_quality:
        highest, _ = iou.max(dim=1)
        pred_inds = torch.nonzero(iou == highest[:, None], as_tuple=True)[1]
        match_labels[pred_inds] = 1
    return matches, match_labels


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    """modeling/sampling.py:9-54 — two torch.randperm calls on the labels' device, in this order."""
    positive = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    negative = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = torch.randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = torch.randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


def get_deltas(src, tgt, weights):
    """modeling/box_regression.py:43-76."""
    sw, sh = src[:, 2] - src[:, 0], src[:, 3] - src[:, 1]
    sx, sy = src[:, 0] + 0.5 * sw, src[:, 1] + 0.5 * sh
    tw, th = tgt[:, 2] - tgt[:, 0], tgt[:, 3] - tgt[:, 1]
    tx, ty = tgt[:, 0] + 0.5 * tw, tgt[:, 1] + 0.5 * th
    wx, wy, ww, wh = weights
    return torch.stack((wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)),
                       dim=1)


def apply_deltas(deltas, boxes, weights, scale_clamp):
    """modeling/box_regression.py:78-116."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    w, h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    cx, cy = boxes[:, 0] + 0.5 * w, boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
    dw, dh = deltas[:, 2::4] / ww, deltas[:, 3::4] / wh
    dw = torch.clamp(dw, max=scale_clamp)
    dh = torch.clamp(dh, max=scale_clamp)
    pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
    pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
    out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
    return out.reshape(deltas.shape)


def batched_nms(boxes, scores, idxs, thr):
    """layers/nms.py:9-21 -> torchvision.ops.boxes.batched_nms. Restated in its class-by-class
    ("vanilla", torchvision/ops/boxes.py::_batched_nms_vanilla) form: the coordinate-offset variant
    torchvision picks for small inputs differs only by fp32 rounding of the shifted coordinates."""
    boxes = boxes.float()
    keep_mask = torch.zeros_like(scores, dtype=torch.bool)
    for cid in torch.unique(idxs):
        cur = torch.where(idxs == cid)[0]
        keep_mask[cur[tv_nms(boxes[cur], scores[cur], thr)]] = True
    keep = torch.where(keep_mask)[0]
    return keep[scores[keep].sort(descending=True)[1]]


def assign_levels(boxes, min_level=2, max_level=5, canonical_size=224, canonical_level=4):
    """modeling/poolers.py:23-59 assign_boxes_to_levels (returns level - min_level)."""
    sizes = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lv = torch.floor(canonical_level + torch.log2(sizes / canonical_size + 1e-8))
    lv = torch.clamp(lv, min=min_level, max=max_level)
    return lv.to(torch.int64) - min_level


def roi_pool(feats, boxes_per_image, out_size, scales=(0.25, 0.125, 0.0625, 0.03125)):
    """modeling/poolers.py:206-263 ROIPooler.forward with ROIAlignV2 (layers/roi_align.py:49-65:
    torchvision roi_align, sampling_ratio=0, aligned=True)."""
    n = sum(len(b) for b in boxes_per_image)
    C = feats[0].shape[1]
    if n == 0:
        return feats[0].new_zeros((0, C, out_size, out_size))
    rois = torch.cat([torch.cat([b.new_full((len(b), 1), i), b], dim=1) for i, b in enumerate(boxes_per_image)])
    lv = assign_levels(rois[:, 1:])
    out = feats[0].new_zeros((n, C, out_size, out_size))
    for l, (f, s) in enumerate(zip(feats, scales)):
        inds = torch.nonzero(lv == l, as_tuple=True)[0]
        out.index_put_((inds,), tv_roi_align(f, rois[inds].to(f.dtype), (out_size, out_size), s, 0, True))
    return out


def crop_and_resize_masks(masks_bool, boxes, mask_size):
    """structures/masks.py:191-222 BitMasks.crop_and_resize."""
    M = len(boxes)
    rois = torch.cat([torch.arange(M, dtype=boxes.dtype)[:, None], boxes], dim=1)
    out = tv_roi_align(masks_bool.to(torch.float32)[:, None], rois, (mask_size, mask_size), 1.0, 0, True).squeeze(1)
    return out >= 0.5


def paste_masks_in_image(masks, boxes, image_shape, threshold=0.5):
    """layers/mask_ops.py:74-147 (+ _do_paste_mask :17-69): for every pixel centre of the output,
    bilinear-sample the MxM mask in box coordinates (grid_sample, align_corners=False, zeros)."""
    N = masks.shape[0]
    H, W = image_shape
    if N == 0:
        return masks.new_empty((0, H, W), dtype=torch.bool if threshold >= 0 else torch.uint8)
    x0, y0, x1, y1 = torch.split(boxes, 1, dim=1)
    img_y = torch.arange(0, H, dtype=torch.float32) + 0.5
    img_x = torch.arange(0, W, dtype=torch.float32) + 0.5
    img_y = (img_y - y0) / (y1 - y0) * 2 - 1
    img_x = (img_x - x0) / (x1 - x0) * 2 - 1
    gx = img_x[:, None, :].expand(N, H, W)
    gy = img_y[:, :, None].expand(N, H, W)
    grid = torch.stack([gx, gy], dim=3)
    img = F.grid_sample(masks[:, None].float(), grid, align_corners=False)[:, 0]
    return img >= threshold if threshold >= 0 else (img * 255).to(torch.uint8)


# ----------------------------------------------------------------------------------------------
# network pieces (functional, parameters by reference state_dict name)
# ----------------------------------------------------------------------------------------------
class Net:
    def __init__(self, params: Dict[str, torch.Tensor], cfg: DetCfg, training: bool, calibrate: bool = False):
        self.p, self.cfg, self.training, self.calibrate = params, cfg, training, calibrate

    def conv(self, x, name, stride=1, padding=0):
        return F.conv2d(x, self.p[name + ".weight"], self.p.get(name + ".bias"), stride=stride, padding=padding)

    def bn(self, x, name):
        """layers/batch_norm.py:187 nn.SyncBatchNorm; with one process = BatchNorm2d. Training uses batch
        statistics (running stats are not updated here: the oracle is stateless)."""
        p = self.p
        if self.calibrate:  # write this batch's statistics into the running buffers (fixture conditioning only)
            p[name + ".running_mean"] = x.mean((0, 2, 3)).detach()
            p[name + ".running_var"] = x.var((0, 2, 3), unbiased=False).detach()
            return F.batch_norm(x, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"],
                                p[name + ".bias"], training=False, eps=self.cfg.bn_eps)
        return F.batch_norm(x, p[name + ".running_mean"].clone(), p[name + ".running_var"].clone(),
                            p[name + ".weight"], p[name + ".bias"], training=self.training, momentum=0.1,
                            eps=self.cfg.bn_eps)

    def conv_bn(self, x, name, stride=1, padding=0, relu=False):
        x = self.bn(self.conv(x, name, stride, padding), name + ".norm")
        return F.relu_(x) if relu else x

    # backbone/resnet.py:330-359 BasicStem, :100-210 BottleneckBlock (stride in the 3x3), :435-458 ResNet.forward
    def resnet(self, x):
        x = self.conv_bn(x, "backbone.bottom_up.stem.conv1", 2, 3, relu=True)
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        outs = {}
        for stage, nblocks in (("res2", 3), ("res3", 4), ("res4", 6), ("res5", 3)):
            for i in range(nblocks):
                pre = "backbone.bottom_up.%s.%d" % (stage, i)
                stride = 2 if (i == 0 and stage != "res2") else 1
                out = self.conv_bn(x, pre + ".conv1", 1, 0, relu=True)
                out = self.conv_bn(out, pre + ".conv2", stride, 1, relu=True)
                out = self.conv_bn(out, pre + ".conv3", 1, 0)
                sc = self.conv_bn(x, pre + ".shortcut", stride, 0) if (pre + ".shortcut.weight") in self.p else x
                x = F.relu_(out + sc)
            outs[stage] = x
        return outs

    # backbone/fpn.py:126-167 FPN.forward, :188-200 LastLevelMaxPool
    def fpn(self, c):
        prev = self.conv_bn(c["res5"], "backbone.fpn_lateral5")
        res = {"p5": self.conv_bn(prev, "backbone.fpn_output5", 1, 1)}
        for lvl in (4, 3, 2):
            td = F.interpolate(prev, scale_factor=2.0, mode="nearest")
            prev = self.conv_bn(c["res%d" % lvl], "backbone.fpn_lateral%d" % lvl) + td
            res["p%d" % lvl] = self.conv_bn(prev, "backbone.fpn_output%d" % lvl, 1, 1)
        res["p6"] = F.max_pool2d(res["p5"], kernel_size=1, stride=2, padding=0)
        return res

    # meta_arch/semantic_seg.py:246-253 SemSegFPNHead.layers (ctor :188-215)
    def sem_seg_layers(self, f):
        total = None
        for name, nconv in (("p2", 1), ("p3", 1), ("p4", 2), ("p5", 3)):
            x = f[name]
            for k in range(nconv):
                pre = "sem_seg_head.%s.%d" % (name, 2 * k)
                x = self.conv(x, pre, 1, 1)
                x = F.relu_(F.group_norm(x, 32, self.p[pre + ".norm.weight"], self.p[pre + ".norm.bias"], 1e-5))
                if name != "p2":
                    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
            total = x if total is None else total + x
        return self.conv(total, "sem_seg_head.predictor")

    # proposal_generator/rpn.py:158-177 StandardRPNHead.forward
    def rpn_head(self, f):
        logits, deltas = [], []
        for name in ("p2", "p3", "p4", "p5", "p6"):
            t = F.relu_(self.conv(f[name], "proposal_generator.rpn_head.conv", 1, 1))
            logits.append(self.conv(t, "proposal_generator.rpn_head.objectness_logits"))
            deltas.append(self.conv(t, "proposal_generator.rpn_head.anchor_deltas"))
        return logits, deltas

    # roi_heads/box_head.py:94-97 + fast_rcnn.py:288-305
    def box_stage(self, x, k):
        x = x.flatten(1)
        x = F.relu(F.linear(x, self.p["roi_heads.box_head.%d.fc1.weight" % k], self.p["roi_heads.box_head.%d.fc1.bias" % k]))
        x = F.relu(F.linear(x, self.p["roi_heads.box_head.%d.fc2.weight" % k], self.p["roi_heads.box_head.%d.fc2.bias" % k]))
        scores = F.linear(x, self.p["roi_heads.box_predictor.%d.cls_score.weight" % k],
                          self.p["roi_heads.box_predictor.%d.cls_score.bias" % k])
        deltas = F.linear(x, self.p["roi_heads.box_predictor.%d.bbox_pred.weight" % k],
                          self.p["roi_heads.box_predictor.%d.bbox_pred.bias" % k])
        return scores, deltas

    # roi_heads/mask_head.py:287-290 (ctor :222-269)
    def mask_head(self, x):
        for k in range(1, 5):
            x = F.relu(self.conv(x, "roi_heads.mask_head.mask_fcn%d" % k, 1, 1))
        x = F.relu(F.conv_transpose2d(x, self.p["roi_heads.mask_head.deconv.weight"],
                                      self.p["roi_heads.mask_head.deconv.bias"], stride=2))
        return self.conv(x, "roi_heads.mask_head.predictor")


def preprocess(images: List[torch.Tensor], cfg: DetCfg, pad_value=0.0):
    """meta_arch/rcnn.py:223-234 + structures/image_list.py:59-129."""
    mean = torch.tensor(cfg.pixel_mean).view(-1, 1, 1)
    std = torch.tensor(cfg.pixel_std).view(-1, 1, 1)
    imgs = [(x.float() - mean) / std for x in images]
    sizes = [(int(im.shape[-2]), int(im.shape[-1])) for im in imgs]
    s = cfg.size_divisibility
    mh = (max(h for h, _ in sizes) + s - 1) // s * s
    mw = (max(w for _, w in sizes) + s - 1) // s * s
    out = imgs[0].new_full((len(imgs), imgs[0].shape[0], mh, mw), pad_value)
    for i, im in enumerate(imgs):
        out[i, :, :im.shape[-2], :im.shape[-1]].copy_(im)
    return out, sizes


def pad_sem_seg(gts: List[torch.Tensor], cfg: DetCfg):
    """meta_arch/panoptic_fpn.py:120-126."""
    s = cfg.size_divisibility
    mh = (max(int(g.shape[-2]) for g in gts) + s - 1) // s * s
    mw = (max(int(g.shape[-1]) for g in gts) + s - 1) // s * s
    out = gts[0].new_full((len(gts), mh, mw), cfg.sem_ignore)
    for i, g in enumerate(gts):
        out[i, :g.shape[-2], :g.shape[-1]].copy_(g)
    return out


def make_anchors(feat_shapes, cfg: DetCfg):
    """modeling/anchor_generator.py:165-231 (offset 0)."""
    out = []
    for (h, w), size, stride in zip(feat_shapes, cfg.anchor_sizes, (4, 8, 16, 32, 64)):
        cell = []
        for ar in cfg.anchor_ratios:
            area = size ** 2.0
            ww = math.sqrt(area / ar)
            hh = ar * ww
            cell.append([-ww / 2.0, -hh / 2.0, ww / 2.0, hh / 2.0])
        cell = torch.tensor(cell)
        sx = torch.arange(0, w * stride, step=stride, dtype=torch.float32)
        sy = torch.arange(0, h * stride, step=stride, dtype=torch.float32)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
        out.append((shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4))
    return out


def rpn_forward(net: Net, feats, image_sizes, gt_boxes_list):
    """proposal_generator/rpn.py:431-480 RPN.forward."""
    cfg = net.cfg
    names = ("p2", "p3", "p4", "p5", "p6")
    anchors = make_anchors([feats[n].shape[-2:] for n in names], cfg)
    logits, deltas = net.rpn_head(feats)
    N = logits[0].shape[0]
    logits = [x.permute(0, 2, 3, 1).flatten(1) for x in logits]
    deltas = [x.view(N, -1, 4, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2) for x in deltas]
    losses = {}
    if net.training:
        # rpn.py:307-363 label_and_sample_anchors
        all_anchors = torch.cat(anchors)
        gt_labels, matched_gt = [], []
        with torch.no_grad():
            for gtb in gt_boxes_list:
                iou = pairwise_iou(gtb, all_anchors)
                midx, lab = matcher(iou, cfg.rpn_iou_thresholds, (0, -1, 1), True)
                pos, neg = subsample_labels(lab, cfg.rpn_batch, cfg.rpn_pos_fraction, 0)
                lab.fill_(-1)
                lab.scatter_(0, pos, 1)
                lab.scatter_(0, neg, 0)
                matched_gt.append(torch.zeros_like(all_anchors) if len(gtb) == 0 else gtb[midx])
                gt_labels.append(lab)
        # rpn.py:366-429 losses (+ box_regression.py:310-345, smooth_l1 with beta 0 == L1)
        gl = torch.stack(gt_labels)
        pos_mask = gl == 1
        gt_deltas = torch.stack([get_deltas(all_anchors, k, (1.0, 1.0, 1.0, 1.0)) for k in matched_gt])
        loc = torch.abs(torch.cat(deltas, dim=1)[pos_mask] - gt_deltas[pos_mask]).sum()
        valid = gl >= 0
        obj = F.binary_cross_entropy_with_logits(torch.cat(logits, dim=1)[valid], gl[valid].to(torch.float32),
                                                 reduction="sum")
        norm = cfg.rpn_batch * N
        losses = {"loss_rpn_cls": obj / norm, "loss_rpn_loc": loc / norm}
    # rpn.py:482-533 predict_proposals + proposal_utils.py:22-135 find_top_rpn_proposals
    with torch.no_grad():
        pre = cfg.rpn_pre_topk_train if net.training else cfg.rpn_pre_topk_test
        post = cfg.rpn_post_topk_train if net.training else cfg.rpn_post_topk_test
        tk_scores, tk_boxes, lvl_ids = [], [], []
        bidx = torch.arange(N)
        for lid, (a, lg, dl) in enumerate(zip(anchors, logits, deltas)):
            props = apply_deltas(dl.reshape(-1, 4), a.unsqueeze(0).expand(N, -1, -1).reshape(-1, 4),
                                 (1.0, 1.0, 1.0, 1.0), cfg.scale_clamp).view(N, -1, 4)
            k = min(lg.shape[1], pre)
            sc, idx = lg.topk(k, dim=1)
            tk_scores.append(sc)
            tk_boxes.append(props[bidx[:, None], idx])
            lvl_ids.append(torch.full((k,), lid, dtype=torch.int64))
        tk_scores, tk_boxes, lvl_ids = torch.cat(tk_scores, 1), torch.cat(tk_boxes, 1), torch.cat(lvl_ids)
        proposals = []
        for n, isz in enumerate(image_sizes):
            boxes, sc, lv = tk_boxes[n], tk_scores[n], lvl_ids
            valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(sc)
            if not valid.all():
                if net.training:
                    raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
                boxes, sc, lv = boxes[valid], sc[valid], lv[valid]
            boxes = clip_boxes(boxes, isz)
            keep = nonempty(boxes, 0.0)
            if keep.sum().item() != len(boxes):
                boxes, sc, lv = boxes[keep], sc[keep], lv[keep]
            keep = batched_nms(boxes, sc, lv, cfg.rpn_nms)[:post]
            proposals.append((boxes[keep], sc[keep]))
    return proposals, losses


def sample_proposals_for_roi_heads(cfg: DetCfg, proposals, gt_boxes_list, gt_classes_list):
    """roi_heads/roi_heads.py:220-302 (+ proposal_utils.py:138-205, roi_heads.py:181-217)."""
    out = []
    for (pb, _), gtb, gtc in zip(proposals, gt_boxes_list, gt_classes_list):
        boxes = torch.cat([pb, gtb])                       # proposals first, then GT
        iou = pairwise_iou(gtb, boxes)
        midx, mlab = matcher(iou, (0.5,), (0, 1), False)
        if len(gtb) > 0:
            cls = gtc[midx]
            cls[mlab == 0] = cfg.num_classes
            cls[mlab == -1] = -1
        else:
            cls = torch.zeros_like(midx) + cfg.num_classes
        fg, bg = subsample_labels(cls, cfg.roi_batch, cfg.roi_pos_fraction, cfg.num_classes)
        sidx = torch.cat([fg, bg], dim=0)
        d = {"proposal_boxes": boxes[sidx], "gt_classes": cls[sidx]}
        if len(gtb) > 0:
            d["matched"] = midx[sidx]
            d["gt_boxes"] = gtb[midx[sidx]]
        out.append(d)
    return out


class _ScaleGradient(torch.autograd.Function):
    """roi_heads/cascade_rcnn.py:20-28: identity in forward, gradient x scale in backward."""

    @staticmethod
    def forward(ctx, input, scale):
        ctx.scale = scale
        return input

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output * ctx.scale, None


def l1_box_loss(cfg, proposal_boxes, gt_boxes, pred_deltas, gt_classes, weights):
    """roi_heads/fast_rcnn.py:424-463 (cls-agnostic) with smooth_l1_beta 0."""
    fg = torch.nonzero((gt_classes >= 0) & (gt_classes < cfg.num_classes), as_tuple=True)[0]
    tgt = get_deltas(proposal_boxes[fg], gt_boxes[fg], weights)
    loss = torch.abs(pred_deltas[fg] - tgt).sum()
    return loss / max(gt_classes.numel(), 1.0)


def roi_heads_train(net: Net, feats, image_sizes, proposals, gt_boxes_list, gt_classes_list, gt_masks_list):
    """roi_heads/cascade_rcnn.py:137-150,153-185 (training branch) + roi_heads.py:818-846 + mask_head.py:33-112."""
    cfg = net.cfg
    with torch.no_grad():
        samples = sample_proposals_for_roi_heads(cfg, proposals, gt_boxes_list, gt_classes_list)
    flist = [feats[n] for n in ("p2", "p3", "p4", "p5")]
    losses = {}
    cur = [dict(s) for s in samples]
    for k in range(3):
        if k > 0:
            with torch.no_grad():
                nxt = []
                for b, isz, gtb, gtc in zip(prev_boxes, image_sizes, gt_boxes_list, gt_classes_list):
                    b = clip_boxes(b.detach(), isz)
                    b = b[nonempty(b)]
                    iou = pairwise_iou(gtb, b)
                    midx, lab = matcher(iou, (cfg.cascade_ious[k],), (0, 1), False)
                    if len(gtb) > 0:
                        cls = gtc[midx]
                        cls[lab == 0] = cfg.num_classes
                        g = gtb[midx]
                    else:
                        cls = torch.zeros_like(midx) + cfg.num_classes
                        g = gtb.new_zeros((len(b), 4))
                    nxt.append({"proposal_boxes": b, "gt_classes": cls, "gt_boxes": g})
                cur = nxt
        x = roi_pool(flist, [c["proposal_boxes"] for c in cur], cfg.box_pool)
        x = _ScaleGradient.apply(x, 1.0 / 3)                     # cascade_rcnn.py:271-272 (training only)
        scores, deltas = net.box_stage(x, k)
        pb = torch.cat([c["proposal_boxes"] for c in cur])
        gcls = torch.cat([c["gt_classes"] for c in cur])
        gbx = torch.cat([c.get("gt_boxes", c["proposal_boxes"]) for c in cur])
        losses["loss_cls_stage%d" % k] = F.cross_entropy(scores, gcls, reduction="mean")
        losses["loss_box_reg_stage%d" % k] = l1_box_loss(cfg, pb, gbx, deltas, gcls, cfg.cascade_weights[k])
        pred = apply_deltas(deltas, pb, cfg.cascade_weights[k], cfg.scale_clamp)
        prev_boxes = pred.split([len(c["proposal_boxes"]) for c in cur])
    # mask branch on the stage-0 samples (select_foreground_proposals roi_heads.py:46-75)
    fg_boxes, fg_cls, gt_m = [], [], []
    for s, masks in zip(samples, gt_masks_list):
        sel = (s["gt_classes"] != -1) & (s["gt_classes"] != cfg.num_classes)
        idx = sel.nonzero().squeeze(1)
        fg_boxes.append(s["proposal_boxes"][idx])
        fg_cls.append(s["gt_classes"][idx])
        if len(idx):
            gt_m.append(crop_and_resize_masks(masks[s["matched"][idx]], s["proposal_boxes"][idx], 2 * cfg.mask_pool))
    xm = roi_pool(flist, fg_boxes, cfg.mask_pool)
    mlog = net.mask_head(xm)
    if len(gt_m) == 0:
        losses["loss_mask"] = mlog.sum() * 0
    else:
        gtm = torch.cat(gt_m, dim=0)
        gc = torch.cat(fg_cls, dim=0)
        sel_logits = mlog[torch.arange(mlog.shape[0]), gc]
        losses["loss_mask"] = F.binary_cross_entropy_with_logits(sel_logits, gtm.to(torch.float32), reduction="mean")
    return losses


def forward_train(params, cfg: DetCfg, images, gt_boxes_list, gt_classes_list, gt_masks_list, sem_seg_list):
    """meta_arch/panoptic_fpn.py:90-138. Returns the ten losses in the reference's key order."""
    net = Net(params, cfg, True)
    x, sizes = preprocess(images, cfg)
    feats = net.fpn(net.resnet(x))
    gt_sem = pad_sem_seg(sem_seg_list, cfg)
    logits = net.sem_seg_layers(feats).float()
    logits = F.interpolate(logits, scale_factor=cfg.sem_common_stride, mode="bilinear", align_corners=False)
    losses = {"loss_sem_seg": F.cross_entropy(logits, gt_sem, reduction="mean", ignore_index=cfg.sem_ignore)
              * cfg.sem_loss_weight}
    proposals, rl = rpn_forward(net, feats, sizes, gt_boxes_list)
    losses.update(rl)
    losses.update(roi_heads_train(net, feats, sizes, proposals, gt_boxes_list, gt_classes_list, gt_masks_list))
    return losses


def combine_semantic_and_instance_outputs(inst, semantic, overlap_threshold, stuff_area_thresh, inst_thresh):
    """meta_arch/panoptic_fpn.py:184-269."""
    panoptic = torch.zeros_like(semantic, dtype=torch.int32)
    order = torch.argsort(-inst["scores"])
    cur = 0
    info = []
    masks = inst["pred_masks"].to(dtype=torch.bool)
    for i in order:
        score = inst["scores"][i].item()
        if score < inst_thresh:
            break
        m = masks[i]
        area = m.sum().item()
        if area == 0:
            continue
        inter = (m > 0) & (panoptic > 0)
        if inter.sum().item() * 1.0 / area > overlap_threshold:
            continue
        if inter.sum().item() > 0:
            m = m & (panoptic == 0)
        cur += 1
        panoptic[m] = cur
        info.append({"id": cur, "isthing": True, "score": score, "category_id": inst["pred_classes"][i].item(),
                     "instance_id": i.item()})
    for lab in torch.unique(semantic).cpu().tolist():
        if lab == 0:
            continue
        m = (semantic == lab) & (panoptic == 0)
        area = m.sum().item()
        if area < stuff_area_thresh:
            continue
        cur += 1
        panoptic[m] = cur
        info.append({"id": cur, "isthing": False, "category_id": lab, "area": area})
    return panoptic, info


@torch.no_grad()
def forward_inference(params, cfg: DetCfg, images, out_sizes=None):
    """meta_arch/panoptic_fpn.py:140-181 + cascade_rcnn.py:186-206 + fast_rcnn.py:118-171 +
    mask_head.py:115-158 + modeling/postprocessing.py:9-100."""
    net = Net(params, cfg, False)
    x, sizes = preprocess(images, cfg)
    feats = net.fpn(net.resnet(x))
    sem = F.interpolate(net.sem_seg_layers(feats), scale_factor=cfg.sem_common_stride, mode="bilinear",
                        align_corners=False)
    proposals, _ = rpn_forward(net, feats, sizes, None)
    flist = [feats[n] for n in ("p2", "p3", "p4", "p5")]
    cur = [p[0] for p in proposals]
    probs = []
    for k in range(3):
        if k > 0:
            cur = [clip_boxes(b, isz) for b, isz in zip(prev_boxes, sizes)]
        xk = roi_pool(flist, cur, cfg.box_pool)
        scores, deltas = net.box_stage(xk, k)
        pb = torch.cat(cur)
        probs.append(F.softmax(scores, dim=-1).split([len(c) for c in cur]))
        prev_boxes = apply_deltas(deltas, pb, cfg.cascade_weights[k], cfg.scale_clamp).split([len(c) for c in cur])
    results = []
    for n, isz in enumerate(sizes):
        sc = sum(p[n] for p in probs) * (1.0 / 3)
        bx = prev_boxes[n]
        valid = torch.isfinite(bx).all(dim=1) & torch.isfinite(sc).all(dim=1)
        bx, sc = bx[valid], sc[valid]
        sc = sc[:, :-1]
        bx = clip_boxes(bx.reshape(-1, 4), isz).view(-1, 1, 4)
        fmask = sc > cfg.score_thresh
        finds = fmask.nonzero()
        b = bx[finds[:, 0], 0]
        s = sc[fmask]
        keep = batched_nms(b, s, finds[:, 1], cfg.nms_test)[:cfg.dets_per_image]
        results.append({"pred_boxes": b[keep], "scores": s[keep], "pred_classes": finds[keep][:, 1]})
    xm = roi_pool(flist, [r["pred_boxes"] for r in results], cfg.mask_pool)
    mlog = net.mask_head(xm)
    cls = torch.cat([r["pred_classes"] for r in results])
    mprob = mlog[torch.arange(mlog.shape[0]), cls][:, None].sigmoid().split([len(r["pred_boxes"]) for r in results])
    out = []
    for n, (r, isz) in enumerate(zip(results, sizes)):
        oh, ow = out_sizes[n] if out_sizes is not None else isz
        sem_r = F.interpolate(sem[n][:, :isz[0], :isz[1]].expand(1, -1, -1, -1), size=(oh, ow), mode="bilinear",
                              align_corners=False)[0]
        sx, sy = ow / isz[1], oh / isz[0]
        b = r["pred_boxes"].clone()
        b[:, 0::2] *= sx
        b[:, 1::2] *= sy
        b = clip_boxes(b, (oh, ow))
        ne = nonempty(b)
        inst = {"pred_boxes": b[ne], "scores": r["scores"][ne], "pred_classes": r["pred_classes"][ne],
                "mask_probs": mprob[n][ne]}
        inst["pred_masks"] = paste_masks_in_image(inst["mask_probs"][:, 0], inst["pred_boxes"], (oh, ow), 0.5)
        pan, info = combine_semantic_and_instance_outputs(inst, sem_r.argmax(dim=0), cfg.combine_overlap,
                                                          cfg.combine_stuff_area, cfg.combine_inst_thresh)
        out.append({"instances": inst, "sem_seg": sem_r, "panoptic_seg": (pan, info)})
    return out


@torch.no_grad()
def eval_fixture_params(cfg: DetCfg, images, seed=0, cls_gain=60.0):
    """Well-conditioned random parameters for inference tests: BN running statistics calibrated on
    `images` (random running stats make eval-mode activations explode) and a larger cls_score gain so
    that some detections pass the 0.05 score threshold. Deterministic given (cfg, images, seed)."""
    p = init_params(cfg, seed)
    net = Net(p, cfg, False, calibrate=True)
    x, _ = preprocess(images, cfg)
    net.fpn(net.resnet(x))
    for k in range(3):
        p["roi_heads.box_predictor.%d.cls_score.weight" % k] *= cls_gain
    return p


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synthetic_batch(n_images, H, W, num_classes, sem_classes, seed, G=20, min_size=32, max_size=512):
    g = torch.Generator().manual_seed(seed)
    images, boxes, classes, masks, sems = [], [], [], [], []
    ys = torch.arange(H, dtype=torch.float32)[:, None] + 0.5
    xs = torch.arange(W, dtype=torch.float32)[None, :] + 0.5
    for _ in range(n_images):
        images.append(torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8))
        cx = torch.rand(G, generator=g) * W
        cy = torch.rand(G, generator=g) * H
        lo, hi = math.log(min(min_size, W / 4)), math.log(min(max_size, W / 2))
        bw = torch.exp(torch.rand(G, generator=g) * (hi - lo) + lo)
        bh = torch.exp(torch.rand(G, generator=g) * (hi - lo) + lo)
        x0 = (cx - bw / 2).clamp(0, W - 8)
        y0 = (cy - bh / 2).clamp(0, H - 8)
        x1 = torch.maximum((cx + bw / 2).clamp(0, W), x0 + 8)
        y1 = torch.maximum((cy + bh / 2).clamp(0, H), y0 + 8)
        b = torch.stack([x0, y0, x1, y1], dim=1)
        boxes.append(b)
        classes.append(torch.randint(0, num_classes, (G,), generator=g))
        ecx, ecy = (b[:, 0] + b[:, 2]) / 2, (b[:, 1] + b[:, 3]) / 2
        rx, ry = (b[:, 2] - b[:, 0]) / 2, (b[:, 3] - b[:, 1]) / 2
        m = (((xs[None] - ecx[:, None, None]) / rx[:, None, None]) ** 2
             + ((ys[None] - ecy[:, None, None]) / ry[:, None, None]) ** 2) <= 1.0
        masks.append(m)
        blk = 64 if H >= 256 else 16
        coarse = torch.randint(0, sem_classes, ((H + blk - 1) // blk, (W + blk - 1) // blk), generator=g)
        sem = coarse.repeat_interleave(blk, 0).repeat_interleave(blk, 1)[:H, :W].clone()
        sem[torch.rand(H, W, generator=g) < 0.05] = 255
        sems.append(sem.long())
    return images, boxes, classes, masks, sems


def init_params(cfg: DetCfg, seed=0):
    """Random-init parameters with the reference's names/shapes (scaled normal; the exact reference
    initialisers are irrelevant for parity: both sides load the same state_dict)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def conv(name, cout, cin, k, bias):
        fan = cin * k * k
        p[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan)
        if bias:
            p[name + ".bias"] = torch.randn(cout, generator=g) * 0.01

    def bn(name, c):
        p[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        p[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
        p[name + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
        p[name + ".running_var"] = 1.0 + 0.1 * torch.rand(c, generator=g)
        p[name + ".num_batches_tracked"] = torch.tensor(0)

    conv("backbone.bottom_up.stem.conv1", 64, 3, 7, False)
    bn("backbone.bottom_up.stem.conv1.norm", 64)
    cin = 64
    for stage, nb, mid, cout in (("res2", 3, 64, 256), ("res3", 4, 128, 512), ("res4", 6, 256, 1024), ("res5", 3, 512, 2048)):
        for i in range(nb):
            pre = "backbone.bottom_up.%s.%d" % (stage, i)
            if cin != cout:
                conv(pre + ".shortcut", cout, cin, 1, False)
                bn(pre + ".shortcut.norm", cout)
            conv(pre + ".conv1", mid, cin, 1, False)
            bn(pre + ".conv1.norm", mid)
            conv(pre + ".conv2", mid, mid, 3, False)
            bn(pre + ".conv2.norm", mid)
            conv(pre + ".conv3", cout, mid, 1, False)
            bn(pre + ".conv3.norm", cout)
            cin = cout
    for lvl, c in ((2, 256), (3, 512), (4, 1024), (5, 2048)):
        conv("backbone.fpn_lateral%d" % lvl, 256, c, 1, False)
        bn("backbone.fpn_lateral%d.norm" % lvl, 256)
        conv("backbone.fpn_output%d" % lvl, 256, 256, 3, False)
        bn("backbone.fpn_output%d.norm" % lvl, 256)
    conv("proposal_generator.rpn_head.conv", 256, 256, 3, True)
    conv("proposal_generator.rpn_head.objectness_logits", 3, 256, 1, True)
    conv("proposal_generator.rpn_head.anchor_deltas", 12, 256, 1, True)
    p["proposal_generator.rpn_head.objectness_logits.weight"] *= 0.05
    p["proposal_generator.rpn_head.anchor_deltas.weight"] *= 0.05
    for k in range(3):
        for nm, o, i in (("fc1", 1024, 256 * 49), ("fc2", 1024, 1024)):
            p["roi_heads.box_head.%d.%s.weight" % (k, nm)] = torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)
            p["roi_heads.box_head.%d.%s.bias" % (k, nm)] = torch.randn(o, generator=g) * 0.01
        p["roi_heads.box_predictor.%d.cls_score.weight" % k] = torch.randn(cfg.num_classes + 1, 1024, generator=g) * 0.01
        p["roi_heads.box_predictor.%d.cls_score.bias" % k] = torch.zeros(cfg.num_classes + 1)
        p["roi_heads.box_predictor.%d.bbox_pred.weight" % k] = torch.randn(4, 1024, generator=g) * 0.001
        p["roi_heads.box_predictor.%d.bbox_pred.bias" % k] = torch.zeros(4)
    for k in range(1, 5):
        conv("roi_heads.mask_head.mask_fcn%d" % k, 256, 256, 3, True)
    p["roi_heads.mask_head.deconv.weight"] = torch.randn(256, 256, 2, 2, generator=g) * math.sqrt(2.0 / 1024)
    p["roi_heads.mask_head.deconv.bias"] = torch.randn(256, generator=g) * 0.01
    conv("roi_heads.mask_head.predictor", cfg.num_classes, 256, 1, True)
    p["roi_heads.mask_head.predictor.weight"] *= 0.05
    for name, idxs, in (("p2", (0,)), ("p3", (0,)), ("p4", (0, 2)), ("p5", (0, 2, 4))):
        for j, ix in enumerate(idxs):
            pre = "sem_seg_head.%s.%d" % (name, ix)
            conv(pre, 128, 256 if j == 0 else 128, 3, False)
            p[pre + ".norm.weight"] = 1.0 + 0.1 * torch.randn(128, generator=g)
            p[pre + ".norm.bias"] = 0.1 * torch.randn(128, generator=g)
    conv("sem_seg_head.predictor", cfg.sem_classes, 128, 1, True)
    return p
