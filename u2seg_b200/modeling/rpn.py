"""RPN — mirror of detectron2/modeling/{anchor_generator.py, box_regression.py, sampling.py,
proposal_generator/rpn.py, proposal_generator/proposal_utils.py}. Same module tree / state_dict names
(proposal_generator.rpn_head.{conv,objectness_logits,anchor_deltas}.*), same config keys.

B200-side differences (results identical): anchors are cached per feature-map shape; anchor labelling
uses the fused IoU+Matcher kernel (no G x 261,888 matrix); NMS runs on the device incl. its scan.
"""
import math
from typing import Dict, List

import torch
import torch.nn.functional as F
from torch import nn

from ..layers import Matcher, batched_nms
from ..registry import ANCHOR_GENERATOR_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, RPN_HEAD_REGISTRY
from ..structures import Boxes, Instances
from .backbone import Conv2d

_DEFAULT_SCALE_CLAMP = math.log(1000.0 / 16)

# torch.randperm indirection: parity tests swap in a CPU-generator version to reproduce the oracle's sampling
_randperm = torch.randperm


class Box2BoxTransform:
    """box_regression.py:21-116."""

    def __init__(self, weights, scale_clamp=_DEFAULT_SCALE_CLAMP):
        self.weights, self.scale_clamp = tuple(weights), scale_clamp

    def get_deltas(self, src_boxes, target_boxes):
        sw, sh = src_boxes[:, 2] - src_boxes[:, 0], src_boxes[:, 3] - src_boxes[:, 1]
        sx, sy = src_boxes[:, 0] + 0.5 * sw, src_boxes[:, 1] + 0.5 * sh
        tw, th = target_boxes[:, 2] - target_boxes[:, 0], target_boxes[:, 3] - target_boxes[:, 1]
        tx, ty = target_boxes[:, 0] + 0.5 * tw, target_boxes[:, 1] + 0.5 * th
        wx, wy, ww, wh = self.weights
        return torch.stack((wx * (tx - sx) / sw, wy * (ty - sy) / sh, ww * torch.log(tw / sw), wh * torch.log(th / sh)),
                           dim=1)

    def apply_deltas(self, deltas, boxes):
        deltas = deltas.float()
        boxes = boxes.to(deltas.dtype)
        w, h = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        cx, cy = boxes[:, 0] + 0.5 * w, boxes[:, 1] + 0.5 * h
        wx, wy, ww, wh = self.weights
        dx, dy = deltas[:, 0::4] / wx, deltas[:, 1::4] / wy
        dw = torch.clamp(deltas[:, 2::4] / ww, max=self.scale_clamp)
        dh = torch.clamp(deltas[:, 3::4] / wh, max=self.scale_clamp)
        pcx, pcy = dx * w[:, None] + cx[:, None], dy * h[:, None] + cy[:, None]
        pw, ph = torch.exp(dw) * w[:, None], torch.exp(dh) * h[:, None]
        out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
        return out.reshape(deltas.shape)


def subsample_labels(labels, num_samples, positive_fraction, bg_label):
    """sampling.py:9-54."""
    positive = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    negative = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = _randperm(positive.numel(), device=positive.device)[:num_pos]
    perm2 = _randperm(negative.numel(), device=negative.device)[:num_neg]
    return positive[perm1], negative[perm2]


@ANCHOR_GENERATOR_REGISTRY.register()
class DefaultAnchorGenerator(nn.Module):
    """anchor_generator.py:86-231. forward(features) -> list[Boxes]; grids cached per (shape, device)."""
    box_dim = 4

    def __init__(self, cfg=None, input_shape=None, *, sizes=None, aspect_ratios=None, strides=None, offset=0.0):
        super().__init__()
        if cfg is not None:
            sizes, aspect_ratios = cfg.MODEL.ANCHOR_GENERATOR.SIZES, cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS
            strides, offset = [x.stride for x in input_shape], cfg.MODEL.ANCHOR_GENERATOR.OFFSET
        self.strides = list(strides)
        n = len(self.strides)
        sizes = list(sizes) * n if len(sizes) == 1 else list(sizes)
        aspect_ratios = list(aspect_ratios) * n if len(aspect_ratios) == 1 else list(aspect_ratios)
        assert len(sizes) == n and len(aspect_ratios) == n
        self.cell_anchors = [self.generate_cell_anchors(s, a) for s, a in zip(sizes, aspect_ratios)]
        self.offset = offset
        self._cache = {}

    @property
    def num_anchors(self):
        return [len(c) for c in self.cell_anchors]

    @staticmethod
    def generate_cell_anchors(sizes, aspect_ratios):
        anchors = []
        for size in sizes:
            area = size ** 2.0
            for ar in aspect_ratios:
                w = math.sqrt(area / ar)
                h = ar * w
                anchors.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        return torch.tensor(anchors)

    def forward(self, features: List[torch.Tensor]):
        key = (tuple(tuple(f.shape[-2:]) for f in features), features[0].device)
        if key not in self._cache:
            out = []
            for f, stride, base in zip(features, self.strides, self.cell_anchors):
                h, w = f.shape[-2:]
                dev = f.device
                sx = torch.arange(self.offset * stride, w * stride, step=stride, dtype=torch.float32, device=dev)
                sy = torch.arange(self.offset * stride, h * stride, step=stride, dtype=torch.float32, device=dev)
                yy, xx = torch.meshgrid(sy, sx, indexing="ij")
                xx, yy = xx.reshape(-1), yy.reshape(-1)
                shifts = torch.stack((xx, yy, xx, yy), dim=1)
                out.append(Boxes((shifts.view(-1, 1, 4) + base.to(dev).view(1, -1, 4)).reshape(-1, 4)))
            self._cache[key] = out
        return self._cache[key]


@RPN_HEAD_REGISTRY.register()
class StandardRPNHead(nn.Module):
    """rpn.py:67-177."""

    def __init__(self, cfg=None, input_shape=None, *, in_channels=None, num_anchors=None, box_dim=4):
        super().__init__()
        if cfg is not None:
            in_channels = input_shape[0].channels
            ag = ANCHOR_GENERATOR_REGISTRY.get(cfg.MODEL.ANCHOR_GENERATOR.NAME)(cfg, input_shape)
            num_anchors, box_dim = ag.num_anchors[0], ag.box_dim
            assert cfg.MODEL.RPN.CONV_DIMS == [-1] or tuple(cfg.MODEL.RPN.CONV_DIMS) == (-1,)
        self.conv = Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1, activation=F.relu_)
        self.objectness_logits = Conv2d(in_channels, num_anchors, kernel_size=1, stride=1)
        self.anchor_deltas = Conv2d(in_channels, num_anchors * box_dim, kernel_size=1, stride=1)
        for layer in (self.conv, self.objectness_logits, self.anchor_deltas):
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)

    def forward(self, features: List[torch.Tensor]):
        logits, deltas = [], []
        for x in features:
            t = self.conv(x)
            logits.append(self.objectness_logits(t))
            deltas.append(self.anchor_deltas(t))
        return logits, deltas


def find_top_rpn_proposals(proposals, pred_objectness_logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_size, training):
    """proposal_utils.py:22-135."""
    num_images = len(image_sizes)
    device = proposals[0].device
    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(num_images, device=device)
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, pred_objectness_logits)):
        k = min(logits_i.shape[1], pre_nms_topk)
        sc, idx = logits_i.topk(k, dim=1)
        topk_proposals.append(proposals_i[batch_idx[:, None], idx])
        topk_scores.append(sc)
        level_ids.append(torch.full((k,), level_id, dtype=torch.int64, device=device))
    topk_scores, topk_proposals, level_ids = torch.cat(topk_scores, 1), torch.cat(topk_proposals, 1), torch.cat(level_ids)
    # one fused validity read for the whole batch instead of one per image (proposal_utils.py:103-118)
    finite = torch.isfinite(topk_proposals).all(dim=2) & torch.isfinite(topk_scores)
    if training and not bool(finite.all()):
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    results = []
    for n, image_size in enumerate(image_sizes):
        boxes, scores, lvl = Boxes(topk_proposals[n]), topk_scores[n], level_ids
        if not training:
            v = finite[n]
            boxes, scores, lvl = boxes[v], scores[v], lvl[v]
        boxes.clip(image_size)
        keep = boxes.nonempty(threshold=min_box_size)
        boxes, scores, lvl = boxes[keep], scores[keep], lvl[keep]     # no-op when all are kept
        keep = batched_nms(boxes.tensor, scores, lvl, nms_thresh, max_keep=post_nms_topk)
        res = Instances(image_size)
        res.proposal_boxes = boxes[keep]
        res.objectness_logits = scores[keep]
        results.append(res)
    return results


@PROPOSAL_GENERATOR_REGISTRY.register()
class RPN(nn.Module):
    """rpn.py:181-533."""

    def __init__(self, cfg, input_shape: Dict[str, "ShapeSpec"]):
        super().__init__()
        c = cfg.MODEL.RPN
        self.in_features = list(c.IN_FEATURES)
        shapes = [input_shape[f] for f in self.in_features]
        self.rpn_head = RPN_HEAD_REGISTRY.get(c.HEAD_NAME)(cfg, shapes)
        self.anchor_generator = ANCHOR_GENERATOR_REGISTRY.get(cfg.MODEL.ANCHOR_GENERATOR.NAME)(cfg, shapes)
        self.anchor_matcher = Matcher(c.IOU_THRESHOLDS, c.IOU_LABELS, allow_low_quality_matches=True)
        self.box2box_transform = Box2BoxTransform(weights=c.BBOX_REG_WEIGHTS)
        self.batch_size_per_image, self.positive_fraction = c.BATCH_SIZE_PER_IMAGE, c.POSITIVE_FRACTION
        self.pre_nms_topk = {True: c.PRE_NMS_TOPK_TRAIN, False: c.PRE_NMS_TOPK_TEST}
        self.post_nms_topk = {True: c.POST_NMS_TOPK_TRAIN, False: c.POST_NMS_TOPK_TEST}
        self.nms_thresh, self.min_box_size = c.NMS_THRESH, float(cfg.MODEL.PROPOSAL_GENERATOR.MIN_SIZE)
        self.anchor_boundary_thresh = c.BOUNDARY_THRESH
        assert self.anchor_boundary_thresh < 0
        self.loss_weight = {"loss_rpn_cls": c.LOSS_WEIGHT, "loss_rpn_loc": c.BBOX_REG_LOSS_WEIGHT * c.LOSS_WEIGHT}
        assert c.BBOX_REG_LOSS_TYPE == "smooth_l1"
        self.smooth_l1_beta = c.SMOOTH_L1_BETA

    @torch.no_grad()
    def label_and_sample_anchors(self, anchors: List[Boxes], gt_instances: List[Instances]):
        """rpn.py:307-363."""
        anchors_t = Boxes.cat(anchors).tensor
        gt_labels, matched_gt_boxes = [], []
        for inst in gt_instances:
            gt = inst.gt_boxes.tensor
            matched_idxs, labels = self.anchor_matcher.match_boxes(gt, anchors_t)
            pos_idx, neg_idx = subsample_labels(labels, self.batch_size_per_image, self.positive_fraction, 0)
            labels.fill_(-1)
            labels.scatter_(0, pos_idx, 1)
            labels.scatter_(0, neg_idx, 0)
            matched_gt_boxes.append(torch.zeros_like(anchors_t) if len(gt) == 0 else gt[matched_idxs])
            gt_labels.append(labels)
        return gt_labels, matched_gt_boxes

    def losses(self, anchors, pred_objectness_logits, gt_labels, pred_anchor_deltas, gt_boxes):
        """rpn.py:366-429 (+ box_regression.py:310-345 with smooth_l1 beta=0 -> L1)."""
        num_images = len(gt_labels)
        gt_labels = torch.stack(gt_labels)
        pos_mask = gt_labels == 1
        anchors_t = Boxes.cat(anchors).tensor
        gt_anchor_deltas = torch.stack([self.box2box_transform.get_deltas(anchors_t, k) for k in gt_boxes])
        # masked sums instead of boolean-mask gathers (rpn.py:405-418): same terms, no data-dependent shapes / host syncs
        n = (torch.cat(pred_anchor_deltas, dim=1).float() - gt_anchor_deltas).abs()
        if self.smooth_l1_beta >= 1e-5:
            n = torch.where(n < self.smooth_l1_beta, 0.5 * n ** 2 / self.smooth_l1_beta, n - 0.5 * self.smooth_l1_beta)
        loc = torch.where(pos_mask[..., None], n, torch.zeros((), dtype=n.dtype, device=n.device)).sum()
        valid = gt_labels >= 0
        obj = F.binary_cross_entropy_with_logits(torch.cat(pred_objectness_logits, dim=1).float(),
                                                 gt_labels.to(torch.float32), weight=valid.to(torch.float32),
                                                 reduction="sum")
        normalizer = self.batch_size_per_image * num_images
        losses = {"loss_rpn_cls": obj / normalizer, "loss_rpn_loc": loc / normalizer}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def forward(self, images, features: Dict[str, torch.Tensor], gt_instances=None):
        """rpn.py:431-480."""
        feats = [features[f] for f in self.in_features]
        anchors = self.anchor_generator(feats)
        logits, deltas = self.rpn_head(feats)
        logits = [s.permute(0, 2, 3, 1).flatten(1) for s in logits]     # (N, Hi*Wi*A): free on NHWC storage
        deltas = [x.view(x.shape[0], -1, 4, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2) for x in deltas]
        if self.training:
            assert gt_instances is not None, "RPN requires gt_instances in training!"
            gt_labels, gt_boxes = self.label_and_sample_anchors(anchors, gt_instances)
            losses = self.losses(anchors, logits, gt_labels, deltas, gt_boxes)
        else:
            losses = {}
        proposals = self.predict_proposals(anchors, logits, deltas, images.image_sizes)
        return proposals, losses

    @torch.no_grad()
    def predict_proposals(self, anchors, pred_objectness_logits, pred_anchor_deltas, image_sizes):
        """rpn.py:482-533."""
        N = pred_anchor_deltas[0].shape[0]
        proposals = []
        for a, d in zip(anchors, pred_anchor_deltas):
            B = a.tensor.size(1)
            p = self.box2box_transform.apply_deltas(d.reshape(-1, B), a.tensor.unsqueeze(0).expand(N, -1, -1).reshape(-1, B))
            proposals.append(p.view(N, -1, B))
        return find_top_rpn_proposals(proposals, [l.float() for l in pred_objectness_logits], image_sizes, self.nms_thresh,
                                      self.pre_nms_topk[self.training], self.post_nms_topk[self.training],
                                      self.min_box_size, self.training)


def build_proposal_generator(cfg, input_shape):
    """proposal_generator/build.py:17-24."""
    return PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(cfg, input_shape)
