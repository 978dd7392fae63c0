"""ROI heads — mirror of detectron2/modeling/roi_heads/{roi_heads.py, cascade_rcnn.py, box_head.py,
fast_rcnn.py, mask_head.py} and proposal_utils.add_ground_truth_to_proposals, for the configuration
u2seg uses (CascadeROIHeads, FastRCNNConvFCHead 2xFC, cls-agnostic box regression, class-specific
MaskRCNNConvUpsampleHead). Same module tree / state_dict names:
  roi_heads.box_head.{k}.fc{1,2}.*, roi_heads.box_predictor.{k}.{cls_score,bbox_pred}.*,
  roi_heads.mask_head.{mask_fcn1..4,deconv,predictor}.*

B200-side differences (results identical): pooling is one fused multi-level launch; proposal <-> GT
matching is the fused IoU+Matcher kernel; GT masks are never gathered per proposal
(roi_heads.py:286-288 materialises 512 x H x W bools per image): the matched GT index is carried
instead and the 28x28 targets are cropped straight from the G bit masks.
"""
import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..layers import FeatureTap, Matcher, ROIPooler, batched_nms, crop_and_resize_masks
from ..registry import ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY, ROI_MASK_HEAD_REGISTRY
from ..structures import Boxes, Instances
from .backbone import Conv2d, ShapeSpec, c2_msra_fill, c2_xavier_fill
from .rpn import Box2BoxTransform, subsample_labels


def add_ground_truth_to_proposals(gt, proposals):
    """proposal_utils.py:138-205: proposals first, then the GT boxes with logit log((1-1e-10)/1e-10)."""
    out = []
    gt_logit_value = math.log((1.0 - 1e-10) / (1 - (1.0 - 1e-10)))
    for g, p in zip(gt, proposals):
        dev = p.objectness_logits.device
        gp = Instances(p.image_size)
        gp.proposal_boxes = g.gt_boxes
        gp.objectness_logits = gt_logit_value * torch.ones(len(g), device=dev)
        out.append(Instances.cat([p, gp]))
    return out


class _ScaleGradient(torch.autograd.Function):
    """cascade_rcnn.py:20-28."""

    @staticmethod
    def forward(ctx, input, scale):
        ctx.scale = scale
        return input

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output * ctx.scale, None


@ROI_BOX_HEAD_REGISTRY.register()
class FastRCNNConvFCHead(nn.Sequential):
    """box_head.py:26-110 (NUM_CONV 0 on this path)."""

    def __init__(self, cfg, input_shape: ShapeSpec):
        super().__init__()
        c = cfg.MODEL.ROI_BOX_HEAD
        assert c.NUM_CONV == 0 and c.NUM_FC > 0
        self._output_size = (input_shape.channels, input_shape.height, input_shape.width)
        self.fcs = []
        for k in range(c.NUM_FC):
            fc = nn.Linear(int(np.prod(self._output_size)), c.FC_DIM)
            self.add_module("fc{}".format(k + 1), fc)
            self.fcs.append(fc)
            self._output_size = c.FC_DIM
        for layer in self.fcs:
            c2_xavier_fill(layer)

    def forward(self, x):
        x = torch.flatten(x, start_dim=1)
        if x.is_cuda and torch.is_autocast_enabled("cuda"):
            from . import conv_tc, ops
            dt = torch.get_autocast_dtype("cuda")
            if ops.TCGEN05_CONV_POLICY == "all" and ops.USE_TCGEN05_CONV and conv_tc.linear_eligible(x.to(dt), self.fcs[0].weight):
                x = x.to(dt)
                for fc in self.fcs:      # fc + ReLU in one tcgen05 launch (box_head.py:94-97)
                    x = conv_tc.linear(x, fc.weight, fc.bias, relu=True)
                return x
        for fc in self.fcs:
            x = F.relu(fc(x))
        return x

    @property
    def output_shape(self):
        return ShapeSpec(channels=self._output_size)


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """fast_rcnn.py:118-171."""
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    boxes, scores = boxes[valid], scores[valid]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    b = Boxes(boxes.reshape(-1, 4))
    b.clip(image_shape)
    boxes = b.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    boxes = boxes[filter_inds[:, 0], 0] if num_bbox_reg_classes == 1 else boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh, max_keep=topk_per_image if topk_per_image >= 0 else None)
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.pred_boxes = Boxes(boxes)
    result.scores = scores
    result.pred_classes = filter_inds[:, 1]
    return result, filter_inds[:, 0]


class FastRCNNOutputLayers(nn.Module):
    """fast_rcnn.py:174-569 (softmax CE, smooth-L1 beta -> L1, cls-agnostic or per-class deltas)."""

    def __init__(self, cfg, input_shape: ShapeSpec, box2box_transform):
        super().__init__()
        input_size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
        self.num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        cls_agnostic = cfg.MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG
        self.cls_score = nn.Linear(input_size, self.num_classes + 1)
        self.bbox_pred = nn.Linear(input_size, (1 if cls_agnostic else self.num_classes) * 4)
        nn.init.normal_(self.cls_score.weight, std=0.01)
        nn.init.normal_(self.bbox_pred.weight, std=0.001)
        for l in (self.cls_score, self.bbox_pred):
            nn.init.constant_(l.bias, 0)
        self.box2box_transform = box2box_transform
        self.smooth_l1_beta = cfg.MODEL.ROI_BOX_HEAD.SMOOTH_L1_BETA
        self.test_score_thresh = cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST
        self.test_nms_thresh = cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST
        self.test_topk_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
        assert cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_TYPE == "smooth_l1"
        w = cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_LOSS_WEIGHT
        self.loss_weight = {"loss_cls": 1.0, "loss_box_reg": w}

    def forward(self, x):
        if x.dim() > 2:
            x = torch.flatten(x, start_dim=1)
        if x.is_cuda and torch.is_autocast_enabled("cuda") and x.shape[0] > 0:
            from . import conv_tc, ops
            dt = torch.get_autocast_dtype("cuda")
            if ops.TCGEN05_CONV_POLICY == "all" and ops.USE_TCGEN05_CONV and conv_tc.USE_CONV2 and x.shape[1] % 128 == 0:
                # fast_rcnn.py:236-239 as ONE tcgen05 GEMM: the (K+1)-way classifier and the box regressor share their
                # input, so their filters are stacked and zero-padded to a multiple of 64 rows (801 + 4 -> 832); the library
                # would run the 801-wide GEMM on its unaligned legacy path
                nc, nb = self.cls_score.weight.shape[0], self.bbox_pred.weight.shape[0]
                pad = (-(nc + nb)) % 64
                w = torch.cat([self.cls_score.weight, self.bbox_pred.weight,
                               self.cls_score.weight.new_zeros((pad, x.shape[1]))])
                b = torch.cat([self.cls_score.bias, self.bbox_pred.bias, self.cls_score.bias.new_zeros((pad,))])
                y = conv_tc.linear(x.to(dt), w, b, relu=False)
                return y[:, :nc], y[:, nc:nc + nb]
        return self.cls_score(x), self.bbox_pred(x)

    def losses(self, predictions, proposals):
        """fast_rcnn.py:307-352."""
        scores, proposal_deltas = predictions
        gt_classes = torch.cat([p.gt_classes for p in proposals], dim=0) if len(proposals) else torch.empty(0)
        if len(proposals):
            proposal_boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
            gt_boxes = torch.cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
        else:
            proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
        loss_cls = F.cross_entropy(scores.float(), gt_classes, reduction="mean") if scores.numel() else scores.sum() * 0.0
        losses = {"loss_cls": loss_cls,
                  "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes)}
        return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    def box_reg_loss(self, proposal_boxes, gt_boxes, pred_deltas, gt_classes):
        """fast_rcnn.py:424-463."""
        box_dim = proposal_boxes.shape[1]
        fg = (gt_classes >= 0) & (gt_classes < self.num_classes)
        if pred_deltas.shape[1] == box_dim:
            pred = pred_deltas
        else:
            idx = gt_classes.clamp(0, self.num_classes - 1)
            pred = pred_deltas.view(-1, self.num_classes, box_dim)[torch.arange(len(idx), device=idx.device), idx]
        # masked sum over the foreground rows instead of nonzero() + gather (no host sync); background rows may hold
        # inf/NaN targets (degenerate or absent GT) and are replaced, not multiplied, by zero
        n = (pred.float() - self.box2box_transform.get_deltas(proposal_boxes, gt_boxes)).abs()
        if self.smooth_l1_beta >= 1e-5:
            n = torch.where(n < self.smooth_l1_beta, 0.5 * n ** 2 / self.smooth_l1_beta, n - 0.5 * self.smooth_l1_beta)
        loss = torch.where(fg[:, None], n, torch.zeros((), dtype=n.dtype, device=n.device)).sum()
        return loss / max(gt_classes.numel(), 1.0)

    def predict_boxes(self, predictions, proposals):
        """fast_rcnn.py:523-547."""
        if not len(proposals):
            return []
        _, proposal_deltas = predictions
        num_prop_per_image = [len(p) for p in proposals]
        proposal_boxes = torch.cat([p.proposal_boxes.tensor for p in proposals], dim=0)
        return self.box2box_transform.apply_deltas(proposal_deltas, proposal_boxes).split(num_prop_per_image)

    def predict_probs(self, predictions, proposals):
        scores, _ = predictions
        return F.softmax(scores.float(), dim=-1).split([len(p) for p in proposals], dim=0)


def mask_rcnn_loss(pred_mask_logits, instances, gt_masks_per_image):
    """mask_head.py:33-112. `instances[i].gt_mask_index` indexes `gt_masks_per_image[i]` (G,H,W) bool."""
    cls_agnostic = pred_mask_logits.size(1) == 1
    total = pred_mask_logits.size(0)
    side = pred_mask_logits.size(2)
    gt_classes, gt_masks = [], []
    for inst, masks in zip(instances, gt_masks_per_image):
        if len(inst) == 0:
            continue
        if not cls_agnostic:
            gt_classes.append(inst.gt_classes.to(dtype=torch.int64))
        gt_masks.append(crop_and_resize_masks(masks, inst.proposal_boxes.tensor, side, gt_index=inst.gt_mask_index))
    if len(gt_masks) == 0:
        return pred_mask_logits.sum() * 0
    gt_masks = torch.cat(gt_masks, dim=0)
    if cls_agnostic:
        pred = pred_mask_logits[:, 0]
    else:
        pred = pred_mask_logits[torch.arange(total, device=pred_mask_logits.device), torch.cat(gt_classes, dim=0)]
    return F.binary_cross_entropy_with_logits(pred.float(), gt_masks.to(torch.float32), reduction="mean")


def mask_rcnn_loss_selected(mask_head, x, instances, gt_masks_per_image):
    """mask_rcnn_loss(mask_head(x), ...) without materialising the logits of the classes the loss never reads."""
    if x.shape[0] == 0:
        return mask_head(x).sum() * 0
    classes = torch.cat([i.gt_classes.to(torch.int64) for i in instances if len(i) > 0])
    pred = mask_head.forward_selected(x, classes)
    side = pred.shape[-1]
    gt_masks = torch.cat([crop_and_resize_masks(m, i.proposal_boxes.tensor, side, gt_index=i.gt_mask_index)
                          for i, m in zip(instances, gt_masks_per_image) if len(i) > 0])
    return F.binary_cross_entropy_with_logits(pred.float(), gt_masks.to(torch.float32), reduction="mean")


def mask_rcnn_inference(pred_mask_logits, pred_instances):
    """mask_head.py:115-158."""
    if pred_mask_logits.size(1) == 1:
        probs = pred_mask_logits.sigmoid()
    else:
        n = pred_mask_logits.shape[0]
        cls = torch.cat([i.pred_classes for i in pred_instances])
        probs = pred_mask_logits[torch.arange(n, device=cls.device), cls][:, None].float().sigmoid()
    for prob, inst in zip(probs.split([len(i) for i in pred_instances], dim=0), pred_instances):
        inst.pred_masks = prob


@ROI_MASK_HEAD_REGISTRY.register()
class MaskRCNNConvUpsampleHead(nn.Sequential):
    """mask_head.py:215-290."""

    def __init__(self, cfg, input_shape: ShapeSpec):
        super().__init__()
        c = cfg.MODEL.ROI_MASK_HEAD
        num_classes = 1 if c.CLS_AGNOSTIC_MASK else cfg.MODEL.ROI_HEADS.NUM_CLASSES
        assert c.NORM == ""
        conv_dims = [c.CONV_DIM] * (c.NUM_CONV + 1)
        self.conv_norm_relus = []
        cur = input_shape.channels
        for k, dim in enumerate(conv_dims[:-1]):
            conv = Conv2d(cur, dim, kernel_size=3, stride=1, padding=1, bias=True, activation=F.relu_)
            self.add_module("mask_fcn{}".format(k + 1), conv)
            self.conv_norm_relus.append(conv)
            cur = dim
        self.deconv = nn.ConvTranspose2d(cur, conv_dims[-1], kernel_size=2, stride=2, padding=0)
        self.add_module("deconv_relu", nn.ReLU())
        cur = conv_dims[-1]
        self.predictor = Conv2d(cur, num_classes, kernel_size=1, stride=1, padding=0)
        for layer in self.conv_norm_relus + [self.deconv]:
            c2_msra_fill(layer)
        nn.init.normal_(self.predictor.weight, std=0.001)
        nn.init.constant_(self.predictor.bias, 0)

    def features(self, x):
        """everything before the predictor: mask_fcn1..n (+ReLU), deconv + ReLU (mask_head.py:242-262). Under CUDA
        autocast with the tcgen05 policy the transposed convolution runs on the 2-CTA kernel, ReLU fused."""
        from . import conv_tc, ops
        skip_relu = False
        for layer in self:
            if layer is self.predictor:
                break
            if skip_relu and isinstance(layer, nn.ReLU):
                skip_relu = False
                continue
            if (layer is self.deconv and x.is_cuda and ops.USE_TCGEN05_CONV and ops.TCGEN05_CONV_POLICY == "all"
                    and x.dtype in (torch.float16, torch.bfloat16)):
                y = conv_tc.deconv2x2(x, layer, relu=True)
                if y is not None:
                    x, skip_relu = y, True
                    continue
            x = layer(x)
        return x

    def forward_selected(self, x, classes):
        """Logits of ONE class per ROI, (R, S, S): exactly the entries that mask_head.py:95-97 (training, gt class)
        and mask_head.py:141-143 (inference, predicted class) read from the (R, num_classes, S, S) predictor output.
        With 800 pseudo-classes that output is 800x larger than what is used (321 MB per step at 256 ROIs), so the
        1x1 predictor runs as a per-ROI dot product with the selected filter instead; unselected filters get a zero
        gradient either way."""
        x = self.features(x)
        R, C, S, _ = x.shape
        w = self.predictor.weight.view(-1, C)
        if w.shape[0] == 1:
            classes = torch.zeros_like(classes)
        rows = x.permute(0, 2, 3, 1).reshape(R, S * S, C)            # NHWC storage: a view
        wsel = w[classes].to(rows.dtype)
        out = torch.bmm(rows, wsel.unsqueeze(2)).squeeze(2) + self.predictor.bias[classes].to(rows.dtype)[:, None]
        return out.view(R, S, S)


@ROI_HEADS_REGISTRY.register()
class CascadeROIHeads(nn.Module):
    """cascade_rcnn.py:32-299 on top of roi_heads.py:122-302,530-846."""

    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        rh = cfg.MODEL.ROI_HEADS
        self.batch_size_per_image, self.positive_fraction = rh.BATCH_SIZE_PER_IMAGE, rh.POSITIVE_FRACTION
        self.num_classes, self.proposal_append_gt = rh.NUM_CLASSES, rh.PROPOSAL_APPEND_GT
        self.proposal_matcher = Matcher(rh.IOU_THRESHOLDS, rh.IOU_LABELS, allow_low_quality_matches=False)
        self.box_in_features = self.mask_in_features = list(rh.IN_FEATURES)
        scales = tuple(1.0 / input_shape[k].stride for k in self.box_in_features)
        in_channels = input_shape[self.box_in_features[0]].channels
        bh = cfg.MODEL.ROI_BOX_HEAD
        assert bh.POOLER_TYPE == "ROIAlignV2" and bh.POOLER_SAMPLING_RATIO == 0 and bh.CLS_AGNOSTIC_BBOX_REG
        cascade_w, cascade_ious = cfg.MODEL.ROI_BOX_CASCADE_HEAD.BBOX_REG_WEIGHTS, cfg.MODEL.ROI_BOX_CASCADE_HEAD.IOUS
        assert len(cascade_w) == len(cascade_ious) and cascade_ious[0] == rh.IOU_THRESHOLDS[0]
        self.num_cascade_stages = len(cascade_ious)
        import os
        # round-2 draft (off by default): pool straight into the (K, C*P*P) layout the FC layers flatten to
        self.box_pooler = ROIPooler(bh.POOLER_RESOLUTION, scales, 0, "ROIAlignV2",
                                    chw_output=os.environ.get("U2B_ROI_CHW", "0") == "1")
        pooled = ShapeSpec(channels=in_channels, width=bh.POOLER_RESOLUTION, height=bh.POOLER_RESOLUTION)
        heads, predictors, matchers = [], [], []
        for iou, w in zip(cascade_ious, cascade_w):
            head = ROI_BOX_HEAD_REGISTRY.get(bh.NAME)(cfg, pooled)
            heads.append(head)
            predictors.append(FastRCNNOutputLayers(cfg, head.output_shape, Box2BoxTransform(weights=w)))
            matchers.append(Matcher([iou], [0, 1], allow_low_quality_matches=False))
        self.box_head, self.box_predictor = nn.ModuleList(heads), nn.ModuleList(predictors)
        self.proposal_matchers = matchers
        self.mask_on = cfg.MODEL.MASK_ON
        if self.mask_on:
            mh = cfg.MODEL.ROI_MASK_HEAD
            assert mh.POOLER_TYPE == "ROIAlignV2"
            self.mask_pooler = ROIPooler(mh.POOLER_RESOLUTION, scales, 0, "ROIAlignV2")
            self.mask_head = ROI_MASK_HEAD_REGISTRY.get(mh.NAME)(
                cfg, ShapeSpec(channels=in_channels, width=mh.POOLER_RESOLUTION, height=mh.POOLER_RESOLUTION))

    # ---- roi_heads.py:181-302 ----
    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes):
        if gt_classes.numel() > 0:
            gt_classes = gt_classes[matched_idxs]
            gt_classes[matched_labels == 0] = self.num_classes
            gt_classes[matched_labels == -1] = -1
        else:
            gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
        fg, bg = subsample_labels(gt_classes, self.batch_size_per_image, self.positive_fraction, self.num_classes)
        sampled = torch.cat([fg, bg], dim=0)
        return sampled, gt_classes[sampled]

    @torch.no_grad()
    def label_and_sample_proposals(self, proposals, targets):
        if self.proposal_append_gt:
            proposals = add_ground_truth_to_proposals(targets, proposals)
        out = []
        for p, t in zip(proposals, targets):
            has_gt = len(t) > 0
            midx, mlab = self.proposal_matcher.match_boxes(t.gt_boxes.tensor, p.proposal_boxes.tensor)
            sampled, gt_classes = self._sample_proposals(midx, mlab, t.gt_classes)
            p = p[sampled]
            p.gt_classes = gt_classes
            if has_gt:
                st = midx[sampled]
                p.gt_boxes = t.gt_boxes[st]
                p.gt_mask_index = st          # instead of gt_masks[st] (roi_heads.py:286-288)
            out.append(p)
        return out

    # ---- cascade_rcnn.py ----
    def forward(self, images, features, proposals, targets=None):
        self._tap = FeatureTap([features[f] for f in self.box_in_features]) if self.training else None
        if self.training:
            proposals = self.label_and_sample_proposals(proposals, targets)
            losses = self._forward_box(features, proposals, targets)
            if self.mask_on:
                losses.update(self._forward_mask(features, proposals, targets))
            return proposals, losses
        pred_instances = self._forward_box(features, proposals)
        pred_instances = self.forward_with_given_boxes(features, pred_instances)
        return pred_instances, {}

    def _forward_box(self, features, proposals, targets=None):
        feats = [features[f] for f in self.box_in_features]
        head_outputs = []
        prev_pred_boxes = None
        image_sizes = [x.image_size for x in proposals]
        for k in range(self.num_cascade_stages):
            if k > 0:
                proposals = self._create_proposals_from_boxes(prev_pred_boxes, image_sizes)
                if self.training:
                    proposals = self._match_and_label_boxes(proposals, k, targets)
            predictions = self._run_stage(feats, proposals, k)
            prev_pred_boxes = self.box_predictor[k].predict_boxes(predictions, proposals)
            head_outputs.append((self.box_predictor[k], predictions, proposals))
        if self.training:
            losses = {}
            for stage, (predictor, predictions, props) in enumerate(head_outputs):
                sl = predictor.losses(predictions, props)
                losses.update({k + "_stage{}".format(stage): v for k, v in sl.items()})
            return losses
        scores_per_stage = [h[0].predict_probs(h[1], h[2]) for h in head_outputs]
        scores = [sum(list(s)) * (1.0 / self.num_cascade_stages) for s in zip(*scores_per_stage)]
        predictor, predictions, props = head_outputs[-1]
        boxes = predictor.predict_boxes(predictions, props)
        return [fast_rcnn_inference_single_image(b, s, sz, predictor.test_score_thresh, predictor.test_nms_thresh,
                                                 predictor.test_topk_per_image)[0]
                for b, s, sz in zip(boxes, scores, image_sizes)]

    @torch.no_grad()
    def _match_and_label_boxes(self, proposals, stage, targets):
        for p, t in zip(proposals, targets):
            midx, lab = self.proposal_matchers[stage].match_boxes(t.gt_boxes.tensor, p.proposal_boxes.tensor)
            if len(t) > 0:
                gt_classes = t.gt_classes[midx]
                gt_classes[lab == 0] = self.num_classes
                gt_boxes = t.gt_boxes[midx]
            else:
                gt_classes = torch.zeros_like(midx) + self.num_classes
                gt_boxes = Boxes(t.gt_boxes.tensor.new_zeros((len(p), 4)))
            p.gt_classes, p.gt_boxes = gt_classes, gt_boxes
        return proposals

    def _run_stage(self, feats, proposals, stage):
        # cascade_rcnn.py:283 _ScaleGradient(1/num_stages) on the pooled features: folded into the pooler's backward
        x = self.box_pooler(feats, [p.proposal_boxes for p in proposals], tap=self._tap,
                            grad_scale=1.0 / self.num_cascade_stages if self.training else 1.0)
        return self.box_predictor[stage](self.box_head[stage](x))

    def _create_proposals_from_boxes(self, boxes, image_sizes):
        out = []
        for b, sz in zip(boxes, image_sizes):
            b = Boxes(b.detach())
            b.clip(sz)
            if self.training:
                b = b[b.nonempty()]
            p = Instances(sz)
            p.proposal_boxes = b
            out.append(p)
        return out

    # ---- roi_heads.py:753-846 ----
    def forward_with_given_boxes(self, features, instances):
        assert not self.training
        if self.mask_on:
            feats = [features[f] for f in self.mask_in_features]
            x = self.mask_pooler(feats, [i.pred_boxes for i in instances])
            if x.shape[0] > 0:     # mask_head.py:141-143: only the predicted class's mask is read
                cls = torch.cat([i.pred_classes for i in instances])
                probs = self.mask_head.forward_selected(x, cls)[:, None].float().sigmoid()
                for prob, inst in zip(probs.split([len(i) for i in instances], dim=0), instances):
                    inst.pred_masks = prob
            else:
                mask_rcnn_inference(self.mask_head(x), instances)
        return instances

    def _forward_mask(self, features, instances, targets):
        # select_foreground_proposals (roi_heads.py:46-75)
        fg = []
        for inst in instances:
            sel = (inst.gt_classes != -1) & (inst.gt_classes != self.num_classes)
            fg.append(inst[sel.nonzero().squeeze(1)])
        feats = [features[f] for f in self.mask_in_features]
        x = self.mask_pooler(feats, [i.proposal_boxes for i in fg], tap=self._tap)
        masks = [t.gt_masks.tensor if len(t) else None for t in targets]
        return {"loss_mask": mask_rcnn_loss_selected(self.mask_head, x, fg, masks)}


def build_roi_heads(cfg, input_shape):
    """roi_heads/roi_heads.py:38-43."""
    return ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, input_shape)
