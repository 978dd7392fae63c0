"""Static-shape training forward of PanopticFPN: the same computation as panoptic_fpn.PanopticFPN.forward
(reference: detectron2/modeling/meta_arch/panoptic_fpn.py:90-138 and everything it calls), restructured around
fixed-capacity device buffers so that it contains NO host synchronisation and NO data-dependent shape:

  * proposals: per image 4000 slots + a device-side count (NMS runs with validity masks, proposal_utils.py:103-122);
  * anchor / proposal sampling (sampling.py:9-54): random keys + top-k instead of nonzero + randperm — the same
    uniform-random-subset semantics, with counts kept on the device;
  * sampled ROIs: 512 slots per image (foreground first), mask branch: 128 slots per image, each with a validity
    mask; losses are masked sums divided by device-side counts (identical values to the reference's means);
  * cascade stages keep all rows and carry the `nonempty` filter (cascade_rcnn.py:292-295) as a mask.

This is what lets engine.Trainer capture forward + backward + optimizer in ONE CUDA graph (the reference's step has
>= 25 device->host syncs). GT is passed as padded tensors: gt_boxes (N,G,4), gt_classes (N,G),
gt_valid (N,G) bool, gt_masks (N,G,H,W) bool, sem_seg (N,H,W); all images of a batch share one size.
"""
import math

import os

import torch
import torch.nn.functional as F

from ..layers import FeatureTap, batched_nms_static, crop_and_resize_masks
from ..structures import Boxes

# sampling keys: uniform random numbers by default; tests swap in a deterministic key function
# RPN / box-head losses with closed-form gradients, decode + clip of the selected anchors, cascade relabelling as one
# kernel each (csrc/det_losses.cu); U2B_FUSED_DET_LOSSES=0 falls back to the torch formulas
FUSED_DET_LOSSES = os.environ.get("U2B_FUSED_DET_LOSSES", "1") == "1"
# mask predictor + BCE + their backward as two kernels (csrc/mask_loss.cu); half-precision activations only
FUSED_MASK_LOSS = os.environ.get("U2B_FUSED_MASK_LOSS", "1") == "1"

_rand_keys = lambda mask: torch.rand(mask.shape, dtype=torch.float32, device=mask.device)  # noqa: E731

# Independent branches of the forward pass are issued on side streams (captured as parallel branches of the step's CUDA
# graph; autograd replays each branch's backward on the stream of its forward): the proposal path (top-k, decode,
# per-image NMS with its single-CTA scan) next to the RPN losses, the mask branch next to cascade stages 0-2, the
# semantic head next to everything. Most of these kernels fill a few SMs only.
MULTI_STREAM = os.environ.get("U2B_MULTI_STREAM", "1") == "1"
FORCE_SINGLE_STREAM = False    # bench.py's in-step kernel timing: one stream, so a kernel's events bracket that kernel alone
_streams = {}


def _side(name):
    dev = torch.cuda.current_device()
    st = _streams.get((dev, name))
    if st is None:
        st = _streams[(dev, name)] = torch.cuda.Stream()
    return st


class _Fork:
    """with _Fork(name): ... runs the body on a side stream ordered after the current stream's work so far;
    .join() makes the current stream wait for it. Tensors that cross are kept alive by the caller until the join."""

    def __init__(self, name, enabled=True):
        self.enabled = enabled and MULTI_STREAM and not FORCE_SINGLE_STREAM and torch.cuda.is_available()
        self.name = name

    def __enter__(self):
        if self.enabled:
            self.main = torch.cuda.current_stream()
            self.side = _side(self.name)
            self.side.wait_stream(self.main)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.enabled:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.enabled:
            torch.cuda.current_stream().wait_stream(self.side)


def _topk_select(mask, k):
    """indices (k,) of up to k True entries of `mask` chosen uniformly at random, and a bool (k,) telling which of the
    k slots are real (sampling.py:49-53 `randperm(n)[:k]`)."""
    key = torch.where(mask, _rand_keys(mask), torch.full((), 2.0, dtype=torch.float32, device=mask.device))
    vals, idx = torch.topk(key, min(k, key.numel()), largest=False, sorted=True)
    return idx, vals < 2.0


def subsample_static(is_pos, is_neg, num_samples, positive_fraction):
    """sampling.py:9-54 with fixed shapes: returns idx (num_samples,), ok (num_samples,) bool, is_fg (num_samples,)
    bool; positives first. num_pos = min(#pos, int(num_samples*fraction)), num_neg = min(#neg, num_samples-num_pos)."""
    max_pos = int(num_samples * positive_fraction)
    pidx, pok = _topk_select(is_pos, max_pos)
    num_pos = pok.sum()
    nidx, nok = _topk_select(is_neg, num_samples)
    nok = nok & (torch.arange(nidx.numel(), device=nidx.device) < (num_samples - num_pos))
    idx = torch.cat([pidx, nidx])
    ok = torch.cat([pok, nok])
    fg = torch.cat([torch.ones_like(pok), torch.zeros_like(nok)])
    order = torch.sort((~ok).to(torch.int8), stable=True)[1][:num_samples]     # compact: real slots first, order kept
    return idx[order], ok[order], (fg & ok)[order]


def _masked_l1(pred, target, mask):
    n = (pred.float() - target).abs()
    return torch.where(mask[..., None], n, torch.zeros((), dtype=n.dtype, device=n.device)).sum()


def _nonempty(b):
    return ((b[:, 2] - b[:, 0]) > 0) & ((b[:, 3] - b[:, 1]) > 0)


def _clip(b, size):
    h, w = size
    return torch.stack((b[:, 0].clamp(0, w), b[:, 1].clamp(0, h), b[:, 2].clamp(0, w), b[:, 3].clamp(0, h)), dim=-1)


def decode_topk_level(box2box, anchors_l, logits_l, deltas_l, k):
    """Scores (N,k) and decoded boxes (N,k,4) of the k best-scoring anchors of one pyramid level
    (rpn.py:497-533 _decode_proposals + proposal_utils.py:67-84 per-level top-k). Only the selected anchors are
    decoded: apply_deltas is row-wise, so this is bit-identical to decoding all A*H*W anchors and gathering, at 1/30
    to 1/100 of the elementwise work (the reference decodes 261,888 boxes per image to keep <= 9,000)."""
    N = logits_l.shape[0]
    sc, idx = logits_l.float().topk(k, dim=1)
    dsel = torch.gather(deltas_l, 1, idx[:, :, None].expand(-1, -1, 4))
    asel = anchors_l[idx]
    boxes = box2box.apply_deltas(dsel.reshape(-1, 4), asel.reshape(-1, 4)).view(N, k, 4)
    return sc, boxes


def rpn_static(rpn, images_size, features, gt_boxes, gt_valid, flags):
    """proposal_generator/rpn.py:431-533 -> (proposals (N,P,4), prop_valid (N,P) bool, losses)."""
    feats = [features[f] for f in rpn.in_features]
    anchors = rpn.anchor_generator(feats)
    logits, deltas = rpn.rpn_head(feats)
    N = logits[0].shape[0]
    logits = [s.permute(0, 2, 3, 1).flatten(1) for s in logits]
    deltas = [x.view(N, -1, 4, x.shape[-2], x.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2) for x in deltas]
    anchors_t = Boxes.cat(anchors).tensor
    A = anchors_t.shape[0]
    fork = _Fork("rpn_proposals")
    with fork:
        proposals, prop_valid = _rpn_proposals_static(rpn, images_size, anchors, anchors_t, [lg.detach() for lg in logits],
                                                      [d.detach() for d in deltas], flags)
    # ---- label_and_sample_anchors (rpn.py:307-363) ----
    with torch.no_grad():
        labels_all, matched_all, midx_all = [], [], []
        for n in range(N):
            midx, lab = rpn.anchor_matcher.match_boxes(gt_boxes[n], anchors_t, gt_valid=gt_valid[n])
            idx, ok, fg = subsample_static(lab == 1, lab == 0, rpn.batch_size_per_image, rpn.positive_fraction)
            out = torch.full((A + 1,), -1, dtype=torch.int8, device=lab.device)
            # unselected slots are redirected to a scratch element (index A) so they can never collide with a selected anchor
            out.scatter_(0, torch.where(ok, idx, torch.full_like(idx, A)), fg.to(torch.int8))
            labels_all.append(out[:A])
            midx_all.append(midx)
            if not FUSED_DET_LOSSES:
                matched_all.append(gt_boxes[n][midx])
        gt_labels = torch.stack(labels_all)
        if not FUSED_DET_LOSSES:
            gt_anchor_deltas = torch.stack([rpn.box2box_transform.get_deltas(anchors_t, k) for k in matched_all])
    if FUSED_DET_LOSSES:     # one kernel for both RPN losses and their gradients (csrc/det_losses.cu)
        from .fused_losses import rpn_losses
        deltas_cat = torch.cat(deltas, dim=1)
        obj, loc = rpn_losses(torch.cat(logits, dim=1), deltas_cat, anchors_t, gt_labels,
                              torch.stack(midx_all), gt_boxes, rpn.box2box_transform.weights)
    else:
        pos_mask = gt_labels == 1
        loc = _masked_l1(torch.cat(deltas, dim=1), gt_anchor_deltas, pos_mask)
        valid = gt_labels >= 0
        obj = F.binary_cross_entropy_with_logits(torch.cat(logits, dim=1).float(), gt_labels.to(torch.float32),
                                                 weight=valid.to(torch.float32), reduction="sum")
    normalizer = rpn.batch_size_per_image * N
    losses = {"loss_rpn_cls": obj / normalizer * rpn.loss_weight["loss_rpn_cls"],
              "loss_rpn_loc": loc / normalizer * rpn.loss_weight["loss_rpn_loc"]}
    fork.join()
    return proposals, prop_valid, losses


def _nms_per_image(boxes, scores, lvl_ids, valid, thresh, post):
    """batched NMS of every image (proposal_utils.py:112-122), images on separate streams: the scan is one CTA each."""
    N = boxes.shape[0]
    out_boxes, out_valid, forks = [None] * N, [None] * N, []
    for n in range(N):
        f = _Fork("rpn_nms%d" % n, enabled=n > 0)
        with f:
            keep, cnt = batched_nms_static(boxes[n], scores[n], lvl_ids, thresh, post, valid=valid[n])
            out_boxes[n] = boxes[n][keep]
            out_valid[n] = torch.arange(post, device=keep.device) < cnt
        forks.append(f)
    for f in forks:
        f.join()
    return torch.stack(out_boxes), torch.stack(out_valid)


@torch.no_grad()
def _rpn_proposals_static(rpn, images_size, anchors, anchors_t, logits, deltas, flags):
    """predict_proposals + find_top_rpn_proposals (rpn.py:482-533, proposal_utils.py:22-135) on fixed-capacity
    buffers -> proposals (N,post,4), valid (N,post)."""
    pre, post = rpn.pre_nms_topk[True], rpn.post_nms_topk[True]
    if FUSED_DET_LOSSES:   # decode + clip + validity of the selected anchors in one kernel
        from .fused_losses import rpn_decode_selected
        deltas_cat = torch.cat(deltas, dim=1)
        sel, scs, lvl_ids, off = [], [], [], 0
        for lid, lg in enumerate(logits):
            k = min(lg.shape[1], pre)
            sc, idx = lg.float().topk(k, dim=1)
            sel.append(idx + off)
            scs.append(sc)
            lvl_ids.append(torch.full((k,), lid, dtype=torch.int64, device=anchors_t.device))
            off += lg.shape[1]
        tk_scores, lvl_ids = torch.cat(scs, 1), torch.cat(lvl_ids)
        boxes_all, valid_all, nonfin = rpn_decode_selected(deltas_cat, anchors_t, torch.cat(sel, 1), tk_scores,
                                                           rpn.box2box_transform, images_size, rpn.min_box_size)
        flags.append(nonfin)
        return _nms_per_image(boxes_all, tk_scores, lvl_ids, valid_all, rpn.nms_thresh, post)
    tk_scores, tk_boxes, lvl_ids = [], [], []
    for lid, (a, lg, dl) in enumerate(zip(anchors, logits, deltas)):
        k = min(lg.shape[1], pre)
        sc, boxes = decode_topk_level(rpn.box2box_transform, a.tensor, lg, dl, k)
        tk_scores.append(sc)
        tk_boxes.append(boxes)
        lvl_ids.append(torch.full((k,), lid, dtype=torch.int64, device=anchors_t.device))
    tk_scores, tk_boxes, lvl_ids = torch.cat(tk_scores, 1), torch.cat(tk_boxes, 1), torch.cat(lvl_ids)
    finite = torch.isfinite(tk_boxes).all(dim=2) & torch.isfinite(tk_scores)
    flags.append(~finite.all())      # proposal_utils.py:105-110 FloatingPointError, checked off the critical path
    N = tk_boxes.shape[0]
    b = torch.stack([_clip(tk_boxes[n], images_size) for n in range(N)])
    v = finite & ((b[..., 2] - b[..., 0]) > rpn.min_box_size) & ((b[..., 3] - b[..., 1]) > rpn.min_box_size)
    return _nms_per_image(b, tk_scores, lvl_ids, v, rpn.nms_thresh, post)


def roi_heads_static(rh, images_size, features, proposals, prop_valid, gt_boxes, gt_classes, gt_valid, gt_masks):
    """roi_heads/cascade_rcnn.py:137-299 + roi_heads.py:220-302,818-846 + mask_head.py:33-112, fixed shapes."""
    N, K = proposals.shape[0], rh.num_classes
    dev = proposals.device
    R = rh.batch_size_per_image
    feats = [features[f] for f in rh.box_in_features]
    tap = FeatureTap(feats, prealloc=MULTI_STREAM)
    dummy = torch.cat([torch.zeros(2, device=dev), torch.ones(2, device=dev)])     # placeholder box of dead slots
    # ---- label_and_sample_proposals ----
    with torch.no_grad():
        boxes0, cls0, ok0, fg0, gidx0, gtb0 = [], [], [], [], [], []
        for n in range(N):
            cand = torch.cat([proposals[n], gt_boxes[n]])                       # proposals first, then GT
            cv = torch.cat([prop_valid[n], gt_valid[n]])
            midx, mlab = rh.proposal_matcher.match_boxes(gt_boxes[n], cand, gt_valid=gt_valid[n])
            cls = gt_classes[n][midx]
            cls = torch.where(mlab == 0, torch.full_like(cls, K), cls)
            has_gt = gt_valid[n].any()
            cls = torch.where(has_gt, cls, torch.full_like(cls, K))
            idx, ok, fg = subsample_static(cv & (cls < K), cv & (cls == K), R, rh.positive_fraction)
            b = torch.where(ok[:, None], cand[idx], dummy)
            boxes0.append(b)
            cls0.append(torch.where(ok, cls[idx], torch.full_like(idx, -100)))
            ok0.append(ok)
            fg0.append(fg)
            gidx0.append(midx[idx])
            gtb0.append(gt_boxes[n][midx[idx]])
    losses = {}
    mask_fork = _Fork("mask_branch")
    with mask_fork:
        loss_mask = _mask_branch_static(rh, feats, tap, boxes0, cls0, fg0, gidx0, gt_masks, K, R)
    cur_boxes, cur_cls, cur_ok, cur_gtb = boxes0, cls0, ok0, gtb0
    for k in range(rh.num_cascade_stages):
        if k > 0 and FUSED_DET_LOSSES:     # round-2 draft: one kernel relabels every slot of every image
            from .fused_losses import cascade_relabel
            with torch.no_grad():
                rb, rc, rok, rgb = cascade_relabel(torch.stack([b.detach() for b in prev_boxes]), torch.stack(cur_ok),
                                                   gt_boxes, gt_classes, gt_valid, images_size,
                                                   rh.proposal_matchers[k].thresholds[1], K)
                cur_boxes, cur_cls, cur_ok, cur_gtb = list(rb), list(rc), list(rok), list(rgb)
        elif k > 0:
            with torch.no_grad():
                nb, nc, nok, ngb = [], [], [], []
                for n in range(N):
                    b = _clip(prev_boxes[n].detach(), images_size)
                    ok = cur_ok[n] & _nonempty(b)                                # cascade_rcnn.py:292-295
                    b = torch.where(ok[:, None], b, dummy)
                    midx, lab = rh.proposal_matchers[k].match_boxes(gt_boxes[n], b, gt_valid=gt_valid[n])
                    cls = gt_classes[n][midx]
                    cls = torch.where(lab == 0, torch.full_like(cls, K), cls)
                    cls = torch.where(gt_valid[n].any(), cls, torch.full_like(cls, K))
                    nb.append(b)
                    nc.append(torch.where(ok, cls, torch.full_like(cls, -100)))
                    nok.append(ok)
                    ngb.append(gt_boxes[n][midx])
                cur_boxes, cur_cls, cur_ok, cur_gtb = nb, nc, nok, ngb
        x = rh.box_pooler(feats, cur_boxes, tap=tap, grad_scale=1.0 / rh.num_cascade_stages)   # cascade_rcnn.py:20-28,283
        scores, deltas = rh.box_predictor[k](rh.box_head[k](x))
        cls_all, ok_all = torch.cat(cur_cls), torch.cat(cur_ok)
        pb, gb = torch.cat(cur_boxes), torch.cat(cur_gtb)
        count = ok_all.sum().clamp(min=1).to(torch.float32)
        # fast_rcnn.py:307-352: mean CE over the sampled rows; L1 over the foreground rows / #rows
        if FUSED_DET_LOSSES:   # round-2 draft: CE + L1 + refined boxes in one kernel (dead slots carry class -100)
            from .fused_losses import box_losses
            ce, l1, refined = box_losses(scores, deltas, cls_all, pb, gb, K, rh.box_predictor[k].box2box_transform)
            losses["loss_cls_stage%d" % k] = ce / count
            losses["loss_box_reg_stage%d" % k] = l1 / count * rh.box_predictor[k].loss_weight["loss_box_reg"]
            prev_boxes = refined.split(R)
            continue
        ce = F.cross_entropy(scores.float(), cls_all, reduction="sum", ignore_index=-100)
        losses["loss_cls_stage%d" % k] = ce / count
        fg = ok_all & (cls_all >= 0) & (cls_all < K)
        tgt = rh.box_predictor[k].box2box_transform.get_deltas(pb, gb)
        losses["loss_box_reg_stage%d" % k] = _masked_l1(deltas, tgt, fg) / count * rh.box_predictor[k].loss_weight["loss_box_reg"]
        prev_boxes = rh.box_predictor[k].box2box_transform.apply_deltas(deltas, pb).split(R)
    mask_fork.join()
    losses["loss_mask"] = loss_mask
    return losses


def _mask_branch_static(rh, feats, tap, boxes0, cls0, fg0, gidx0, gt_masks, K, R):
    """mask branch on the stage-0 foreground slots (first quarter of every image's slots): roi_heads.py:818-846 +
    mask_head.py:33-112."""
    N = len(boxes0)
    M = int(R * rh.positive_fraction)
    mb = [b[:M] for b in boxes0]
    mok = torch.cat([f[:M] for f in fg0])
    mcls = torch.cat([c[:M] for c in cls0]).clamp(0, K - 1)
    xm = rh.mask_pooler(feats, mb, tap=tap)
    if FUSED_MASK_LOSS:
        from .fused_losses import mask_loss_selected, mask_loss_supported
        head = rh.mask_head
        xf = head.features(xm)                                   # mask_fcn1..4, deconv + ReLU; the predictor is fused below
        if mask_loss_supported(xf) and head.predictor.weight.shape[0] > 1:
            side = xf.shape[-1]
            with torch.no_grad():
                tgt = torch.cat([crop_and_resize_masks(gt_masks[n], mb[n], side, gt_index=gidx0[n][:M]) for n in range(N)])
            total = mask_loss_selected(xf, head.predictor.weight, head.predictor.bias, mcls, tgt, mok)
            denom = (mok.sum() * side * side).clamp(min=1).to(torch.float32)
            return total / denom
    sel = rh.mask_head.forward_selected(xm, mcls).float()        # (N*M, S, S): the gt-class logits only
    side = sel.shape[-1]
    with torch.no_grad():
        tgt = torch.cat([crop_and_resize_masks(gt_masks[n], mb[n], side, gt_index=gidx0[n][:M]) for n in range(N)])
    bce = F.binary_cross_entropy_with_logits(sel, tgt.to(torch.float32), reduction="none")
    denom = (mok.sum() * side * side).clamp(min=1).to(torch.float32)
    return (bce * mok[:, None, None].to(bce.dtype)).sum() / denom


class _ScaleGrad(torch.autograd.Function):
    """cascade_rcnn.py:20-28."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x

    @staticmethod
    def backward(ctx, g):
        return g * ctx.scale, None


def forward_train_static(model, images_u8, gt_boxes, gt_classes, gt_valid, gt_masks, sem_seg):
    """PanopticFPN.forward (training) on padded tensors. images_u8: (N,3,H,W) uint8 on the device. Returns
    (dict of the 10 losses in the reference's key order, nonfinite flag tensor)."""
    N, _, H, W = images_u8.shape
    s = model.backbone.size_divisibility
    Hp, Wp = (H + s - 1) // s * s, (W + s - 1) // s * s
    from . import ops
    if ops.PREPROCESS_KERNEL and images_u8.is_cuda and images_u8.dtype == torch.uint8:
        # normalise + pad + layout in one kernel; under autocast straight to the stem's input dtype (same single rounding
        # autocast's cast would apply to the fp32 tensor)
        odt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
        x = ops.preprocess_u8(images_u8, model._pixel_mean_host, model._pixel_std_host, s, odt)
    else:
        x = ((images_u8.float() - model.pixel_mean) / model.pixel_std)
        if (Hp, Wp) != (H, W):
            x = F.pad(x, (0, Wp - W, 0, Hp - H))
        x = x.contiguous(memory_format=torch.channels_last)
    if (Hp, Wp) != (H, W):
        sem_seg = F.pad(sem_seg, (0, Wp - W, 0, Hp - H), value=model.sem_seg_head.ignore_value)
    if model._bn_counters:
        torch._foreach_add_(model._bn_counters, 1)
    features = model.backbone(x)
    sem_fork = _Fork("sem_seg_head", enabled=True)
    sem_fork.enabled = torch.cuda.is_available() and not FORCE_SINGLE_STREAM   # forked also with U2B_MULTI_STREAM=0 (round 1)
    with sem_fork:
        _, sem_losses = model.sem_seg_head(features, sem_seg)
    flags = []
    proposals, prop_valid, rpn_losses = rpn_static(model.proposal_generator, (H, W), features, gt_boxes, gt_valid, flags)
    det_losses = roi_heads_static(model.roi_heads, (H, W), features, proposals, prop_valid, gt_boxes, gt_classes,
                                  gt_valid, gt_masks)
    sem_fork.join()
    losses = dict(sem_losses)
    losses.update(rpn_losses)
    losses.update(det_losses)
    return losses, torch.stack(flags).any()


def pack_batch(batched_inputs, device, g_max=None):
    """list[dict] (reference input format, same-size images) -> the padded tensors forward_train_static takes."""
    imgs = torch.stack([d["image"] for d in batched_inputs]).to(device, non_blocking=True)
    imgs = imgs.contiguous(memory_format=torch.channels_last)      # NHWC bytes: the normalisation below stays NHWC
    G = g_max or max(1, max(len(d["instances"]) for d in batched_inputs))
    N, H, W = len(batched_inputs), imgs.shape[-2], imgs.shape[-1]
    gb = torch.zeros((N, G, 4), dtype=torch.float32, device=device)
    gc = torch.zeros((N, G), dtype=torch.int64, device=device)
    gv = torch.zeros((N, G), dtype=torch.bool, device=device)
    gm = torch.zeros((N, G, H, W), dtype=torch.bool, device=device)
    for n, d in enumerate(batched_inputs):
        inst = d["instances"]
        g = len(inst)
        if g:
            gb[n, :g] = inst.gt_boxes.tensor.to(device, non_blocking=True)
            gc[n, :g] = inst.gt_classes.to(device, non_blocking=True)
            gv[n, :g] = True
            gm[n, :g] = inst.gt_masks.tensor.to(device, non_blocking=True)
    sem = torch.stack([d["sem_seg"] for d in batched_inputs]).to(device, non_blocking=True)
    return imgs, gb, gc, gv, gm, sem
