"""Operator layer — host-side mirror of detectron2/layers + the box/mask ops of detectron2/structures
used on the u2seg hot path. Same Python signatures as the reference ops; every device op goes through
libu2b200.so (include/u2b200.h). Tensors are logical NCHW in torch's channels_last memory format
(physical NHWC), which is what the kernels read.

There is no CPU path: calling these on CPU tensors raises.
"""
import ctypes

import torch
from torch import nn

from . import _lib

_DTYPE_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s: libu2b200 ops need CUDA tensors (no CPU fallback)" % what)


def _aligned(t, dtype=None):
    """contiguous, 16-byte aligned, optionally cast."""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


def to_nhwc(x):
    """logical NCHW tensor -> channels_last storage (no copy if it already is)."""
    return x.contiguous(memory_format=torch.channels_last)


def _arr(ctype, values):
    return (ctype * len(values))(*values)


# --------------------------------------------------------------------------------------
# ROIAlign / multi-level pooling
# --------------------------------------------------------------------------------------
def assign_boxes_to_levels_rois(rois5, min_level, max_level, canonical_box_size, canonical_level):
    """poolers.py:23-59 on (K,5) rois; returns int32 (K,) level - min_level."""
    L = _lib.lib()
    K = rois5.shape[0]
    levels = torch.empty((K,), dtype=torch.int32, device=rois5.device)
    _lib.check(L.u2b_assign_levels(_lib.ptr(rois5), K, min_level, max_level, float(canonical_box_size),
                                   canonical_level, _lib.ptr(levels), _lib.stream_ptr()), "u2b_assign_levels")
    _lib.count_launches(1 if K else 0)
    return levels


class _GradHolder:
    """fp32 NHWC gradient maps shared by every pooling call of one forward pass (see FeatureTap)."""

    def __init__(self):
        self.bufs = None

    def get(self, shapes, device):
        if self.bufs is None:
            self.bufs = [torch.zeros((s[0], s[2], s[3], s[1]), dtype=torch.float32, device=device) for s in shapes]
        return self.bufs


class _PoolTap(torch.autograd.Function):
    """Identity 'tap' on the pyramid: returns a scalar token every pooling call of the step depends on. The
    pooling backward passes scatter into the shared fp32 maps of `holder` and send only a scalar gradient to the
    token; autograd therefore runs this node's backward after ALL of them, where the maps are cast once and
    handed to the feature maps. (The reference zero-fills, scatters and casts a full pyramid per call: 4x/step.)"""

    @staticmethod
    def forward(ctx, holder, *feats):
        ctx.holder = holder
        ctx.meta = [(tuple(f.shape), f.dtype) for f in feats]
        return feats[0].new_zeros(())

    @staticmethod
    def backward(ctx, gtoken):
        bufs = ctx.holder.bufs
        ctx.holder.bufs = None
        if bufs is None:
            return (None,) + tuple(None for _ in ctx.meta)
        return (None,) + tuple(b.permute(0, 3, 1, 2).to(dt) for b, (_, dt) in zip(bufs, ctx.meta))


class FeatureTap:
    """Create once per forward pass over the list of pyramid levels, pass as `tap=` to every ROIPooler call."""

    def __init__(self, feats, prealloc=False):
        self.holder = _GradHolder()
        self.feats = [f.detach() for f in feats]
        self.token = _PoolTap.apply(self.holder, *feats) if any(f.requires_grad for f in feats) else None
        if prealloc and self.token is not None:
            # pooling calls issued from several CUDA streams (box cascade on one, mask branch on another) scatter into the
            # same maps in backward: the zero-fill must be ordered before ALL of them, i.e. happen here, before the fork
            self.holder.get([tuple(f.shape) for f in feats], feats[0].device)


class _MultiLevelROIAlign(torch.autograd.Function):
    """out[K, C, P, P] (channels_last) = ROIAlignV2 of rois on their assigned pyramid level."""

    @staticmethod
    def forward(ctx, rois5, levels, P, scales, token, holder, grad_scale, chw, *feats):
        ctx.holder = holder
        ctx.grad_scale = float(grad_scale)
        ctx.chw = bool(chw)
        L = _lib.lib()
        f0 = feats[0]
        _need_cuda(f0, "roi_align")
        dt = f0.dtype
        feats_cl = [to_nhwc(f) for f in feats]
        C = f0.shape[1]
        K = rois5.shape[0]
        if ctx.chw:    # round-2 draft: (K,C,P,P) contiguous, what the box head's flatten reads
            assert L.u2b_roi_align_chw_supported(C, P), "channel-major ROIAlign: C=%d P=%d not supported" % (C, P)
            out = torch.empty((K, C, P, P), dtype=dt, device=f0.device)
        else:
            out = torch.empty((K, P, P, C), dtype=dt, device=f0.device).permute(0, 3, 1, 2)   # (K,C,P,P), NHWC storage
        if K > 0:
            ptrs = _arr(ctypes.c_void_p, [f.data_ptr() for f in feats_cl])
            hs = _arr(ctypes.c_int32, [f.shape[2] for f in feats_cl])
            ws = _arr(ctypes.c_int32, [f.shape[3] for f in feats_cl])
            sc = _arr(ctypes.c_float, list(scales))
            fwd = L.u2b_roi_align_fwd_chw if ctx.chw else L.u2b_roi_align_fwd
            _lib.check(fwd(_DTYPE_CODE[dt], len(feats_cl), ptrs, hs, ws, sc, C, _lib.ptr(rois5),
                           _lib.ptr(levels) if levels is not None else None, K, P,
                           ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr()), "u2b_roi_align_fwd")
            _lib.count_launches(1)
        ctx.save_for_backward(rois5, levels if levels is not None else torch.empty(0))
        ctx.has_levels = levels is not None
        ctx.meta = (P, tuple(scales), [tuple(f.shape) for f in feats], dt)
        return out

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        rois5, levels = ctx.saved_tensors
        P, scales, shapes, dt = ctx.meta
        K = rois5.shape[0]
        C = shapes[0][1]
        shared = ctx.holder is not None
        if shared:
            grads = ctx.holder.get(shapes, gout.device)
        else:
            grads = [torch.zeros((s[0], s[2], s[3], s[1]), dtype=torch.float32, device=gout.device) for s in shapes]
        if K > 0:
            g = gout.to(dt).contiguous() if ctx.chw else to_nhwc(gout.to(dt))
            ptrs = _arr(ctypes.c_void_p, [t.data_ptr() for t in grads])
            hs = _arr(ctypes.c_int32, [s[2] for s in shapes])
            ws = _arr(ctypes.c_int32, [s[3] for s in shapes])
            sc = _arr(ctypes.c_float, list(scales))
            bwd = L.u2b_roi_align_bwd_chw if ctx.chw else L.u2b_roi_align_bwd
            _lib.check(bwd(_DTYPE_CODE[dt], len(shapes), ptrs, hs, ws, sc, C, _lib.ptr(rois5),
                           _lib.ptr(levels) if ctx.has_levels else None, K, P,
                           ctypes.c_void_p(g.data_ptr()), ctx.grad_scale, _lib.stream_ptr()), "u2b_roi_align_bwd")
            _lib.count_launches(1)
        if shared:   # gradients reach the features through _PoolTap; the token only orders the backward passes
            return (None, None, None, None, gout.new_zeros(()), None, None, None) + tuple(None for _ in shapes)
        outs = [t.permute(0, 3, 1, 2).to(dt) for t in grads]   # logical NCHW views of the NHWC grads
        return (None, None, None, None, None, None, None, None) + tuple(outs)


class ROIAlign(nn.Module):
    """detectron2/layers/roi_align.py:8-74. Only the ROIAlignV2 configuration the u2seg configs use
    is implemented in CUDA: aligned=True, sampling_ratio=0."""

    def __init__(self, output_size, spatial_scale, sampling_ratio, aligned=True):
        super().__init__()
        self.output_size = output_size if isinstance(output_size, int) else output_size[0]
        if not isinstance(output_size, int):
            assert output_size[0] == output_size[1], "square outputs only"
        self.spatial_scale = float(spatial_scale)
        if sampling_ratio != 0 or not aligned:
            raise NotImplementedError("libu2b200 implements ROIAlignV2 (aligned=True, sampling_ratio=0) only")
        self.sampling_ratio, self.aligned = sampling_ratio, aligned

    def forward(self, input, rois):
        assert rois.dim() == 2 and rois.size(1) == 5
        rois = _aligned(rois, torch.float32)
        return _MultiLevelROIAlign.apply(rois, None, self.output_size, (self.spatial_scale,), None, None, 1.0, False, input)


def convert_boxes_to_pooler_format(box_tensors):
    """poolers.py:72-98: list of (Ni,4) -> (sum Ni, 5) with the batch index in column 0."""
    boxes = torch.cat(box_tensors, dim=0)
    idx = torch.cat([torch.full((len(b), 1), float(i), dtype=boxes.dtype, device=boxes.device)
                     for i, b in enumerate(box_tensors)], dim=0)     # device fills only: CUDA-graph capturable
    return torch.cat([idx, boxes], dim=1)


class ROIPooler(nn.Module):
    """detectron2/modeling/poolers.py:114-263 (pooler_type 'ROIAlignV2'). One fused launch over all
    levels instead of per-level nonzero + roi_align + index_put_."""

    def __init__(self, output_size, scales, sampling_ratio=0, pooler_type="ROIAlignV2", canonical_box_size=224,
                 canonical_level=4, chw_output=False):
        super().__init__()
        self.chw_output = bool(chw_output)     # round-2 draft: (M,C,P,P) contiguous instead of channels_last
        import math
        assert pooler_type == "ROIAlignV2" and sampling_ratio == 0
        self.output_size = output_size if isinstance(output_size, int) else output_size[0]
        self.scales = tuple(float(s) for s in scales)
        min_level, max_level = -math.log2(scales[0]), -math.log2(scales[-1])
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level, self.max_level = int(min_level), int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1
        self.canonical_level, self.canonical_box_size = canonical_level, canonical_box_size

    def forward(self, x, box_lists, tap=None, grad_scale=1.0):
        """x: list of (N,C,Hl,Wl); box_lists: list (per image) of (Ni,4) tensors (or objects with .tensor).
        tap: optional FeatureTap over the same `x` shared by all pooling calls of the step.
        grad_scale: factor applied to the gradient flowing back into the features (cascade_rcnn.py:20-28
        _ScaleGradient fused into the backward kernel)."""
        boxes = [b.tensor if hasattr(b, "tensor") else b for b in box_lists]
        assert len(x) == len(self.scales) and len(boxes) == x[0].size(0)
        rois = _aligned(convert_boxes_to_pooler_format(boxes).float())
        levels = None
        if len(x) > 1:
            levels = assign_boxes_to_levels_rois(rois, self.min_level, self.max_level, self.canonical_box_size,
                                                 self.canonical_level)
        if tap is not None and tap.token is not None:
            return _MultiLevelROIAlign.apply(rois, levels, self.output_size, self.scales, tap.token, tap.holder,
                                             grad_scale, self.chw_output, *tap.feats)
        return _MultiLevelROIAlign.apply(rois, levels, self.output_size, self.scales, None, None, grad_scale,
                                         self.chw_output, *x)


# --------------------------------------------------------------------------------------
# masks
# --------------------------------------------------------------------------------------
def paste_masks_in_image(masks, boxes, image_shape, threshold=0.5):
    """detectron2/layers/mask_ops.py:74-147. masks (B,M,M) float, boxes (B,4) -> bool (B,H,W)."""
    L = _lib.lib()
    _need_cuda(masks, "paste_masks_in_image")
    assert masks.shape[-1] == masks.shape[-2], "Only square mask predictions are supported"
    if hasattr(boxes, "tensor"):
        boxes = boxes.tensor
    N, M = masks.shape[0], masks.shape[-1]
    H, W = int(image_shape[0]), int(image_shape[1])
    if threshold < 0:
        raise NotImplementedError("soft (uint8) pasting is not on the u2seg path")
    out = torch.empty((N, H, W), dtype=torch.uint8, device=masks.device)
    if N > 0:
        m = _aligned(masks.reshape(N, M, M), torch.float32)
        b = _aligned(boxes, torch.float32)
        _lib.check(L.u2b_paste_masks(_lib.ptr(m), _lib.ptr(b), N, M, H, W, float(threshold), _lib.ptr(out),
                                     _lib.stream_ptr()), "u2b_paste_masks")
        _lib.count_launches(1)
    return out.view(torch.bool)


def crop_and_resize_masks(masks, boxes, mask_size, gt_index=None, return_values=False):
    """structures/masks.py:191-222 BitMasks.crop_and_resize, reading `masks[gt_index[i]]` in place.
    masks (G,H,W) bool; boxes (M,4); -> bool (M, S, S)."""
    L = _lib.lib()
    _need_cuda(masks, "crop_and_resize_masks")
    M = boxes.shape[0]
    H, W = masks.shape[-2:]
    mk = masks.contiguous().view(torch.uint8) if masks.dtype == torch.bool else masks.to(torch.uint8).contiguous()
    out = torch.empty((M, mask_size, mask_size), dtype=torch.uint8, device=masks.device)
    val = torch.empty((M, mask_size, mask_size), dtype=torch.float32, device=masks.device) if return_values else None
    if M > 0:
        b = _aligned(boxes, torch.float32)
        gi = gt_index.to(torch.int64).contiguous() if gt_index is not None else None
        _lib.check(L.u2b_crop_resize_masks(_lib.ptr(mk), _lib.ptr(gi), _lib.ptr(b), M, H, W, mask_size,
                                           _lib.ptr(out), _lib.ptr(val), _lib.stream_ptr()), "u2b_crop_resize_masks")
        _lib.count_launches(1)
    return (out.view(torch.bool), val) if return_values else out.view(torch.bool)


# --------------------------------------------------------------------------------------
# semantic-segmentation loss: bilinear upsampling + cross-entropy, fused (forward and backward in one kernel)
# --------------------------------------------------------------------------------------
class _UpsampleCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, scale, ignore_index):
        L = _lib.lib()
        N, C, h, w = logits.shape
        z = logits if (logits.is_contiguous(memory_format=torch.channels_last) and logits.stride(1) == 1) else \
            logits.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        tg = targets.to(torch.int64).contiguous()
        npart = L.u2b_upsample_ce_num_partials(N, h * scale, w * scale)
        partials = torch.empty((npart, 2), dtype=torch.float32, device=logits.device)
        need = ctx.needs_input_grad[0]
        dz = torch.zeros((N, h, w, C), dtype=torch.float32, device=logits.device) if need else None
        assert z.stride(1) == 1 and z.permute(0, 2, 3, 1).is_contiguous()
        _lib.check(L.u2b_upsample_ce(_DTYPE_CODE[z.dtype], ctypes.c_void_p(z.data_ptr()), _lib.ptr(tg), N, h, w, C, scale, ignore_index,
                                     _lib.ptr(dz), _lib.ptr(partials), _lib.stream_ptr()), "u2b_upsample_ce")
        _lib.count_launches(1)
        tot = partials.sum(0)
        ctx.save_for_backward(dz, tot[1])
        ctx.in_dtype = logits.dtype
        return tot[0] / tot[1]

    @staticmethod
    def backward(ctx, g):
        dz, count = ctx.saved_tensors
        gz = (dz * (g / count)).to(ctx.in_dtype).permute(0, 3, 1, 2)        # logical NCHW, NHWC storage
        return gz, None, None, None


def upsample_cross_entropy(logits, targets, scale, ignore_index):
    """mean over non-ignored pixels of CE(bilinear_upsample(logits.float(), scale), targets)
    (semantic_seg.py:255-267) without materialising anything of full resolution. logits (N,C,h,w); targets
    (N, h*scale, w*scale) integer."""
    _need_cuda(logits, "upsample_cross_entropy")
    assert targets.shape == (logits.shape[0], logits.shape[2] * scale, logits.shape[3] * scale), \
        "targets %s do not match logits %s x%d" % (tuple(targets.shape), tuple(logits.shape), scale)
    return _UpsampleCE.apply(logits, targets, int(scale), int(ignore_index))


def upsample_cross_entropy_supported(logits, scale):
    return logits.is_cuda and logits.dtype in _DTYPE_CODE and float(scale) == int(scale) and \
        bool(_lib.lib().u2b_upsample_ce_supported(logits.shape[1], int(scale)))


# --------------------------------------------------------------------------------------
# bilinear upsampling on NHWC activations (round-2 draft, csrc/upsample.cu)
# --------------------------------------------------------------------------------------
class _UpsampleBilinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        L = _lib.lib()
        xc = to_nhwc(x)
        N, C, h, w = xc.shape
        y = torch.empty((N, h * scale, w * scale, C), dtype=xc.dtype, device=xc.device).permute(0, 3, 1, 2)
        _lib.check(L.u2b_upsample_bilinear(_DTYPE_CODE[xc.dtype], 0, ctypes.c_void_p(xc.data_ptr()),
                                           ctypes.c_void_p(y.data_ptr()), N, h, w, C, scale, _lib.stream_ptr()),
                   "u2b_upsample_bilinear")
        _lib.count_launches(1)
        ctx.meta = (N, C, h, w, scale, xc.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.lib()
        N, C, h, w, scale, dt = ctx.meta
        g = to_nhwc(gy.to(dt))
        dx = torch.empty((N, h, w, C), dtype=dt, device=g.device).permute(0, 3, 1, 2)
        _lib.check(L.u2b_upsample_bilinear(_DTYPE_CODE[dt], 1, ctypes.c_void_p(g.data_ptr()),
                                           ctypes.c_void_p(dx.data_ptr()), N, h, w, C, scale, _lib.stream_ptr()),
                   "u2b_upsample_bilinear")
        _lib.count_launches(1)
        return dx, None


def upsample_bilinear(x, scale):
    """F.interpolate(x, scale_factor=scale, mode="bilinear", align_corners=False) for NHWC CUDA tensors."""
    _need_cuda(x, "upsample_bilinear")
    return _UpsampleBilinear.apply(x, int(scale))


def upsample_bilinear_supported(x, scale):
    return (x.is_cuda and x.dim() == 4 and x.dtype in _DTYPE_CODE and scale is not None and float(scale) == int(scale)
            and bool(_lib.lib().u2b_upsample_bilinear_supported(x.shape[1], int(scale))))


# --------------------------------------------------------------------------------------
# boxes: fused IoU + Matcher, NMS
# --------------------------------------------------------------------------------------
class Matcher:
    """detectron2/modeling/matcher.py:8-127, fused with pairwise_iou (structures/boxes.py:336):
    call with the two box sets instead of the G x A quality matrix."""

    def __init__(self, thresholds, labels, allow_low_quality_matches=False):
        thresholds = list(thresholds)
        assert thresholds[0] > 0
        thresholds = [-float("inf")] + thresholds + [float("inf")]
        assert all(l <= h for l, h in zip(thresholds[:-1], thresholds[1:]))
        assert all(l in (-1, 0, 1) for l in labels) and len(labels) == len(thresholds) - 1
        self.thresholds, self.labels = thresholds, list(labels)
        self.allow_low_quality_matches = allow_low_quality_matches
        self._dev = {}

    def _consts(self, device):
        if device not in self._dev:
            self._dev[device] = (torch.tensor(self.thresholds, dtype=torch.float32, device=device),
                                 torch.tensor(self.labels, dtype=torch.int32, device=device))
        return self._dev[device]

    def match_boxes(self, gt_boxes, pred_boxes, gt_valid=None):
        """-> (matches int64 (A,), match_labels int8 (A,)) == Matcher()(pairwise_iou(gt, pred)).
        gt_valid: optional bool (G,) for fixed-capacity GT buffers (padded rows are never matched)."""
        L = _lib.lib()
        _need_cuda(pred_boxes, "Matcher")
        G, A = gt_boxes.shape[0], pred_boxes.shape[0]
        dev = pred_boxes.device
        if G == 0:  # matcher.py:80-88
            return (torch.zeros((A,), dtype=torch.int64, device=dev),
                    torch.full((A,), self.labels[0], dtype=torch.int8, device=dev))
        thr, lab = self._consts(dev)
        matches = torch.empty((A,), dtype=torch.int64, device=dev)
        vals = torch.empty((A,), dtype=torch.float32, device=dev)
        out = torch.empty((A,), dtype=torch.int8, device=dev)
        scratch = torch.empty((G,), dtype=torch.int32, device=dev) if self.allow_low_quality_matches else None
        if A > 0:
            g, p = _aligned(gt_boxes, torch.float32), _aligned(pred_boxes, torch.float32)
            gv = gt_valid.to(torch.uint8).contiguous() if gt_valid is not None else None
            _lib.check(L.u2b_iou_match(_lib.ptr(g), G, _lib.ptr(gv), _lib.ptr(p), A, _lib.ptr(thr), _lib.ptr(lab),
                                       len(self.thresholds), int(self.allow_low_quality_matches), _lib.ptr(matches),
                                       _lib.ptr(vals), _lib.ptr(out), _lib.ptr(scratch), _lib.stream_ptr()),
                       "u2b_iou_match")
            _lib.count_launches(2)
        return matches, out


def batched_nms_static(boxes, scores, idxs, iou_threshold, max_keep, valid=None):
    """batched_nms with fixed-capacity outputs and NO host synchronisation: returns (keep int64 (max_keep,) — rows
    beyond the count are 0 — and count int32 (1,) on the device). valid: optional bool (n,)."""
    L = _lib.lib()
    _need_cuda(boxes, "batched_nms")
    n = boxes.shape[0]
    dev = boxes.device
    keep = torch.zeros((max(n, max_keep),), dtype=torch.int64, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    if n == 0:
        return keep[:max_keep], cnt
    b = _aligned(boxes, torch.float32)
    sc = scores.float()
    if valid is not None:
        sc = torch.where(valid, sc, torch.full_like(sc, float("-inf")))
    order = torch.sort(sc, descending=True, stable=True)[1]
    ws_bytes = int(L.u2b_nms_workspace_bytes(n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    cats = idxs.to(torch.int64).contiguous() if idxs is not None else None
    vv = valid.to(torch.uint8).contiguous() if valid is not None else None
    _lib.check(L.u2b_batched_nms(_lib.ptr(b), _lib.ptr(cats), _lib.ptr(order), _lib.ptr(vv), n, float(iou_threshold),
                                 int(max_keep), _lib.ptr(keep), _lib.ptr(cnt), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()),
               "u2b_batched_nms")
    _lib.count_launches(3)
    return keep[:max_keep], cnt


def batched_nms(boxes, scores, idxs, iou_threshold, max_keep=None):
    """detectron2/layers/nms.py:9-21. Returns kept indices sorted by decreasing score (int64).
    The suppression scan runs on the device; the only host sync is reading the kept count."""
    L = _lib.lib()
    _need_cuda(boxes, "batched_nms")
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    b = _aligned(boxes, torch.float32)
    order = torch.sort(scores.float(), descending=True, stable=True)[1]
    keep = torch.empty((n,), dtype=torch.int64, device=boxes.device)
    cnt = torch.empty((1,), dtype=torch.int32, device=boxes.device)
    ws_bytes = int(L.u2b_nms_workspace_bytes(n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=boxes.device)
    cats = idxs.to(torch.int64).contiguous() if idxs is not None else None
    _lib.check(L.u2b_batched_nms(_lib.ptr(b), _lib.ptr(cats), _lib.ptr(order), None, n, float(iou_threshold),
                                 -1 if max_keep is None else int(max_keep), _lib.ptr(keep), _lib.ptr(cnt), _lib.ptr(ws),
                                 ws_bytes, _lib.stream_ptr()), "u2b_batched_nms")
    _lib.count_launches(3)
    return keep[:int(cnt.item())]


def nms(boxes, scores, iou_threshold):
    return batched_nms(boxes, scores, None, iou_threshold)
