"""Fused detection losses over libu2b200 (csrc/det_losses.cu): one kernel computes the summed loss and its closed-form
gradient with respect to the head outputs; autograd only multiplies by the upstream scalar.

Used by the static-shape training step (modeling/static_train.py, `U2B_FUSED_DET_LOSSES=0` falls back to the torch
formulas); tests/test_fused_losses_gpu.py compares every kernel with the torch restatements below."""
import ctypes

import torch
import torch.nn.functional as F

from .. import _lib

_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _w4(weights):
    return (ctypes.c_float * 4)(*[float(w) for w in weights])


class _RPNLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, deltas, anchors, labels, matched, gt_boxes, weights):
        L = _lib.lib()
        N, A = logits.shape
        lg, dl = logits.contiguous(), deltas.contiguous()
        assert dl.dtype == lg.dtype and dl.shape == (N, A, 4)
        need_l, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_l = torch.empty(lg.shape, dtype=torch.float32, device=lg.device) if need_l else None
        g_d = torch.empty(dl.shape, dtype=torch.float32, device=lg.device) if need_d else None
        parts = torch.empty((int(L.u2b_rpn_losses_num_partials(N * A)), 2), dtype=torch.float32, device=lg.device)
        # converted operands are bound to names for the duration of the call: a temporary handed to _p() inline is freed
        # before the launch and its block may be recycled by the NEXT inline conversion (stream-ordered, so its kernel
        # would overwrite this operand before ours reads it)
        an, lb, mt, gb = (anchors.float().contiguous(), labels.to(torch.int8).contiguous(),
                          matched.to(torch.int64).contiguous(), gt_boxes.float().contiguous())
        _lib.check(L.u2b_rpn_losses(_CODE[lg.dtype], _p(lg), _p(dl), _p(an), _p(lb), _p(mt), _p(gb), N, A, gt_boxes.shape[1],
                                    _w4(weights),
                                    _p(g_l), _p(g_d), _p(parts), _lib.stream_ptr()), "u2b_rpn_losses")
        _lib.count_launches(1)
        tot = parts.sum(0)
        ctx.save_for_backward(g_l if need_l else torch.empty(0), g_d if need_d else torch.empty(0))
        ctx.need = (need_l, need_d, lg.dtype)
        return tot[0], tot[1]

    @staticmethod
    def backward(ctx, g_cls, g_loc):
        g_l, g_d = ctx.saved_tensors
        need_l, need_d, dt = ctx.need
        return ((g_l * g_cls).to(dt) if need_l else None, (g_d * g_loc).to(dt) if need_d else None,
                None, None, None, None, None)


def rpn_losses(logits, deltas, anchors, labels, matched, gt_boxes, weights):
    """(sum of BCE over labels >= 0, sum of L1 over labels == 1): rpn.py:365-429 before the 1/normalizer and weights.
    logits (N,A), deltas (N,A,4), anchors (A,4), labels (N,A) in {-1,0,1}, matched (N,A) index into gt_boxes (N,G,4)."""
    return _RPNLosses.apply(logits, deltas, anchors, labels, matched, gt_boxes, tuple(weights))


def rpn_losses_reference(logits, deltas, anchors, labels, matched, gt_boxes, box2box):
    """torch restatement (what static_train.rpn_static computes today)."""
    N = logits.shape[0]
    tgt = torch.stack([box2box.get_deltas(anchors, gt_boxes[n][matched[n]]) for n in range(N)])
    pos = labels == 1
    d = (deltas.float() - tgt).abs()
    loc = torch.where(pos[..., None], d, torch.zeros((), dtype=d.dtype, device=d.device)).sum()
    valid = labels >= 0
    cls = F.binary_cross_entropy_with_logits(logits.float(), labels.to(torch.float32), weight=valid.to(torch.float32),
                                             reduction="sum")
    return cls, loc


class _BoxLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, deltas, classes, proposals, gt_boxes, num_fg_classes, weights, scale_clamp):
        L = _lib.lib()
        R, C = scores.shape
        sc, dl = scores.contiguous(), deltas.contiguous()
        assert dl.dtype == sc.dtype and dl.shape == (R, 4)
        need_s, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_s = torch.empty(sc.shape, dtype=torch.float32, device=sc.device) if need_s else None
        g_d = torch.empty(dl.shape, dtype=torch.float32, device=sc.device) if need_d else None
        refined = torch.empty((R, 4), dtype=torch.float32, device=sc.device)
        parts = torch.empty((int(L.u2b_box_losses_num_partials(R)), 2), dtype=torch.float32, device=sc.device)
        cl, pr, gb = classes.to(torch.int64).contiguous(), proposals.float().contiguous(), gt_boxes.float().contiguous()
        _lib.check(L.u2b_box_losses(_CODE[sc.dtype], _p(sc), _p(cl), _p(dl), _p(pr), _p(gb), R, C,
                                    int(num_fg_classes), _w4(weights), float(scale_clamp), _p(g_s), _p(g_d), _p(refined),
                                    _p(parts), _lib.stream_ptr()), "u2b_box_losses")
        _lib.count_launches(1)
        tot = parts.sum(0)
        ctx.save_for_backward(g_s if need_s else torch.empty(0), g_d if need_d else torch.empty(0))
        ctx.need = (need_s, need_d, sc.dtype)
        ctx.mark_non_differentiable(refined)
        return tot[0], tot[1], refined

    @staticmethod
    def backward(ctx, g_ce, g_l1, _g_refined):
        g_s, g_d = ctx.saved_tensors
        need_s, need_d, dt = ctx.need
        return ((g_s * g_ce).to(dt) if need_s else None, (g_d * g_l1).to(dt) if need_d else None,
                None, None, None, None, None, None)


def box_losses(scores, deltas, classes, proposals, gt_boxes, num_fg_classes, box2box):
    """(sum of CE over rows with class != -100, sum of L1 over foreground rows, refined boxes): fast_rcnn.py:307-352 +
    cascade_rcnn.py:271-299, class-agnostic regression. The refined boxes carry no gradient (the cascade detaches
    them, cascade_rcnn.py:292)."""
    return _BoxLosses.apply(scores, deltas, classes, proposals, gt_boxes, num_fg_classes, tuple(box2box.weights),
                            box2box.scale_clamp)


def box_losses_reference(scores, deltas, classes, proposals, gt_boxes, num_fg_classes, box2box):
    ce = F.cross_entropy(scores.float(), classes, reduction="sum", ignore_index=-100)
    fg = (classes >= 0) & (classes < num_fg_classes)
    d = (deltas.float() - box2box.get_deltas(proposals, gt_boxes)).abs()
    l1 = torch.where(fg[:, None], d, torch.zeros((), dtype=d.dtype, device=d.device)).sum()
    return ce, l1, box2box.apply_deltas(deltas, proposals)


def rpn_decode_selected(deltas, anchors, sel, scores, box2box, image_size, min_size):
    """Decode + clip + validity of the anchors selected by the per-level top-k (rpn.py:497-533,
    proposal_utils.py:85-121) in one kernel. deltas (N,A,4), anchors (A,4), sel (N,Ksel) int64, scores (N,Ksel).
    Returns boxes (N,Ksel,4) fp32, valid (N,Ksel) bool, nonfinite (0-dim bool tensor)."""
    L = _lib.lib()
    N, A = deltas.shape[0], deltas.shape[1]
    Ksel = sel.shape[1]
    dl = deltas.contiguous()
    boxes = torch.empty((N, Ksel, 4), dtype=torch.float32, device=dl.device)
    valid = torch.empty((N, Ksel), dtype=torch.uint8, device=dl.device)
    nonfinite = torch.zeros((), dtype=torch.int32, device=dl.device)
    h, w = image_size
    an, se, sc = anchors.float().contiguous(), sel.to(torch.int64).contiguous(), scores.float().contiguous()
    _lib.check(L.u2b_rpn_decode_selected(_CODE[dl.dtype], _p(dl), _p(an), _p(se), _p(sc), N, A, Ksel,
                                         _w4(box2box.weights), float(box2box.scale_clamp), float(h), float(w),
                                         float(min_size), _p(boxes), _p(valid), _p(nonfinite), _lib.stream_ptr()),
               "u2b_rpn_decode_selected")
    _lib.count_launches(1)
    return boxes, valid.view(torch.bool), nonfinite != 0


def rpn_decode_selected_reference(deltas, anchors, sel, scores, box2box, image_size, min_size):
    N, Ksel = sel.shape
    dsel = torch.gather(deltas, 1, sel[:, :, None].expand(-1, -1, 4))
    b = box2box.apply_deltas(dsel.reshape(-1, 4), anchors[sel].reshape(-1, 4)).view(N, Ksel, 4)
    finite = torch.isfinite(b).all(dim=2) & torch.isfinite(scores)
    h, w = image_size
    b = torch.stack((b[..., 0].clamp(0, w), b[..., 1].clamp(0, h), b[..., 2].clamp(0, w), b[..., 3].clamp(0, h)), dim=-1)
    valid = finite & ((b[..., 2] - b[..., 0]) > min_size) & ((b[..., 3] - b[..., 1]) > min_size)
    return b, valid, ~finite.all()


def cascade_relabel(refined, ok_prev, gt_boxes, gt_classes, gt_valid, image_size, iou_thr, num_classes):
    """cascade_rcnn.py:193-236,271-299 for stage k > 0 on fixed-capacity slots, all images in one launch.
    refined (N,R,4), ok_prev (N,R) bool, gt_boxes (N,G,4), gt_classes (N,G), gt_valid (N,G) bool ->
    boxes (N,R,4), classes (N,R) int64 (K background, -100 dead), ok (N,R) bool, matched GT boxes (N,R,4)."""
    L = _lib.lib()
    N, R = refined.shape[0], refined.shape[1]
    G = gt_boxes.shape[1]
    dev = refined.device
    boxes = torch.empty((N, R, 4), dtype=torch.float32, device=dev)
    classes = torch.empty((N, R), dtype=torch.int64, device=dev)
    ok = torch.empty((N, R), dtype=torch.uint8, device=dev)
    gtb = torch.empty((N, R, 4), dtype=torch.float32, device=dev)
    h, w = image_size
    rf, okp, gb, gc, gv = (refined.float().contiguous(), ok_prev.to(torch.uint8).contiguous(), gt_boxes.float().contiguous(),
                           gt_classes.to(torch.int64).contiguous(), gt_valid.to(torch.uint8).contiguous())
    _lib.check(L.u2b_cascade_relabel(_p(rf), _p(okp), _p(gb), _p(gc), _p(gv), N, R, G, float(h), float(w), float(iou_thr),
                                     int(num_classes), _p(boxes), _p(classes), _p(ok), _p(gtb), _lib.stream_ptr()),
               "u2b_cascade_relabel")
    _lib.count_launches(1)
    return boxes, classes, ok.view(torch.bool), gtb


def cascade_relabel_reference(refined, ok_prev, gt_boxes, gt_classes, gt_valid, image_size, matcher, num_classes):
    """the per-image torch formulas of static_train.roi_heads_static (stage k > 0)."""
    from .static_train import _clip, _nonempty
    K = num_classes
    dev = refined.device
    dummy = torch.cat([torch.zeros(2, device=dev), torch.ones(2, device=dev)])
    nb, nc, nok, ngb = [], [], [], []
    for n in range(refined.shape[0]):
        b = _clip(refined[n], image_size)
        ok = ok_prev[n] & _nonempty(b)
        b = torch.where(ok[:, None], b, dummy)
        midx, lab = matcher.match_boxes(gt_boxes[n], b, gt_valid=gt_valid[n])
        cls = gt_classes[n][midx]
        cls = torch.where(lab == 0, torch.full_like(cls, K), cls)
        cls = torch.where(gt_valid[n].any(), cls, torch.full_like(cls, K))
        nb.append(b)
        nc.append(torch.where(ok, cls, torch.full_like(cls, -100)))
        nok.append(ok)
        ngb.append(gt_boxes[n][midx])
    return torch.stack(nb), torch.stack(nc), torch.stack(nok), torch.stack(ngb)


class _MaskLossSelected(torch.autograd.Function):
    """sum over live ROIs and pixels of BCE-with-logits of the GT-class mask logits (mask_head.py:33-112), predictor
    included: csrc/mask_loss.cu. Returns the loss SUM (the caller divides by the live pixel count)."""

    @staticmethod
    def forward(ctx, x, weight, bias, classes, target, ok):
        L = _lib.lib()
        R, C, S, _ = x.shape
        P = S * S
        xr = x.permute(0, 2, 3, 1)                                   # NHWC storage -> (R, S, S, C) rows
        if not xr.is_contiguous():
            xr = xr.contiguous()
        K = weight.shape[0]
        w = weight.detach().reshape(K, C).to(x.dtype).contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        cls = classes.to(torch.int64).contiguous()
        tg = target.reshape(R, P).to(torch.uint8).contiguous()
        okb = ok.to(torch.uint8).contiguous()
        g = torch.empty((R, P), dtype=torch.float32, device=x.device)
        per = torch.empty((R, int(L.u2b_mask_loss_num_partials())), dtype=torch.float32, device=x.device)
        _lib.check(L.u2b_mask_loss_fwd(_CODE[x.dtype], _p(xr), _p(w), _p(b), _p(cls), _p(tg), _p(okb), R, P, C, _p(g), _p(per),
                                       _lib.stream_ptr()), "u2b_mask_loss_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(xr, w, cls, g)
        ctx.meta = (weight.shape, weight.dtype, bias is not None, bias.dtype if bias is not None else None, K)
        return per.sum()

    @staticmethod
    def backward(ctx, gout):
        L = _lib.lib()
        xr, w, cls, g = ctx.saved_tensors
        wshape, wdt, has_bias, bdt, K = ctx.meta
        R, S, _, C = xr.shape
        P = S * S
        dx = torch.empty_like(xr)
        dw = torch.zeros((K, C), dtype=torch.float32, device=xr.device)
        db = torch.zeros((K,), dtype=torch.float32, device=xr.device)
        ws = torch.empty((R * (C + 1),), dtype=torch.float32, device=xr.device)
        up = gout.detach().float().reshape(1).contiguous()
        _lib.check(L.u2b_mask_loss_bwd(_CODE[xr.dtype], _p(xr), _p(w), _p(cls), _p(g), _p(up), R, P, C, _p(dx), _p(dw), _p(db),
                                       _p(ws), _lib.stream_ptr()), "u2b_mask_loss_bwd")
        _lib.count_launches(2)
        return (dx.permute(0, 3, 1, 2), dw.to(wdt).reshape(wshape), db.to(bdt) if has_bias else None, None, None, None)


def mask_loss_supported(x):
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16) and x.shape[0] > 0
            and bool(_lib.lib().u2b_mask_loss_supported(int(x.shape[1]))))


def mask_loss_selected(x, weight, bias, classes, target, ok):
    """x (R,C,S,S) features entering the mask predictor (channels_last), weight (K,C,1,1), bias (K), classes (R),
    target (R,S,S) bool, ok (R) bool -> sum over live ROIs / pixels of BCE(logit of the ROI's class, target)."""
    return _MaskLossSelected.apply(x, weight, bias, classes, target, ok)


def mask_loss_selected_reference(x, weight, bias, classes, target, ok):
    """the torch formulas of static_train._mask_branch_static (forward_selected + BCE + masking)."""
    R, C, S, _ = x.shape
    rows = x.permute(0, 2, 3, 1).reshape(R, S * S, C).float()
    w = weight.reshape(-1, C).to(x.dtype).float()[classes]
    z = torch.bmm(rows, w.unsqueeze(2)).squeeze(2) + bias.float()[classes][:, None]
    bce = F.binary_cross_entropy_with_logits(z.view(R, S, S), target.to(torch.float32), reduction="none")
    return (bce * ok[:, None, None].to(bce.dtype)).sum()
