"""GPU parity of the fused detection-loss kernels (csrc/det_losses.cu, csrc/mask_loss.cu, csrc/upsample.cu, the 1-CTA
weight-gradient kernel csrc/conv_wgrad_tc.cu) against their torch restatements (= the formulas of static_train.py,
rpn.py:365-429, fast_rcnn.py:307-352, mask_head.py:33-112)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _boxes(n, g, lo=8.0, hi=200.0, size=640.0):
    c = torch.rand(n, 2, generator=g) * size
    wh = torch.rand(n, 2, generator=g) * (hi - lo) + lo
    return torch.cat([c - wh / 2, c + wh / 2], 1)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_rpn_losses_match_torch(dtype, tol):
    from u2seg_b200.modeling.fused_losses import rpn_losses, rpn_losses_reference
    from u2seg_b200.modeling.rpn import Box2BoxTransform
    g = torch.Generator().manual_seed(1)
    N, A, G = 2, 20000, 7
    b2b = Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0))
    anchors, gt = _boxes(A, g).cuda(), torch.stack([_boxes(G, g), _boxes(G, g)]).cuda()
    labels = torch.randint(-1, 2, (N, A), generator=g).to(torch.int8).cuda()
    matched = torch.randint(0, G, (N, A), generator=g).cuda()
    lg = (torch.randn(N, A, generator=g) * 3).to(dtype).cuda()
    dl = (torch.randn(N, A, 4, generator=g) * 0.5).to(dtype).cuda()
    la, da = lg.clone().requires_grad_(True), dl.clone().requires_grad_(True)
    lb, db = lg.clone().requires_grad_(True), dl.clone().requires_grad_(True)
    c1, l1 = rpn_losses(la, da, anchors, labels, matched, gt, b2b.weights)
    c2, l2 = rpn_losses_reference(lb, db, anchors, labels, matched, gt, b2b)
    assert abs(float(c1) - float(c2)) <= 1e-4 * abs(float(c2)) and abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l2))
    (c1 * 0.3 + l1 * 0.7).backward()
    (c2 * 0.3 + l2 * 0.7).backward()
    for got, want in ((la.grad, lb.grad), (da.grad, db.grad)):
        assert float((got.float() - want.float()).abs().max()) <= tol * float(want.float().abs().max()) + 1e-7


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_box_losses_match_torch(dtype, tol):
    from u2seg_b200.modeling.fused_losses import box_losses, box_losses_reference
    from u2seg_b200.modeling.rpn import Box2BoxTransform
    g = torch.Generator().manual_seed(2)
    R, K = 1024, 800
    b2b = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
    props, gtb = _boxes(R, g).cuda(), _boxes(R, g).cuda()
    classes = torch.randint(0, K + 1, (R,), generator=g)
    classes[torch.rand(R, generator=g) < 0.2] = -100
    classes = classes.cuda()
    sc = (torch.randn(R, K + 1, generator=g) * 2).to(dtype).cuda()
    dl = (torch.randn(R, 4, generator=g) * 0.5).to(dtype).cuda()
    sa, da = sc.clone().requires_grad_(True), dl.clone().requires_grad_(True)
    sb, db = sc.clone().requires_grad_(True), dl.clone().requires_grad_(True)
    ce1, l11, ref1 = box_losses(sa, da, classes, props, gtb, K, b2b)
    ce2, l12, ref2 = box_losses_reference(sb, db, classes, props, gtb, K, b2b)
    assert abs(float(ce1) - float(ce2)) <= 1e-4 * abs(float(ce2)) and abs(float(l11) - float(l12)) <= 1e-4 * abs(float(l12))
    assert torch.allclose(ref1, ref2.float(), rtol=1e-5, atol=1e-3)
    (ce1 * 0.5 + l11 * 2.0).backward()
    (ce2 * 0.5 + l12 * 2.0).backward()
    for got, want in ((sa.grad, sb.grad), (da.grad, db.grad)):
        assert float((got.float() - want.float()).abs().max()) <= tol * float(want.float().abs().max()) + 1e-7


def test_static_step_with_fused_losses_equals_torch_losses(monkeypatch):
    """forward_train_static with the fused RPN / box-head loss kernels == the torch formulas: same 10 losses (1e-4) and
    the same parameter gradients (1e-3 of each tensor's largest entry), fp32, deterministic samplers."""
    from oracle import detector_oracle as do
    from test_model_gpu import _build, _make_batch
    from u2seg_b200.modeling import static_train
    K, S = 800, 28
    params = do.init_params(do.DetCfg(K, S), 0)
    data = do.synthetic_batch(2, 192, 256, K, S, seed=13, G=5, min_size=20, max_size=120)
    monkeypatch.setattr(static_train, "_rand_keys",
                        lambda mask: torch.arange(mask.numel(), device=mask.device, dtype=torch.float32) / (mask.numel() + 1))
    batch = _make_batch(data)
    out = []
    for fused in (False, True):
        monkeypatch.setattr(static_train, "FUSED_DET_LOSSES", fused)
        model = _build(K, params, True)
        packed = static_train.pack_batch(batch, torch.device("cuda"), g_max=8)
        losses, flag = static_train.forward_train_static(model, *packed)
        assert not bool(flag)
        sum(losses.values()).backward()
        out.append(({k: float(v) for k, v in losses.items()}, {n: p.grad.clone() for n, p in model.named_parameters()}))
    (la, ga), (lb, gb) = out
    for k in la:
        assert abs(la[k] - lb[k]) <= 1e-4 * abs(la[k]) + 1e-6, k
    for n in ga:
        d = float(ga[n].abs().max()) + 1e-12
        assert float((ga[n] - gb[n]).abs().max()) <= 1e-3 * d + 1e-7, n


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
def test_roi_pooler_channel_major_layout_equals_nhwc_layout(dtype, tol):
    """ROIPooler(chw_output=True) (csrc/roi_align.cu *_chw kernels) == the channels_last pooler: same values (identical
    arithmetic per bin), same feature gradients up to the order of the fp32 atomics."""
    from u2seg_b200.layers import ROIPooler
    g = torch.Generator().manual_seed(4)
    scales = (1 / 4, 1 / 8, 1 / 16, 1 / 32)
    feats = [torch.randn(2, 256, 256 // 2 ** i, 320 // 2 ** i, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
             for i in range(4)]
    boxes = [_boxes(300, g, 8, 600, 1000).clamp(0, 1000).cuda() for _ in range(2)]
    outs = []
    for chw in (False, True):
        fs = [f.clone().requires_grad_(True) for f in feats]
        pooler = ROIPooler(7, scales, 0, "ROIAlignV2", chw_output=chw)
        y = pooler(fs, boxes, grad_scale=1 / 3)
        assert y.shape == (600, 256, 7, 7) and y.is_contiguous() == chw
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).to(dtype).cuda()
        y.backward(gy)
        outs.append((y.detach().float(), [f.grad.float() for f in fs]))
    (ya, ga), (yb, gb) = outs
    assert torch.equal(ya, yb)
    for a, b in zip(ga, gb):
        assert float((a - b).abs().max()) <= max(tol, 1e-5) * float(a.abs().max()) + 1e-7


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-6), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("N,C,h,w,scale", [(2, 128, 33, 47, 2), (1, 64, 16, 16, 2), (1, 8, 5, 3, 4)])
def test_upsample_bilinear_matches_interpolate(N, C, h, w, scale, dtype, tol):
    """csrc/upsample.cu against F.interpolate(bilinear, align_corners=False) in fp32: values and input gradient."""
    import torch.nn.functional as F
    from u2seg_b200.layers import upsample_bilinear
    g = torch.Generator().manual_seed(C + h)
    x = torch.randn(N, C, h, w, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    xa = x.clone().requires_grad_(True)
    xb = x.float().clone().requires_grad_(True)
    ya = upsample_bilinear(xa, scale)
    yb = F.interpolate(xb, scale_factor=scale, mode="bilinear", align_corners=False)
    assert ya.shape == yb.shape and ya.dtype == dtype
    assert float((ya.float() - yb).abs().max()) <= tol * float(yb.abs().max()) + 1e-7
    gy = torch.randn(yb.shape, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    ya.backward(gy)
    yb.backward(gy.float())
    assert float((xa.grad.float() - xb.grad).abs().max()) <= tol * float(xb.grad.abs().max()) + 1e-6


@pytest.mark.parametrize("N,Cin,H,W,Cout,k,stride", [(2, 256, 64, 64, 256, 3, 1), (2, 128, 37, 51, 128, 3, 1), (1, 64, 32, 32, 128, 1, 1),
                                                     (2, 128, 64, 64, 128, 3, 2), (256, 256, 14, 14, 256, 3, 1)])
def test_tcgen05_wgrad_matches_library(N, Cin, H, W, Cout, k, stride):
    """csrc/conv_wgrad_tc.cu against aten.convolution_backward (fp32 reference on the same bf16 operands)."""
    from u2seg_b200.modeling.conv_tc import conv2d_nhwc_wgrad
    g = torch.Generator().manual_seed(Cin + H)
    pad = k // 2
    x = torch.randn(N, Cin, H, W, generator=g).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    gy = torch.randn(N, Cout, OH, OW, generator=g).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    w = torch.zeros(Cout, Cin, k, k, device="cuda")
    _, want, _ = torch.ops.aten.convolution_backward(gy.float(), x.float(), w, None, [stride, stride], [pad, pad], [1, 1],
                                                     False, [0, 0], 1, [False, True, False])
    got = conv2d_nhwc_wgrad(x, gy, k, k, stride, pad)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-3 * float(want.abs().max())       # fp32 accumulation, split-K order


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rpn_decode_selected_matches_torch(dtype):
    from u2seg_b200.modeling.fused_losses import rpn_decode_selected, rpn_decode_selected_reference
    from u2seg_b200.modeling.rpn import Box2BoxTransform
    g = torch.Generator().manual_seed(6)
    N, A, K = 2, 30000, 4000
    b2b = Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0))
    anchors = _boxes(A, g, 4, 400, 1024).cuda()
    deltas = (torch.randn(N, A, 4, generator=g) * 0.7).to(dtype).cuda()
    deltas[0, 5] = float("inf")
    sel = torch.stack([torch.randperm(A, generator=g)[:K] for _ in range(N)]).cuda()
    sel[0, 0] = 5                                            # a non-finite box among the selected ones
    scores = torch.randn(N, K, generator=g).cuda()
    b1, v1, f1 = rpn_decode_selected(deltas, anchors, sel, scores, b2b, (800, 1024), 0.0)
    b2, v2, f2 = rpn_decode_selected_reference(deltas, anchors, sel, scores, b2b, (800, 1024), 0.0)
    assert bool(f1) and bool(f2) and torch.equal(v1, v2)
    ok = v2
    assert torch.allclose(b1[ok], b2[ok], rtol=1e-6, atol=1e-3)


def test_cascade_relabel_matches_torch_formulas():
    """csrc/det_losses.cu cascade_relabel_kernel == the per-image clip / nonempty / Matcher / class-assignment ops of
    static_train.roi_heads_static (INT outputs bit exact, boxes exact)."""
    from u2seg_b200.layers import Matcher
    from u2seg_b200.modeling.fused_losses import cascade_relabel, cascade_relabel_reference
    g = torch.Generator().manual_seed(8)
    N, R, G, K = 2, 512, 20, 800
    gt = torch.stack([_boxes(G, g, 16, 300, 900).clamp(0, 1024) for _ in range(N)]).cuda()
    gt_classes = torch.randint(0, K, (N, G), generator=g).cuda()
    gt_valid = (torch.rand(N, G, generator=g) < 0.7).cuda()
    gt_valid[1] = False                                        # an image without ground truth
    src = gt[:, torch.randint(0, G, (R,), generator=g)]        # proposals near GT boxes -> a mix of fg / bg
    refined = (src.cpu() + torch.randn(N, R, 4, generator=g) * 25).cuda()
    refined[0, :7, 2] = refined[0, :7, 0] - 5                  # empty boxes
    ok_prev = (torch.rand(N, R, generator=g) < 0.9).cuda()
    for thr in (0.6, 0.7):
        m = Matcher([thr], [0, 1], allow_low_quality_matches=False)
        got = cascade_relabel(refined, ok_prev, gt, gt_classes, gt_valid, (1024, 1024), thr, K)
        want = cascade_relabel_reference(refined, ok_prev, gt, gt_classes, gt_valid, (1024, 1024), m, K)
        for j, what in enumerate(("boxes", "classes", "ok")):
            if not torch.equal(got[j], want[j]):
                a, b = got[j].reshape(got[j].shape[0] * got[j].shape[1], -1), want[j].reshape(want[j].shape[0] * want[j].shape[1], -1)
                bad = (a != b).any(dim=1).nonzero().flatten()
                raise AssertionError("%s differ in %d slots, e.g. slot %d: kernel %s vs torch %s (refined %s, ok_prev %s)"
                                     % (what, bad.numel(), int(bad[0]), a[bad[0]].tolist(), b[bad[0]].tolist(),
                                        refined.reshape(-1, 4)[bad[0]].tolist(), bool(ok_prev.reshape(-1)[bad[0]])))
        live = want[2] & (want[1] != K)                        # the matched GT box only matters for foreground slots
        assert torch.equal(got[3][live], want[3][live])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_mask_loss_selected_matches_torch(dtype):
    """csrc/mask_loss.cu (predictor of the GT class + BCE-with-logits + closed-form backward) vs the torch formulas
    (bmm with the gathered filter rows, F.binary_cross_entropy_with_logits, autograd)."""
    from u2seg_b200.modeling.fused_losses import mask_loss_selected, mask_loss_selected_reference
    g = torch.Generator().manual_seed(11)
    R, C, S, K = 37, 256, 28, 800
    x = (torch.randn(R, C, S, S, generator=g) * 0.5).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, C, 1, 1, generator=g) * 0.05).cuda()
    b = (torch.randn(K, generator=g) * 0.1).cuda()
    cls = torch.randint(0, K, (R,), generator=g).cuda()
    cls[5] = cls[2]
    cls[9] = cls[2]                                            # several ROIs of one class: their filter gradients add up
    tgt = (torch.rand(R, S, S, generator=g) < 0.4).cuda()
    ok = (torch.rand(R, generator=g) < 0.8).cuda()
    outs = []
    for fn in (mask_loss_selected, mask_loss_selected_reference):
        xx, ww, bb = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        loss = fn(xx, ww, bb, cls, tgt, ok)
        (loss / 1000.0).backward()
        outs.append((loss.detach().float(), xx.grad.float(), ww.grad.float(), bb.grad.float()))
    (l1, gx1, gw1, gb1), (l2, gx2, gw2, gb2) = outs
    assert abs(float(l1) - float(l2)) <= 2e-4 * abs(float(l2))
    assert float((gx1 - gx2).abs().max()) <= 1e-2 * float(gx2.abs().max())            # dX is stored in bf16 / fp16
    assert float((gw1 - gw2).abs().max()) <= 2e-3 * float(gw2.abs().max())
    assert float((gb1 - gb2).abs().max()) <= 2e-3 * float(gb2.abs().max()) + 1e-7
    assert float(gx1[~ok].abs().max()) == 0.0                                          # dead slots: exactly zero
