"""GPU parity of the detector ops (through the C ABI) vs reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import detector_oracle as do

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(golden_dir):
    return np.load(os.path.join(golden_dir, "detector_ops.npz"))


def T(a, dev="cuda"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def cl(x):
    return x.contiguous(memory_format=torch.channels_last)


def test_roialign_reference_known_answer():
    # tests/layers/test_roi_align.py:14-47 (aligned=True expectation), channels padded to the vector width
    from u2seg_b200.layers import ROIAlign
    inp = torch.arange(25).reshape(1, 1, 5, 5).float().repeat(1, 4, 1, 1).cuda()
    rois = torch.tensor([[0, 1, 1, 3, 3.0]]).cuda()
    out = ROIAlign((4, 4), 1.0, 0, aligned=True)(cl(inp), rois)
    want = torch.tensor([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
    for c in range(4):
        assert torch.allclose(out[0, c].cpu(), want)


@pytest.mark.parametrize("P,key", [(7, "pool_out7"), (14, "pool_out14")])
def test_pooler_matches_reference_golden(ops, P, key):
    from u2seg_b200.layers import ROIPooler, assign_boxes_to_levels_rois, convert_boxes_to_pooler_format
    feats = [cl(T(ops["pool_feat%d" % i])) for i in range(4)]
    boxes = [T(ops["pool_boxes0"]), T(ops["pool_boxes1"])]
    rois = convert_boxes_to_pooler_format(boxes)
    lv = assign_boxes_to_levels_rois(rois.contiguous(), 2, 5, 224, 4)
    assert np.array_equal(lv.cpu().numpy().astype(np.int64), ops["pool_levels"])             # INT: bit exact
    out = ROIPooler(P, (0.25, 0.125, 0.0625, 0.03125), 0, "ROIAlignV2")(feats, boxes)
    np.testing.assert_allclose(out.cpu().numpy(), ops[key], rtol=1e-3, atol=1e-4)   # north_star: within 1e-3 fp32


def test_levels_at_power_of_two_boundaries():
    from u2seg_b200.layers import assign_boxes_to_levels_rois
    sizes = [224.0 * 2 ** k for k in (-3, -2, -1, 0, 1, 2)] + [111.99999, 112.00001, 223.9999, 224.0001, 448.0, 0.0, 1e-3]
    rois = torch.tensor([[0, 3.0, 5.0, 3.0 + s, 5.0 + s] for s in sizes])
    want = do.assign_levels(rois[:, 1:])
    got = assign_boxes_to_levels_rois(rois.cuda(), 2, 5, 224, 4).cpu().long()
    assert torch.equal(got, want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float16, 1e-3), (torch.bfloat16, 8e-3)])
def test_pooler_hot_path_shapes_fwd_bwd(dtype, tol):
    """(2,256,H,W) pyramid of a 512x512 image, K=256 rois, 7x7: forward and backward vs torchvision CPU."""
    from u2seg_b200.layers import ROIPooler
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(2, 256, 512 // s, 512 // s, generator=g).to(dtype).float() for s in (4, 8, 16, 32)]
    boxes = []
    for _ in range(2):
        c = torch.rand(128, 2, generator=g) * 512
        wh = torch.exp(torch.rand(128, 2, generator=g) * 5 + 1.5)
        b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 512)
        boxes.append(b)
    fr = [f.clone().requires_grad_(True) for f in feats]
    want = do.roi_pool(fr, boxes, 7)
    gout = torch.randn(want.shape, generator=g).to(dtype).float()
    want.backward(gout)
    fg = [cl(f.to(dtype).cuda()).requires_grad_(True) for f in feats]
    out = ROIPooler(7, (0.25, 0.125, 0.0625, 0.03125), 0, "ROIAlignV2")(fg, [b.cuda() for b in boxes])
    assert out.dtype == dtype and out.shape == want.shape
    np.testing.assert_allclose(out.detach().float().cpu().numpy(), want.detach().numpy(), rtol=tol, atol=tol * 4)
    out.backward(gout.to(dtype).cuda())
    for a, b in zip(fg, fr):
        scale = float(b.grad.abs().max())
        assert float((a.grad.float().cpu() - b.grad).abs().max()) <= max(tol * 8 * scale, 1e-4)


@pytest.mark.parametrize("P", [7, 14])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_roi_align_per_roi_kernels_equal_per_bin_kernels(P, dtype):
    """csrc/roi_align.cu: the per-ROI factorised kernels (separable weight tables, one atomic per footprint pixel) against the
    per-sample kernels (torchvision's order) - same values up to fp32 summation order. Boxes: ordinary, tiny (bins < 1 px),
    elongated beyond the 64-pixel table (fallback inside the launch), partly / fully outside the image, degenerate."""
    from u2seg_b200 import _lib
    from u2seg_b200.layers import ROIPooler
    g = torch.Generator().manual_seed(5)
    feats = [cl(torch.randn(2, 64, 256 // s, 256 // s, generator=g).to(dtype).cuda()) for s in (4, 8, 16, 32)]
    c = torch.rand(150, 2, generator=g) * 256
    wh = torch.exp(torch.rand(150, 2, generator=g) * 5.5)
    b = torch.cat([c - wh / 2, c + wh / 2], 1)
    special = torch.tensor([[0, 0, 256, 256.0], [-40, -30, 20, 25], [250, 240, 300, 290], [400, 400, 420, 430], [10, 10, 10, 10],
                            [3, 100, 253, 104], [100, 2, 103, 255], [5.2, 7.9, 6.1, 8.3]])
    boxes = [torch.cat([b[:75], special]).cuda(), b[75:].cuda()]
    res = []
    for impl in (0, 3):
        _lib.check(_lib.lib().u2b_roi_align_set_impl(impl), "set_impl")
        try:
            fs = [f.clone().requires_grad_(True) for f in feats]
            out = ROIPooler(P, (0.25, 0.125, 0.0625, 0.03125), 0, "ROIAlignV2")(fs, boxes)
            gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(1)).to(dtype).cuda()
            out.backward(gout)
            res.append((out.detach().float(), [f.grad.float() for f in fs]))
        finally:
            _lib.check(_lib.lib().u2b_roi_align_set_impl(2), "set_impl")
    (o0, g0), (o1, g1) = res
    tol = 1e-5 if dtype == torch.float32 else 8e-3          # bf16: the OUTPUT rounding of a value that differs in the last fp32 bits
    assert torch.allclose(o1, o0, rtol=tol, atol=tol), float((o1 - o0).abs().max())
    for a, bb in zip(g1, g0):
        assert float((a - bb).abs().max()) <= (2e-5 if dtype == torch.float32 else 1e-2) * max(1.0, float(bb.abs().max()))


def test_feature_tap_shared_backward_equals_separate_calls():
    """4 pooling calls through one FeatureTap (one zero-fill + one cast per step) == 4 independent calls."""
    from u2seg_b200.layers import FeatureTap, ROIPooler
    g = torch.Generator().manual_seed(3)
    feats = [cl(torch.randn(2, 64, 128 // s, 128 // s, generator=g).cuda()) for s in (4, 8, 16, 32)]
    boxes = [[(torch.rand(20, 2, generator=g) * 100).repeat(1, 2).add(torch.tensor([0, 0, 20.0, 28.0])).cuda() for _ in range(2)]
             for _ in range(4)]
    pool = [ROIPooler(7, (0.25, 0.125, 0.0625, 0.03125)), ROIPooler(14, (0.25, 0.125, 0.0625, 0.03125))]
    gouts = None
    res = []
    for shared in (False, True):
        fs = [f.clone().requires_grad_(True) for f in feats]
        tap = FeatureTap(fs) if shared else None
        outs = [pool[i % 2](fs, boxes[i], tap=tap) for i in range(4)]
        if gouts is None:
            gouts = [torch.randn_like(o) for o in outs]
        torch.autograd.backward(outs, gouts)
        res.append([f.grad.clone() for f in fs])
    for a, b in zip(*res):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_roialign_empty_inputs():
    # tests/layers/test_roi_align.py:111-128 (empty boxes)
    from u2seg_b200.layers import ROIAlign
    inp = cl(torch.randn(1, 8, 10, 10).cuda()).requires_grad_(True)
    out = ROIAlign(7, 1.0, 0, aligned=True)(inp, torch.zeros(0, 5).cuda())
    assert out.shape == (0, 8, 7, 7)
    out.sum().backward()
    assert float(inp.grad.abs().sum()) == 0.0


def test_paste_masks_matches_reference(ops):
    from u2seg_b200.layers import paste_masks_in_image
    m, b = T(ops["paste_masks"]), T(ops["paste_boxes"])
    got = paste_masks_in_image(m, b, (100, 150), 0.5).cpu()
    want = torch.from_numpy(np.unpackbits(ops["paste_out"])[:9 * 100 * 150].reshape(9, 100, 150).astype(bool))
    diff = (got != want)
    if diff.any():   # FP->bool: exact except where |p - 0.5| < eps
        soft = do.paste_masks_in_image(m.cpu(), b.cpu(), (100, 150), -1)  # uint8 soft values
        assert diff.sum() <= 3 and bool(((soft[diff].int() - 127).abs() <= 1).all())


def test_paste_masks_full_size_properties():
    """config-5 size (N=100, 800x1333): pixels far outside every box are 0; a constant mask of ones fills
    exactly the pixel centres inside the box."""
    from u2seg_b200.layers import paste_masks_in_image
    N, H, W = 100, 800, 1333
    g = torch.Generator().manual_seed(1)
    c = torch.rand(N, 2, generator=g) * torch.tensor([W, H])
    wh = torch.rand(N, 2, generator=g) * 300 + 8
    boxes = torch.cat([c - wh / 2, c + wh / 2], 1)
    out = paste_masks_in_image(torch.ones(N, 28, 28).cuda(), boxes.cuda(), (H, W), 0.5).cpu()
    ys = (torch.arange(H) + 0.5)[None, :, None]
    xs = (torch.arange(W) + 0.5)[None, None, :]
    b = boxes[:, :, None, None]
    # bilinear with zero padding: 1 at least half a mask-pixel inside the box, 0 more than half a mask-pixel outside
    hw, hh = (b[:, 2] - b[:, 0]) / 56, (b[:, 3] - b[:, 1]) / 56
    inside = (xs >= b[:, 0] + hw + 1e-3) & (xs <= b[:, 2] - hw - 1e-3) & (ys >= b[:, 1] + hh + 1e-3) & (ys <= b[:, 3] - hh - 1e-3)
    far = (xs < b[:, 0] - hw - 1e-3) | (xs > b[:, 2] + hw + 1e-3) | (ys < b[:, 1] - hh - 1e-3) | (ys > b[:, 3] + hh + 1e-3)
    assert bool(out[inside].all()) and not bool(out[far].any())


def test_crop_and_resize_matches_reference(ops):
    from u2seg_b200.layers import crop_and_resize_masks
    gm = T(np.unpackbits(ops["crop_masks"])[:6 * 100 * 150].reshape(6, 100, 150).astype(bool))
    got, val = crop_and_resize_masks(gm, T(ops["crop_boxes"]), 28, return_values=True)
    want = np.unpackbits(ops["crop_out"])[:6 * 28 * 28].reshape(6, 28, 28).astype(bool)
    diff = got.cpu().numpy() != want
    assert diff.sum() <= 2 and bool((np.abs(val.cpu().numpy()[diff] - 0.5) < 1e-5).all())
    # gather indirection == explicit gather
    idx = torch.tensor([5, 0, 0, 3, 2, 1, 4, 4]).cuda()
    bx = T(ops["crop_boxes"])[torch.tensor([0, 1, 2, 3, 4, 5, 0, 1])]
    assert torch.equal(crop_and_resize_masks(gm, bx, 28, gt_index=idx), crop_and_resize_masks(gm[idx], bx, 28))


def test_iou_matcher_matches_reference(ops):
    from u2seg_b200.layers import Matcher
    gt, an = T(ops["iou_gt"]), T(ops["iou_an"])
    m, l = Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True).match_boxes(gt, an)
    assert np.array_equal(m.cpu().numpy(), ops["match_rpn_idx"]) and np.array_equal(l.cpu().numpy(), ops["match_rpn_lab"])
    m, l = Matcher([0.5], [0, 1], allow_low_quality_matches=False).match_boxes(gt, an)
    assert np.array_equal(m.cpu().numpy(), ops["match_roi_idx"]) and np.array_equal(l.cpu().numpy(), ops["match_roi_lab"])


def test_iou_matcher_full_anchor_set_vs_oracle():
    from u2seg_b200.layers import Matcher
    cfg = do.DetCfg()
    anchors = torch.cat(do.make_anchors([(256 // s, 256 // s) for s in (1, 2, 4, 8, 16)], cfg))   # 1024^2 image: 261,888
    assert anchors.shape[0] == 261888
    _, boxes, _, _, _ = do.synthetic_batch(1, 64, 64, 800, 28, seed=3, G=20)
    gt = boxes[0] * 16
    iou = do.pairwise_iou(gt, anchors)
    wm, wl = do.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
    m, l = Matcher([0.3, 0.7], [0, -1, 1], True).match_boxes(gt.cuda(), anchors.cuda())
    assert torch.equal(m.cpu(), wm) and torch.equal(l.cpu(), wl)
    # empty GT (matcher.py:80-88)
    m, l = Matcher([0.3, 0.7], [0, -1, 1], True).match_boxes(torch.zeros(0, 4).cuda(), anchors[:100].cuda())
    assert int(m.sum()) == 0 and bool((l == 0).all())


@pytest.mark.parametrize("thr,key", [(0.65, "nms_keep_065"), (0.5, "nms_keep_050")])
def test_batched_nms_matches_reference(ops, thr, key):
    from u2seg_b200.layers import batched_nms
    keep = batched_nms(T(ops["nms_boxes"]), T(ops["nms_scores"]), T(ops["nms_idxs"]), thr)
    assert np.array_equal(keep.cpu().numpy(), ops[key])                        # INT: bit exact, same order


def test_batched_nms_rpn_size_vs_oracle():
    from u2seg_b200.layers import batched_nms
    g = torch.Generator().manual_seed(4)
    n = 9000
    c = torch.rand(n, 2, generator=g) * 1024
    wh = torch.exp(torch.rand(n, 2, generator=g) * 4 + 2)
    b = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1024)
    s = torch.randn(n, generator=g)
    s[100:200] = s[0]            # score ties: stable order
    lv = torch.randint(0, 5, (n,), generator=g)
    want = do.batched_nms(b, s, lv, 0.65)
    got = batched_nms(b.cuda(), s.cuda(), lv.cuda(), 0.65).cpu()
    assert set(got.tolist()) == set(want.tolist())
    assert torch.equal(s[got], s[want])      # same score order (ties may permute among equal scores)
    assert batched_nms(b[:0].cuda(), s[:0].cuda(), lv[:0].cuda(), 0.5).numel() == 0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("C,HW,res,relu", [(64, (33, 47), False, True), (256, (16, 16), True, True), (2048, (4, 6), True, False),
                                           (512, (20, 12), False, False)])
def test_fused_bn_act_matches_torch(dtype, tol, C, HW, res, relu):
    """u2b_bn_* (stats, finalize, apply, bwd_reduce, bwd_apply) vs F.batch_norm(training) + add + relu in fp32."""
    import torch.nn.functional as F
    from u2seg_b200.modeling.backbone import SyncBatchNorm
    from u2seg_b200.modeling.fused_bn import bn_act
    g = torch.Generator().manual_seed(C + HW[0])
    N = 2
    x = (torch.randn(N, C, *HW, generator=g) * 2 + 0.5).to(dtype)
    r = torch.randn(N, C, *HW, generator=g).to(dtype) if res else None
    gy = torch.randn(N, C, *HW, generator=g).to(dtype)
    bn = SyncBatchNorm(C).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g))
        bn.bias.copy_(0.2 * torch.randn(C, generator=g))
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    xg = cl(x.cuda()).requires_grad_(True)
    rg = cl(r.cuda()).requires_grad_(True) if res else None
    y = bn_act(xg, bn, rg, relu)
    y.backward(cl(gy.cuda()))
    # fp32 reference on the same (rounded) inputs
    xr = x.float().cuda().requires_grad_(True)
    rr = r.float().cuda().requires_grad_(True) if res else None
    w, b = bn.weight.detach().clone().requires_grad_(True), bn.bias.detach().clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    yr = F.batch_norm(xr, rm, rv, w, b, True, 0.1, bn.eps)
    if res:
        yr = yr + rr
    if relu:
        yr = F.relu(yr)
    yr.backward(gy.float().cuda())
    sc = lambda t: float(t.abs().max()) + 1e-12
    assert float((y.float() - yr).abs().max()) <= tol * sc(yr)
    assert float((xg.grad.float() - xr.grad).abs().max()) <= tol * 2 * sc(xr.grad)
    if res:
        assert float((rg.grad.float() - rr.grad).abs().max()) <= tol * sc(rr.grad)
    assert float((bn.weight.grad - w.grad).abs().max()) <= tol * 2 * sc(w.grad)
    assert float((bn.bias.grad - b.grad).abs().max()) <= tol * 2 * sc(b.grad)
    assert torch.allclose(bn.running_mean, rm, rtol=1e-4, atol=1e-5) and torch.allclose(bn.running_var, rv, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("N,C,h,w,scale,dtype", [(2, 28, 37, 50, 4, torch.float32), (1, 28, 64, 64, 4, torch.bfloat16),
                                                 (2, 5, 33, 17, 2, torch.float32), (1, 40, 8, 8, 8, torch.float32)])
def test_upsample_cross_entropy_matches_interpolate_plus_cross_entropy(N, C, h, w, scale, dtype):
    """semantic_seg.py:255-267 fused: loss and d(loss)/d(logits) against F.interpolate + F.cross_entropy (fp32)."""
    from u2seg_b200.layers import upsample_cross_entropy
    g = torch.Generator().manual_seed(7)
    z = (torch.randn(N, C, h, w, generator=g) * 3).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, C, (N, h * scale, w * scale), generator=g)
    t[torch.rand(t.shape, generator=g) < 0.2] = 255          # ignored pixels
    t = t.cuda()
    za = z.clone().requires_grad_(True)
    zb = z.clone().requires_grad_(True)
    got = upsample_cross_entropy(za, t, scale, 255)
    up = F.interpolate(zb.float(), scale_factor=scale, mode="bilinear", align_corners=False)
    want = F.cross_entropy(up, t, reduction="mean", ignore_index=255)
    assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want)) + 1e-6           # FLOAT: 1e-5 relative
    (got * 0.5).backward()
    (want * 0.5).backward()
    tol = 1e-5 if dtype == torch.float32 else 1e-2         # bf16 gradient is rounded to bf16 on both sides
    d = float((za.grad.float() - zb.grad.float()).abs().max())
    assert d <= tol * float(zb.grad.float().abs().max()) + 1e-9, d
    assert za.grad.shape == z.shape and za.grad.dtype == dtype


def test_upsample_cross_entropy_all_ignored_is_nan_like_reference():
    from u2seg_b200.layers import upsample_cross_entropy
    z = torch.randn(1, 4, 8, 8).cuda().contiguous(memory_format=torch.channels_last)
    t = torch.full((1, 32, 32), 255, dtype=torch.int64).cuda()
    got = upsample_cross_entropy(z, t, 4, 255)
    want = F.cross_entropy(F.interpolate(z, scale_factor=4, mode="bilinear", align_corners=False), t, ignore_index=255)
    assert torch.isnan(got) and torch.isnan(want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("N,C,G,HW,relu", [(2, 128, 32, (33, 47), True), (3, 64, 32, (16, 16), False), (1, 256, 32, (8, 12), True)])
def test_group_norm_relu_matches_torch(N, C, G, HW, relu, dtype, tol):
    """fused NHWC GroupNorm(+ReLU) (csrc/batchnorm.cu gn_* + per-image BN kernels) against F.group_norm in fp32:
    output, dx, dgamma, dbeta. FLOAT tolerance: 2e-4 fp32, 2e-2 bf16 (activation dtype rounding)."""
    from u2seg_b200.modeling.fused_bn import gn_act, gn_supported
    g = torch.Generator().manual_seed(C + N)
    x = cl((torch.randn(N, C, *HW, generator=g) * 2 + 0.5).cuda().to(dtype))
    gn = torch.nn.GroupNorm(G, C).cuda()
    with torch.no_grad():
        gn.weight.copy_(torch.rand(C, generator=g).cuda() + 0.5)
        gn.bias.copy_(torch.randn(C, generator=g).cuda() * 0.2)
    assert gn_supported(x, gn)
    xa = x.clone().requires_grad_(True)
    y = gn_act(xa, gn, relu)
    xr = x.float().clone().requires_grad_(True)
    w2, b2 = gn.weight.detach().clone().requires_grad_(True), gn.bias.detach().clone().requires_grad_(True)
    yr = F.group_norm(xr, G, w2, b2, gn.eps)
    yr = F.relu(yr) if relu else yr
    assert y.dtype == dtype and float((y.float() - yr).abs().max()) <= tol * float(yr.abs().max())
    gy = cl(torch.randn(yr.shape, generator=g).cuda().to(dtype))
    y.backward(gy)
    yr.backward(gy.float())
    for got, want in ((xa.grad, xr.grad), (gn.weight.grad, w2.grad), (gn.bias.grad, b2.grad)):
        assert float((got.float() - want).abs().max()) <= 2 * tol * float(want.abs().max()) + 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 64, 64, 96), (1, 64, 33, 47), (2, 16, 8, 8)])
def test_maxpool3x3s2_matches_torch_including_ties(shape, dtype):
    """csrc/pool.cu vs F.max_pool2d (resnet.py:358) on post-ReLU-like data with many exact ties (zeros and repeated
    bf16 values): pooled values bit-exact, gradients bit-exact (first maximum in scan order receives the gradient)."""
    import torch.nn.functional as F
    from u2seg_b200.modeling.ops import max_pool_3x3_s2
    g = torch.Generator().manual_seed(5)
    N, C, H, W = shape
    x = (torch.randn(N, C, H, W, generator=g).clamp(min=0) * 4).round() / 4        # coarse grid: plenty of equal neighbours
    x = x.to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = max_pool_3x3_s2(xa)
    yb = F.max_pool2d(xb, kernel_size=3, stride=2, padding=1)
    assert torch.equal(ya, yb)
    gy = torch.randn(yb.shape, generator=g).to(dtype).cuda().contiguous(memory_format=torch.channels_last)
    ya.backward(gy)
    yb.backward(gy)
    assert torch.allclose(xa.grad.float(), xb.grad.float(), rtol=1e-2, atol=1e-3)      # <= 4 bf16 terms summed in fp32 vs ATen
    assert torch.equal(xa.grad != 0, xb.grad != 0)                                     # same routing, also on ties


def test_fpn_lateral_sum_folded_into_bn_matches_unfused():
    """ops.lateral_add_upsample: lateral 1x1 conv + SyncBN + nearest-x2(prev) in the BN apply pass (csrc/batchnorm.cu
    bn_apply_resup, csrc/pool.cu sum2x2) vs the unfused torch composition, forward and backward."""
    import torch.nn.functional as F
    from u2seg_b200.modeling import ops
    from u2seg_b200.modeling.backbone import Conv2d
    torch.manual_seed(0)
    lat = Conv2d(512, 256, kernel_size=1, bias=False, norm=torch.nn.BatchNorm2d(256)).cuda().train()
    feat = torch.randn(2, 512, 32, 48, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    prev = torch.randn(2, 256, 16, 24, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    outs = []
    for fused in (True, False):
        ops.FUSED_FPN_SUM = fused
        try:
            f, p = feat.clone().requires_grad_(True), prev.clone().requires_grad_(True)
            lat.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = ops.lateral_add_upsample(lat, f, p)
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).to(y.dtype)
            y.backward(gy)
            outs.append((y.float(), f.grad.float(), p.grad.float(), lat.weight.grad.float().clone()))
        finally:
            ops.FUSED_FPN_SUM = True
    for a, b in zip(*outs):
        assert float((a - b).abs().max()) <= 2e-2 * float(b.abs().max()) + 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_preprocess_u8_matches_reference_formula(dtype):
    """csrc/pool.cu preprocess_u8_kernel == rcnn.py:223-234 ((x - mean) / std in fp32) + image_list.py zero padding."""
    from u2seg_b200.modeling.ops import preprocess_u8
    g = torch.Generator().manual_seed(2)
    img = torch.randint(0, 256, (2, 3, 50, 70), generator=g, dtype=torch.uint8).cuda().contiguous(memory_format=torch.channels_last)
    mean, std = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]
    got = preprocess_u8(img, mean, std, 32, dtype)
    m = torch.tensor(mean, device="cuda").view(-1, 1, 1)
    s = torch.tensor(std, device="cuda").view(-1, 1, 1)
    want = torch.nn.functional.pad((img.float() - m) / s, (0, 96 - 70, 0, 64 - 50)).to(dtype)
    assert got.shape == (2, 3, 64, 96) and got.dtype == dtype and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(got, want)                      # bit-exact: same op order, IEEE division, one rounding
