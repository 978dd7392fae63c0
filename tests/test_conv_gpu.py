"""GPU numerics of the tcgen05 implicit-GEMM convolution vs a plain PyTorch fp32 reference of the same
op on identical (fp16/bf16-rounded) inputs. Tolerance: 1e-3 relative to the output scale (fp32 accumulate)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (N, Cin, H, W, Cout, k, stride) — hot-path shapes of Appendix A at reduced spatial size + ragged edges
SHAPES = [
    (2, 256, 64, 64, 256, 3, 1),     # FPN output / RPN conv / mask_fcn class (BN=256, 4 stages)
    (2, 256, 32, 48, 128, 3, 1),     # sem-seg head (BN=128)
    (1, 64, 56, 56, 64, 3, 1),       # res2 conv2 (BN=64)
    (2, 64, 40, 40, 256, 1, 1),      # res2 conv3 1x1
    (2, 512, 16, 16, 2048, 1, 1),    # res5 conv3: 8 N-tiles
    (2, 1024, 20, 28, 256, 1, 1),    # fpn lateral4, ragged 20x28
    (2, 128, 33, 47, 128, 3, 2),     # stride-2 3x3 on odd sizes (res3.0.conv2)
    (2, 256, 32, 32, 512, 1, 2),     # stride-2 shortcut
    (3, 256, 14, 14, 256, 3, 1),     # mask head 14x14 ROIs (BW=16 tile with masked columns)
    (1, 256, 200, 136, 256, 3, 1),   # inference-like ragged map
]


def _ref(x, w, b, stride, pad, res, relu):
    torch.backends.cudnn.allow_tf32 = False
    y = F.conv2d(x.float(), w.float(), b, stride, pad)
    if res is not None:
        y = y + res.float()
    return F.relu(y) if relu else y


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", SHAPES)
def test_conv_forward_matches_fp32_reference(shape, dtype, cluster):
    from u2seg_b200.modeling.conv_tc import conv2d_nhwc, set_cluster
    set_cluster(cluster)
    N, Cin, H, W, Cout, k, stride = shape
    g = torch.Generator(device="cuda").manual_seed(hash(shape) % 1000)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).to(dtype)
    b = torch.randn(Cout, device="cuda", generator=g)
    pad = k // 2
    for bias, res, relu in ((None, False, False), (b, False, True), (b, True, True)):
        OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        r = torch.randn(N, Cout, OH, OW, device="cuda", generator=g).to(dtype).contiguous(memory_format=torch.channels_last) if res else None
        y = conv2d_nhwc(x, w.permute(0, 2, 3, 1).contiguous(), stride, pad, bias, r, relu)
        want = _ref(x, w, bias, stride, pad, r, relu)
        assert y.shape == want.shape and y.dtype == dtype
        err = (y.float() - want).abs().max().item()
        scale = want.abs().max().item()
        tol = (1e-3 if dtype == torch.float16 else 8e-3) * scale      # output rounding of the storage dtype
        assert err <= tol, (shape, dtype, bias is not None, res, relu, err, scale)


def test_conv_autograd_and_linear():
    from u2seg_b200.modeling import conv_tc
    from u2seg_b200.modeling.backbone import Conv2d
    torch.manual_seed(0)
    m = Conv2d(256, 256, 3, padding=1, bias=True, activation=F.relu_).cuda()
    x = torch.randn(2, 256, 24, 40, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = conv_tc.try_conv(x, m)
    assert y is not None
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = m.weight.detach().bfloat16().float().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, m.bias.detach(), 1, 1))
    yr.backward(gy.float())
    sc = lambda t: t.abs().max().item()
    assert (y.float() - yr).abs().max().item() <= 8e-3 * sc(yr)
    assert (x.grad.float() - xr.grad).abs().max().item() <= 2e-2 * sc(xr.grad)
    assert (m.weight.grad.float() - wr.grad).abs().max().item() <= 2e-2 * sc(wr.grad)
    # Linear via the same kernel (box head fc1: 12544 -> 1024)
    lin = torch.nn.Linear(12544, 1024).cuda()
    a = torch.randn(300, 12544, device="cuda").bfloat16()
    got = conv_tc.linear(a, lin.weight, lin.bias, relu=True)
    want = F.relu(F.linear(a.float(), lin.weight.bfloat16().float(), lin.bias))
    assert (got.float() - want).abs().max().item() <= 8e-3 * sc(want)


@pytest.mark.parametrize("N,H,W", [(2, 64, 96), (1, 37, 51), (2, 256, 320), (1, 8, 8)])
def test_stem_conv_forward_and_weight_gradient(N, H, W):
    """csrc/stem_conv.cu (resnet.py:338-362 BasicStem.conv1, 7x7/2 pad 3, 3->64) against F.conv2d in fp32 on the
    same bf16-rounded operands. FLOAT tolerance: output is rounded to bf16 (2^-8 relative); the weight gradient is
    accumulated in fp32 (1e-3 of its largest entry)."""
    import torch.nn.functional as F
    from u2seg_b200.modeling.conv_tc import _StemConv
    g = torch.Generator().manual_seed(N * 1000 + H)
    x = torch.randn(N, 3, H, W, generator=g).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(64, 3, 7, 7, generator=g) * 0.1).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    w.requires_grad_(True)
    y = _StemConv.apply(x, w)
    wr = w.detach().float().requires_grad_(True)
    yr = F.conv2d(x.float(), wr, None, 2, 3)
    assert y.shape == yr.shape and y.dtype == torch.bfloat16
    assert float((y.float() - yr).abs().max()) <= 2 ** -7 * float(yr.abs().max()) + 1e-6
    gy = torch.randn(yr.shape, generator=g).bfloat16().cuda().contiguous(memory_format=torch.channels_last)
    y.backward(gy)
    yr.backward(gy.float())
    assert w.grad.shape == w.shape
    d = float((w.grad.float() - wr.grad).abs().max())
    assert d <= 1e-2 * float(wr.grad.abs().max()), d      # w.grad itself is rounded to bf16 (2^-8)


# ---- second-generation kernel: cta_group::2 pair tiles, TMA-store epilogue, fused BN statistics (csrc/conv2.cu) ----
SHAPES2 = SHAPES + [
    (2, 256, 24, 8, 256, 3, 1),      # 3 m-tiles: odd tile count -> padding CTA in the last pair
    (1, 64, 8, 16, 64, 1, 1),        # a single 128-pixel tile: one pair, the peer only pads
    (1, 1024, 1, 300, 1024, 1, 1),   # Linear as a (1,1,M,K) image, ragged M
    (2, 2048, 10, 12, 512, 1, 1),    # res5 conv1: K = 2048, small M -> narrow tiles
]


@pytest.mark.parametrize("tile_n", [0, 64])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", SHAPES2)
def test_conv2_forward_and_bn_statistics(shape, dtype, tile_n):
    from u2seg_b200.modeling.conv_tc import conv2_nhwc, set_tile_n
    N, Cin, H, W, Cout, k, stride = shape
    set_tile_n(tile_n)
    try:
        g = torch.Generator(device="cuda").manual_seed(hash(shape) % 1000)
        x = torch.randn(N, Cin, H, W, device="cuda", generator=g).to(dtype)
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)                 # real NHWC strides (also for H == 1)
        w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).to(dtype)
        b = torch.randn(Cout, device="cuda", generator=g)
        pad = k // 2
        for bias, relu in ((None, False), (b, True)):
            y, st = conv2_nhwc(x, w.permute(0, 2, 3, 1).contiguous(), stride, pad, bias, relu, want_stats=True)
            want = _ref(x, w, bias, stride, pad, None, relu)
            assert y.shape == want.shape and y.dtype == dtype
            err = (y.float() - want).abs().max().item()
            scale = want.abs().max().item()
            tol = (1e-3 if dtype == torch.float16 else 8e-3) * scale
            assert err <= tol, (shape, dtype, bias is not None, relu, err, scale)
            # statistics are those of the tensor as stored (rounded), summed in fp32: compare with a double sum
            yd = y.double()
            s1, s2 = yd.sum(dim=(0, 2, 3)), (yd * yd).sum(dim=(0, 2, 3))
            got1, got2 = st[:, :Cout].double().sum(0), st[:, Cout:].double().sum(0)
            n = y.numel() / Cout
            assert (got1 - s1).abs().max().item() <= 1e-4 * n ** 0.5 * max(1.0, yd.abs().max().item())
            assert ((got2 - s2).abs() / (s2.abs() + 1e-6)).max().item() <= 1e-4
    finally:
        set_tile_n(0)


# ---- 2-CTA tcgen05 weight gradient (csrc/conv_wgrad2.cu) ----
WGRAD2_SHAPES = [
    # N, Cin, H, W, Cout, k, stride
    (2, 256, 64, 64, 256, 3, 1),      # FPN output / RPN / mask_fcn / res4 conv2 class: 9 taps x split-K
    (2, 256, 40, 24, 128, 3, 1),      # sem-seg head: only Cin is a multiple of 256 -> swapped operands (D = dW^T)
    (2, 512, 16, 16, 512, 3, 1),      # res5 conv2
    (2, 256, 33, 47, 256, 3, 2),      # stride 2 on odd sizes (res4.0 conv2)
    (2, 256, 32, 32, 1024, 1, 1),     # res4 conv3 1x1
    (2, 1024, 20, 28, 256, 1, 1),     # fpn lateral4 / res4 conv1, ragged
    (2, 512, 32, 32, 1024, 1, 2),     # stride-2 shortcut
    (2, 128, 40, 40, 512, 1, 1),      # res3 conv3: Cb = 128 -> BN 128
    (3, 256, 14, 14, 256, 3, 1),      # 14x14 ROI maps
    (1, 1024, 1, 300, 1024, 1, 1),    # Linear fc2 as a (1,1,M,K) image, ragged M
    (1, 12544, 1, 512, 1024, 1, 1),   # Linear fc1
]


@pytest.mark.parametrize("shape", WGRAD2_SHAPES)
def test_conv_wgrad2_matches_fp32_reference(shape):
    from u2seg_b200.modeling.conv_tc import conv_wgrad2, wgrad2_supported
    N, Cin, H, W, Cout, k, stride = shape
    pad = k // 2
    g = torch.Generator(device="cuda").manual_seed(hash(shape) % 1000)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g).bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    gy = torch.randn(N, Cout, OH, OW, device="cuda", generator=g).bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    assert wgrad2_supported(x, Cout, k, k, stride, pad)
    w = torch.zeros(Cout, Cin, k, k, device="cuda")
    torch.backends.cudnn.allow_tf32 = False
    _, want, _ = torch.ops.aten.convolution_backward(gy.float(), x.float(), w, None, [stride, stride], [pad, pad], [1, 1],
                                                     False, [0, 0], 1, [False, True, False])
    got = conv_wgrad2(x, gy, k, k, stride, pad, torch.float32)
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-3 * float(want.abs().max())       # fp32 accumulation, split-K order
    got16 = conv_wgrad2(x, gy, k, k, stride, pad, torch.bfloat16)
    assert float((got16.float() - want).abs().max()) <= 8e-3 * float(want.abs().max())


@pytest.mark.parametrize("shape", [s for s in SHAPES2 if s[6] == 1 and s[1] % 128 == 0])
def test_conv2_dgrad_reads_the_forward_filter_in_place(shape):
    """dX = conv(dY, rot180(W)^T) on the 2-CTA kernel with the forward (Cout,R,S,Cin) filter as an MN-major operand."""
    from u2seg_b200.modeling.conv_tc import conv2_nhwc_dgrad
    N, Cin, H, W, Cout, k, stride = shape
    pad = k // 2
    g = torch.Generator(device="cuda").manual_seed(hash(shape) % 1000 + 1)
    gy = torch.randn(N, Cout, H, W, device="cuda", generator=g).bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cout * k * k) ** 0.5).bfloat16()
    torch.backends.cudnn.allow_tf32 = False
    want = torch.ops.aten.convolution_backward(gy.float(), torch.zeros(N, Cin, H, W, device="cuda"), w.float(), None, [1, 1],
                                               [pad, pad], [1, 1], False, [0, 0], 1, [True, False, False])[0]
    got = conv2_nhwc_dgrad(gy, w.permute(0, 2, 3, 1).contiguous(), pad)
    assert got.shape == want.shape
    assert float((got.float() - want).abs().max()) <= 8e-3 * float(want.abs().max())


def test_deconv2x2_forward_and_backward_match_torch():
    """ConvTranspose2d(256, 256, 2, stride 2) + ReLU of the mask head (mask_head.py:256-262) on the tcgen05 kernels:
    forward (four interleaved GEMMs), input gradient (2x2 / stride-2 conv), weight gradient (wgrad2), bias gradient."""
    from u2seg_b200.modeling import conv_tc
    torch.manual_seed(0)
    m = torch.nn.ConvTranspose2d(256, 256, 2, stride=2).cuda().to(memory_format=torch.channels_last)
    x = torch.randn(5, 256, 14, 14, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = conv_tc.deconv2x2(x, m, relu=True)
    assert y is not None and y.shape == (5, 256, 28, 28)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = m.weight.detach().bfloat16().float().requires_grad_(True)
    br = m.bias.detach().clone().requires_grad_(True)
    yr = F.relu(F.conv_transpose2d(xr, wr, br, stride=2))
    yr.backward(gy.float())
    sc = lambda t: t.abs().max().item()      # noqa: E731
    assert (y.float() - yr).abs().max().item() <= 8e-3 * sc(yr)
    assert (x.grad.float() - xr.grad).abs().max().item() <= 2e-2 * sc(xr.grad)
    assert (m.weight.grad.float() - wr.grad).abs().max().item() <= 2e-2 * sc(wr.grad)
    assert (m.bias.grad.float() - br.grad).abs().max().item() <= 2e-2 * sc(br.grad)
