"""Host wrapper of the tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) for backbone.Conv2d modules
and nn.Linear. Forward (and the stride-1 input gradient, which is the same kernel on the flipped,
transposed filter) run in libu2b200; the weight gradient still uses the library kernel in round 1."""
import ctypes

import torch
import torch.nn.functional as F

from .. import _lib

_CODE = {torch.float16: 1, torch.bfloat16: 2}


def conv2d_nhwc(x, w_ohwi, stride, pad, bias=None, residual=None, relu=False):
    """x: logical (N,Cin,H,W) channels_last half tensor; w_ohwi: (Cout,R,S,Cin) contiguous, same dtype."""
    L = _lib.lib()
    N, Cin, H, W = x.shape
    Cout, R, S, _ = w_ohwi.shape
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    out = torch.empty((N, OH, OW, Cout), dtype=x.dtype, device=x.device).permute(0, 3, 1, 2)   # NHWC storage
    if out.numel() == 0:
        return out
    _lib.check(L.u2b_conv2d_nhwc_fwd(_CODE[x.dtype], ctypes.c_void_p(x.data_ptr()), N, H, W, Cin,
                                     ctypes.c_void_p(w_ohwi.data_ptr()), Cout, R, S, stride, pad,
                                     _lib.ptr(bias), ctypes.c_void_p(residual.data_ptr()) if residual is not None else None,
                                     int(relu), ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr()), "u2b_conv2d_nhwc_fwd")
    _lib.count_launches(1)
    return out


# bench.py's in-step roofline: when TIMING is a list, every tcgen05 launch below is bracketed by CUDA events on the
# launching stream and recorded as (kind, shape key, flop, start event, end event)
TIMING = None
TIMING_EXTERNAL = False    # True while the measurement step is being captured into a CUDA graph (event record NODES)


class _Timed:
    def __init__(self, kind, key, flop):
        self.on = TIMING is not None
        if self.on:
            self.rec = (kind, key, float(flop), torch.cuda.Event(enable_timing=True, external=TIMING_EXTERNAL),
                        torch.cuda.Event(enable_timing=True, external=TIMING_EXTERNAL))

    def __enter__(self):
        if self.on:
            self.rec[3].record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.rec[4].record()
            TIMING.append(self.rec)
        return False


def conv2_stats_rows(x_shape, R, S, stride, pad):
    N, _, H, W = x_shape
    return int(_lib.lib().u2b_conv2_stats_rows(N, H, W, R, S, stride, pad))


def conv2_nhwc(x, w_ohwi, stride, pad, bias=None, relu=False, want_stats=False):
    """2-CTA tcgen05 kernel (csrc/conv2.cu). x: logical (N,Cin,H,W) channels_last half tensor; w_ohwi: (Cout,R,S,Cin).
    Returns y, or (y, stats) with stats (tiles, 2*Cout) fp32 partial [sum | sumsq] rows of the rounded outputs."""
    L = _lib.lib()
    N, Cin, H, W = x.shape
    Cout, R, S, _ = w_ohwi.shape
    OH, OW = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    out = torch.empty((N, OH, OW, Cout), dtype=x.dtype, device=x.device).permute(0, 3, 1, 2)   # NHWC storage
    stats = None
    if want_stats:
        stats = torch.empty((conv2_stats_rows(x.shape, R, S, stride, pad), 2 * Cout), dtype=torch.float32, device=x.device)
    if out.numel():
        with _Timed("fwd", (N, H, W, Cin, Cout, R, stride), 2.0 * N * OH * OW * Cout * Cin * R * S):
            _lib.check(L.u2b_conv2_nhwc_fwd(_CODE[x.dtype], ctypes.c_void_p(x.data_ptr()), N, H, W, Cin,
                                            ctypes.c_void_p(w_ohwi.data_ptr()), Cout, R, S, stride, pad, _lib.ptr(bias),
                                            int(relu), ctypes.c_void_p(out.data_ptr()), _lib.ptr(stats), _lib.stream_ptr()),
                       "u2b_conv2_nhwc_fwd")
        _lib.count_launches(1)
    return (out, stats) if want_stats else out


def conv2_dgrad_supported(gy, weight, stride, pad):
    Cout, Cin, R, S = weight.shape
    return (USE_CONV2 and gy.is_cuda and gy.dtype in _CODE
            and bool(_lib.lib().u2b_conv2_dgrad_supported(Cin, Cout, R, S, stride, pad)))


def conv2_nhwc_dgrad(gy, w_ohwi, pad):
    """dX of a stride-1 'same' conv on the 2-CTA kernel, reading the FORWARD filter (Cout,R,S,Cin) in place.
    gy: logical (N,Cout,H,W) channels_last half tensor. Returns logical (N,Cin,H,W), NHWC storage."""
    L = _lib.lib()
    N, Cout, H, W = gy.shape
    _, R, S, Cin = w_ohwi.shape
    dx = torch.empty((N, H, W, Cin), dtype=gy.dtype, device=gy.device).permute(0, 3, 1, 2)
    if dx.numel():
        with _Timed("dgrad", (N, H, W, Cin, Cout, R, 1), 2.0 * N * H * W * Cout * Cin * R * S):
            _lib.check(L.u2b_conv2_nhwc_dgrad(_CODE[gy.dtype], ctypes.c_void_p(gy.data_ptr()), N, H, W, Cout,
                                              ctypes.c_void_p(w_ohwi.data_ptr()), Cin, R, S, pad,
                                              ctypes.c_void_p(dx.data_ptr()), _lib.stream_ptr()), "u2b_conv2_nhwc_dgrad")
        _lib.count_launches(1)
    return dx


def set_tile_n(bn):
    _lib.check(_lib.lib().u2b_conv2_set_tile_n(int(bn)), "u2b_conv2_set_tile_n")


TC_WGRAD = __import__("os").environ.get("U2B_TC_WGRAD", "0") == "1"     # round-2 draft (csrc/conv_wgrad_tc.cu)


def wgrad_supported(x, weight, stride, pad):
    Cout, Cin, R, S = weight.shape
    return (x.is_cuda and x.dtype in _CODE and
            bool(_lib.lib().u2b_conv2d_wgrad_supported(Cin, Cout, R, S, stride, pad)))


def conv2d_nhwc_wgrad(x, gy, R, S, stride, pad):
    """dW of conv(x, W) given dY, on tcgen05 (round-2 draft). x (N,Cin,H,W), gy (N,Cout,OH,OW), both channels_last
    half tensors. Returns fp32 logical (Cout,Cin,R,S) with OHWI (channels_last) storage."""
    L = _lib.lib()
    N, Cin, H, W = x.shape
    Cout = gy.shape[1]
    ks = int(L.u2b_conv2d_wgrad_ksplit(N, H, W, Cin, Cout, R, S, stride, pad))
    parts = torch.empty((ks, Cout, R, S, Cin), dtype=torch.float32, device=x.device)
    _lib.check(L.u2b_conv2d_nhwc_wgrad(_CODE[x.dtype], ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(gy.data_ptr()), N, H, W,
                                       Cin, Cout, R, S, stride, pad, ctypes.c_void_p(parts.data_ptr()), _lib.stream_ptr()),
               "u2b_conv2d_nhwc_wgrad")
    _lib.count_launches(1)
    return parts.sum(0).permute(0, 3, 1, 2)


USE_WGRAD2 = __import__("os").environ.get("U2B_WGRAD2", "1") == "1"   # 2-CTA tcgen05 weight gradient (csrc/conv_wgrad2.cu)
_OUT_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


# The 2-CTA weight-gradient kernel wins where the GEMM-K (pixel) axis is long: 1.17x the library on the 154.6 GFLOP FPN /
# RPN layers, about level at 40-80 GFLOP, behind it on the small late-stage layers whose K is a few thousand pixels (two
# launches and a split-K partial round trip against ~15 us of work). Below this size the library kernel is used.
WGRAD2_MIN_GFLOP = float(__import__("os").environ.get("U2B_WGRAD2_MIN_GF", "35"))


def wgrad2_supported(x, Cout, R, S, stride, pad, gy_hw=None):
    if not (USE_WGRAD2 and x.is_cuda and x.dtype in _CODE
            and bool(_lib.lib().u2b_conv_wgrad2_supported(int(x.shape[1]), int(Cout), R, S, stride, pad))):
        return False
    if gy_hw is not None:
        gf = 2.0 * x.shape[0] * gy_hw[0] * gy_hw[1] * Cout * x.shape[1] * R * S / 1e9
        return gf >= WGRAD2_MIN_GFLOP
    return True


def conv_wgrad2(x, gy, R, S, stride, pad, out_dtype=torch.float32):
    """dW of conv(x, W) given dY on the 2-CTA tcgen05 kernel. x (N,Cin,H,W), gy (N,Cout,OH,OW): channels_last half
    tensors. Returns logical (Cout,Cin,R,S) in out_dtype with OHWI (channels_last) storage."""
    L = _lib.lib()
    N, Cin, H, W = x.shape
    Cout = gy.shape[1]
    nws = int(L.u2b_conv_wgrad2_workspace_floats(N, H, W, Cin, Cout, R, S, stride, pad))
    ws = torch.empty((nws,), dtype=torch.float32, device=x.device)
    dw = torch.empty((Cout, R, S, Cin), dtype=out_dtype, device=x.device)
    with _Timed("wgrad", (N, H, W, Cin, Cout, R, stride), 2.0 * N * gy.shape[2] * gy.shape[3] * Cout * Cin * R * S):
        _lib.check(L.u2b_conv_wgrad2(_CODE[x.dtype], ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(gy.data_ptr()), N, H, W,
                                     Cin, Cout, R, S, stride, pad, ctypes.c_void_p(ws.data_ptr()), _OUT_CODE[out_dtype],
                                     ctypes.c_void_p(dw.data_ptr()), _lib.stream_ptr()), "u2b_conv_wgrad2")
    _lib.count_launches(2)
    return dw.permute(0, 3, 1, 2)


def set_cluster(cl):
    """thread-block cluster size of the conv kernel (1 = no multicast, 2, 4)."""
    _lib.check(_lib.lib().u2b_conv2d_set_cluster(int(cl)), "u2b_conv2d_set_cluster")


def _nhwc(x):
    # size-1 spatial dims make is_contiguous(channels_last) ambiguous: force real NHWC strides
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and x.stride(1) == 1:
        return x
    N, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


USE_CONV2 = __import__("os").environ.get("U2B_CONV2", "1") == "1"    # 2-CTA kernel (csrc/conv2.cu) for forward and dgrad


def _conv_fwd(xc, w_ohwi, stride, pad, bias, relu, want_stats=False):
    if USE_CONV2:
        return conv2_nhwc(xc, w_ohwi, stride, pad, bias, relu, want_stats)
    y = conv2d_nhwc(xc, w_ohwi, stride, pad, bias, None, relu)
    return (y, None) if want_stats else y


class _ConvTC(torch.autograd.Function):
    """y = [relu](conv(x, W) + b) on the tcgen05 kernels; with want_stats also the per-tile BN statistics partials of y
    (non-differentiable; consumed by fused_bn.bn_act so that the SyncBN after the conv needs no reduction pass)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, relu, want_stats=False):
        dt = x.dtype
        xc = _nhwc(x)
        w = weight.detach().to(dt).permute(0, 2, 3, 1).contiguous()              # (Cout,R,S,Cin)
        b = bias.detach().float().contiguous() if bias is not None else None
        stats = None
        if want_stats:
            y, stats = _conv_fwd(xc, w, stride, pad, b, relu, True)
        else:
            y = _conv_fwd(xc, w, stride, pad, b, relu)
        ctx.save_for_backward(xc, weight, y if relu else torch.empty(0))
        ctx.meta = (stride, pad, relu, bias is not None)
        if want_stats and stats is not None:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return (y, None) if want_stats else y

    @staticmethod
    def backward(ctx, gy, _gstats=None):
        xc, weight, y = ctx.saved_tensors
        stride, pad, relu, has_bias = ctx.meta
        dt = xc.dtype
        if relu:
            gy = torch.ops.aten.threshold_backward(gy, y, 0)       # ReLU backward, one kernel
        gy = _nhwc(gy.to(dt))
        gx = gw = gb = None
        R = weight.shape[2]
        if ctx.needs_input_grad[0] and stride == 1:
            # dX = conv(dY, rot180(W)^T), same padding for 1x1/3x3 'same' convs: the forward kernel again
            if conv2_dgrad_supported(gy, weight, stride, pad):
                # the forward filter (Cout,R,S,Cin) as it is: read MN-major, taps flipped inside the kernel
                gx = conv2_nhwc_dgrad(gy, weight.detach().to(dt).permute(0, 2, 3, 1).contiguous(), pad)
            else:
                wt = weight.detach().to(dt).flip(2, 3).permute(1, 2, 3, 0).contiguous() if R > 1 else \
                    weight.detach().to(dt).permute(1, 2, 3, 0).contiguous()           # (Cin,R,S,Cout)
                gx = _conv_fwd(gy, wt, 1, R - 1 - pad, None, False)
        need_gx_lib = ctx.needs_input_grad[0] and gx is None
        if ctx.needs_input_grad[1] and wgrad2_supported(xc, weight.shape[0], R, weight.shape[3], stride, pad, gy.shape[2:]):
            gw = conv_wgrad2(xc, gy, R, weight.shape[3], stride, pad, weight.dtype)
        elif TC_WGRAD and ctx.needs_input_grad[1] and wgrad_supported(xc, weight, stride, pad):
            gw = conv2d_nhwc_wgrad(xc, gy, weight.shape[2], weight.shape[3], stride, pad).to(weight.dtype)   # 1-CTA draft
        mask = [need_gx_lib, ctx.needs_input_grad[1] and gw is None, False]
        if mask[0] or mask[1]:
            w_dt = weight.detach().to(dt)
            g_in, g_w, _ = torch.ops.aten.convolution_backward(gy, xc, w_dt, None, [stride, stride], [pad, pad], [1, 1],
                                                               False, [0, 0], 1, mask)
            if need_gx_lib:
                gx = g_in
            if mask[1]:
                gw = g_w.to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            from .fused_bn import channel_sum
            gb = channel_sum(gy)
        return gx, gw, gb, None, None, None, None


class _StemConv(torch.autograd.Function):
    """backbone/resnet.py:338-362 BasicStem.conv1 (7x7/2, 3 -> 64) on csrc/stem_conv.cu: forward and weight
    gradient as mma.sync implicit GEMMs that stream the 64-channel side once."""

    @staticmethod
    def forward(ctx, x, weight):
        L = _lib.lib()
        xc = _nhwc(x.to(torch.bfloat16))
        N, _, H, W = xc.shape
        w = weight.detach().to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()     # (64,7,7,3)
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, OH, OW, 64), dtype=torch.bfloat16, device=x.device).permute(0, 3, 1, 2)
        _lib.check(L.u2b_stem_conv_fwd(ctypes.c_void_p(xc.data_ptr()), N, H, W, ctypes.c_void_p(w.data_ptr()),
                                       ctypes.c_void_p(y.data_ptr()), _lib.stream_ptr()), "u2b_stem_conv_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(xc)
        ctx.wmeta = (weight.dtype,)
        return y

    @staticmethod
    def backward(ctx, gy):
        (xc,) = ctx.saved_tensors
        if not ctx.needs_input_grad[1]:
            return None, None
        L = _lib.lib()
        N, _, H, W = xc.shape
        g = _nhwc(gy.to(torch.bfloat16))
        nparts = int(L.u2b_stem_conv_wgrad_num_partials(N, H, W))
        parts = torch.empty((nparts, 64, 160), dtype=torch.float32, device=xc.device)
        _lib.check(L.u2b_stem_conv_wgrad(ctypes.c_void_p(xc.data_ptr()), ctypes.c_void_p(g.data_ptr()), N, H, W,
                                         ctypes.c_void_p(parts.data_ptr()), _lib.stream_ptr()), "u2b_stem_conv_wgrad")
        _lib.count_launches(1)
        gw = parts.sum(0)[:, :147].reshape(64, 7, 7, 3).to(ctx.wmeta[0]).permute(0, 3, 1, 2)   # channels_last (64,3,7,7)
        return None, gw


def stem_eligible(x, m):
    return (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and m.bias is None and m.groups == 1
            and m.dilation == (1, 1) and not x.requires_grad
            and bool(_lib.lib().u2b_stem_conv_supported(m.in_channels, m.out_channels, m.kernel_size[0], m.kernel_size[1],
                                                        m.stride[0], m.padding[0])))


def eligible(x, m):
    if not (x.is_cuda and x.dtype in _CODE and x.dim() == 4):
        return False
    if m.groups != 1 or m.dilation != (1, 1) or m.stride[0] != m.stride[1] or m.padding[0] != m.padding[1]:
        return False
    R, S = m.kernel_size
    return bool(_lib.lib().u2b_conv2d_supported(m.in_channels, m.out_channels, R, S, m.stride[0], m.padding[0]))


def try_conv(x, m, residual=None, residual_up2x=False):
    """act(norm(conv(x)) [+ residual]) for a backbone.Conv2d module, or None when the shape is not covered."""
    from . import ops
    if not eligible(x, m):
        return None
    is_relu = m.activation in (F.relu, F.relu_)
    fuse_relu = m.norm is None and residual is None and is_relu
    if (USE_CONV2 and ops.FUSED_BN and isinstance(m.norm, torch.nn.BatchNorm2d) and (m.activation is None or is_relu)):
        from . import fused_bn
        if fused_bn.supported(x, m.norm):
            # SyncBN statistics come out of the conv epilogue (csrc/conv2.cu): no reduction pass over the activation
            y, stats = _ConvTC.apply(x, m.weight, m.bias, m.stride[0], m.padding[0], False, True)
            if m.norm.num_batches_tracked is not None and not getattr(m.norm, "_counter_batched", False):
                m.norm.num_batches_tracked.add_(1)
            return fused_bn.bn_act(y, m.norm, residual, is_relu, partials=stats, residual_up2x=residual_up2x)
    if residual_up2x:
        return None
    y = _ConvTC.apply(x, m.weight, m.bias, m.stride[0], m.padding[0], fuse_relu)
    if fuse_relu:
        return y
    return ops._norm_act(y, m, residual)


class _LinearTC(torch.autograd.Function):
    """y = [relu](x W^T + b) for nn.Linear (roi_heads/box_head.py:70,94-97): forward and input gradient on the 2-CTA
    tcgen05 kernel, the (M, K) rows viewed as a (1,1,M,K) NHWC image under a 1x1 convolution; the weight gradient
    dY^T X is a plain GEMM (library call until the tcgen05 wgrad kernel covers it)."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        M, K = x.shape
        Nout = weight.shape[0]
        dt = x.dtype
        x2 = x.contiguous()
        w = weight.detach().to(dt).contiguous()
        b = bias.detach().float().contiguous() if bias is not None else None
        y = conv2_nhwc(x2.view(1, 1, M, K).permute(0, 3, 1, 2), w.view(Nout, 1, 1, K), 1, 0, b, relu)
        y2 = y.permute(0, 2, 3, 1).reshape(M, Nout)                     # NHWC storage: a view
        ctx.save_for_backward(x2, weight, y2 if relu else torch.empty(0))
        ctx.meta = (relu, bias is not None)
        return y2

    @staticmethod
    def backward(ctx, gy):
        x2, weight, y2 = ctx.saved_tensors
        relu, has_bias = ctx.meta
        dt = x2.dtype
        M, K = x2.shape
        Nout = weight.shape[0]
        if relu:
            gy = torch.ops.aten.threshold_backward(gy, y2, 0)
        gy = gy.to(dt).contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            w4 = weight.detach().to(dt).contiguous().view(Nout, 1, 1, K)
            g4 = gy.view(1, 1, M, Nout).permute(0, 3, 1, 2)
            if K % 128 == 0:      # dX = dY W: the weight as it is, read MN-major by the kernel (no transposed copy)
                g = conv2_nhwc_dgrad(g4, w4, 0)
            else:
                g = conv2_nhwc(g4, weight.detach().to(dt).t().contiguous().view(K, 1, 1, Nout), 1, 0, None, False)
            gx = g.permute(0, 2, 3, 1).reshape(M, K)
        if ctx.needs_input_grad[1]:
            x4 = x2.view(1, 1, M, K).permute(0, 3, 1, 2)
            if wgrad2_supported(x4, Nout, 1, 1, 1, 0, (1, M)):   # dW = dY^T X with the rows as the GEMM-K (pixel) axis
                gw = conv_wgrad2(x4, gy.view(1, 1, M, Nout).permute(0, 3, 1, 2), 1, 1, 1, 0, weight.dtype).reshape(Nout, K)
            else:
                gw = torch.mm(gy.t(), x2).to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0, dtype=torch.float32)
        return gx, gw, gb, None


class _Deconv2x2(torch.autograd.Function):
    """[relu](ConvTranspose2d(k=2, s=2)(x) + b) on the tcgen05 kernels (mask_head.py:256-262): forward = four interleaved
    1x1 GEMMs (u2b_deconv2x2_nhwc_fwd); input gradient = the 2x2 / stride-2 convolution of dY with the same weight;
    weight gradient = that convolution's wgrad. The channels_last weight (physical (Cin,2,2,Cout)) is used in place."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        L = _lib.lib()
        dt = x.dtype
        xc = _nhwc(x)
        N, Cin, H, W = xc.shape
        Cout = weight.shape[1]
        w = weight.detach().to(dt).permute(0, 2, 3, 1).contiguous()             # (Cin, 2, 2, Cout)
        b = bias.detach().float().contiguous() if bias is not None else None
        y = torch.empty((N, 2 * H, 2 * W, Cout), dtype=dt, device=x.device).permute(0, 3, 1, 2)
        if y.numel():
            with _Timed("fwd", (N, H, W, Cin, 4 * Cout, 1, 1), 2.0 * N * H * W * Cin * 4 * Cout):
                _lib.check(L.u2b_deconv2x2_nhwc_fwd(_CODE[dt], ctypes.c_void_p(xc.data_ptr()), N, H, W, Cin,
                                                    ctypes.c_void_p(w.data_ptr()), Cout, _lib.ptr(b), int(relu),
                                                    ctypes.c_void_p(y.data_ptr()), _lib.stream_ptr()), "u2b_deconv2x2_nhwc_fwd")
            _lib.count_launches(4)
        ctx.save_for_backward(xc, weight, y if relu else torch.empty(0))
        ctx.meta = (relu, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        xc, weight, y = ctx.saved_tensors
        relu, has_bias = ctx.meta
        dt = xc.dtype
        if relu:
            gy = torch.ops.aten.threshold_backward(gy, y, 0)
        gy = _nhwc(gy.to(dt))
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            w = weight.detach().to(dt).permute(0, 2, 3, 1).contiguous()         # OHWI filter of the 2x2 / s2 conv: (Cin,2,2,Cout)
            gx = conv2_nhwc(gy, w, 2, 0, None, False)
        if ctx.needs_input_grad[1]:
            if wgrad2_supported(gy, weight.shape[0], 2, 2, 2, 0, xc.shape[2:]):
                gw = conv_wgrad2(gy, xc, 2, 2, 2, 0, weight.dtype)               # logical (Cin, Cout, 2, 2), channels_last storage
            else:
                gw = torch.ops.aten.convolution_backward(gy, xc, weight.detach().to(dt), None, [2, 2], [0, 0], [1, 1], True,
                                                         [0, 0], 1, [False, True, False])[1].to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            from .fused_bn import channel_sum
            gb = channel_sum(gy)
        return gx, gw, gb, None


def deconv2x2(x, m, relu=False):
    """nn.ConvTranspose2d(kernel 2, stride 2) module `m` (+ ReLU) on the tcgen05 kernels, or None if not covered."""
    if not (USE_CONV2 and x.is_cuda and x.dim() == 4 and x.dtype in _CODE and m.kernel_size == (2, 2) and m.stride == (2, 2)
            and m.padding == (0, 0) and m.output_padding == (0, 0) and m.groups == 1 and m.dilation == (1, 1)
            and bool(_lib.lib().u2b_deconv2x2_supported(m.in_channels, m.out_channels))):
        return None
    return _Deconv2x2.apply(x, m.weight, m.bias, relu)


def linear_eligible(x, weight):
    return (USE_CONV2 and x.is_cuda and x.dim() == 2 and x.dtype in _CODE and x.shape[1] % 64 == 0
            and weight.shape[0] % 64 == 0 and x.shape[0] > 0)


def linear(x, weight, bias, relu=False):
    """nn.Linear on the tcgen05 kernel when the shape allows (K, Nout multiples of 64), else the library."""
    if not linear_eligible(x, weight):
        y = F.linear(x, weight.to(x.dtype), bias.to(x.dtype) if bias is not None else None)
        return F.relu(y) if relu else y
    return _LinearTC.apply(x, weight, bias, relu)
