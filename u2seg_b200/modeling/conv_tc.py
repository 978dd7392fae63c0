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
        _lib.check(L.u2b_conv2_nhwc_fwd(_CODE[x.dtype], ctypes.c_void_p(x.data_ptr()), N, H, W, Cin,
                                        ctypes.c_void_p(w_ohwi.data_ptr()), Cout, R, S, stride, pad, _lib.ptr(bias), int(relu),
                                        ctypes.c_void_p(out.data_ptr()), _lib.ptr(stats), _lib.stream_ptr()),
                   "u2b_conv2_nhwc_fwd")
        _lib.count_launches(1)
    return (out, stats) if want_stats else out


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


def set_cluster(cl):
    """thread-block cluster size of the conv kernel (1 = no multicast, 2, 4)."""
    _lib.check(_lib.lib().u2b_conv2d_set_cluster(int(cl)), "u2b_conv2d_set_cluster")


def _nhwc(x):
    # size-1 spatial dims make is_contiguous(channels_last) ambiguous: force real NHWC strides
    if x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and x.stride(1) == 1:
        return x
    N, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


class _ConvTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, relu):
        dt = x.dtype
        xc = _nhwc(x)
        w = weight.detach().to(dt).permute(0, 2, 3, 1).contiguous()              # (Cout,R,S,Cin)
        b = bias.detach().float().contiguous() if bias is not None else None
        y = conv2d_nhwc(xc, w, stride, pad, b, None, relu)
        ctx.save_for_backward(xc, weight, y if relu else torch.empty(0))
        ctx.meta = (stride, pad, relu, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
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
            wt = weight.detach().to(dt).flip(2, 3).permute(1, 2, 3, 0).contiguous()  # (Cin,R,S,Cout)
            gx = conv2d_nhwc(gy, wt, 1, R - 1 - pad, None, None, False)
        need_gx_lib = ctx.needs_input_grad[0] and gx is None
        if TC_WGRAD and ctx.needs_input_grad[1] and wgrad_supported(xc, weight, stride, pad):
            gw = conv2d_nhwc_wgrad(xc, gy, weight.shape[2], weight.shape[3], stride, pad).to(weight.dtype)   # r2 draft
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
        return gx, gw, gb, None, None, None


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


def try_conv(x, m, residual=None):
    """act(norm(conv(x)) [+ residual]) for a backbone.Conv2d module, or None when the shape is not covered."""
    from . import ops
    if not eligible(x, m):
        return None
    is_relu = m.activation in (F.relu, F.relu_)
    fuse_relu = m.norm is None and residual is None and is_relu
    y = _ConvTC.apply(x, m.weight, m.bias, m.stride[0], m.padding[0], fuse_relu)
    if fuse_relu:
        return y
    return ops._norm_act(y, m, residual)


def linear(x, weight, bias, relu=False):
    """nn.Linear through the same kernel: (M,K) rows as a (1,1,M,K) NHWC image, 1x1 conv."""
    M, K = x.shape
    if not (x.is_cuda and x.dtype in _CODE and K % 64 == 0 and weight.shape[0] % 64 == 0):
        y = F.linear(x, weight.to(x.dtype), bias.to(x.dtype) if bias is not None else None)
        return F.relu(y) if relu else y
    x4 = x.contiguous().view(1, 1, M, K).permute(0, 3, 1, 2)         # logical (1,K,1,M), NHWC storage
    y = _ConvTC.apply(x4, weight.view(weight.shape[0], K, 1, 1), bias, 1, 0, relu)
    return y.permute(0, 2, 3, 1).reshape(M, weight.shape[0])
