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
            gy = gy * (y > 0).to(gy.dtype)
        gy = _nhwc(gy.to(dt))
        gx = gw = gb = None
        R = weight.shape[2]
        if ctx.needs_input_grad[0] and stride == 1:
            # dX = conv(dY, rot180(W)^T), same padding for 1x1/3x3 'same' convs: the forward kernel again
            wt = weight.detach().to(dt).flip(2, 3).permute(1, 2, 3, 0).contiguous()  # (Cin,R,S,Cout)
            gx = conv2d_nhwc(gy, wt, 1, R - 1 - pad, None, None, False)
        need_gx_lib = ctx.needs_input_grad[0] and gx is None
        mask = [need_gx_lib, ctx.needs_input_grad[1], False]
        if mask[0] or mask[1]:
            w_dt = weight.detach().to(dt)
            g_in, g_w, _ = torch.ops.aten.convolution_backward(gy, xc, w_dt, None, [stride, stride], [pad, pad], [1, 1],
                                                               False, [0, 0], 1, mask)
            if need_gx_lib:
                gx = g_in
            if mask[1]:
                gw = g_w.to(weight.dtype)
        if has_bias and ctx.needs_input_grad[2]:
            gb = gy.float().sum(dim=(0, 2, 3))
        return gx, gw, gb, None, None, None


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
