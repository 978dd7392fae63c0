"""Compute dispatch for the dense layers of the model (conv / norm / activation / residual).

Round-1 state: the large 3x3 convolutions (forward + input gradient) run on the hand-written tcgen05
implicit-GEMM kernel of libu2b200 (csrc/conv_tc.cu) when activations are fp16/bf16; the remaining
convolutions, all weight gradients and the batch-norm statistics go through the library kernels of this image
(cuDNN via F.conv2d, ATen batch_norm; counted as library calls, like cuBLAS), in channels_last. ROI pooling,
NMS, matching, mask ops and k-means always run in libu2b200.
"""
import os

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

# Which convolutions run on the hand-written tcgen05 kernel (csrc/conv_tc.cu) instead of the library (cuDNN):
#   "large3x3" (round-1 default): 3x3 stride-1 convs with Cin >= 128 — FPN outputs, RPN head conv, mask head, sem-seg head:
#                         57% of the step's forward MACs; forward and input-gradient (dgrad) both use conv_tc
#                         (measured 1.35 PFLOP/s on the 256->256 3x3 at 2x256x256, 0.89x cuDNN); weight gradients
#                         stay in the library this round;
#   "all" (default since round 2, 2-CTA kernel csrc/conv2.cu): every shape the kernels support - 1x1 and 3x3, stride 1
#          and 2, Cin and Cout multiples of 64 - forward and stride-1 input gradient, plus the box-head Linear layers;
#   "none": library only.
TCGEN05_CONV_POLICY = os.environ.get("U2B_CONV_POLICY", "all")
USE_TCGEN05_CONV = True


def _use_tc(x, m):
    if not USE_TCGEN05_CONV or TCGEN05_CONV_POLICY == "none":
        return False
    if TCGEN05_CONV_POLICY == "all":
        return True
    return m.kernel_size == (3, 3) and m.stride == (1, 1) and m.in_channels >= 128 and m.out_channels >= 128


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def batch_norm(x, bn, relu=False):
    """nn.SyncBatchNorm semantics (layers/batch_norm.py:187): batch statistics over the whole
    data-parallel group in training, running statistics in eval."""
    if bn.training:
        if bn.num_batches_tracked is not None and not getattr(bn, "_counter_batched", False):
            bn.num_batches_tracked.add_(1)
        if _world() > 1:
            from torch.nn.modules._functions import SyncBatchNorm as _SBN
            y = _SBN.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum,
                           dist.group.WORLD, _world())
        else:
            y = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
    else:
        y = F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
    return F.relu_(y) if relu else y


FUSED_BN = True     # SyncBN + residual + ReLU through libu2b200 (csrc/batchnorm.cu) in training mode
FUSED_GN = True     # GroupNorm + ReLU (semantic head) through the same NHWC kernels, per image
UPSAMPLE_KERNEL = os.environ.get("U2B_UPSAMPLE_KERNEL", "1") == "1"   # NHWC bilinear x2 forward / backward (csrc/upsample.cu)
STEM_KERNEL = True  # csrc/stem_conv.cu for the 7x7/2 3->64 stem (bf16 autocast only)


def _norm_act(y, m, residual):
    """act(norm(y) [+ residual])"""
    if m.norm is not None:
        if isinstance(m.norm, nn.BatchNorm2d):
            if FUSED_BN:
                from . import fused_bn
                is_relu = m.activation in (F.relu, F.relu_)
                if fused_bn.supported(y, m.norm) and (m.activation is None or is_relu):
                    if m.norm.num_batches_tracked is not None and not getattr(m.norm, "_counter_batched", False):
                        m.norm.num_batches_tracked.add_(1)
                    return fused_bn.bn_act(y, m.norm, residual, is_relu)
            y = batch_norm(y, m.norm)
        elif FUSED_GN and isinstance(m.norm, nn.GroupNorm) and residual is None and \
                (m.activation is None or m.activation in (F.relu, F.relu_)):
            from . import fused_bn
            if fused_bn.gn_supported(y, m.norm):
                return fused_bn.gn_act(y, m.norm, m.activation is not None)
            y = m.norm(y)
        else:
            y = m.norm(y)
    if residual is not None:
        y = y + residual
    if m.activation is not None:
        y = m.activation(y)
    return y


def conv_norm_act(x, m, residual=None):
    """y = act(norm(conv(x)) [+ residual]) for a backbone.Conv2d module `m`."""
    if _use_tc(x, m):
        from . import conv_tc
        y = conv_tc.try_conv(x, m, residual)
        if y is not None:
            return y
    if STEM_KERNEL and m.in_channels == 3 and x.is_cuda:
        from . import conv_tc
        xa = x.to(torch.get_autocast_dtype("cuda")) if torch.is_autocast_enabled("cuda") else x
        if conv_tc.stem_eligible(xa, m):
            return _norm_act(conv_tc._StemConv.apply(xa, m.weight), m, residual)
    y = F.conv2d(x, m.weight, m.bias, m.stride, m.padding, m.dilation, m.groups)
    return _norm_act(y, m, residual)


def interpolate(x, **kw):
    """F.interpolate in the activation dtype. Under CUDA autocast the upsampling ops are promoted to fp32, which
    turns the FPN top-down path and the semantic head's level sum into fp32 tensors (2x the bytes of every later
    add / cast); the kernels accumulate in fp32 internally either way, so only the stored result is bf16."""
    if x.is_cuda and x.dtype in (torch.bfloat16, torch.float16) and torch.is_autocast_enabled("cuda"):
        with torch.autocast("cuda", enabled=False):
            return F.interpolate(x, **kw)
    return F.interpolate(x, **kw)


class Upsample(nn.Upsample):
    """nn.Upsample through `interpolate` above (no parameters: state_dict unchanged)."""

    def forward(self, x):
        if UPSAMPLE_KERNEL and self.mode == "bilinear" and self.align_corners is False and self.size is None:
            from ..layers import upsample_bilinear, upsample_bilinear_supported    # round-2 draft (csrc/upsample.cu)
            if upsample_bilinear_supported(x, self.scale_factor):
                return upsample_bilinear(x, self.scale_factor)
        return interpolate(x, size=self.size, scale_factor=self.scale_factor, mode=self.mode,
                           align_corners=self.align_corners)


FUSED_FPN_SUM = os.environ.get("U2B_FUSED_FPN_SUM", "1") == "1"
MAXPOOL_KERNEL = os.environ.get("U2B_MAXPOOL_KERNEL", "1") == "1"


def lateral_add_upsample(lateral, feat, prev):
    """fpn.py:153-156: lateral(feat) + F.interpolate(prev, scale_factor=2, mode='nearest'). With the tcgen05 conv +
    fused SyncBN path the sum is folded into the norm's apply pass (the upsampled map is never materialised) and the
    gradient of the upsampling is one 2x2 fold kernel."""
    if (FUSED_FPN_SUM and _use_tc(feat, lateral) and lateral.activation is None
            and prev.shape[2] * 2 == feat.shape[2] and prev.shape[3] * 2 == feat.shape[3]):
        from . import conv_tc
        y = conv_tc.try_conv(feat, lateral, prev, residual_up2x=True)
        if y is not None:
            return y
    td = interpolate(prev, scale_factor=2.0, mode="nearest")
    return lateral(feat) + td


class _MaxPool3x3S2(torch.autograd.Function):
    """F.max_pool2d(x, 3, 2, 1) on NHWC half tensors (resnet.py:358): csrc/pool.cu."""

    @staticmethod
    def forward(ctx, x):
        import ctypes
        from .. import _lib
        L = _lib.lib()
        xc = x if (x.is_contiguous(memory_format=torch.channels_last) and x.stride(1) == 1) else \
            x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        N, C, H, W = xc.shape
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, OH, OW, C), dtype=xc.dtype, device=xc.device).permute(0, 3, 1, 2)
        idx = torch.empty((N, OH, OW, C), dtype=torch.uint8, device=xc.device)
        code = 2 if xc.dtype == torch.bfloat16 else 1
        _lib.check(L.u2b_maxpool3x3s2_fwd(code, ctypes.c_void_p(xc.data_ptr()), N, H, W, C, ctypes.c_void_p(y.data_ptr()),
                                          ctypes.c_void_p(idx.data_ptr()), _lib.stream_ptr()), "u2b_maxpool3x3s2_fwd")
        _lib.count_launches(1)
        ctx.save_for_backward(idx)
        ctx.meta = (N, C, H, W, code, xc.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        import ctypes
        from .. import _lib
        (idx,) = ctx.saved_tensors
        N, C, H, W, code, dt = ctx.meta
        g = gy.to(dt)
        if not (g.is_contiguous(memory_format=torch.channels_last) and g.stride(1) == 1):
            g = g.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        dx = torch.empty((N, H, W, C), dtype=dt, device=g.device).permute(0, 3, 1, 2)
        _lib.check(_lib.lib().u2b_maxpool3x3s2_bwd(code, ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(idx.data_ptr()), N, H, W,
                                                   C, ctypes.c_void_p(dx.data_ptr()), _lib.stream_ptr()), "u2b_maxpool3x3s2_bwd")
        _lib.count_launches(1)
        return dx


def max_pool_3x3_s2(x):
    """F.max_pool2d(x, kernel_size=3, stride=2, padding=1)."""
    if MAXPOOL_KERNEL and x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16) and x.shape[1] % 8 == 0:
        return _MaxPool3x3S2.apply(x)
    return F.max_pool2d(x, kernel_size=3, stride=2, padding=1)


PREPROCESS_KERNEL = os.environ.get("U2B_PREPROCESS_KERNEL", "1") == "1"


def preprocess_u8(images_u8, mean, std, size_divisibility, out_dtype=torch.float32):
    """rcnn.py:223-234 for a (N,3,H,W) uint8 batch stored channels_last: normalise, zero-pad to the backbone's size
    divisibility, keep NHWC (csrc/pool.cu preprocess_u8_kernel). mean / std: sequences of 3 python floats."""
    import ctypes
    from .. import _lib
    N, C, H, W = images_u8.shape
    assert C == 3 and images_u8.dtype == torch.uint8 and images_u8.is_cuda
    x = images_u8 if (images_u8.is_contiguous(memory_format=torch.channels_last) and images_u8.stride(1) == 1) else \
        images_u8.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    s = size_divisibility
    Hp, Wp = (H + s - 1) // s * s, (W + s - 1) // s * s
    out = torch.empty((N, Hp, Wp, 3), dtype=out_dtype, device=x.device).permute(0, 3, 1, 2)
    m3, s3 = (ctypes.c_float * 3)(*[float(v) for v in mean]), (ctypes.c_float * 3)(*[float(v) for v in std])
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}[out_dtype]
    _lib.check(_lib.lib().u2b_preprocess_u8_nhwc(ctypes.c_void_p(x.data_ptr()), N, H, W, Hp, Wp, m3, s3, code,
                                                 ctypes.c_void_p(out.data_ptr()), _lib.stream_ptr()), "u2b_preprocess_u8_nhwc")
    _lib.count_launches(1)
    return out
