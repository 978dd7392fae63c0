"""ResNet-50 + FPN backbone — mirror of detectron2/modeling/backbone/{resnet.py,fpn.py,backbone.py} and
detectron2/layers/{wrappers.py:87-134 Conv2d, batch_norm.py:169-197 get_norm}, same module tree and
state_dict names (backbone.bottom_up.stem.conv1.{weight,norm.*}, backbone.bottom_up.res{2..5}.{i}.*,
backbone.fpn_lateral{2..5}.*, backbone.fpn_output{2..5}.*).

Activations are channels_last (physical NHWC). Conv + norm (+ReLU, +residual) is one fused call into
ops.conv_bn_act so the implementation underneath (tcgen05 implicit GEMM with the SyncBN statistics
in its epilogue) can change without touching this file.
"""
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F
from torch import nn

from ..registry import BACKBONE_REGISTRY
from . import ops


@dataclass
class ShapeSpec:
    channels: Optional[int] = None
    height: Optional[int] = None
    width: Optional[int] = None
    stride: Optional[int] = None


class SyncBatchNorm(nn.BatchNorm2d):
    """layers/batch_norm.py:187 ("SyncBN" -> nn.SyncBatchNorm). Same parameters/buffers; statistics are
    reduced over the data-parallel group when one is initialised (ops.batch_norm)."""

    def forward(self, x):
        return ops.batch_norm(x, self, relu=False)


def get_norm(norm, out_channels):
    """layers/batch_norm.py:169-197."""
    if norm is None or norm == "":
        return None
    if norm in ("SyncBN", "BN"):
        return SyncBatchNorm(out_channels)
    if norm == "GN":
        return nn.GroupNorm(32, out_channels)
    raise NotImplementedError("norm %r is not on the u2seg hot path" % (norm,))


class Conv2d(nn.Conv2d):
    """layers/wrappers.py:87-134: conv (+ norm) (+ activation)."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x, residual=None):
        return ops.conv_norm_act(x, self, residual)


def c2_msra_fill(m):
    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


def c2_xavier_fill(m):
    nn.init.kaiming_uniform_(m.weight, a=1)
    if m.bias is not None:
        nn.init.constant_(m.bias, 0)


class Backbone(nn.Module):
    """backbone/backbone.py:11-74."""
    size_divisibility_ = 0

    @property
    def size_divisibility(self):
        return self.size_divisibility_

    @property
    def padding_constraints(self):
        return {}

    def output_shape(self):
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n])
                for n in self._out_features}


class BasicStem(nn.Module):
    """resnet.py:330-359."""

    def __init__(self, in_channels=3, out_channels=64, norm="BN"):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, 4
        self.conv1 = Conv2d(in_channels, out_channels, kernel_size=7, stride=2, padding=3, bias=False,
                            norm=get_norm(norm, out_channels), activation=F.relu_)
        c2_msra_fill(self.conv1)

    def forward(self, x):
        x = self.conv1(x)
        return ops.max_pool_3x3_s2(x)


class BottleneckBlock(nn.Module):
    """resnet.py:100-210."""

    def __init__(self, in_channels, out_channels, *, bottleneck_channels, stride=1, num_groups=1, norm="BN",
                 stride_in_1x1=False, dilation=1):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        if in_channels != out_channels:
            self.shortcut = Conv2d(in_channels, out_channels, kernel_size=1, stride=stride, bias=False,
                                   norm=get_norm(norm, out_channels))
        else:
            self.shortcut = None
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, stride=s1, bias=False,
                            norm=get_norm(norm, bottleneck_channels), activation=F.relu_)
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, stride=s3, padding=dilation,
                            bias=False, groups=num_groups, dilation=dilation,
                            norm=get_norm(norm, bottleneck_channels), activation=F.relu_)
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False,
                            norm=get_norm(norm, out_channels), activation=F.relu_)   # relu after the residual add
        for layer in (self.conv1, self.conv2, self.conv3, self.shortcut):
            if layer is not None:
                c2_msra_fill(layer)

    def forward(self, x):
        out = self.conv2(self.conv1(x))
        shortcut = self.shortcut(x) if self.shortcut is not None else x
        return self.conv3(out, residual=shortcut)   # relu(norm(conv3(out)) + shortcut), resnet.py:205-209


class ResNet(Backbone):
    """resnet.py:362-458."""

    def __init__(self, stem, stages, out_features):
        super().__init__()
        self.stem = stem
        cur_stride, cur_ch = stem.stride, stem.out_channels
        self._out_feature_strides, self._out_feature_channels = {"stem": cur_stride}, {"stem": cur_ch}
        self.stage_names, self.stages = [], []
        for i, blocks in enumerate(stages):
            name = "res" + str(i + 2)
            stage = nn.Sequential(*blocks)
            self.add_module(name, stage)
            self.stage_names.append(name)
            self.stages.append(stage)
            cur_stride = int(cur_stride * math.prod([b.stride for b in blocks]))
            cur_ch = blocks[-1].out_channels
            self._out_feature_strides[name], self._out_feature_channels[name] = cur_stride, cur_ch
        self._out_features = out_features

    def forward(self, x):
        outputs = {}
        x = self.stem(x)
        if "stem" in self._out_features:
            outputs["stem"] = x
        for name, stage in zip(self.stage_names, self.stages):
            x = stage(x)
            if name in self._out_features:
                outputs[name] = x
        return outputs


@BACKBONE_REGISTRY.register()
def build_resnet_backbone(cfg, input_shape):
    """resnet.py:614-694 (depth 50/101/152 bottleneck variants; no deformable conv on this path)."""
    norm = cfg.MODEL.RESNETS.NORM
    stem = BasicStem(in_channels=input_shape.channels, out_channels=cfg.MODEL.RESNETS.STEM_OUT_CHANNELS, norm=norm)
    depth = cfg.MODEL.RESNETS.DEPTH
    num_blocks = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}[depth]
    out_features = cfg.MODEL.RESNETS.OUT_FEATURES
    bott = cfg.MODEL.RESNETS.NUM_GROUPS * cfg.MODEL.RESNETS.WIDTH_PER_GROUP
    in_ch, out_ch = cfg.MODEL.RESNETS.STEM_OUT_CHANNELS, cfg.MODEL.RESNETS.RES2_OUT_CHANNELS
    stages = []
    for idx, stage_idx in enumerate(range(2, 6)):
        first_stride = 1 if idx == 0 else 2
        blocks = []
        for i in range(num_blocks[idx]):
            blocks.append(BottleneckBlock(in_ch, out_ch, bottleneck_channels=bott, stride=first_stride if i == 0 else 1,
                                          num_groups=cfg.MODEL.RESNETS.NUM_GROUPS, norm=norm,
                                          stride_in_1x1=cfg.MODEL.RESNETS.STRIDE_IN_1X1))
            in_ch = out_ch
        out_ch *= 2
        bott *= 2
        stages.append(blocks)
    assert cfg.MODEL.BACKBONE.FREEZE_AT == 0, "u2seg configs train the whole backbone (FREEZE_AT: 0)"
    return ResNet(stem, stages, out_features=out_features)


class LastLevelMaxPool(nn.Module):
    """fpn.py:188-200."""

    def __init__(self):
        super().__init__()
        self.num_levels, self.in_feature = 1, "p5"

    def forward(self, x):
        return [F.max_pool2d(x, kernel_size=1, stride=2, padding=0)]


class FPN(Backbone):
    """fpn.py:17-167."""

    def __init__(self, bottom_up, in_features, out_channels, norm="", top_block=None, fuse_type="sum"):
        super().__init__()
        shapes = bottom_up.output_shape()
        strides = [shapes[f].stride for f in in_features]
        in_channels = [shapes[f].channels for f in in_features]
        lateral_convs, output_convs = [], []
        use_bias = norm == ""
        for idx, ch in enumerate(in_channels):
            lateral = Conv2d(ch, out_channels, kernel_size=1, bias=use_bias, norm=get_norm(norm, out_channels))
            output = Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=use_bias,
                            norm=get_norm(norm, out_channels))
            c2_xavier_fill(lateral)
            c2_xavier_fill(output)
            stage = int(math.log2(strides[idx]))
            self.add_module("fpn_lateral{}".format(stage), lateral)
            self.add_module("fpn_output{}".format(stage), output)
            lateral_convs.append(lateral)
            output_convs.append(output)
        self.lateral_convs, self.output_convs = lateral_convs[::-1], output_convs[::-1]   # top-down order
        self.top_block, self.in_features, self.bottom_up = top_block, tuple(in_features), bottom_up
        self._out_feature_strides = {"p{}".format(int(math.log2(s))): s for s in strides}
        if top_block is not None:
            stage = int(math.log2(strides[-1]))
            for s in range(stage, stage + top_block.num_levels):
                self._out_feature_strides["p{}".format(s + 1)] = 2 ** (s + 1)
        self._out_features = list(self._out_feature_strides.keys())
        self._out_feature_channels = {k: out_channels for k in self._out_features}
        self.size_divisibility_ = strides[-1]
        assert fuse_type == "sum"

    def forward(self, x):
        bottom_up = self.bottom_up(x)
        results = []
        prev = self.lateral_convs[0](bottom_up[self.in_features[-1]])
        results.append(self.output_convs[0](prev))
        for idx, (lateral, output) in enumerate(zip(self.lateral_convs, self.output_convs)):
            if idx > 0:
                feat = bottom_up[self.in_features[-idx - 1]]
                prev = ops.lateral_add_upsample(lateral, feat, prev)   # lateral(feat) + nearest x2 of prev (fpn.py:153-156)
                results.insert(0, output(prev))
        if self.top_block is not None:
            top_in = results[self._out_features.index(self.top_block.in_feature)]
            results.extend(self.top_block(top_in))
        return {f: r for f, r in zip(self._out_features, results)}


@BACKBONE_REGISTRY.register()
def build_resnet_fpn_backbone(cfg, input_shape):
    """fpn.py:225-245."""
    bottom_up = build_resnet_backbone(cfg, input_shape)
    return FPN(bottom_up=bottom_up, in_features=cfg.MODEL.FPN.IN_FEATURES, out_channels=cfg.MODEL.FPN.OUT_CHANNELS,
               norm=cfg.MODEL.FPN.NORM, top_block=LastLevelMaxPool(), fuse_type=cfg.MODEL.FPN.FUSE_TYPE)


def build_backbone(cfg, input_shape=None):
    """backbone/build.py:20-33."""
    if input_shape is None:
        input_shape = ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN))
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)
