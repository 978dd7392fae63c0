"""SemSegFPNHead — mirror of detectron2/modeling/meta_arch/semantic_seg.py:143-267; state_dict names
sem_seg_head.{p2.0,p3.0,p4.0,p4.2,p5.0,p5.2,p5.4}.{weight,norm.*}, sem_seg_head.predictor.*."""
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ..registry import SEM_SEG_HEADS_REGISTRY
from . import ops
from .backbone import Conv2d, ShapeSpec, c2_msra_fill, get_norm


@SEM_SEG_HEADS_REGISTRY.register()
class SemSegFPNHead(nn.Module):
    def __init__(self, cfg, input_shape: Dict[str, ShapeSpec]):
        super().__init__()
        c = cfg.MODEL.SEM_SEG_HEAD
        shapes = sorted(((k, v) for k, v in input_shape.items() if k in c.IN_FEATURES), key=lambda x: x[1].stride)
        self.in_features = [k for k, _ in shapes]
        self.ignore_value, self.common_stride, self.loss_weight = c.IGNORE_VALUE, c.COMMON_STRIDE, c.LOSS_WEIGHT
        self.scale_heads = []
        for name, spec in shapes:
            head_ops = []
            head_length = max(1, int(np.log2(spec.stride) - np.log2(self.common_stride)))
            for k in range(head_length):
                norm_module = get_norm(c.NORM, c.CONVS_DIM)
                conv = Conv2d(spec.channels if k == 0 else c.CONVS_DIM, c.CONVS_DIM, kernel_size=3, stride=1, padding=1,
                              bias=not c.NORM, norm=norm_module, activation=F.relu_)
                c2_msra_fill(conv)
                head_ops.append(conv)
                if spec.stride != self.common_stride:
                    head_ops.append(ops.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
            self.scale_heads.append(nn.Sequential(*head_ops))
            self.add_module(name, self.scale_heads[-1])
        self.predictor = Conv2d(c.CONVS_DIM, c.NUM_CLASSES, kernel_size=1, stride=1, padding=0)
        c2_msra_fill(self.predictor)

    def layers(self, features):
        x = None
        for i, f in enumerate(self.in_features):
            y = self.scale_heads[i](features[f])
            x = y if x is None else x + y
        return self.predictor(x)

    def forward(self, features, targets=None):
        x = self.layers(features)
        if self.training:
            return None, self.losses(x, targets)
        x = F.interpolate(x, scale_factor=self.common_stride, mode="bilinear", align_corners=False)
        return x, {}

    def losses(self, predictions, targets):
        """semantic_seg.py:255-267: fp32 logits, bilinear x4, CE(mean, ignore) * weight."""
        from ..layers import upsample_cross_entropy, upsample_cross_entropy_supported
        if upsample_cross_entropy_supported(predictions, self.common_stride):
            loss = upsample_cross_entropy(predictions, targets, self.common_stride, self.ignore_value)
            return {"loss_sem_seg": loss * self.loss_weight}
        predictions = predictions.float()
        predictions = F.interpolate(predictions, scale_factor=self.common_stride, mode="bilinear", align_corners=False)
        loss = F.cross_entropy(predictions, targets, reduction="mean", ignore_index=self.ignore_value)
        return {"loss_sem_seg": loss * self.loss_weight}


def build_sem_seg_head(cfg, input_shape):
    """semantic_seg.py:134-139."""
    return SEM_SEG_HEADS_REGISTRY.get(cfg.MODEL.SEM_SEG_HEAD.NAME)(cfg, input_shape)
