"""PanopticFPN meta-architecture — mirror of detectron2/modeling/meta_arch/{build.py:16-25, rcnn.py:25-234,
panoptic_fpn.py:21-181}. Registered as "PanopticFPN" in META_ARCH_REGISTRY; model(batched_inputs) has the
reference's I/O contract (list[dict] in; dict of 10 losses in training, list[dict] in eval)."""
from typing import Dict, List

import torch
from torch import nn

from ..registry import META_ARCH_REGISTRY
from ..structures import ImageList
from .backbone import build_backbone
from .postprocessing import combine_semantic_and_instance_outputs, detector_postprocess, sem_seg_postprocess
from .roi_heads import build_roi_heads
from .rpn import build_proposal_generator
from .semantic_seg import build_sem_seg_head


@META_ARCH_REGISTRY.register()
class PanopticFPN(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.backbone = build_backbone(cfg)
        shapes = self.backbone.output_shape()
        self.proposal_generator = build_proposal_generator(cfg, shapes)
        self.roi_heads = build_roi_heads(cfg, shapes)
        self.sem_seg_head = build_sem_seg_head(cfg, shapes)
        self.register_buffer("pixel_mean", torch.tensor(cfg.MODEL.PIXEL_MEAN).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(cfg.MODEL.PIXEL_STD).view(-1, 1, 1), False)
        # host copies (fp32-rounded like the buffers) for the fused preprocessing kernel's scalar arguments
        self._pixel_mean_host = [float(torch.tensor(v, dtype=torch.float32)) for v in cfg.MODEL.PIXEL_MEAN]
        self._pixel_std_host = [float(torch.tensor(v, dtype=torch.float32)) for v in cfg.MODEL.PIXEL_STD]
        c = cfg.MODEL.PANOPTIC_FPN.COMBINE
        self.combine_overlap_thresh = c.OVERLAP_THRESH
        self.combine_stuff_area_thresh = c.STUFF_AREA_LIMIT
        self.combine_instances_score_thresh = c.INSTANCES_CONFIDENCE_THRESH
        assert cfg.MODEL.PANOPTIC_FPN.INSTANCE_LOSS_WEIGHT == 1.0
        self.input_format = cfg.INPUT.FORMAT
        self._side_stream = None
        # num_batches_tracked of the 61 BN layers: one foreach add per step instead of 61 tiny launches
        self._bn_counters = []
        for m in self.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.num_batches_tracked is not None:
                m._counter_batched = True
                self._bn_counters.append(m.num_batches_tracked)

    @property
    def device(self):
        return self.pixel_mean.device

    def run_backbone(self, x):
        """backbone(x); replays the CUDA graph captured by engine.Trainer for this input shape when there is one
        (the backbone + FPN is the static-shape part of the step: ~1200 of its ~3700 kernel launches)."""
        g = getattr(self, "_graphed_backbone", None)
        if g is not None and self.training and tuple(x.shape) == g[2]:
            return dict(zip(g[1], g[0](x)))
        return self.backbone(x)

    def preprocess_image(self, batched_inputs: List[Dict[str, torch.Tensor]]):
        """rcnn.py:223-234: H2D, (x - mean) / std, zero-pad to a multiple of size_divisibility, batch.
        Output is channels_last (what every conv kernel below reads)."""
        images = [x["image"].to(self.device, non_blocking=True) for x in batched_inputs]
        images = [(x.float() - self.pixel_mean) / self.pixel_std for x in images]
        il = ImageList.from_tensors(images, self.backbone.size_divisibility)
        il.tensor = il.tensor.contiguous(memory_format=torch.channels_last)
        return il

    def forward(self, batched_inputs):
        """panoptic_fpn.py:90-138."""
        if not self.training:
            return self.inference(batched_inputs)
        images = self.preprocess_image(batched_inputs)
        if self._bn_counters:
            torch._foreach_add_(self._bn_counters, 1)
        features = self.run_backbone(images.tensor)
        assert "sem_seg" in batched_inputs[0]
        gt_sem_seg = [x["sem_seg"].to(self.device, non_blocking=True) for x in batched_inputs]
        gt_sem_seg = ImageList.from_tensors(gt_sem_seg, self.backbone.size_divisibility,
                                            self.sem_seg_head.ignore_value).tensor
        # The semantic head has no host synchronisation; the proposal / ROI-sampling path below has ~30
        # (data-dependent shapes). Run the former on a side stream so the GPU stays busy across those bubbles.
        main = torch.cuda.current_stream()
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream()
        self._side_stream.wait_stream(main)
        with torch.cuda.stream(self._side_stream):
            _, sem_seg_losses = self.sem_seg_head(features, gt_sem_seg)
        gt_instances = [x["instances"].to(self.device) for x in batched_inputs]
        proposals, proposal_losses = self.proposal_generator(images, features, gt_instances)
        _, detector_losses = self.roi_heads(images, features, proposals, gt_instances)
        main.wait_stream(self._side_stream)
        losses = sem_seg_losses
        losses.update(proposal_losses)
        losses.update(detector_losses)
        return losses

    @torch.no_grad()
    def inference(self, batched_inputs, do_postprocess=True):
        """panoptic_fpn.py:140-181."""
        images = self.preprocess_image(batched_inputs)
        features = self.backbone(images.tensor)
        sem_seg_results, _ = self.sem_seg_head(features, None)
        proposals, _ = self.proposal_generator(images, features, None)
        detector_results, _ = self.roi_heads(images, features, proposals, None)
        if not do_postprocess:
            return detector_results, sem_seg_results
        out = []
        for sem, det, inp, image_size in zip(sem_seg_results, detector_results, batched_inputs, images.image_sizes):
            h, w = inp.get("height", image_size[0]), inp.get("width", image_size[1])
            sem_r = sem_seg_postprocess(sem.float(), image_size, h, w)
            det_r = detector_postprocess(det, h, w)
            pan = combine_semantic_and_instance_outputs(det_r, sem_r.argmax(dim=0), self.combine_overlap_thresh,
                                                        self.combine_stuff_area_thresh,
                                                        self.combine_instances_score_thresh)
            out.append({"sem_seg": sem_r, "instances": det_r, "panoptic_seg": pan})
        return out


def build_model(cfg):
    """meta_arch/build.py:16-25."""
    model = META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model
