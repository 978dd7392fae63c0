"""Model layer: registries + modules with the reference's names (detectron2/modeling/__init__.py:4-55)."""
from ..registry import (ANCHOR_GENERATOR_REGISTRY, BACKBONE_REGISTRY, META_ARCH_REGISTRY,  # noqa: F401
                        PROPOSAL_GENERATOR_REGISTRY, ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY,
                        ROI_MASK_HEAD_REGISTRY, RPN_HEAD_REGISTRY, SEM_SEG_HEADS_REGISTRY)
from .backbone import FPN, ResNet, ShapeSpec, build_backbone, build_resnet_fpn_backbone  # noqa: F401
from .rpn import RPN, DefaultAnchorGenerator, StandardRPNHead, build_proposal_generator  # noqa: F401
from .roi_heads import (CascadeROIHeads, FastRCNNConvFCHead, FastRCNNOutputLayers,  # noqa: F401
                        MaskRCNNConvUpsampleHead, build_roi_heads)
from .semantic_seg import SemSegFPNHead, build_sem_seg_head  # noqa: F401
from .panoptic_fpn import PanopticFPN, build_model  # noqa: F401
