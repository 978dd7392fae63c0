"""Drop-in registration into a live Detectron2 / U2Seg installation.

    import u2seg_b200.plugin; u2seg_b200.plugin.register()
    python tools/train_net.py --config-file configs/COCO-PanopticSegmentation/u2seg_R50_800.yaml \\
        MODEL.META_ARCHITECTURE B200PanopticFPN

Detectron2's registries assert name uniqueness (fvcore Registry._do_register), so the B200 classes are added
under "B200"-prefixed names, selected by one CLI/yaml override (tools/train_net.py:119 merges `opts`).
`replace=True` instead overwrites the stock entries in `Registry._obj_map`, so unmodified configs resolve to
the B200 implementations. The B200 modules read the same cfg keys as the reference (a detectron2 CfgNode is
accepted as is), take the same `list[dict]` inputs (detectron2 Instances/Boxes/BitMasks are used through the
attributes both share) and keep the reference's state_dict names, so DetectionCheckpointer files load.
"""


def register(prefix="B200", replace=False):
    from detectron2.modeling import (ANCHOR_GENERATOR_REGISTRY, BACKBONE_REGISTRY, META_ARCH_REGISTRY,
                                     PROPOSAL_GENERATOR_REGISTRY, ROI_BOX_HEAD_REGISTRY, ROI_HEADS_REGISTRY,
                                     ROI_MASK_HEAD_REGISTRY, RPN_HEAD_REGISTRY, SEM_SEG_HEADS_REGISTRY)

    from . import modeling as M
    from .modeling import backbone as B
    from .modeling import roi_heads as RH
    from .modeling import rpn as R
    table = [
        (META_ARCH_REGISTRY, "PanopticFPN", M.PanopticFPN),
        (BACKBONE_REGISTRY, "build_resnet_fpn_backbone", B.build_resnet_fpn_backbone),
        (BACKBONE_REGISTRY, "build_resnet_backbone", B.build_resnet_backbone),
        (PROPOSAL_GENERATOR_REGISTRY, "RPN", R.RPN),
        (RPN_HEAD_REGISTRY, "StandardRPNHead", R.StandardRPNHead),
        (ANCHOR_GENERATOR_REGISTRY, "DefaultAnchorGenerator", R.DefaultAnchorGenerator),
        (ROI_HEADS_REGISTRY, "CascadeROIHeads", RH.CascadeROIHeads),
        (ROI_BOX_HEAD_REGISTRY, "FastRCNNConvFCHead", RH.FastRCNNConvFCHead),
        (ROI_MASK_HEAD_REGISTRY, "MaskRCNNConvUpsampleHead", RH.MaskRCNNConvUpsampleHead),
        (SEM_SEG_HEADS_REGISTRY, "SemSegFPNHead", M.SemSegFPNHead),
    ]
    names = []
    for reg, name, obj in table:
        if replace:
            reg._obj_map[name] = obj
            names.append(name)
        else:
            new = prefix + name
            if new not in reg._obj_map:
                reg._obj_map[new] = obj
            names.append(new)
    return names
