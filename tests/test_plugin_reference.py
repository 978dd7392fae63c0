"""Build container only (needs /root/reference + oracle/ref_stubs): the B200 classes register into the
reference's registries and the reference's own build_model / config entry constructs them from the
unmodified u2seg yaml, with the reference's state_dict names."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.reference
def test_register_and_build_through_reference_entry():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_stubs"), "/root/reference"]
    from detectron2.config import get_cfg
    from detectron2.modeling import META_ARCH_REGISTRY, build_model

    import u2seg_b200.plugin as plugin
    names = plugin.register()
    assert "B200PanopticFPN" in names and "B200PanopticFPN" in META_ARCH_REGISTRY
    cfg = get_cfg()
    cfg.merge_from_file("/root/reference/configs/COCO-PanopticSegmentation/u2seg_R50_800.yaml")
    cfg.merge_from_list(["MODEL.META_ARCHITECTURE", "B200PanopticFPN", "MODEL.DEVICE", "cpu"])   # train_net.py:119
    model = build_model(cfg)
    from u2seg_b200.modeling import PanopticFPN
    assert isinstance(model, PanopticFPN)
    ref_cfg = get_cfg()
    ref_cfg.merge_from_file("/root/reference/configs/COCO-PanopticSegmentation/u2seg_R50_800.yaml")
    ref_cfg.MODEL.DEVICE = "cpu"
    ref = build_model(ref_cfg)
    a, b = model.state_dict(), ref.state_dict()
    assert set(a) == set(b) and all(a[k].shape == b[k].shape for k in a)
    model.load_state_dict(b)          # a reference checkpoint loads into the B200 model
