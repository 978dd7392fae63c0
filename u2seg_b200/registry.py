"""Registries with the reference's names (the drop-in surface):
detectron2/modeling/meta_arch/build.py:7, backbone/build.py:7, proposal_generator/build.py:4,
proposal_generator/rpn.py:21, anchor_generator.py:13, roi_heads/roi_heads.py:25, box_head.py:14,
mask_head.py:23, meta_arch/semantic_seg.py:26."""


class Registry:
    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        assert name not in self._obj_map, "An object named '%s' was already registered in '%s' registry!" % (name, self._name)
        self._obj_map[name] = obj

    def register(self, obj=None):
        if obj is None:
            def deco(o):
                self._do_register(o.__name__, o)
                return o
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError("No object named '%s' found in '%s' registry!" % (name, self._name))
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
PROPOSAL_GENERATOR_REGISTRY = Registry("PROPOSAL_GENERATOR")
RPN_HEAD_REGISTRY = Registry("RPN_HEAD")
ANCHOR_GENERATOR_REGISTRY = Registry("ANCHOR_GENERATOR")
ROI_HEADS_REGISTRY = Registry("ROI_HEADS")
ROI_BOX_HEAD_REGISTRY = Registry("ROI_BOX_HEAD")
ROI_MASK_HEAD_REGISTRY = Registry("ROI_MASK_HEAD")
SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")
