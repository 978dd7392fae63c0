"""Config entry: yaml files with `_BASE_` inheritance + dotted CLI overrides, the subset of
detectron2/config (config.py:12-93, defaults.py) the u2seg hot path reads. The reference yaml files
configs/COCO-PanopticSegmentation/u2seg_R50_{300,800}.yaml parse unmodified (unknown keys are kept)."""
import copy
import os
from ast import literal_eval

import yaml


class CfgNode(dict):
    def __init__(self, d=None):
        super().__init__()
        for k, v in (d or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v
        self.__dict__["_frozen"] = False

    def __getattr__(self, k):
        if k in self:
            return self[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if self.__dict__.get("_frozen"):
            raise AttributeError("config is frozen")
        self[k] = v

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def clone(self):
        return copy.deepcopy(self)

    def merge_from_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(self.get(k), CfgNode):
                self[k].merge_from_dict(v)
            elif isinstance(v, dict):
                self[k] = CfgNode(v)
            else:
                old = self.get(k)
                if isinstance(old, tuple) and isinstance(v, list):
                    v = tuple(v)
                if isinstance(old, float) and isinstance(v, int):
                    v = float(v)
                if isinstance(v, str) and isinstance(old, (tuple, list)):
                    try:
                        v = type(old)(literal_eval(v))
                    except (ValueError, SyntaxError):
                        pass
                self[k] = v

    @staticmethod
    def load_yaml_with_base(filename):
        with open(filename) as f:
            cfg = yaml.safe_load(f) or {}
        base = cfg.pop("_BASE_", None)
        if base is None:
            return cfg
        if not os.path.isabs(base):
            base = os.path.join(os.path.dirname(filename), base)
        out = CfgNode.load_yaml_with_base(base)

        def merge(a, b):
            for k, v in a.items():
                if isinstance(v, dict) and isinstance(b.get(k), dict):
                    merge(v, b[k])
                else:
                    b[k] = v
        merge(cfg, out)
        return out

    def merge_from_file(self, filename):
        self.merge_from_dict(CfgNode.load_yaml_with_base(filename))

    def merge_from_list(self, opts):
        assert len(opts) % 2 == 0
        for k, v in zip(opts[0::2], opts[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            if isinstance(v, str):
                try:
                    v = literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            node.merge_from_dict({parts[-1]: v})


_DEFAULTS = {
    "VERSION": 2,
    "MODEL": {
        "META_ARCHITECTURE": "GeneralizedRCNN", "DEVICE": "cuda", "MASK_ON": False, "WEIGHTS": "",
        "PIXEL_MEAN": [103.530, 116.280, 123.675], "PIXEL_STD": [1.0, 1.0, 1.0],
        "BACKBONE": {"NAME": "build_resnet_backbone", "FREEZE_AT": 2},
        "FPN": {"IN_FEATURES": [], "OUT_CHANNELS": 256, "NORM": "", "FUSE_TYPE": "sum"},
        "RESNETS": {"DEPTH": 50, "OUT_FEATURES": ["res4"], "NUM_GROUPS": 1, "NORM": "FrozenBN", "WIDTH_PER_GROUP": 64,
                    "STRIDE_IN_1X1": True, "RES5_DILATION": 1, "RES2_OUT_CHANNELS": 256, "STEM_OUT_CHANNELS": 64},
        "PROPOSAL_GENERATOR": {"NAME": "RPN", "MIN_SIZE": 0},
        "ANCHOR_GENERATOR": {"NAME": "DefaultAnchorGenerator", "SIZES": [[32, 64, 128, 256, 512]],
                             "ASPECT_RATIOS": [[0.5, 1.0, 2.0]], "OFFSET": 0.0},
        "RPN": {"HEAD_NAME": "StandardRPNHead", "IN_FEATURES": ["res4"], "BOUNDARY_THRESH": -1,
                "IOU_THRESHOLDS": [0.3, 0.7], "IOU_LABELS": [0, -1, 1], "BATCH_SIZE_PER_IMAGE": 256,
                "POSITIVE_FRACTION": 0.5, "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                "BBOX_REG_WEIGHTS": (1.0, 1.0, 1.0, 1.0), "SMOOTH_L1_BETA": 0.0, "LOSS_WEIGHT": 1.0,
                "PRE_NMS_TOPK_TRAIN": 12000, "PRE_NMS_TOPK_TEST": 6000, "POST_NMS_TOPK_TRAIN": 2000,
                "POST_NMS_TOPK_TEST": 1000, "NMS_THRESH": 0.7, "CONV_DIMS": [-1]},
        "ROI_HEADS": {"NAME": "Res5ROIHeads", "NUM_CLASSES": 80, "IN_FEATURES": ["res4"], "IOU_THRESHOLDS": [0.5],
                      "IOU_LABELS": [0, 1], "BATCH_SIZE_PER_IMAGE": 512, "POSITIVE_FRACTION": 0.25,
                      "SCORE_THRESH_TEST": 0.05, "NMS_THRESH_TEST": 0.5, "PROPOSAL_APPEND_GT": True},
        "ROI_BOX_HEAD": {"NAME": "", "BBOX_REG_LOSS_TYPE": "smooth_l1", "BBOX_REG_LOSS_WEIGHT": 1.0,
                         "BBOX_REG_WEIGHTS": (10.0, 10.0, 5.0, 5.0), "SMOOTH_L1_BETA": 0.0, "POOLER_RESOLUTION": 14,
                         "POOLER_SAMPLING_RATIO": 0, "POOLER_TYPE": "ROIAlignV2", "NUM_FC": 0, "FC_DIM": 1024,
                         "NUM_CONV": 0, "CONV_DIM": 256, "NORM": "", "CLS_AGNOSTIC_BBOX_REG": False,
                         "TRAIN_ON_PRED_BOXES": False},
        "ROI_BOX_CASCADE_HEAD": {"BBOX_REG_WEIGHTS": ((10.0, 10.0, 5.0, 5.0), (20.0, 20.0, 10.0, 10.0),
                                                      (30.0, 30.0, 15.0, 15.0)), "IOUS": (0.5, 0.6, 0.7)},
        "ROI_MASK_HEAD": {"NAME": "MaskRCNNConvUpsampleHead", "POOLER_RESOLUTION": 14, "POOLER_SAMPLING_RATIO": 0,
                          "NUM_CONV": 0, "CONV_DIM": 256, "NORM": "", "CLS_AGNOSTIC_MASK": False,
                          "POOLER_TYPE": "ROIAlignV2"},
        "SEM_SEG_HEAD": {"NAME": "SemSegFPNHead", "IN_FEATURES": ["p2", "p3", "p4", "p5"], "IGNORE_VALUE": 255,
                         "NUM_CLASSES": 54, "CONVS_DIM": 128, "COMMON_STRIDE": 4, "NORM": "GN", "LOSS_WEIGHT": 1.0},
        "PANOPTIC_FPN": {"INSTANCE_LOSS_WEIGHT": 1.0,
                         "COMBINE": {"ENABLED": True, "OVERLAP_THRESH": 0.5, "STUFF_AREA_LIMIT": 4096,
                                     "INSTANCES_CONFIDENCE_THRESH": 0.5}},
    },
    "INPUT": {"MIN_SIZE_TRAIN": (800,), "MAX_SIZE_TRAIN": 1333, "MIN_SIZE_TEST": 800, "MAX_SIZE_TEST": 1333,
              "FORMAT": "BGR", "MASK_FORMAT": "polygon"},
    "SOLVER": {"IMS_PER_BATCH": 16, "BASE_LR": 0.001, "MOMENTUM": 0.9, "NESTEROV": False, "WEIGHT_DECAY": 0.0001,
               "WEIGHT_DECAY_NORM": 0.0, "WEIGHT_DECAY_BIAS": None, "BIAS_LR_FACTOR": 1.0, "GAMMA": 0.1,
               "STEPS": (30000,), "MAX_ITER": 40000, "WARMUP_FACTOR": 0.001, "WARMUP_ITERS": 1000,
               "WARMUP_METHOD": "linear", "CHECKPOINT_PERIOD": 5000,
               "CLIP_GRADIENTS": {"ENABLED": False, "CLIP_TYPE": "value", "CLIP_VALUE": 1.0, "NORM_TYPE": 2.0},
               "AMP": {"ENABLED": False}},
    "TEST": {"DETECTIONS_PER_IMAGE": 100, "EVAL_PERIOD": 0},
    "DATASETS": {"TRAIN": (), "TEST": ()},
    "DATALOADER": {"NUM_WORKERS": 4, "FILTER_EMPTY_ANNOTATIONS": True},
    "OUTPUT_DIR": "./output",
    "SEED": -1,
}

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def get_cfg():
    return CfgNode(copy.deepcopy(_DEFAULTS))


def get_u2seg_cfg(num_classes=800):
    """The shipped equivalent of configs/COCO-PanopticSegmentation/u2seg_R50_{300,800}.yaml."""
    cfg = get_cfg()
    cfg.merge_from_file(os.path.join(CONFIG_DIR, "u2seg_R50_%d.yaml" % num_classes))
    return cfg
