"""fvcore.common.config.CfgNode stand-in: yacs CfgNode + `_BASE_` yaml inheritance."""
import os

import yaml
from yacs.config import CfgNode as _CfgNode

BASE_KEY = "_BASE_"


class CfgNode(_CfgNode):
    @classmethod
    def _open_cfg(cls, filename):
        return open(filename, "r")

    @classmethod
    def load_yaml_with_base(cls, filename, allow_unsafe=False):
        with cls._open_cfg(filename) as f:
            cfg = yaml.safe_load(f)

        def merge_a_into_b(a, b):
            for k, v in a.items():
                if isinstance(v, dict) and k in b:
                    assert isinstance(b[k], dict), k
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        def _load_with_base(base_cfg_file):
            if base_cfg_file.startswith("~"):
                base_cfg_file = os.path.expanduser(base_cfg_file)
            if not any(map(base_cfg_file.startswith, ["/", "https://", "http://"])):
                base_cfg_file = os.path.join(os.path.dirname(filename), base_cfg_file)
            return cls.load_yaml_with_base(base_cfg_file, allow_unsafe=allow_unsafe)

        if BASE_KEY in cfg:
            if isinstance(cfg[BASE_KEY], list):
                base_cfg = {}
                for b in cfg[BASE_KEY]:
                    merge_a_into_b(_load_with_base(b), base_cfg)
            else:
                base_cfg = _load_with_base(cfg[BASE_KEY])
            del cfg[BASE_KEY]
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename, allow_unsafe=False):
        loaded_cfg = self.load_yaml_with_base(cfg_filename, allow_unsafe=allow_unsafe)
        loaded_cfg = type(self)(loaded_cfg)
        self.merge_from_other_cfg(loaded_cfg)

    def merge_from_other_cfg(self, cfg_other):
        assert BASE_KEY not in cfg_other
        return super().merge_from_other_cfg(cfg_other)

    def merge_from_list(self, cfg_list):
        keys = set(cfg_list[0::2])
        assert BASE_KEY not in keys
        return super().merge_from_list(cfg_list)

    def __setattr__(self, name, val):
        if name.startswith("COMPUTED_"):
            if name in self:
                old_val = self[name]
                if old_val == val:
                    return
                raise KeyError(name)
            self[name] = val
        else:
            super().__setattr__(name, val)
