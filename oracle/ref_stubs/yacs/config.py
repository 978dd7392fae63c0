"""Minimal yacs.config.CfgNode stand-in (attribute dict + merge/freeze), enough for detectron2.config."""
import copy
from ast import literal_eval

import yaml

_VALID_TYPES = {tuple, list, str, int, float, bool, type(None)}


class CfgNode(dict):
    IMMUTABLE = "__immutable__"
    DEPRECATED_KEYS = "__deprecated_keys__"
    RENAMED_KEYS = "__renamed_keys__"
    NEW_ALLOWED = "__new_allowed__"

    def __init__(self, init_dict=None, key_list=None, new_allowed=False):
        init_dict = {} if init_dict is None else init_dict
        key_list = [] if key_list is None else key_list
        init_dict = self._create_config_tree_from_dict(init_dict, key_list)
        super().__init__(init_dict)
        self.__dict__[CfgNode.IMMUTABLE] = False
        self.__dict__[CfgNode.DEPRECATED_KEYS] = set()
        self.__dict__[CfgNode.RENAMED_KEYS] = {}
        self.__dict__[CfgNode.NEW_ALLOWED] = new_allowed

    @classmethod
    def _create_config_tree_from_dict(cls, dic, key_list):
        dic = copy.deepcopy(dic)
        for k, v in dic.items():
            if isinstance(v, dict) and not isinstance(v, CfgNode):
                dic[k] = cls(v, key_list=key_list + [k])
        return dic

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.is_frozen():
            raise AttributeError("Attempted to set {} to {}, but CfgNode is immutable".format(name, value))
        assert name not in self.__dict__, name
        self[name] = value

    def __str__(self):
        return self.dump()

    def __repr__(self):
        return "{}({})".format(self.__class__.__name__, super().__repr__())

    def _to_dict(self):
        def conv(n):
            if isinstance(n, CfgNode):
                return {k: conv(v) for k, v in n.items()}
            if isinstance(n, tuple):
                return list(n)
            return n
        return conv(self)

    def dump(self, **kwargs):
        return yaml.safe_dump(self._to_dict(), **kwargs)

    def merge_from_file(self, cfg_filename):
        with open(cfg_filename, "r") as f:
            cfg = self.load_cfg(f)
        self.merge_from_other_cfg(cfg)

    def merge_from_other_cfg(self, cfg_other):
        _merge_a_into_b(cfg_other, self, self, [])

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0, cfg_list
        root = self
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            key_list = full_key.split(".")
            d = self
            for subkey in key_list[:-1]:
                assert subkey in d, "Non-existent key: {}".format(full_key)
                d = d[subkey]
            subkey = key_list[-1]
            assert subkey in d, "Non-existent key: {}".format(full_key)
            value = self._decode_cfg_value(v)
            value = _check_and_coerce_cfg_value_type(value, d[subkey], subkey, full_key)
            d[subkey] = value

    def freeze(self):
        self._immutable(True)

    def defrost(self):
        self._immutable(False)

    def is_frozen(self):
        return self.__dict__[CfgNode.IMMUTABLE]

    def _immutable(self, is_immutable):
        self.__dict__[CfgNode.IMMUTABLE] = is_immutable
        for v in self.__dict__.values():
            if isinstance(v, CfgNode):
                v._immutable(is_immutable)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._immutable(is_immutable)

    def clone(self):
        return copy.deepcopy(self)

    def is_new_allowed(self):
        return self.__dict__[CfgNode.NEW_ALLOWED]

    def set_new_allowed(self, is_new_allowed):
        self.__dict__[CfgNode.NEW_ALLOWED] = is_new_allowed
        for v in self.values():
            if isinstance(v, CfgNode):
                v.set_new_allowed(is_new_allowed)

    def key_is_deprecated(self, full_key):
        return full_key in self.__dict__[CfgNode.DEPRECATED_KEYS]

    def key_is_renamed(self, full_key):
        return full_key in self.__dict__[CfgNode.RENAMED_KEYS]

    @classmethod
    def load_cfg(cls, cfg_file_obj_or_str):
        if isinstance(cfg_file_obj_or_str, str):
            return cls(yaml.safe_load(cfg_file_obj_or_str))
        return cls(yaml.safe_load(cfg_file_obj_or_str))

    @classmethod
    def _decode_cfg_value(cls, value):
        if isinstance(value, dict):
            return cls(value)
        if not isinstance(value, str):
            return value
        try:
            value = literal_eval(value)
        except (ValueError, SyntaxError):
            pass
        return value


def _check_and_coerce_cfg_value_type(replacement, original, key, full_key):
    original_type, replacement_type = type(original), type(replacement)
    if replacement_type == original_type:
        return replacement
    if original is None or replacement is None:
        return replacement
    casts = [(list, tuple), (tuple, list), (int, float)]
    for (from_type, to_type) in casts:
        if replacement_type == from_type and original_type == to_type:
            return to_type(replacement)
    raise ValueError("Type mismatch ({} vs. {}) with values ({} vs. {}) for config key: {}".format(
        original_type, replacement_type, original, replacement, full_key))


def _merge_a_into_b(a, b, root, key_list):
    for k, v_ in a.items():
        full_key = ".".join(key_list + [k])
        v = copy.deepcopy(v_)
        v = b._decode_cfg_value(v)
        if k in b:
            v = _check_and_coerce_cfg_value_type(v, b[k], k, full_key)
            if isinstance(v, CfgNode):
                _merge_a_into_b(v, b[k], root, key_list + [k])
            else:
                b[k] = v
        elif b.is_new_allowed():
            b[k] = v
        else:
            if root.key_is_deprecated(full_key) or root.key_is_renamed(full_key):
                continue
            raise KeyError("Non-existent config key: {}".format(full_key))
