class Transform:
    def _set_attributes(self, params=None):
        if params:
            for k, v in params.items():
                if k != "self" and not k.startswith("_"):
                    setattr(self, k, v)

    @classmethod
    def register_type(cls, data_type, func=None):
        if func is None:
            def wrapper(decorated_func):
                cls.register_type(data_type, decorated_func)
                return decorated_func
            return wrapper
        setattr(cls, "apply_" + data_type, func)


class TransformList(Transform):
    def __init__(self, transforms):
        self.transforms = list(transforms)


class _T(Transform):
    def __init__(self, *a, **k):
        pass


class BlendTransform(_T): pass
class CropTransform(_T): pass
class PadTransform(_T): pass
class GridSampleTransform(_T): pass
class HFlipTransform(_T): pass
class VFlipTransform(_T): pass
class NoOpTransform(_T): pass
class ScaleTransform(_T): pass


__all__ = ["BlendTransform", "CropTransform", "PadTransform", "GridSampleTransform", "HFlipTransform",
           "VFlipTransform", "NoOpTransform", "ScaleTransform", "Transform", "TransformList"]
