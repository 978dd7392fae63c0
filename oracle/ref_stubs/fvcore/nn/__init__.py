from . import weight_init  # noqa
from .smooth_l1_loss import smooth_l1_loss  # noqa


def _unavailable(*a, **k):
    raise NotImplementedError("fvcore stub")


giou_loss = sigmoid_focal_loss = sigmoid_focal_loss_jit = _unavailable
activation_count = flop_count = parameter_count = parameter_count_table = _unavailable


class FlopCountAnalysis:
    pass
