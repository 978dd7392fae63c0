class ParamScheduler:
    WHERE_EPSILON = 1e-6

    def __call__(self, where):
        raise NotImplementedError


class _Named(ParamScheduler):
    def __init__(self, *a, **k):
        self.a, self.k = a, k


class CompositeParamScheduler(_Named): pass
class ConstantParamScheduler(_Named): pass
class LinearParamScheduler(_Named): pass
class CosineParamScheduler(_Named): pass
class MultiStepParamScheduler(_Named): pass
class StepParamScheduler(_Named): pass
class StepWithFixedGammaParamScheduler(_Named): pass
class ExponentialParamScheduler(_Named): pass
class PolynomialDecayParamScheduler(_Named): pass
