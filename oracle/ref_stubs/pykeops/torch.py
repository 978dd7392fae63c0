"""Dense stand-in for pykeops.torch.LazyTensor (pykeops needs a GPU + nvrtc and is not installed; no network).

* nn_utils.KMeans(force_no_lazy_tensor=True) - the reference's dense branch - never calls LazyTensor: importing the name
  is all it needs.
* nn_utils.kNN builds D_ij = ((LazyTensor(x_i) - LazyTensor(x_j)) ** 2).sum(-1) and calls D_ij.Kmin_argKmin(K, dim=1):
  LazyTensor returns a torch.Tensor subclass, so the reference's own expression is evaluated densely by torch, and
  Kmin_argKmin is the K smallest entries along `dim`, ascending - what the KeOps reduction returns."""
import torch


class _Dense(torch.Tensor):
    def Kmin_argKmin(self, K, dim=1, backend=None):
        v, i = torch.topk(self.as_subclass(torch.Tensor), K, dim=dim, largest=False, sorted=True)
        return v, i


def LazyTensor(x):
    return x.as_subclass(_Dense) if isinstance(x, torch.Tensor) else x
