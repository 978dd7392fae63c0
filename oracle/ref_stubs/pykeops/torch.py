def LazyTensor(x):
    """identity: enables nn_utils.KMeans(force_no_lazy_tensor=True) (the reference's dense branch) on CPU."""
    return x
