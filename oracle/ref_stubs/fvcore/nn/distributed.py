def differentiable_all_reduce(x):
    return x
