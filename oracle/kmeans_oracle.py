"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference k-means.

Follows u2seg/Instance_Clustering/shared/utils/nn_utils.py:304-379 line by line:
  :309-311 torch.manual_seed(seed)            :316 r = torch.randperm(N)[:K] (CPU generator)
  :338     c = x[r].clone()                   :353 D_ij = ((x_i - c_j)**2).sum(-1)  (fp32)
  :355     cl = D_ij.argmin(dim=1)            :359-360 c.zero_(); c.scatter_add_(0, cl[:,None].repeat(1,D), x)
  :363-364 Ncl = bincount(cl, minlength=K); c /= Ncl   (in place: next E-step sees the new c; 0/0 = NaN)
The (N,K,D) broadcast is chunked over N (the reference's dense branch materialises it whole and
its KeOps branch never does). Empty clusters: the reference's production (KeOps) reduction is
comparison based and never selects a NaN centroid; this oracle pins that behaviour (NaN -> +inf
before argmin) — torch.argmin of the dense branch would pick the NaN instead (SURVEY §7 hard part 9).

Pinned against the reference itself: oracle/make_golden.py runs the unmodified nn_utils.KMeans
(dense branch, pykeops stubbed) and tests/test_kmeans_oracle.py compares bit-exactly.
"""
import torch


def assign_oracle(x, c, chunk=4096):
    """labels (int64) = argmin_j sum_d (x_i - c_j)^2, fp32, first minimum, NaN never selected."""
    N = x.shape[0]
    out = torch.empty((N,), dtype=torch.int64)
    for s in range(0, N, chunk):
        d = ((x[s:s + chunk, None, :] - c[None, :, :]) ** 2).sum(-1)
        d = torch.where(torch.isnan(d), torch.full_like(d, float("inf")), d)
        out[s:s + chunk] = d.argmin(dim=1)
    return out


def update_oracle(x, cl, K):
    D = x.shape[1]
    c = torch.zeros((K, D), dtype=x.dtype)
    c.scatter_add_(0, cl[:, None].repeat(1, D), x)
    Ncl = torch.bincount(cl, minlength=K).type_as(c).view(K, 1)
    c /= Ncl
    return c, Ncl.view(-1)


def kmeans_oracle(x, seed, K=10, Niter=10, init_inds=None, chunk=4096):
    """x: (N, D) fp32 CPU tensor (the fp16 embeddings upcast). Returns (cl int64, c fp32)."""
    x = x.float().contiguous()
    N, D = x.shape
    if seed is not None:
        torch.manual_seed(seed)
    if init_inds is None:
        r = torch.randperm(N)[:K]
    else:
        r = torch.randperm(init_inds.shape[0])[:K]
        r = init_inds[r]
    c = x[r, :].clone()
    cl = None
    for _ in range(Niter):
        cl = assign_oracle(x, c, chunk)
        c.zero_()
        c.scatter_add_(0, cl[:, None].repeat(1, D), x)
        Ncl = torch.bincount(cl, minlength=K).type_as(c).view(K, 1)
        c /= Ncl
    return cl, c


def make_mixture(N, D, modes, seed, spread=4.0):
    """Synthetic embeddings of SURVEY §8(d): L2-normalised Gaussian mixture, returned as fp16."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(modes, D, generator=g)
    which = torch.randint(0, modes, (N,), generator=g)
    x = torch.randn(N, D, generator=g) + spread * centres[which]
    x = torch.nn.functional.normalize(x, dim=1)
    return x.half()
