"""GPU parity of the exact kNN (csrc/knn.cu + u2seg_b200/clustering.kNN) with the reference's recorded outputs
(tests/golden/knn_*.npz, nn_utils.kNN run unmodified through the dense pykeops stub) and with the oracle.

Distances are FP (fp32 sum of squares; a different summation order than torch's dense formula: 1e-5 relative). Indices
are INT and must be exact - except among equal or rounding-close distances, where the reference itself accepts any index
with the right distance (nn_utils.py:279-293): a position may differ only if the two distances agree to 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle.kmeans_oracle import make_mixture
from oracle.knn_oracle import knn_oracle

pytestmark = pytest.mark.gpu


def _check(ind, d, want_ind, want_d, x_train, x_test):
    ind, d = ind.cpu(), d.cpu()
    assert ind.shape == want_ind.shape and d.shape == want_d.shape and ind.dtype == torch.int64
    assert torch.allclose(d, want_d, rtol=1e-5, atol=1e-6)
    assert bool((d[:, 1:] >= d[:, :-1]).all())                    # ascending
    diff = ind != want_ind
    if diff.any():                                                # only rounding-level ties may permute
        r, c = torch.where(diff)
        exact = ((x_test[r] - x_train[ind[r, c]]) ** 2).sum(-1)
        assert torch.allclose(exact, want_d[r, c], rtol=1e-5, atol=1e-6)
        assert diff.float().mean() < 0.05          # tie permutations only (duplicated rows in the small cases)


@pytest.mark.parametrize("name", ["knn_n3000_d128_k20", "knn_self_n2500_d384_k20"])
def test_knn_matches_reference_golden(golden_dir, name):
    from u2seg_b200.clustering import kNN
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n_train, n_test, D, K, seed = [int(v) for v in g["meta"]]
    xt = make_mixture(n_train, D, 60, seed=seed, spread=1.0).float()
    xq = xt if n_test == n_train else make_mixture(n_test, D, 60, seed=seed + 100, spread=1.0).float()
    ind, d = kNN(xt.cuda(), xq.cuda() if xq is not xt else xt.cuda(), K=K)
    _check(ind, d, torch.from_numpy(g["ind"]), torch.from_numpy(g["dist"]), xt, xq)


@pytest.mark.parametrize("n_train,n_test,D,K", [(4097, 300, 256, 20), (161, 129, 128, 5), (20000, 1000, 384, 20),
                                                (1000, 77, 384, 40)])
def test_knn_matches_oracle(n_train, n_test, D, K):
    from u2seg_b200.clustering import kNN
    xt = make_mixture(n_train, D, 40, seed=n_train, spread=1.0).float()
    xq = make_mixture(n_test, D, 40, seed=n_test, spread=1.0).float()
    xq[:5] = xt[:5]                                               # exact hits (distance 0)
    xt[10] = xt[11]                                               # duplicate train rows: an exact tie
    want_ind, want_d = knn_oracle(xt, xq, K)
    ind, d, stats = kNN(xt.cuda(), xq.cuda(), K=K, return_stats=True)
    _check(ind, d, want_ind, want_d, xt, xq)


def test_knn_uncertifiable_rows_fall_back_to_exhaustive_search():
    """Many train rows at (almost) the same distance from a query: 48 candidates cannot be certified, the flagged rows
    are recomputed exhaustively and the answer is still exact."""
    from u2seg_b200.clustering import kNN
    g = torch.Generator().manual_seed(4)
    D, K = 128, 20
    base = torch.nn.functional.normalize(torch.randn(1, D, generator=g), dim=1)
    xt = base + 1e-4 * torch.randn(600, D, generator=g)           # 600 near-identical train rows
    xt = torch.cat([xt, torch.randn(2000, D, generator=g)])
    xq = base + 1e-4 * torch.randn(64, D, generator=g)
    want_ind, want_d = knn_oracle(xt, xq, K)
    ind, d, stats = kNN(xt.cuda(), xq.cuda(), K=K, return_stats=True)
    assert stats["uncertified_rows"] > 0 and stats["exhaustive_rows"] > 0
    assert torch.allclose(d.cpu(), want_d, rtol=1e-4, atol=1e-9)


def test_knn_dense_neighbourhood_certified_by_second_pass():
    """100 train rows within the fp16 rounding bound of each other: the 48-candidate lists cannot be certified, the fp32 second
    pass (128 candidates, bound 4 D 2^-24 |x||y|) can; no exhaustive search, exact answer."""
    from u2seg_b200.clustering import kNN
    g = torch.Generator().manual_seed(11)
    D, K = 128, 20
    base = torch.nn.functional.normalize(torch.randn(1, D, generator=g), dim=1)
    xt = torch.cat([base + 2e-3 * torch.randn(100, D, generator=g),
                    torch.nn.functional.normalize(torch.randn(3000, D, generator=g), dim=1)])
    xq = base + 2e-3 * torch.randn(64, D, generator=g)
    want_ind, want_d = knn_oracle(xt, xq, K)
    ind, d, stats = kNN(xt.cuda(), xq.cuda(), K=K, return_stats=True)
    print("dense neighbourhood:", stats)
    assert stats["uncertified_rows"] > 0 and stats["exhaustive_rows"] == 0
    assert torch.allclose(d.cpu(), want_d, rtol=1e-4, atol=1e-9)
    assert float((ind.cpu() != want_ind).float().mean()) < 0.02      # fp32 summation-order ties only


def test_partitioned_knn_self_search_properties():
    """nn_utils.partitioned_kNN semantics at a size the oracle cannot enumerate: every row's nearest neighbour is itself
    at distance 0, distances ascend, and 512 sampled rows agree with an exhaustive fp32 search."""
    from u2seg_b200.clustering import _knn_exhaustive, partitioned_kNN
    N, D, K = 120000, 384, 20
    x = torch.nn.functional.normalize(make_mixture(N, D, 500, seed=3, spread=1.0).float(), dim=1).cuda()
    d, ind = partitioned_kNN(x, K=K, partitions_size=50000)
    assert d.shape == (N, K) and ind.shape == (N, K)
    assert float(d[:, 0].abs().max()) <= 1e-6 and bool((ind[:, 0] == torch.arange(N)).float().mean() > 0.999)
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    rows = torch.randperm(N, generator=torch.Generator().manual_seed(0))[:512]
    d_e, i_e = _knn_exhaustive(x[rows.cuda()], x, K)
    assert torch.allclose(d[rows], d_e.cpu(), rtol=1e-5, atol=1e-6)
