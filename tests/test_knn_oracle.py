"""CPU: the kNN oracle (oracle/knn_oracle.py, restatement of nn_utils.py:203-224) against the fixtures recorded from the
UNMODIFIED reference (`python oracle/make_golden.py knn`: nn_utils.kNN through the dense pykeops stub)."""
import os

import numpy as np
import pytest
import torch

from oracle.kmeans_oracle import make_mixture
from oracle.knn_oracle import knn_oracle


def _case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    n_train, n_test, D, K, seed = [int(v) for v in g["meta"]]
    xt = make_mixture(n_train, D, 60, seed=seed, spread=1.0).float()
    xq = xt if n_test == n_train else make_mixture(n_test, D, 60, seed=seed + 100, spread=1.0).float()
    return g, xt, xq, K


@pytest.mark.parametrize("name", ["knn_n3000_d128_k20", "knn_self_n2500_d384_k20"])
def test_knn_oracle_matches_reference(golden_dir, name):
    g, xt, xq, K = _case(golden_dir, name)
    ind, d = knn_oracle(xt, xq, K)
    assert torch.equal(d, torch.from_numpy(g["dist"]))            # same dense formula, same torch ops: bit-identical
    assert torch.equal(ind, torch.from_numpy(g["ind"]))
