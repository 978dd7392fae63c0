"""Host-side logic of u2seg_b200.clustering that needs no GPU: run_kMeans' load path / file conventions (nn_utils.py:382-405)
and the kNN rounding bound."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from u2seg_b200 import clustering


def test_run_kmeans_recompute_false_loads_saved_files_and_writes_decode_json(tmp_path):
    labels = torch.randint(0, 7, (50,))
    cents = torch.randn(7, 16)
    np.save(tmp_path / "cluster_labels_400_3.npy", labels.numpy())
    np.save(tmp_path / "centroids_400_3.npy", cents.numpy())
    ds = SimpleNamespace(imgs=[("/data/train/n%02d/im_%03d.JPEG" % (i % 5, i), 0) for i in range(50)])
    cl, c = clustering.run_kMeans(None, 7, 400, ds, Niter=3, seed=3, save_dir=str(tmp_path))     # default recompute=False
    assert torch.equal(cl, labels) and torch.equal(c, cents)
    dec = json.load(open(tmp_path / "cluster_labels_decode.json"))
    assert len(dec) == 50 and dec["n04/im_049.JPEG"] == int(labels[49])           # nn_utils.py:87-91: last two path components
    # existing files are never overwritten (nn_utils.py:75-107)
    json.dump({"stale": 1}, open(tmp_path / "cluster_labels_decode.json", "w"))
    clustering.run_kMeans(None, 7, 400, ds, Niter=3, seed=3, save_dir=str(tmp_path))
    assert json.load(open(tmp_path / "cluster_labels_decode.json")) == {"stale": 1}
    with pytest.raises(ValueError):
        clustering.run_kMeans(None, 7, 400, ds, Niter=3, seed=3)                  # nowhere to load from
    with pytest.raises(FileNotFoundError):
        clustering.run_kMeans(None, 7, 400, ds, Niter=3, seed=4, save_dir=str(tmp_path))   # other seed: other file name


def test_knn_rounding_bound_covers_the_fp16_candidate_values():
    """|(|y|^2 - 2 x~.y~) - (|y|^2 - 2 x.y)| <= eps for every pair, with eps from the measured rounding-error norms."""
    g = torch.Generator().manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(300, 128, generator=g), dim=1) * 3.0
    y = torch.nn.functional.normalize(torch.randn(500, 128, generator=g), dim=1) * 0.7
    x16, y16 = x.half(), y.half()
    ex = (x - x16.float()).norm(dim=1).max()
    ey = (y - y16.float()).norm(dim=1).max()
    eps = clustering._knn_eps((x * x).sum(1).max(), (y * y).sum(1).max(), ex, ey, 128)
    exact = -2.0 * (x.double() @ y.double().t())
    approx = -2.0 * (x16.float() @ y16.float().t()).double()       # fp32 accumulation like the tensor-core pass
    worst = float((approx - exact).abs().max())
    assert worst <= eps
    assert eps <= 2.0 ** -9 * 3.0 * 0.7 * 1.25                       # tighter than the 2^-11-per-element worst case used before
