"""GPU parity: libu2b200 k-means (through the C ABI) vs the oracle / reference golden vectors."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.kmeans_oracle import assign_oracle, kmeans_oracle, make_mixture, update_oracle

pytestmark = pytest.mark.gpu


def _load(path):
    g = np.load(path)
    N, D, K, Niter, modes, seed, spread_m = [int(v) for v in g["meta"]]
    x16 = torch.from_numpy(g["x16"]) if "x16" in g.files else make_mixture(N, D, modes, 100 + seed, spread_m / 1000.0)
    return g, x16, (N, D, K, Niter, seed)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_kmeans_matches_reference_golden(golden_dir, idx):
    from u2seg_b200.clustering import KMeans
    path = sorted(glob.glob(os.path.join(golden_dir, "kmeans_*.npz")))[idx]
    g, x16, (N, D, K, Niter, seed) = _load(path)
    cl, c = KMeans(x16, seed, K=K, Niter=Niter, verbose=False)
    assert cl.dtype == torch.int64 and c.dtype == torch.float32
    assert np.array_equal(cl.cpu().numpy(), g["labels"])                 # INT: bit exact
    np.testing.assert_allclose(c.cpu().numpy(), g["centroids"], rtol=1e-3, atol=1e-5)  # FP: 1e-3


@pytest.mark.parametrize("cluster", [1, 2, 4])
@pytest.mark.parametrize("N,D,K", [(1, 64, 1), (127, 64, 3), (129, 128, 160), (5000, 384, 161), (20000, 384, 800),
                                    (4097, 256, 300), (40000, 384, 800)])
def test_assign_bit_exact_vs_oracle(N, D, K, cluster):
    from u2seg_b200.clustering import KMeansState, set_cluster
    set_cluster(cluster)
    x16 = make_mixture(N, D, max(2, K + 7), seed=N + K, spread=1.0)
    g = torch.Generator().manual_seed(5)
    c = x16.float()[torch.randint(0, N, (K,), generator=g)] + 0.01 * torch.randn(K, D, generator=g)
    want = assign_oracle(x16.float(), c, chunk=512)
    st = KMeansState(x16.cuda(), K)
    got = st.assign(c.cuda().contiguous()).cpu().long()
    bad = (got != want).nonzero().flatten()
    if bad.numel():  # only fp32-rounding-level ties may differ: check and bound them
        xd, cd = x16.double(), c.double()
        for i in bad.tolist():
            dg = ((xd[i] - cd[got[i]]) ** 2).sum()
            dw = ((xd[i] - cd[want[i]]) ** 2).sum()
            assert abs(dg - dw) <= 2e-6 * max(1.0, float(dw)), (i, float(dg), float(dw))
        assert bad.numel() <= max(1, N // 100000), "too many near-tie differences: %d" % bad.numel()


def test_assign_full_baseline_size_sampled_rows_vs_oracle():
    """BASELINE config 4 itself (N=1.28 M, D=384, K=800): the E-step over ALL rows, then 65,536 sampled rows against the chunked
    oracle (nn_utils.py:342-355 dense branch). Same acceptance as the small cases: identical labels except fp32-rounding ties,
    each checked in float64."""
    from u2seg_b200.clustering import KMeansState, set_cluster
    set_cluster(2)
    N, D, K, S = 1_280_000, 384, 800, 65_536
    x16 = make_mixture(N, D, 1000, seed=3, spread=1.0)
    g = torch.Generator().manual_seed(9)
    c = x16.float()[torch.randint(0, N, (K,), generator=g)] + 0.01 * torch.randn(K, D, generator=g)
    st = KMeansState(x16.cuda(), K)
    got_all = st.assign(c.cuda().contiguous()).cpu().long()
    assert got_all.shape == (N,) and int(got_all.min()) >= 0 and int(got_all.max()) < K
    rows = torch.randperm(N, generator=g)[:S]
    rows = torch.cat([rows, torch.tensor([0, 127, 128, N - 129, N - 128, N - 1])])   # tile edges of the first / last CTA
    want = assign_oracle(x16[rows].float(), c, chunk=512)
    got = got_all[rows]
    bad = (got != want).nonzero().flatten()
    print("full-size E-step: %d of %d sampled rows differ from the oracle (rounding ties)" % (bad.numel(), rows.numel()))
    if bad.numel():
        xd, cd = x16[rows].double(), c.double()
        for i in bad.tolist():
            dg = ((xd[i] - cd[got[i]]) ** 2).sum()
            dw = ((xd[i] - cd[want[i]]) ** 2).sum()
            assert abs(dg - dw) <= 2e-6 * max(1.0, float(dw)), (i, float(dg), float(dw))
    assert bad.numel() <= 2


def test_assign_duplicate_centroids_first_minimum():
    from u2seg_b200.clustering import KMeansState
    x16 = make_mixture(1000, 64, 10, seed=1, spread=1.0)
    c = x16.float()[:6].clone()
    c[4] = c[1]          # exact duplicate -> the lower index must win (nn_utils.py:355 argmin)
    c[5] = c[0]
    want = assign_oracle(x16.float(), c)
    got = KMeansState(x16.cuda(), 6).assign(c.cuda()).cpu().long()
    assert torch.equal(got, want)
    assert not ((got == 4) | (got == 5)).any()


def test_nan_centroid_never_selected_and_stays_nan():
    from u2seg_b200.clustering import KMeansState
    x16 = make_mixture(3000, 128, 12, seed=2, spread=1.0)
    K = 9
    c = x16.float()[:K].clone()
    c[3] = float("nan")
    want = assign_oracle(x16.float(), c)
    st = KMeansState(x16.cuda(), K)
    cd = c.cuda().contiguous()
    got = st.assign(cd).cpu().long()
    assert torch.equal(got, want) and not (got == 3).any()
    st.accumulate()
    st.finalize(cd)
    wc, wn = update_oracle(x16.float(), want, K)
    assert torch.isnan(cd[3]).all()
    ok = ~torch.isnan(wc)
    np.testing.assert_allclose(cd.cpu()[ok].numpy(), wc[ok].numpy(), rtol=1e-3, atol=1e-5)
    assert torch.equal(st.sums[:, -1].cpu(), wn)


def test_accumulate_matches_oracle_k300_d384():
    from u2seg_b200.clustering import KMeansState
    N, D, K = 30000, 384, 300
    x16 = make_mixture(N, D, 400, seed=9, spread=1.0)
    lab = torch.randint(0, K, (N,), generator=torch.Generator().manual_seed(1))
    st = KMeansState(x16.cuda(), K)
    st.labels.copy_(lab.int().cuda())
    sums = st.accumulate().cpu()
    wc = torch.zeros(K, D).scatter_add_(0, lab[:, None].repeat(1, D), x16.float())
    np.testing.assert_allclose(sums[:, :D].numpy(), wc.numpy(), rtol=1e-3, atol=1e-4)
    assert torch.equal(sums[:, D], torch.bincount(lab, minlength=K).float())


def test_full_size_properties():
    """BASELINE size class (N large): size-independent properties instead of a full CPU oracle."""
    from u2seg_b200.clustering import KMeansState
    N, D, K = 262144, 384, 800
    x16 = make_mixture(N, D, 1000, seed=0, spread=1.0).cuda()
    c = x16[torch.randperm(N, generator=torch.Generator().manual_seed(0))[:K].cuda()].float().contiguous()
    st = KMeansState(x16, K)
    lab = st.assign(c).long()
    assert int(lab.min()) >= 0 and int(lab.max()) < K
    # (1) optimality on a sample, in fp64: the chosen centroid is the nearest up to fp32 rounding
    idx = torch.randperm(N, generator=torch.Generator().manual_seed(1))[:4096].cuda()
    d = torch.cdist(x16[idx].double(), c.double()) ** 2
    chosen = d.gather(1, lab[idx][:, None]).squeeze(1)
    assert bool((chosen <= d.min(1).values * (1 + 2e-6) + 1e-9).all())
    # (2) idempotence: a second assignment with the same centroids gives identical labels
    assert torch.equal(st.assign(c).long(), lab)
    # (3) checksum of checksums: counts sum to N, sums add up to the column sums of X
    sums = st.accumulate()
    assert float(sums[:, -1].sum()) == N
    np.testing.assert_allclose(sums[:, :D].sum(0).cpu().numpy(), x16.float().sum(0).cpu().numpy(), rtol=2e-3, atol=1e-2)
    # (4) a Lloyd step never increases the objective
    def objective(cc, ll):
        return float(((x16.float() - cc[ll]) ** 2).sum())
    before = objective(c, lab)
    st.finalize(c)
    lab2 = st.assign(c).long()
    assert objective(c, lab2) <= before * (1 + 1e-6)


@pytest.mark.parametrize("N,D,K", [(50000, 384, 800), (4097, 128, 37), (20000, 256, 300)])
def test_mstep_sort_by_label_equals_shared_accumulators(N, D, K):
    """The two M-step implementations (csrc/kmeans.cu: counting sort + segment sums vs shared-memory accumulators) agree:
    counts exactly, sums to fp32 summation-order accuracy; also against a float64 scatter-add of the same fp16 data."""
    from u2seg_b200 import _lib
    from u2seg_b200.clustering import KMeansState
    g = torch.Generator().manual_seed(N)
    x16 = torch.randn(N, D, generator=g).half().cuda()
    st = KMeansState(x16, K)
    st.labels.copy_(torch.randint(0, K, (N,), generator=g).int().cuda())
    st.labels[:100] = 3                                   # one long run + some empty clusters possible
    outs = []
    for mode in (1, 0):
        _lib.check(_lib.lib().u2b_kmeans_set_mstep(mode), "u2b_kmeans_set_mstep")
        try:
            outs.append(st.accumulate().clone())
        finally:
            _lib.lib().u2b_kmeans_set_mstep(1)
    a, b = outs
    assert torch.equal(a[:, D], b[:, D])                  # counts
    want = torch.zeros(K, D, dtype=torch.float64, device="cuda").index_add_(0, st.labels.long(), x16.double())
    for got in (a, b):
        assert float((got[:, :D].double() - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max()))
    assert torch.equal(a[:, D].double(), torch.bincount(st.labels.long(), minlength=K).double())


def test_run_kmeans_save_load_and_decode_json(tmp_path):
    """nn_utils.py:382-405: recompute=True saves labels / centroids (never overwriting), recompute=False loads them back,
    cluster_labels_decode.json maps the last two path components of every image to its cluster id."""
    import json
    from types import SimpleNamespace
    from u2seg_b200.clustering import run_kMeans
    x = make_mixture(3000, 64, 12, seed=2, spread=1.0)
    ds = SimpleNamespace(imgs=[("/data/imagenet/train/n%04d/img_%05d.JPEG" % (i % 7, i), 0) for i in range(3000)])
    cl, c = run_kMeans(x, 10, 400, ds, Niter=5, recompute=True, seed=3, save=True, save_dir=str(tmp_path))
    assert (tmp_path / "cluster_labels_400_3.npy").exists() and (tmp_path / "centroids_400_3.npy").exists()
    cl2, c2 = run_kMeans(None, 10, 400, ds, Niter=5, seed=3, save_dir=str(tmp_path))        # reference default: recompute=False
    assert torch.equal(cl, cl2) and torch.equal(c, c2)
    dec = json.load(open(tmp_path / "cluster_labels_decode.json"))
    assert len(dec) == 3000 and dec["n0003/img_00003.JPEG"] == int(cl[3])
    with pytest.raises(ValueError):
        run_kMeans(None, 10, 400, None, Niter=5, seed=3)                                   # nothing to load from
    with pytest.raises(ValueError):
        run_kMeans(x.float() * 1e6, 10, 401, None, Niter=2, recompute=True, seed=3)      # beyond fp16: refused, not inf
