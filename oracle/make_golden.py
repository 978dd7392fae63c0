"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this
container through the dependency stubs in oracle/ref_stubs (test infrastructure only).

    python oracle/make_golden.py [kmeans] [knn] [dino] [ops] [model] [baseline] ...

The fixtures travel to the GPU box; /root/reference does not. Each fixture stores the inputs
(or the seed that regenerates them) and the reference's outputs.
"""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
STUBS = os.path.join(ROOT, "oracle", "ref_stubs")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _kmeans_child():
    """Runs inside selective_labeling/ with USL_MODE set (the reference resolves its config from cwd)."""
    sys.path.insert(0, STUBS)
    sys.path.insert(0, ".")
    import utils  # reference: shared/utils via the selective_labeling/utils symlink
    from oracle.kmeans_oracle import make_mixture
    cases = [
        # name, N, D, K, Niter, modes, seed, spread (chosen so that no cluster empties: the
        # reference's dense branch collapses through torch.argmin-of-NaN once one does)
        ("kmeans_n2048_d128_k40", 2048, 128, 40, 6, 64, 3, 1.0),
        ("kmeans_n4096_d384_k64", 4096, 384, 64, 5, 96, 0, 1.0),
        ("kmeans_n3000_d64_k170", 3000, 64, 170, 4, 300, 1, 0.7),
    ]
    for name, N, D, K, Niter, modes, seed, spread in cases:
        x16 = make_mixture(N, D, modes, seed=100 + seed, spread=spread)
        x = x16.float()
        cl, c = utils.KMeans(x, seed, K=K, Niter=Niter, verbose=False, force_no_lazy_tensor=True)
        torch.manual_seed(seed)
        r = torch.randperm(N)[:K]
        out = dict(labels=cl.numpy().astype(np.int64), centroids=c.numpy(), init=r.numpy(),
                   meta=np.array([N, D, K, Niter, modes, seed, int(spread * 1000)], dtype=np.int64))
        if N * D <= 2048 * 128:
            out["x16"] = x16.numpy()   # small case: store the inputs too
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        assert int((torch.bincount(cl, minlength=K) == 0).sum()) == 0, "fixture must not contain empty clusters"
        print("wrote", name)


def _knn_child():
    """Runs inside selective_labeling/ (like _kmeans_child): the unmodified nn_utils.kNN through the dense pykeops stub."""
    sys.path.insert(0, STUBS)
    sys.path.insert(0, ".")
    import utils
    from oracle.kmeans_oracle import make_mixture
    cases = [("knn_n3000_d128_k20", 3000, 1200, 128, 20, 7), ("knn_self_n2500_d384_k20", 2500, 2500, 384, 20, 8)]
    for name, n_train, n_test, D, K, seed in cases:
        xt = make_mixture(n_train, D, 60, seed=seed, spread=1.0).float()
        xq = xt if n_test == n_train else make_mixture(n_test, D, 60, seed=seed + 100, spread=1.0).float()
        ind, d = utils.kNN(xt, xq, K=K)          # (ind_knn, d_knn): nn_utils.py:224 returns (d_knn, ind_knn) from KeOps swapped
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), ind=ind.numpy().astype(np.int64), dist=d.numpy(),
                            meta=np.array([n_train, n_test, D, K, seed], dtype=np.int64))
        print("wrote", name, tuple(ind.shape), float(d[:, 0].max()))


def gen_knn():
    env = dict(os.environ, USL_MODE="USL", PYTHONPATH=ROOT)
    cwd = os.path.join(REF, "u2seg", "Instance_Clustering", "selective_labeling")
    subprocess.check_call([sys.executable, os.path.abspath(__file__), "_knn_child"], cwd=cwd, env=env)


def gen_kmeans():
    env = dict(os.environ, USL_MODE="USL", PYTHONPATH=ROOT)
    cwd = os.path.join(REF, "u2seg", "Instance_Clustering", "selective_labeling")
    subprocess.check_call([sys.executable, os.path.abspath(__file__), "_kmeans_child"], cwd=cwd, env=env)


DINO_CASES = [
    # name, (patch, embed_dim, depth, heads), param seed, (B, H, W), image seed
    ("dino_vits8_d3_64x96", (8, 384, 3, 6), 11, (2, 64, 96), 21),        # non-square: position table resized to 8 x 12
    ("dino_vits16_d2_224", (16, 384, 2, 6), 12, (1, 224, 224), 22),       # native grid: position table used as it is
    ("dino_vitb8_d2_72", (8, 768, 2, 12), 13, (2, 72, 72), 23),           # ViT-B widths (K = 768 / 3072)
]


def gen_dino():
    """The UNMODIFIED reference VisionTransformer (selective_labeling/dino.py) with oracle.dino_oracle.init_params loaded
    through its own load_state_dict(strict=True); outputs: CLS features (its forward) and the normalised tokens of the
    last layer (get_intermediate_layers(n=1))."""
    import importlib.util
    from functools import partial
    from oracle import dino_oracle as vo
    spec = importlib.util.spec_from_file_location("ref_dino", os.path.join(REF, "u2seg", "Instance_Clustering", "selective_labeling", "dino.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name, (p, D, depth, heads), pseed, (B, H, W), iseed in DINO_CASES:
        cfg = vo.ViTCfg(patch_size=p, embed_dim=D, depth=depth, num_heads=heads)
        torch.manual_seed(0)
        model = ref.VisionTransformer(patch_size=p, embed_dim=D, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=0).eval()
        sd = vo.init_params(cfg, pseed)
        model.load_state_dict(sd, strict=True)
        x = vo.synthetic_images(B, H, W, iseed)
        with torch.no_grad():
            feats = model(x)
            tokens = model.get_intermediate_layers(x, n=1)[0]
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), feats=feats.numpy(), tokens_head=tokens[:, :5].numpy(),
                            tokens_abs_sum=np.array([float(tokens.double().abs().sum())]),
                            meta=np.array([p, D, depth, heads, pseed, B, H, W, iseed], dtype=np.int64))
        print("wrote", name, tuple(feats.shape), float(feats.abs().mean()))


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    what = sys.argv[1:] or ["kmeans", "ops", "model"]
    if "_kmeans_child" in what:
        _kmeans_child()
        sys.exit(0)
    if "_knn_child" in what:
        _knn_child()
        sys.exit(0)
    if "kmeans" in what:
        gen_kmeans()
    if "knn" in what:
        gen_knn()
    if "dino" in what:
        gen_dino()
    if {"ops", "model", "baseline", "baseline64"} & set(what):
        sys.path.insert(0, STUBS)
        sys.path.insert(0, REF)
        from oracle import make_golden_detector as mgd
        if "ops" in what:
            mgd.gen_ops(GOLD)
        if "model" in what:
            mgd.gen_model(GOLD)
        if "baseline" in what:      # BASELINE config 2 (2 x 1024x1024, G=20): ~1 min per fixture on 8 vCPU
            mgd.gen_model_baseline(GOLD)
        if "baseline64" in what:    # float64 run of the same step: the rounding-noise yardstick (several minutes)
            mgd.gen_model_baseline_fp64(GOLD)
