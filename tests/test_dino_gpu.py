"""GPU parity of u2seg_b200.dino (Linear layers on the tcgen05 GEMM through the C ABI) against the reference ViT's goldens
and, at the benchmark's size, against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import dino_oracle as vo

pytestmark = pytest.mark.gpu


def _build(cfg, sd, dtype):
    from functools import partial
    from u2seg_b200.dino import VisionTransformer
    m = VisionTransformer(patch_size=cfg.patch_size, embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads,
                          mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=0,
                          compute_dtype=dtype)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _errors(got, want):
    got, want = got.double(), want.double()
    rel = float((got - want).norm() / want.norm())
    cos = float(torch.nn.functional.cosine_similarity(got, want, dim=-1).min())
    return rel, cos


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 3e-2), (torch.float16, 6e-3)])
@pytest.mark.parametrize("name", vo.GOLDEN_CASES)
def test_features_match_reference_golden(golden_dir, name, dtype, tol):
    """2-3 blocks, ViT-S and ViT-B widths, patch 8 and 16, resized and native position tables: CLS features and the first
    tokens within the rounding of 16-bit GEMM operands (fp32 accumulation, fp32 residual stream)."""
    from u2seg_b200 import _lib
    g, cfg, sd, x = vo.load_golden_case(golden_dir, name)
    m = _build(cfg, sd, dtype)
    l0 = _lib.launch_count
    with torch.no_grad():
        tokens = m.get_intermediate_layers(x.cuda(), n=1)[0].float().cpu()
        feats = m(x.cuda()).float().cpu()
    assert _lib.launch_count - l0 >= 2 * (1 + 4 * cfg.depth), "the Linear layers must run on the tcgen05 kernel"
    assert feats.shape == g["feats"].shape and feats.dtype == torch.float32
    rel, cos = _errors(feats, torch.from_numpy(g["feats"]))
    rel_t, _ = _errors(tokens[:, :5], torch.from_numpy(g["tokens_head"]))
    print("%s %s: CLS relative L2 error %.2e (min cosine %.6f), first tokens %.2e" % (name, dtype, rel, cos, rel_t))
    assert rel <= tol and rel_t <= tol and cos >= 1 - tol


def test_vits8_full_depth_480_matches_oracle():
    """The benchmark's configuration (ViT-S/8, 12 blocks, 480x480 -> 3601 tokens, a row count that is not a multiple of the
    GEMM's 128-row tile) against the CPU oracle with the same random weights."""
    cfg = vo.ViTCfg(patch_size=8, embed_dim=384, depth=12, num_heads=6)
    sd = vo.init_params(cfg, 5)
    x = vo.synthetic_images(1, 480, 480, 6)
    with torch.no_grad():
        want = vo.forward_features(sd, cfg, x)
        got = _build(cfg, sd, torch.bfloat16)(x.cuda()).float().cpu()
        got16 = _build(cfg, sd, torch.float16)(x.cuda()).float().cpu()
    rel, cos = _errors(got, want)
    rel16, cos16 = _errors(got16, want)
    print("ViT-S/8 480x480, 12 blocks: bf16 relative L2 error %.2e (cosine %.6f); fp16 %.2e (%.6f)" % (rel, cos, rel16, cos16))
    assert torch.isfinite(got).all() and rel <= 6e-2 and cos >= 0.995
    assert rel16 <= 1.5e-2 and cos16 >= 0.9995


def test_get_feats_list_keeps_dataset_order():
    from types import SimpleNamespace
    from u2seg_b200.dino import get_feats_list
    cfg = vo.ViTCfg(patch_size=16, embed_dim=384, depth=1, num_heads=6)
    m = _build(cfg, vo.init_params(cfg, 2), torch.bfloat16)
    x = vo.synthetic_images(5, 64, 64, 3)
    loader = [(x[:2], torch.tensor([0, 1])), (x[2:], torch.tensor([2, 3, 4]))]
    feats = get_feats_list(m, loader, feat_dim=384)
    with torch.no_grad():
        whole = m(x.cuda()).float()
    assert feats.shape == (5, 384) and torch.allclose(feats, whole, rtol=2e-2, atol=2e-2)
