"""u2seg_b200.dino host logic on the CPU: module names / state_dict equal the reference's, and the forward wiring (patch
unfolding, position-table resize, head split, residual order) reproduces the reference goldens when the tensor-core GEMM is
replaced by F.linear through the test hook `_forward_tokens(x, linear=...)`. The product entry point itself refuses CPU
tensors (no CPU path)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dino_oracle as vo
CASES, load_case = vo.GOLDEN_CASES, vo.load_golden_case


def _build(cfg, sd):
    from functools import partial
    from u2seg_b200.dino import VisionTransformer
    m = VisionTransformer(patch_size=cfg.patch_size, embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads,
                          mlp_ratio=4, qkv_bias=True, norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=0).eval()
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m


def _cpu_linear(x2d, lin):
    return F.linear(x2d.float(), lin.weight.reshape(lin.weight.shape[0], -1), lin.bias)


@pytest.mark.parametrize("name", CASES)
def test_wiring_matches_reference_golden(golden_dir, name):
    g, cfg, sd, x = load_case(golden_dir, name)
    m = _build(cfg, sd)
    assert list(m.state_dict().keys()) == list(sd.keys())                 # the reference's names, in its order
    with torch.no_grad():
        tokens = m._forward_tokens(x, linear=_cpu_linear)
    np.testing.assert_allclose(tokens[:, 0].numpy(), g["feats"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(tokens[:, :5].numpy(), g["tokens_head"], rtol=1e-4, atol=1e-4)


def test_product_forward_refuses_cpu_tensors(golden_dir):
    g, cfg, sd, x = load_case(golden_dir, CASES[0])
    m = _build(cfg, sd)
    with pytest.raises(RuntimeError):
        m(x)


def test_get_feats_list_load_path(tmp_path):
    from u2seg_b200.dino import get_feats_list
    feats = np.random.RandomState(0).randn(7, 16)
    np.save(tmp_path / "memory_feats_list.npy", feats)
    out = get_feats_list(None, None, recompute=False, save_dir=str(tmp_path))
    assert out.dtype == torch.float32 and np.allclose(out.numpy(), feats.astype(np.float32))
