"""oracle/dino_oracle.py against the UNMODIFIED reference ViT (fixtures from `python oracle/make_golden.py dino`)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import dino_oracle as vo

CASES = vo.GOLDEN_CASES
load_case = vo.load_golden_case


def test_dino_goldens_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "dino_*.npz"))) >= 3


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_vit(golden_dir, name):
    g, cfg, sd, x = load_case(golden_dir, name)
    with torch.no_grad():
        feats, tokens = vo.forward_features(sd, cfg, x, return_tokens=True)
    np.testing.assert_allclose(feats.numpy(), g["feats"], rtol=1e-5, atol=1e-5)          # fp32, same torch CPU ops
    np.testing.assert_allclose(tokens[:, :5].numpy(), g["tokens_head"], rtol=1e-5, atol=1e-5)
    assert abs(float(tokens.double().abs().sum()) - float(g["tokens_abs_sum"][0])) <= 1e-6 * float(g["tokens_abs_sum"][0])


def test_get_feats_list_stacks_batches_in_order(golden_dir):
    g, cfg, sd, x = load_case(golden_dir, CASES[0])
    with torch.no_grad():
        out = vo.get_feats_list(sd, cfg, [x[:1], x[1:]])
    np.testing.assert_allclose(out.numpy(), g["feats"], rtol=1e-5, atol=1e-5)
