"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference's DINO ViT feature extractor.

Follows u2seg/Instance_Clustering/selective_labeling/dino.py (a copy of facebookresearch/dino's vision_transformer.py):
  :141-155 PatchEmbed.forward      x = proj(x).flatten(2).transpose(1, 2)            (Conv2d kernel = stride = patch)
  :198-216 interpolate_pos_encoding bicubic resize of the patch position table when the token grid differs, scale factors
                                    (w0 + 0.1) / sqrt(N), class position kept
  :218-229 prepare_tokens           [cls | patches] + positions
  :96-118  Attention.forward        qkv -> (3, B, heads, N, hd); softmax(q k^T * hd^-0.5) v; proj
  :75-93   Mlp.forward              fc2(gelu(fc1(x)))     (nn.GELU: exact erf form)
  :121-139 Block.forward            x + attn(norm1(x)); x + mlp(norm2(x))    (drop / drop_path are identities in eval)
  :231-236 VisionTransformer.forward / :293-305 ViTFeat.forward   norm(x)[:, 0]  -> (B, embed_dim) CLS features
  :266-270 / :272-276 vit_small (384, 12 blocks, 6 heads) / vit_base (768, 12, 12): qkv_bias=True, LayerNorm eps 1e-6
and shared/utils/nn_utils.py:155-199 get_feats_list (features of every batch stacked in dataset order).

Pinned against the reference itself: oracle/make_golden.py `dino` instantiates the unmodified VisionTransformer, loads
`init_params(...)` below into it (same state_dict names) and stores its outputs; tests/test_dino_oracle.py compares.
"""
import math

import torch
import torch.nn.functional as F


class ViTCfg:
    def __init__(self, patch_size=8, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, img_size=224, eps=1e-6):
        self.patch_size, self.embed_dim, self.depth, self.num_heads = patch_size, embed_dim, depth, num_heads
        self.mlp_ratio, self.img_size, self.eps = mlp_ratio, img_size, eps

    @property
    def num_patches(self):
        return (self.img_size // self.patch_size) ** 2


def init_params(cfg, seed):
    """A reproducible random state_dict with the reference's names and shapes (dino.py:158-196). Unlike the reference's
    initialisation (zero biases, unit LayerNorm) every tensor is random, so that a dropped bias or a swapped norm shows."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    D, p, H = cfg.embed_dim, cfg.patch_size, int(cfg.embed_dim * cfg.mlp_ratio)
    sd = {"cls_token": rnd(1, 1, D), "pos_embed": rnd(1, cfg.num_patches + 1, D, std=0.2),
          "patch_embed.proj.weight": rnd(D, 3, p, p, std=0.05), "patch_embed.proj.bias": rnd(D, std=0.1)}
    for i in range(cfg.depth):
        b = "blocks.%d." % i
        sd[b + "norm1.weight"] = 1.0 + rnd(D, std=0.1)
        sd[b + "norm1.bias"] = rnd(D, std=0.1)
        sd[b + "attn.qkv.weight"] = rnd(3 * D, D, std=0.05)
        sd[b + "attn.qkv.bias"] = rnd(3 * D, std=0.1)
        sd[b + "attn.proj.weight"] = rnd(D, D, std=0.05)
        sd[b + "attn.proj.bias"] = rnd(D, std=0.1)
        sd[b + "norm2.weight"] = 1.0 + rnd(D, std=0.1)
        sd[b + "norm2.bias"] = rnd(D, std=0.1)
        sd[b + "mlp.fc1.weight"] = rnd(H, D, std=0.05)
        sd[b + "mlp.fc1.bias"] = rnd(H, std=0.1)
        sd[b + "mlp.fc2.weight"] = rnd(D, H, std=0.03)
        sd[b + "mlp.fc2.bias"] = rnd(D, std=0.1)
    sd["norm.weight"] = 1.0 + rnd(D, std=0.1)
    sd["norm.bias"] = rnd(D, std=0.1)
    return sd


def synthetic_images(B, H, W, seed):
    """ImageNet-normalised-looking inputs: N(0,1) per channel, (B,3,H,W) fp32."""
    return torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(seed))


def interpolate_pos_encoding(pos_embed, npatch, w, h, patch_size):
    """dino.py:198-216 (w, h are the image's first / second spatial size, as the reference names them)."""
    N = pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return pos_embed
    class_pos, patch_pos = pos_embed[:, 0], pos_embed[:, 1:]
    dim = pos_embed.shape[-1]
    w0, h0 = w // patch_size + 0.1, h // patch_size + 0.1
    s = int(math.sqrt(N))
    patch_pos = F.interpolate(patch_pos.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                              scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((class_pos.unsqueeze(0), patch_pos), dim=1)


def prepare_tokens(sd, cfg, x):
    B, _, w, h = x.shape
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch_size)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["cls_token"].expand(B, -1, -1), t), dim=1)
    return t + interpolate_pos_encoding(sd["pos_embed"], t.shape[1] - 1, w, h, cfg.patch_size)


def block(sd, cfg, i, x):
    b = "blocks.%d." % i
    B, N, C = x.shape
    nh = cfg.num_heads
    y = F.layer_norm(x, (C,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], cfg.eps)
    qkv = F.linear(y, sd[b + "attn.qkv.weight"], sd[b + "attn.qkv.bias"]).reshape(B, N, 3, nh, C // nh).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * (C // nh) ** -0.5).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, C)
    x = x + F.linear(y, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
    y = F.layer_norm(x, (C,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], cfg.eps)
    y = F.gelu(F.linear(y, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]))
    return x + F.linear(y, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])


def forward_features(sd, cfg, x, return_tokens=False):
    """(B,3,H,W) fp32 -> (B, embed_dim) CLS features (dino.py:231-236); optionally the normalised tokens as well."""
    t = prepare_tokens(sd, cfg, x)
    for i in range(cfg.depth):
        t = block(sd, cfg, i, t)
    t = F.layer_norm(t, (cfg.embed_dim,), sd["norm.weight"], sd["norm.bias"], cfg.eps)
    return (t[:, 0], t) if return_tokens else t[:, 0]


def get_feats_list(sd, cfg, batches):
    """nn_utils.py:155-199 with recompute=True: the features of every batch, stacked in order, as a float tensor."""
    return torch.cat([forward_features(sd, cfg, x) for x in batches], dim=0).float()


GOLDEN_CASES = ["dino_vits8_d3_64x96", "dino_vits16_d2_224", "dino_vitb8_d2_72"]


def load_golden_case(golden_dir, name):
    """fixture of oracle/make_golden.py `dino` -> (npz, cfg, the state_dict it was produced with, its input images)"""
    import os

    import numpy as np
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    p, D, depth, heads, pseed, B, H, W, iseed = [int(v) for v in g["meta"]]
    cfg = ViTCfg(patch_size=p, embed_dim=D, depth=depth, num_heads=heads)
    return g, cfg, init_params(cfg, pseed), synthetic_images(B, H, W, iseed)
