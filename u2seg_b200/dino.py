"""DINO ViT feature extractor: the producer of the N x D embeddings that k-means / kNN consume.

Reference: u2seg/Instance_Clustering/selective_labeling/dino.py (VisionTransformer :158-236, vit_small / vit_base :266-276,
ViTFeat :278-308) and shared/utils/nn_utils.py:155-199 (get_feats_list). Same constructor arguments, module names and
state_dict (DINO checkpoints load with strict=True), same outputs: (B, embed_dim) CLS features after the final LayerNorm.

B200 path: every Linear layer - patch embedding (a stride = kernel convolution is a GEMM over unfolded patches), qkv,
attention projection, both MLP layers: > 99 % of the non-attention flop - runs on the 2-CTA tcgen05 GEMM of csrc/conv2.cu
through `modeling.conv_tc.linear` (bf16 or fp16 operands, fp32 accumulation, bias fused). The residual stream and the
LayerNorms stay in fp32. LayerNorm, GELU and softmax attention are library calls (ATen, `scaled_dot_product_attention`)
in this first version of the row. Inference only (the reference extracts features under no_grad); no CPU path.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .modeling import conv_tc


class Mlp(nn.Module):
    """dino.py:75-93"""

    def __init__(self, in_features, hidden_features=None, out_features=None):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)


class Attention(nn.Module):
    """dino.py:96-118"""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class Block(nn.Module):
    """dino.py:121-139"""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio))


class PatchEmbed(nn.Module):
    """dino.py:141-155"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size // patch_size) * (img_size // patch_size)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


def _tc_linear(x2d, weight, bias, dtype):
    """(M, K) @ (Nout, K)^T + bias on the tcgen05 GEMM. Rows are padded to the kernel's 128-row tile so that every tile is
    full (a few zero rows at most), operands are cast to `dtype`; returns (M, Nout) in `dtype`."""
    if not x2d.is_cuda:
        raise RuntimeError("u2seg_b200.dino needs a CUDA device (no CPU path)")
    M = x2d.shape[0]
    Mp = (M + 127) // 128 * 128
    xh = x2d.to(dtype)
    if Mp != M:
        xh = F.pad(xh, (0, 0, 0, Mp - M))
    assert conv_tc.linear_eligible(xh.contiguous(), weight), "ViT layer shapes are multiples of 64: %r x %r" % (tuple(xh.shape), tuple(weight.shape))
    y = conv_tc.linear(xh.contiguous(), weight, bias, False)
    return y[:M]


class VisionTransformer(nn.Module):
    """dino.py:158-236. `compute_dtype`: operand type of the tensor-core GEMMs (bf16 default; fp16 is 8x finer and safe
    for DINO's activation range)."""

    def __init__(self, img_size=[224], patch_size=16, in_chans=3, num_classes=0, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, compute_dtype=torch.bfloat16, **kwargs):
        super().__init__()
        assert num_classes == 0, "feature extractor: the reference builds it with num_classes=0 (dino.py:283-287)"
        self.num_features = self.embed_dim = embed_dim
        self.compute_dtype = compute_dtype
        self.patch_embed = PatchEmbed(img_size=img_size[0], patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                           qk_scale=qk_scale, norm_layer=norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        self.apply(self._init_weights)
        self._w16 = {}

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # ---- the reference's helpers -----------------------------------------------------------------------------------
    def interpolate_pos_encoding(self, x, w, h):
        """dino.py:198-216"""
        npatch = x.shape[1] - 1
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        class_pos_embed, patch_pos_embed = self.pos_embed[:, 0], self.pos_embed[:, 1:]
        dim = x.shape[-1]
        w0, h0 = w // self.patch_embed.patch_size + 0.1, h // self.patch_embed.patch_size + 0.1
        s = int(math.sqrt(N))
        patch_pos_embed = F.interpolate(patch_pos_embed.reshape(1, s, s, dim).permute(0, 3, 1, 2),
                                        scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
        assert int(w0) == patch_pos_embed.shape[-2] and int(h0) == patch_pos_embed.shape[-1]
        patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).reshape(1, -1, dim)
        return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)

    def _weight(self, p):
        """compute-dtype copy of a weight, refreshed when the parameter changes (load_state_dict, optimizer step)"""
        key = id(p)
        hit = self._w16.get(key)
        if hit is None or hit[0] != p._version or hit[1].dtype != self.compute_dtype or hit[1].device != p.device:
            w = p.detach().reshape(p.shape[0], -1).to(self.compute_dtype).contiguous()
            self._w16[key] = hit = (p._version, w)
        return hit[1]

    def _linear(self, x2d, lin):
        return _tc_linear(x2d, self._weight(lin.weight), lin.bias, self.compute_dtype)

    def prepare_tokens(self, x, linear=None):
        """dino.py:218-229; the patch convolution (kernel = stride) as a GEMM over unfolded patches"""
        linear = linear or self._linear
        B, nc, w, h = x.shape
        p = self.patch_embed.patch_size
        gw, gh = w // p, h // p
        patches = x[:, :, :gw * p, :gh * p].reshape(B, nc, gw, p, gh, p).permute(0, 2, 4, 1, 3, 5).reshape(B * gw * gh, nc * p * p)
        t = linear(patches, self.patch_embed.proj).reshape(B, gw * gh, self.embed_dim).float()
        t = torch.cat((self.cls_token.expand(B, -1, -1).float(), t), dim=1)
        return t + self.interpolate_pos_encoding(t, w, h).float()

    def _block(self, blk, x, linear):
        B, N, C = x.shape
        nh = blk.attn.num_heads
        y = F.layer_norm(x, (C,), blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
        qkv = linear(y.reshape(B * N, C), blk.attn.qkv).reshape(B, N, 3, nh, C // nh).permute(2, 0, 3, 1, 4)
        y = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=blk.attn.scale)     # softmax(q k^T * scale) v
        y = y.transpose(1, 2).reshape(B * N, C)
        x = x + linear(y, blk.attn.proj).reshape(B, N, C).float()
        y = F.layer_norm(x, (C,), blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
        y = F.gelu(linear(y.reshape(B * N, C), blk.mlp.fc1))
        return x + linear(y, blk.mlp.fc2).reshape(B, N, C).float()

    def _forward_tokens(self, x, linear=None):
        linear = linear or self._linear
        t = self.prepare_tokens(x, linear)
        for blk in self.blocks:
            t = self._block(blk, t, linear)
        return F.layer_norm(t, (self.embed_dim,), self.norm.weight, self.norm.bias, self.norm.eps)

    def forward(self, x):
        """(B,3,H,W) -> (B, embed_dim) fp32 CLS features (dino.py:231-236)"""
        return self._forward_tokens(x)[:, 0]

    def get_intermediate_layers(self, x, n=1):
        """dino.py:248-256, for n = 1 (the normalised tokens of the last block)"""
        assert n == 1, "only the last layer is kept"
        return [self._forward_tokens(x)]


def vit_small(patch_size=16, **kwargs):
    """dino.py:266-270"""
    from functools import partial
    return VisionTransformer(patch_size=patch_size, embed_dim=384, depth=12, num_heads=6, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


def vit_base(patch_size=16, **kwargs):
    """dino.py:272-276"""
    from functools import partial
    return VisionTransformer(patch_size=patch_size, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                             norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)


class ViTFeat(nn.Module):
    """dino.py:278-308. `pretrained_pth`: a local checkpoint file with the DINO state_dict (the reference downloads it by
    URL; there is no network here), or None for random weights."""

    def __init__(self, pretrained_pth, feat_dim, vit_arch="base", vit_feat="k", patch_size=16, compute_dtype=torch.bfloat16):
        super().__init__()
        build = vit_base if vit_arch == "base" else vit_small
        self.model = build(patch_size=patch_size, num_classes=0, compute_dtype=compute_dtype)
        self.feat_dim, self.vit_feat, self.patch_size = feat_dim, vit_feat, patch_size
        if pretrained_pth is not None:
            self.model.load_state_dict(torch.load(pretrained_pth, map_location="cpu"), strict=True)
            print("Loading weight from {}".format(pretrained_pth))

    def forward(self, x):
        return self.model(x)


@torch.no_grad()
def get_feats_list(model, train_memory_loader, feat_dim=None, recompute=True, save_dir=None, **kwargs):
    """nn_utils.py:155-199 for the DINO branch: features of the whole loader in dataset order as a float tensor (N, feat_dim).
    The reference bounces every batch through .cpu().numpy() into a float64 array; here the batches stay on the device and
    one tensor comes back. recompute=False loads memory_feats_list.npy from save_dir (the reference's cfg.RUN_DIR)."""
    import numpy as np
    path = os.path.join(save_dir, "memory_feats_list.npy") if save_dir is not None else None
    if not recompute:
        assert path is not None, "recompute=False needs save_dir"
        return torch.tensor(np.load(path)).float()
    dev = next(model.parameters()).device
    out, targets = [], []
    for images, t in train_memory_loader:
        out.append(model(images.to(dev, non_blocking=True), **kwargs).float())
        targets.append(torch.as_tensor(t))
    feats = torch.cat(out, dim=0)
    if feat_dim is not None:
        assert feats.shape[1] == feat_dim, (feats.shape, feat_dim)
    ds_targets = getattr(getattr(train_memory_loader, "dataset", None), "targets", None)
    if ds_targets is not None:
        assert torch.equal(torch.cat(targets).cpu().long(), torch.as_tensor(ds_targets).long()), "loader must not shuffle"
    if path is not None and not os.path.exists(path):
        np.save(path, feats.cpu().numpy().astype(np.float64))
    print("feats_list:", feats.shape)
    return feats
