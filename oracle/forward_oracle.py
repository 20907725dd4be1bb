# -*- coding: utf-8 -*-
"""Plain-PyTorch fp32 restatement of the reference forward -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A *functional* re-derivation (state_dict in, dict of head maps out) of
  models/segmentation/cell_segmentation/cellvit.py:153-244   (CellViT.forward, _forward_upsample)
  models/segmentation/cell_segmentation/cellvit.py:586-644   (CellViTSAM.forward)
  models/segmentation/cell_segmentation/utils.py:11-86,149-233 (conv blocks, encoder wrappers)
  models/encoders/VIT/SAM/image_encoder.py:177-392            (block, attention, windows, rel-pos)
  models/encoders/VIT/vits_histo.py:172-188,241-247,377-415   (ViT-S block, pos-embed interpolation)
used as the floating-point checker of the CUDA path (tolerance 1e-3 abs on NP/HV/NT, BASELINE.json).
Pinned against the reference modules imported from /root/reference (tests/test_oracle_vs_reference.py,
exact equality on CPU) and against golden fixtures under tests/golden/. The shared skip decoders
(decoder0..3) are evaluated once, not once per branch -- identical results in eval mode.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

SAM_CFG = {
    "SAM-B": dict(embed_dim=768, depth=12, num_heads=12, global_idx=(2, 5, 8, 11), extract=(3, 6, 9, 12)),
    "SAM-L": dict(embed_dim=1024, depth=24, num_heads=16, global_idx=(5, 11, 17, 23), extract=(6, 12, 18, 24)),
    "SAM-H": dict(embed_dim=1280, depth=32, num_heads=16, global_idx=(7, 15, 23, 31), extract=(8, 16, 24, 32)),
}
VIT256_CFG = dict(embed_dim=384, depth=12, num_heads=6, extract=(3, 6, 9, 12))


def _ln(x, sd, prefix, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


# ----------------------------------------------------------------------------- SAM encoder
def _rel_table(q_size, k_size, rel_pos):
    """image_encoder.py:321-351."""
    L = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != L:
        r = F.interpolate(rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
        r = r.reshape(-1, L).permute(1, 0)
    else:
        r = rel_pos
    qc = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    kc = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    rel = (qc - kc) + (k_size - 1) * max(q_size / k_size, 1.0)
    return r[rel.long()]


def _sam_attention(x, sd, pfx, heads):
    """image_encoder.py:235-260; x [B',h,w,D]."""
    Bp, h, w, D = x.shape
    hd = D // heads
    qkv = _lin(x, sd, pfx + ".qkv").reshape(Bp, h * w, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv.reshape(3, Bp * heads, h * w, hd).unbind(0)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh = _rel_table(h, h, sd[pfx + ".rel_pos_h"])
    Rw = _rel_table(w, w, sd[pfx + ".rel_pos_w"])
    rq = q.reshape(Bp * heads, h, w, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    attn = (attn.view(-1, h, w, h, w) + rel_h[..., :, None] + rel_w[..., None, :]).view(-1, h * w, h * w)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).view(Bp, heads, h, w, hd).permute(0, 2, 3, 1, 4).reshape(Bp, h, w, D)
    return _lin(o, sd, pfx + ".proj")


def _sam_block(x, sd, pfx, heads, window):
    """image_encoder.py:177-193 with window partition :263-318."""
    B, H, W, D = x.shape
    y = _ln(x, sd, pfx + ".norm1")
    if window > 0:
        ph, pw = (window - H % window) % window, (window - W % window) % window
        y = F.pad(y, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        y = y.view(B, Hp // window, window, Wp // window, window, D).permute(0, 1, 3, 2, 4, 5)
        y = y.reshape(-1, window, window, D)
    y = _sam_attention(y, sd, pfx + ".attn", heads)
    if window > 0:
        y = y.view(B, Hp // window, Wp // window, window, window, D).permute(0, 1, 3, 2, 4, 5)
        y = y.reshape(B, Hp, Wp, D)[:, :H, :W, :]
    x = x + y
    m = _lin(F.gelu(_lin(_ln(x, sd, pfx + ".norm2"), sd, pfx + ".mlp.lin1")), sd, pfx + ".mlp.lin2")
    return x + m


def _ln2d(x, w, b, eps=1e-6):
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


def sam_encoder(sd, x, cfg, window=14):
    """utils.py:218-233. Returns (tissue-feature [B,256], skips z1..z4 as [B,D,h,w])."""
    t = F.conv2d(x, sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"], stride=16)
    t = t.permute(0, 2, 3, 1)
    t = t + sd["encoder.pos_embed"][:, : t.shape[1], : t.shape[1], :]
    skips = []
    for i in range(cfg["depth"]):
        t = _sam_block(t, sd, f"encoder.blocks.{i}", cfg["num_heads"], 0 if i in cfg["global_idx"] else window)
        if i + 1 in cfg["extract"]:
            skips.append(t.permute(0, 3, 1, 2))
    n = F.conv2d(t.permute(0, 3, 1, 2), sd["encoder.neck.0.weight"])
    n = _ln2d(n, sd["encoder.neck.1.weight"], sd["encoder.neck.1.bias"])
    n = F.conv2d(n, sd["encoder.neck.2.weight"], padding=1)
    n = _ln2d(n, sd["encoder.neck.3.weight"], sd["encoder.neck.3.bias"])
    return n.flatten(2).mean(-1), skips


# ----------------------------------------------------------------------------- ViT-256 encoder
def _vit_pos(sd, n_tok, w, h, dim):
    """vits_histo.py:377-402 (bicubic resize of the 14x14 grid with the +0.1 scale-factor trick)."""
    pe = sd["encoder.pos_embed"]
    N = pe.shape[1] - 1
    if n_tok == N and w == h:
        return pe
    w0, h0 = w // 16 + 0.1, h // 16 + 0.1
    g = int(math.sqrt(N))
    p = F.interpolate(pe[:, 1:].reshape(1, g, g, dim).permute(0, 3, 1, 2),
                      scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == p.shape[-2] and int(h0) == p.shape[-1]
    return torch.cat((pe[:, :1], p.permute(0, 2, 3, 1).reshape(1, -1, dim)), dim=1)


def vit256_encoder(sd, x, cfg):
    """utils.py:149-174. Returns (tissue logits, skips as [B,D,h,w])."""
    B, _, Hh, Ww = x.shape
    D, heads = cfg["embed_dim"], cfg["num_heads"]
    t = F.conv2d(x, sd["encoder.patch_embed.proj.weight"], sd["encoder.patch_embed.proj.bias"], stride=16)
    gh, gw = t.shape[-2:]
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd["encoder.cls_token"].expand(B, -1, -1), t), dim=1)
    t = t + _vit_pos(sd, t.shape[1] - 1, Hh, Ww, D)
    skips = []
    for i in range(cfg["depth"]):
        p = f"encoder.blocks.{i}"
        y = _ln(t, sd, p + ".norm1")
        N = y.shape[1]
        qkv = _lin(y, sd, p + ".attn.qkv").reshape(B, N, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
        a = ((qkv[0] @ qkv[1].transpose(-2, -1)) * (D // heads) ** -0.5).softmax(dim=-1)
        t = t + _lin((a @ qkv[2]).transpose(1, 2).reshape(B, N, D), sd, p + ".attn.proj")
        t = t + _lin(F.gelu(_lin(_ln(t, sd, p + ".norm2"), sd, p + ".mlp.fc1")), sd, p + ".mlp.fc2")
        if i + 1 in cfg["extract"]:
            skips.append(t[:, 1:, :].transpose(-1, -2).reshape(B, D, gh, gw))
    cls = _ln(t, sd, "encoder.norm")[:, 0]
    return _lin(cls, sd, "encoder.head"), skips


# ----------------------------------------------------------------------------- decoder
def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _conv_block(x, sd, p):  # Conv2DBlock, utils.py:11-43
    return F.relu(_bn(F.conv2d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1), sd, p + ".block.1"))


def _deconv_block(x, sd, p):  # Deconv2DBlock, utils.py:46-86
    x = F.conv_transpose2d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], stride=2)
    return F.relu(_bn(F.conv2d(x, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], padding=1), sd, p + ".block.2"))


def _branch(sd, name, z4, s3, s2, s1, s0, head=True):
    """cellvit.py:212-244 with the shared skips s0..s3 precomputed. ``head=False``: stop at the 64-channel feature map
    (the trunk of the ``*Shared`` variants, cellvit_shared.py:199-231)."""
    b = F.conv_transpose2d(z4, sd[f"{name}.bottleneck_upsampler.weight"], sd[f"{name}.bottleneck_upsampler.bias"], stride=2)
    b = torch.cat([s3, b], dim=1)
    for i in range(3):
        b = _conv_block(b, sd, f"{name}.decoder3_upsampler.{i}")
    b = F.conv_transpose2d(b, sd[f"{name}.decoder3_upsampler.3.weight"], sd[f"{name}.decoder3_upsampler.3.bias"], stride=2)
    for lvl, s in (("decoder2_upsampler", s2), ("decoder1_upsampler", s1)):
        b = torch.cat([s, b], dim=1)
        for i in range(2):
            b = _conv_block(b, sd, f"{name}.{lvl}.{i}")
        b = F.conv_transpose2d(b, sd[f"{name}.{lvl}.2.weight"], sd[f"{name}.{lvl}.2.bias"], stride=2)
    b = torch.cat([s0, b], dim=1)
    for i in range(2):
        b = _conv_block(b, sd, f"{name}.decoder0_header.{i}")
    if not head:
        return b
    return F.conv2d(b, sd[f"{name}.decoder0_header.2.weight"], sd[f"{name}.decoder0_header.2.bias"])


@torch.no_grad()
def cellvit_forward(sd, x, arch: str, retrieve_tokens: bool = False, regression_loss: bool = False):
    """arch in {"ViT256", "SAM-B", "SAM-L", "SAM-H"}; sd: reference-keyed state_dict of fp32 tensors."""
    assert x.shape[-2] % 16 == 0 and x.shape[-1] % 16 == 0
    out = {}
    if arch == "ViT256":
        logits, z = vit256_encoder(sd, x, VIT256_CFG)
        out["tissue_types"] = logits
    else:
        feat, z = sam_encoder(sd, x, SAM_CFG[arch])
        out["tissue_types"] = _lin(feat, sd, "classifier_head")
    z1, z2, z3, z4 = z
    shared = "decoder.bottleneck_upsampler.weight" in sd   # the *Shared variants: one trunk, three 1x1 heads (cellvit_shared.py:147-197)
    sk = [f"decoder.decoder{k}_skip" for k in range(4)] if shared else [f"decoder{k}" for k in range(4)]
    s0 = _conv_block(_conv_block(x, sd, sk[0] + ".0"), sd, sk[0] + ".1")
    s1 = z1
    for i in range(3):
        s1 = _deconv_block(s1, sd, f"{sk[1]}.{i}")
    s2 = _deconv_block(_deconv_block(z2, sd, sk[2] + ".0"), sd, sk[2] + ".1")
    s3 = _deconv_block(z3, sd, sk[3] + ".0")
    if shared:
        up = _branch(sd, "decoder", z4, s3, s2, s1, s0, head=False)
        nb = F.conv2d(up, sd["nuclei_binary_map_decoder.weight"], sd["nuclei_binary_map_decoder.bias"])
        out["hv_map"] = F.conv2d(up, sd["hv_map_decoder.weight"], sd["hv_map_decoder.bias"])
        out["nuclei_type_map"] = F.conv2d(up, sd["nuclei_type_maps_decoder.weight"], sd["nuclei_type_maps_decoder.bias"])
    else:
        nb = _branch(sd, "nuclei_binary_map_decoder", z4, s3, s2, s1, s0)
        out["hv_map"] = _branch(sd, "hv_map_decoder", z4, s3, s2, s1, s0)
        out["nuclei_type_map"] = _branch(sd, "nuclei_type_maps_decoder", z4, s3, s2, s1, s0)
    if regression_loss:
        out["nuclei_binary_map"], out["regression_map"] = nb[:, :2], nb[:, 2:]
    else:
        out["nuclei_binary_map"] = nb
    if retrieve_tokens:
        out["tokens"] = z4
    return out
