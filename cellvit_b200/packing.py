# -*- coding: utf-8 -*-
"""Reference ``state_dict``  ->  packed device tensors of the C-ABI model (``cvb_model_set_param``).

One-time layout plumbing (casts, transposes, BatchNorm folding, channel padding) done with torch tensor ops when
weights are loaded or the tile size changes; nothing here runs per tile. Names on the right are the keys
``csrc/model.cu`` looks up; the table is reproduced in DESIGN.md.

Reference key families: models/segmentation/cell_segmentation/cellvit.py:91-151,246-330,554-572;
models/encoders/VIT/SAM/image_encoder.py:65-113,223-233; models/encoders/VIT/vits_histo.py:314-356.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BRANCHES = {"np": "nuclei_binary_map_decoder", "hv": "hv_map_decoder", "nt": "nuclei_type_maps_decoder"}


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def _h(t):
    return t.detach().to(torch.float16).contiguous()


def _f(t):
    return t.detach().to(torch.float32).contiguous()


def pack_conv3x3(w, splits, n_pad):
    """w [N, sum(real), 3, 3] -> fp16 [n_pad, 9 * sum(padded)], k = tap * Cpad + channel; splits = [(real, padded), ...]."""
    N = w.shape[0]
    cpad = sum(p for _, p in splits)
    out = torch.zeros(n_pad, 3, 3, cpad, dtype=torch.float32, device=w.device)
    src = dst = 0
    for real, padded in splits:
        out[:N, :, :, dst:dst + real] = w[:, src:src + real].permute(0, 2, 3, 1)
        src += real
        dst += padded
    return _h(out.reshape(n_pad, 9 * cpad))


def pack_conv_t(w, b, cin_pad, cout_pad):
    """ConvTranspose2d weight [Cin, Cout, 2, 2] -> fp16 [(dy,dx,co_pad), ci_pad]; bias -> fp32 [cout_pad]."""
    cin, cout = w.shape[:2]
    out = torch.zeros(2, 2, cout_pad, cin_pad, dtype=torch.float32, device=w.device)
    out[:, :, :cout, :cin] = w.permute(2, 3, 1, 0)
    bias = torch.zeros(cout_pad, dtype=torch.float32, device=w.device)
    bias[:cout] = b
    return _h(out.reshape(4 * cout_pad, cin_pad)), bias


def fold_bn(sd, conv_prefix, bn_prefix, n_pad, eps=1e-5):
    """Conv bias + eval-mode BatchNorm -> per-channel (scale, shift); padded channels get (1, 0)."""
    g, b = _f(sd[bn_prefix + ".weight"]), _f(sd[bn_prefix + ".bias"])
    mean, var = _f(sd[bn_prefix + ".running_mean"]), _f(sd[bn_prefix + ".running_var"])
    s = g / torch.sqrt(var + eps)
    sh = b - mean * s + _f(sd[conv_prefix + ".bias"]) * s
    scale = torch.ones(n_pad, dtype=torch.float32, device=g.device)
    shift = torch.zeros(n_pad, dtype=torch.float32, device=g.device)
    scale[: s.numel()], shift[: s.numel()] = s, sh
    return scale, shift


def rel_table(rel_pos, size):
    """image_encoder.py:321-351 for q_size == k_size == size -> fp16 [2*size-1, hd] (row = q - k + size - 1)."""
    L = 2 * size - 1
    if rel_pos.shape[0] != L:
        r = F.interpolate(rel_pos.float().reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=L, mode="linear")
        return _h(r.reshape(-1, L).permute(1, 0))
    return _h(rel_pos)


def vit_pos(pos_embed, h, w, dim):
    """vits_histo.py:377-402 (bicubic, +0.1 scale-factor trick) -> fp32 [h*w+1, dim]."""
    N = pos_embed.shape[1] - 1
    if h * w == N and w == h:
        return _f(pos_embed[0])
    w0, h0 = h + 0.1, w + 0.1
    g = int(math.sqrt(N))
    p = F.interpolate(pos_embed[:, 1:].float().reshape(1, g, g, dim).permute(0, 3, 1, 2),
                      scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
    assert int(w0) == p.shape[-2] and int(h0) == p.shape[-1]
    return torch.cat((pos_embed[0, :1].float(), p.permute(0, 2, 3, 1).reshape(-1, dim)), dim=0).contiguous()


def pack_static(sd, cfg):
    """Size-independent tensors. cfg: dict(sam, embed_dim, depth, num_heads, window, global_idx, skip11, skip12, bott[, shared])."""
    D, depth, sam = cfg["embed_dim"], cfg["depth"], cfg["sam"]
    bott, bt = cfg["bott"], pad64(cfg["bott"])
    s11, s12 = cfg["skip11"], cfg["skip12"]
    P = {}
    P["patch.w"] = _h(sd["encoder.patch_embed.proj.weight"].reshape(D, -1))
    P["patch.b"] = _f(sd["encoder.patch_embed.proj.bias"])
    names = ("attn.qkv", "attn.proj", "mlp.lin1", "mlp.lin2") if sam else ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")
    for i in range(depth):
        r, p = f"encoder.blocks.{i}", f"b{i}"
        for src, dst in (("norm1", "n1"), ("norm2", "n2")):
            P[f"{p}.{dst}.w"], P[f"{p}.{dst}.b"] = _f(sd[f"{r}.{src}.weight"]), _f(sd[f"{r}.{src}.bias"])
        for src, dst in zip(names, ("qkv", "proj", "fc1", "fc2")):
            P[f"{p}.{dst}.w"], P[f"{p}.{dst}.b"] = _h(sd[f"{r}.{src}.weight"]), _f(sd[f"{r}.{src}.bias"])
        if sam and i not in cfg["global_idx"]:
            P[f"{p}.relh"] = rel_table(sd[f"{r}.attn.rel_pos_h"], cfg["window"])
            P[f"{p}.relw"] = rel_table(sd[f"{r}.attn.rel_pos_w"], cfg["window"])
            # both tables as one K-major [64, hd] B operand for the tcgen05 window attention: rows 0.. = rel_h, rows 32.. = rel_w
            cat = torch.zeros(64, P[f"{p}.relh"].shape[1], dtype=P[f"{p}.relh"].dtype, device=P[f"{p}.relh"].device)
            cat[:P[f"{p}.relh"].shape[0]] = P[f"{p}.relh"]
            cat[32:32 + P[f"{p}.relw"].shape[0]] = P[f"{p}.relw"]
            P[f"{p}.relcat"] = cat.contiguous()
    if sam:
        P["neck.0.w"] = _h(sd["encoder.neck.0.weight"].reshape(256, D))
        P["neck.1.w"], P["neck.1.b"] = _f(sd["encoder.neck.1.weight"]), _f(sd["encoder.neck.1.bias"])
        P["neck.2.w"] = pack_conv3x3(sd["encoder.neck.2.weight"].float(), [(256, 256)], 256)
        P["neck.3.w"], P["neck.3.b"] = _f(sd["encoder.neck.3.weight"]), _f(sd["encoder.neck.3.bias"])
        P["cls.w"], P["cls.b"] = _f(sd["classifier_head.weight"]), _f(sd["classifier_head.bias"])
    else:
        P["norm.w"], P["norm.b"] = _f(sd["encoder.norm.weight"]), _f(sd["encoder.norm.bias"])
        P["head.w"], P["head.b"] = _f(sd["encoder.head.weight"]), _f(sd["encoder.head.bias"])

    def conv_block(ref, name, splits, n):
        P[name + ".w"] = pack_conv3x3(sd[ref + ".block.0.weight"].float(), splits, pad64(n))
        P[name + ".scale"], P[name + ".shift"] = fold_bn(sd, ref + ".block.0", ref + ".block.1", pad64(n))

    def deconv_block(ref, name, cin, cout):
        P[name + ".ct.w"], P[name + ".ct.b"] = pack_conv_t(sd[ref + ".block.0.weight"].float(), sd[ref + ".block.0.bias"].float(),
                                                           pad64(cin), pad64(cout))
        P[name + ".conv.w"] = pack_conv3x3(sd[ref + ".block.1.weight"].float(), [(cout, pad64(cout))], pad64(cout))
        P[name + ".conv.scale"], P[name + ".conv.shift"] = fold_bn(sd, ref + ".block.1", ref + ".block.2", pad64(cout))

    # reference module names of the four skip decoders and of the upsampling trunks: CellViT keeps the skips at the top level and
    # one trunk per output (cellvit.py:116-151); the *Shared variants keep everything under ``decoder`` with ONE trunk and a
    # 1x1 head per output (cellvit_shared.py:113-145, 231-330)
    shared = bool(cfg.get("shared"))
    sk = [f"decoder.decoder{k}_skip" for k in range(4)] if shared else [f"decoder{k}" for k in range(4)]
    P["decoder0.0.w"] = _f(sd[sk[0] + ".0.block.0.weight"].reshape(32, 27))
    P["decoder0.0.scale"], P["decoder0.0.shift"] = fold_bn(sd, sk[0] + ".0.block.0", sk[0] + ".0.block.1", 32)
    conv_block(sk[0] + ".1", "decoder0.1", [(32, 64)], 64)
    deconv_block(sk[1] + ".0", "decoder1.0", D, s11)
    deconv_block(sk[1] + ".1", "decoder1.1", s11, s12)
    deconv_block(sk[1] + ".2", "decoder1.2", s12, 128)
    deconv_block(sk[2] + ".0", "decoder2.0", D, s11)
    deconv_block(sk[2] + ".1", "decoder2.1", s11, 256)
    deconv_block(sk[3] + ".0", "decoder3.0", D, bott)
    trunks = {"dec": "decoder"} if shared else BRANCHES
    for n, ref in trunks.items():
        def ct(src, dst, cin, cout):
            P[f"{n}.{dst}.w"], P[f"{n}.{dst}.b"] = pack_conv_t(sd[f"{ref}.{src}.weight"].float(), sd[f"{ref}.{src}.bias"].float(),
                                                               pad64(cin), pad64(cout))
        ct("bottleneck_upsampler", "bottleneck", D, bott)
        conv_block(f"{ref}.decoder3_upsampler.0", f"{n}.d3.0", [(bott, bt), (bott, bt)], bott)
        conv_block(f"{ref}.decoder3_upsampler.1", f"{n}.d3.1", [(bott, bt)], bott)
        conv_block(f"{ref}.decoder3_upsampler.2", f"{n}.d3.2", [(bott, bt)], bott)
        ct("decoder3_upsampler.3", "d3.ct", bott, 256)
        conv_block(f"{ref}.decoder2_upsampler.0", f"{n}.d2.0", [(256, 256), (256, 256)], 256)
        conv_block(f"{ref}.decoder2_upsampler.1", f"{n}.d2.1", [(256, 256)], 256)
        ct("decoder2_upsampler.2", "d2.ct", 256, 128)
        conv_block(f"{ref}.decoder1_upsampler.0", f"{n}.d1.0", [(128, 128), (128, 128)], 128)
        conv_block(f"{ref}.decoder1_upsampler.1", f"{n}.d1.1", [(128, 128)], 128)
        ct("decoder1_upsampler.2", "d1.ct", 128, 64)
        conv_block(f"{ref}.decoder0_header.0", f"{n}.d0.0", [(64, 64), (64, 64)], 64)
        conv_block(f"{ref}.decoder0_header.1", f"{n}.d0.1", [(64, 64)], 64)
    for n, ref in BRANCHES.items():   # the 1x1 heads
        key = ref if shared else f"{ref}.decoder0_header.2"
        hw = sd[key + ".weight"]
        P[f"{n}.head.w"] = _f(hw.reshape(hw.shape[0], 64))
        P[f"{n}.head.b"] = _f(sd[key + ".bias"])
    return P


def pack_for_size(sd, cfg, h, w):
    """Tensors that depend on the token grid (position embeddings, global-attention rel-pos tables)."""
    D, P = cfg["embed_dim"], {}
    if cfg["sam"]:
        P["pos"] = _f(sd["encoder.pos_embed"][0, :h, :w, :].reshape(h * w, D))
        for i in cfg["global_idx"]:
            r = f"encoder.blocks.{i}.attn"
            P[f"b{i}.relh"] = rel_table(sd[r + ".rel_pos_h"], h)
            P[f"b{i}.relw"] = rel_table(sd[r + ".rel_pos_w"], w)
    else:
        pos = vit_pos(sd["encoder.pos_embed"], h, w, D)
        P["pos"] = pos
        P["clspos"] = (_f(sd["encoder.cls_token"]).reshape(D) + pos[0]).contiguous()
    return P
