# -*- coding: utf-8 -*-
"""state_dict layout of the reference CellViT models and a seeded weight generator.

``state_spec`` enumerates every key/shape/kind of the reference modules' ``state_dict`` in the
reference's own order (models/segmentation/cell_segmentation/cellvit.py:91-151,246-330,554-572;
models/encoders/VIT/SAM/image_encoder.py:65-113,160-173,223-233; models/encoders/VIT/vits_histo.py:314-356),
so the drop-in modules ``load_state_dict(strict=True)`` from reference checkpoints. There are no
checkpoints offline, so ``synth_state_dict`` draws every tensor from a generator seeded by
(seed, key) with the reference initialisers' distributions (trunc-normal 0.02 for ViT linears,
PyTorch's kaiming-uniform defaults for convs, non-trivial BatchNorm statistics).
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict

import torch

SAM_CFG = {
    "SAM-B": dict(embed_dim=768, depth=12, num_heads=12, global_idx=(2, 5, 8, 11), extract=(3, 6, 9, 12)),
    "SAM-L": dict(embed_dim=1024, depth=24, num_heads=16, global_idx=(5, 11, 17, 23), extract=(6, 12, 18, 24)),
    "SAM-H": dict(embed_dim=1280, depth=32, num_heads=16, global_idx=(7, 15, 23, 31), extract=(8, 16, 24, 32)),
}
VIT256_CFG = dict(embed_dim=384, depth=12, num_heads=6, extract=(3, 6, 9, 12))


def arch_config(arch: str) -> dict:
    if arch == "ViT256":
        return dict(VIT256_CFG, arch=arch, sam=False)
    if arch.upper() in SAM_CFG:
        return dict(SAM_CFG[arch.upper()], arch=arch.upper(), sam=True)
    raise NotImplementedError("Unknown ViT-SAM backbone structure")


def decoder_dims(embed_dim: int):
    """cellvit.py:106-113 -> (skip_dim_11, skip_dim_12, bottleneck_dim)."""
    return (256, 128, 312) if embed_dim < 512 else (512, 256, 512)


def state_spec(arch: str, num_nuclei_classes: int, num_tissue_classes: int, regression_loss: bool = False,
               embed_dim=None, depth=None, num_heads=None, shared: bool = False):
    """OrderedDict key -> (shape, kind); kind in {w_lin, w_conv, w_convT, bias, ln_w, ln_b, bn_w, bn_b,
    bn_mean, bn_var, bn_nbt, pos, cls, relpos}. ``shared``: the ``*Shared`` variants (cellvit_shared.py:113-145, 231-330) --
    ONE decoder trunk under ``decoder.*`` and a 1x1 convolution per output on its 64-channel feature map."""
    cfg = arch_config(arch)
    D = embed_dim or cfg["embed_dim"]
    depth = depth or cfg["depth"]
    heads = num_heads or cfg["num_heads"]
    hd = D // heads
    S = OrderedDict()

    def lin(p, i, o, bias=True):
        S[p + ".weight"] = ((o, i), "w_lin")
        if bias:
            S[p + ".bias"] = ((o,), "bias")

    def conv(p, i, o, k, bias=True):
        S[p + ".weight"] = ((o, i, k, k), "w_conv")
        if bias:
            S[p + ".bias"] = ((o,), "bias")

    def convT(p, i, o):
        S[p + ".weight"] = ((i, o, 2, 2), "w_convT")
        S[p + ".bias"] = ((o,), "bias")

    def ln(p, n):
        S[p + ".weight"] = ((n,), "ln_w")
        S[p + ".bias"] = ((n,), "ln_b")

    def bn(p, n):
        S[p + ".weight"] = ((n,), "bn_w")
        S[p + ".bias"] = ((n,), "bn_b")
        S[p + ".running_mean"] = ((n,), "bn_mean")
        S[p + ".running_var"] = ((n,), "bn_var")
        S[p + ".num_batches_tracked"] = ((), "bn_nbt")

    def conv_block(p, i, o):
        conv(p + ".block.0", i, o, 3)
        bn(p + ".block.1", o)

    def deconv_block(p, i, o):
        convT(p + ".block.0", i, o)
        conv(p + ".block.1", o, o, 3)
        bn(p + ".block.2", o)

    if cfg["sam"]:
        S["encoder.pos_embed"] = ((1, 64, 64, D), "pos")
        conv("encoder.patch_embed.proj", 3, D, 16)
        for i in range(depth):
            p = f"encoder.blocks.{i}"
            L = 127 if i in cfg["global_idx"] else 27
            ln(p + ".norm1", D)
            S[p + ".attn.rel_pos_h"] = ((L, hd), "relpos")
            S[p + ".attn.rel_pos_w"] = ((L, hd), "relpos")
            lin(p + ".attn.qkv", D, 3 * D)
            lin(p + ".attn.proj", D, D)
            ln(p + ".norm2", D)
            lin(p + ".mlp.lin1", D, 4 * D)
            lin(p + ".mlp.lin2", 4 * D, D)
        conv("encoder.neck.0", D, 256, 1, bias=False)
        ln("encoder.neck.1", 256)
        conv("encoder.neck.2", 256, 256, 3, bias=False)
        ln("encoder.neck.3", 256)
    else:
        S["encoder.cls_token"] = ((1, 1, D), "cls")
        S["encoder.pos_embed"] = ((1, 197, D), "pos")
        conv("encoder.patch_embed.proj", 3, D, 16)
        for i in range(depth):
            p = f"encoder.blocks.{i}"
            ln(p + ".norm1", D)
            lin(p + ".attn.qkv", D, 3 * D)
            lin(p + ".attn.proj", D, D)
            ln(p + ".norm2", D)
            lin(p + ".mlp.fc1", D, 4 * D)
            lin(p + ".mlp.fc2", 4 * D, D)
        ln("encoder.norm", D)
        if num_tissue_classes > 0:
            lin("encoder.head", D, num_tissue_classes)

    s11, s12, bd = decoder_dims(D)
    nb_out = 2 + (2 if regression_loss else 0)
    heads = (("nuclei_binary_map_decoder", nb_out), ("hv_map_decoder", 2), ("nuclei_type_maps_decoder", num_nuclei_classes))

    def skips(names):
        conv_block(names[0] + ".0", 3, 32)
        conv_block(names[0] + ".1", 32, 64)
        deconv_block(names[1] + ".0", D, s11)
        deconv_block(names[1] + ".1", s11, s12)
        deconv_block(names[1] + ".2", s12, 128)
        deconv_block(names[2] + ".0", D, s11)
        deconv_block(names[2] + ".1", s11, 256)
        deconv_block(names[3] + ".0", D, bd)

    def upsampling(name, ncls):
        convT(f"{name}.bottleneck_upsampler", D, bd)
        conv_block(f"{name}.decoder3_upsampler.0", 2 * bd, bd)
        conv_block(f"{name}.decoder3_upsampler.1", bd, bd)
        conv_block(f"{name}.decoder3_upsampler.2", bd, bd)
        convT(f"{name}.decoder3_upsampler.3", bd, 256)
        conv_block(f"{name}.decoder2_upsampler.0", 512, 256)
        conv_block(f"{name}.decoder2_upsampler.1", 256, 256)
        convT(f"{name}.decoder2_upsampler.2", 256, 128)
        conv_block(f"{name}.decoder1_upsampler.0", 256, 128)
        conv_block(f"{name}.decoder1_upsampler.1", 128, 128)
        convT(f"{name}.decoder1_upsampler.2", 128, 64)
        conv_block(f"{name}.decoder0_header.0", 128, 64)
        conv_block(f"{name}.decoder0_header.1", 64, 64)
        if ncls:
            conv(f"{name}.decoder0_header.2", 64, ncls, 1)

    if shared:
        skips([f"decoder.decoder{k}_skip" for k in range(4)])
        upsampling("decoder", 0)
        for name, ncls in heads:
            conv(name, 64, ncls, 1)
    else:
        skips([f"decoder{k}" for k in range(4)])
        for name, ncls in heads:
            upsampling(name, ncls)
    if cfg["sam"] and num_tissue_classes > 0:
        lin("classifier_head", 256, num_tissue_classes)
    return S


def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def synth_tensor(key: str, shape, kind: str, seed: int = 0) -> torch.Tensor:
    g = _gen(seed, key)
    if kind == "bn_nbt":
        return torch.tensor(100, dtype=torch.long)
    if kind in ("w_lin", "pos", "cls", "relpos"):
        t = torch.empty(shape).normal_(0, 0.02, generator=g).clamp_(-0.04, 0.04)
        return t
    if kind in ("w_conv", "w_convT"):
        fan_in = shape[1] * shape[2] * shape[3]
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if kind == "bias":
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if kind in ("ln_w", "bn_w"):
        return 1.0 + (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    if kind in ("ln_b", "bn_b"):
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if kind == "bn_mean":
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if kind == "bn_var":
        return 0.5 + torch.rand(shape, generator=g)
    raise ValueError(kind)


def synth_state_dict(arch: str, num_nuclei_classes: int = 6, num_tissue_classes: int = 19, seed: int = 0,
                     regression_loss: bool = False, shared: bool = False) -> "OrderedDict[str, torch.Tensor]":
    spec = state_spec(arch, num_nuclei_classes, num_tissue_classes, regression_loss, shared=shared)
    return OrderedDict((k, synth_tensor(k, shp, kind, seed)) for k, (shp, kind) in spec.items())
