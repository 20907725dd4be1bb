# -*- coding: utf-8 -*-
"""Seeded synthetic inputs for tests and bench.py (no datasets or checkpoints exist offline).

* ``synthetic_tiles``  -- uint8 RGB tiles, normalised the way the reference's inference
  transform does (``ToTensor`` + ``Normalize(mean=.5, std=.5)``,
  cell_segmentation/inference/cell_detection.py:214-227).
* ``synthetic_nuclei`` -- head maps (NP argmax, HV, NT argmax) of a tile with random elliptical
  nuclei. HV maps follow the recipe the reference uses for its ground truth
  (cell_segmentation/datasets/pannuke.py:335-415 ``gen_instance_hv_map``): per instance,
  centre-of-mass-relative x/y offsets, negative and positive side normalised separately to [-1, 1].
  Random-init networks emit spatially constant argmax maps, so the
  post-processing is exercised and timed on these maps instead.
"""
from __future__ import annotations

import numpy as np


def synthetic_tiles(batch: int, size=1024, seed: int = 0) -> np.ndarray:
    """[B,3,H,W] float32 in [-1,1] from seeded uniform uint8 RGB; ``size`` = edge of a square tile or ``(H, W)``."""
    rng = np.random.default_rng(seed)
    hh, ww = (size, size) if isinstance(size, int) else size
    u8 = rng.integers(0, 256, size=(batch, hh, ww, 3), dtype=np.uint8)
    x = u8.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))


def synthetic_instances(size: int = 1024, n_nuclei: int = 700, seed: int = 0,
                        axes=(6.0, 14.0)) -> np.ndarray:
    """int32 [size,size] instance map; later ellipses overwrite earlier ones (touching clusters)."""
    rng = np.random.default_rng(seed)
    inst = np.zeros((size, size), np.int32)
    for i in range(1, n_nuclei + 1):
        cy, cx = rng.uniform(0, size, 2)
        a, b = rng.uniform(axes[0], axes[1], 2)
        th = rng.uniform(0, np.pi)
        r = int(np.ceil(max(a, b))) + 1
        y0, y1 = max(0, int(cy) - r), min(size, int(cy) + r + 1)
        x0, x1 = max(0, int(cx) - r), min(size, int(cx) + r + 1)
        if y1 <= y0 or x1 <= x0:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1].astype(np.float64)
        dy, dx = yy - cy, xx - cx
        u = dx * np.cos(th) + dy * np.sin(th)
        v = -dx * np.sin(th) + dy * np.cos(th)
        inst[y0:y1, x0:x1][(u / a) ** 2 + (v / b) ** 2 <= 1.0] = i
    return inst


def hv_from_instances(inst: np.ndarray) -> np.ndarray:
    """float32 [2,H,W] (x-map, y-map) by the gen_instance_hv_map recipe."""
    H, W = inst.shape
    hv = np.zeros((2, H, W), np.float32)
    ids = np.unique(inst)
    ids = ids[ids != 0]
    # bounding boxes of every id in one pass
    ys, xs = np.nonzero(inst)
    lab = inst[ys, xs]
    order = np.argsort(lab, kind="stable")
    ys, xs, lab = ys[order], xs[order], lab[order]
    starts = np.searchsorted(lab, ids, "left")
    ends = np.searchsorted(lab, ids, "right")
    for s, e in zip(starts, ends):
        py, px = ys[s:e], xs[s:e]
        r0, r1, c0, c1 = py.min(), py.max() + 1, px.min(), px.max() + 1
        # expand the box by 2 px where possible (pannuke.py:363-370)
        if r0 >= 2: r0 -= 2
        if c0 >= 2: c0 -= 2
        if r1 <= H - 2: r1 += 2
        if c1 <= H - 2: c1 += 2
        if r1 - r0 < 2 or c1 - c0 < 2:
            continue
        ly, lx = py - r0, px - c0
        com_y, com_x = int(ly.mean() + 0.5), int(lx.mean() + 0.5)
        ox = (lx + 1 - com_x).astype(np.float32)
        oy = (ly + 1 - com_y).astype(np.float32)
        for o in (ox, oy):
            neg, pos = o < 0, o > 0
            if neg.any(): o[neg] /= -o[neg].min()
            if pos.any(): o[pos] /= o[pos].max()
        hv[0, py, px] = ox
        hv[1, py, px] = oy
    return hv


def synthetic_nuclei(size: int = 1024, n_nuclei: int = 700, seed: int = 0, n_types: int = 6,
                     noise: float = 0.0):
    """Returns dict(np_bin u8[H,W], hv f32[2,H,W], nt i32[H,W], inst i32[H,W])."""
    inst = synthetic_instances(size, n_nuclei, seed)
    hv = hv_from_instances(inst)
    rng = np.random.default_rng(seed + 7919)
    if noise > 0:
        hv = (hv + rng.normal(0, noise, hv.shape).astype(np.float32)).astype(np.float32)
    types = rng.integers(1, n_types, size=int(inst.max()) + 1).astype(np.int32)
    types[0] = 0
    nt = types[inst]
    # a sprinkle of disagreeing pixels so the majority vote has something to do
    flip = rng.random(inst.shape) < 0.05
    nt = np.where(flip & (inst > 0), rng.integers(0, n_types, inst.shape).astype(np.int32), nt)
    return {"np_bin": (inst > 0).astype(np.uint8), "hv": hv, "nt": nt.astype(np.int32), "inst": inst}


def head_logits_from_maps(np_bin: np.ndarray, nt: np.ndarray, n_types: int = 6, mag: float = 10.0):
    """NP [2,H,W] and NT [n_types,H,W] float32 logits (+-mag / one-hot*mag) whose argmax reproduces
    the given maps -- the form in which bench.py injects synthetic nuclei at the head outputs."""
    npl = np.stack([np.where(np_bin > 0, -mag, mag), np.where(np_bin > 0, mag, -mag)]).astype(np.float32)
    ntl = np.zeros((n_types,) + nt.shape, np.float32)
    for t in range(n_types):
        ntl[t][nt == t] = mag
    return npl, ntl
