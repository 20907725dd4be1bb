"""GPU post-processing (csrc/postproc.cu) vs the CPU oracle (oracle/postproc_oracle.c, pinned to the reference's
cv2/scipy path) -- bit-exact on every stage output, the label map and the instance table -- and vs the golden
fixtures generated from the unmodified reference."""
import glob
import os

import numpy as np
import pytest
import torch

from cellvit_b200 import synth
from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor, ROW_DTYPE
from oracle import postproc_oracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run_gpu(tiles, mag, nr_types=6, gt=False):
    proc = DetectionCellPostProcessor(nr_types=nr_types, magnification=mag, gt=gt)
    np_bin = torch.from_numpy(np.stack([t["np_bin"] for t in tiles])).cuda()
    hv = torch.from_numpy(np.stack([t["hv"] for t in tiles])).cuda()
    nt = torch.from_numpy(np.stack([t["nt"] for t in tiles]).astype(np.int32)).cuda()
    labels, rows, dbg = proc.run_maps(np_bin, hv, nt, debug=True)
    torch.cuda.synchronize()
    return labels.cpu().numpy(), rows, {k: v.cpu().numpy() for k, v in dbg.items()}


def _check_against_oracle(tiles, mag, gt=False):
    labels, rows, dbg = _run_gpu(tiles, mag, gt=gt)
    for b, t in enumerate(tiles):
        olab, inter = po.proc_np_hv(t["np_bin"], t["hv"], mag, gt=gt, want_intermediates=True)
        assert np.array_equal(dbg["blb"][b], inter["blb"]), f"blb tile {b}"
        assert np.array_equal(dbg["dist"][b], inter["dist"]), f"dist tile {b}: {np.abs(dbg['dist'][b] - inter['dist']).max()}"
        assert np.array_equal(dbg["marker"][b], inter["marker"]), f"marker tile {b}"
        assert np.array_equal(labels[b], olab), f"labels tile {b}: {(labels[b] != olab).sum()} px differ"
        orows = po.instance_table(olab, t["nt"], 6)
        assert rows[b].dtype == ROW_DTYPE and len(rows[b]) == len(orows)
        for f in ("id", "rmin", "cmin", "rmax", "cmax", "area", "type", "cx", "cy", "type_prob", "hist"):
            assert np.array_equal(rows[b][f], orows[f]), f
    return labels, rows


@pytest.mark.parametrize("size,n,seed,mag,noise", [(256, 40, 0, 40, 0.0), (256, 60, 1, 20, 0.0), (512, 170, 2, 40, 0.02),
                                                   (1024, 700, 0, 40, 0.0), (1024, 700, 4, 40, 0.05), (1024, 1500, 5, 20, 0.0)])
def test_stages_and_labels_bit_exact(size, n, seed, mag, noise):
    _check_against_oracle([synth.synthetic_nuclei(size, n, seed, noise=noise)], mag)


def test_batch_of_different_tiles():
    tiles = [synth.synthetic_nuclei(512, n, s, noise=nz) for n, s, nz in [(100, 10, 0.0), (250, 11, 0.03), (5, 12, 0.0)]]
    _check_against_oracle(tiles, 40)


def test_gt_parameters():
    _check_against_oracle([synth.synthetic_nuclei(256, 30, 21)], 40, gt=True)


def _blank(size):
    return {"np_bin": np.zeros((size, size), np.uint8), "hv": np.zeros((2, size, size), np.float32), "nt": np.zeros((size, size), np.int32)}


def test_degenerate_tiles():
    size = 96  # not a multiple of 32 pixels wide per warp segment boundary -> generic path
    empty = _blank(size)
    full = _blank(size)
    full["np_bin"][:] = 1
    yy, xx = np.mgrid[0:size, 0:size]
    full["hv"][0] = (xx - size / 2) / size
    full["hv"][1] = (yy - size / 2) / size
    full["nt"][:] = 2
    specks = _blank(size)
    specks["np_bin"][::7, ::5] = 1                      # single-pixel blobs: all removed by the size filter
    specks["np_bin"][40:60, 0:96] = 1                   # a bar touching both vertical borders
    specks["hv"][0, 40:60] = np.linspace(-1, 1, 96, dtype=np.float32)[None]
    border = _blank(size)
    border["np_bin"][0:12, 0:12] = 1; border["np_bin"][-12:, -12:] = 1; border["np_bin"][0:9, -9:] = 1
    border["hv"][:] = np.random.default_rng(0).normal(0, 0.3, (2, size, size)).astype(np.float32)
    _check_against_oracle([empty, full, specks, border], 40)


def test_odd_width_tile():
    t = synth.synthetic_nuclei(200, 30, 8)
    t = {k: np.ascontiguousarray(v[..., :150, :177]) for k, v in t.items() if k != "inst"}
    _check_against_oracle([t], 40)


def test_plateau_ties():
    # identical mirrored nuclei produce exact ties in the flood order: the (value, age, index) order must match
    size = 128
    t = _blank(size)
    yy, xx = np.mgrid[0:size, 0:size]
    for cy, cx in [(40, 40), (40, 62), (62, 40), (62, 62), (100, 90)]:
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= 12 ** 2
        t["np_bin"][m] = 1
        t["hv"][0][m] = ((xx - cx) / 12.0)[m]
        t["hv"][1][m] = ((yy - cy) / 12.0)[m]
        t["nt"][m] = 3
    _check_against_oracle([t], 40)


def test_float_entry_equals_maps_entry_and_dict_matches_oracle():
    t = synth.synthetic_nuclei(256, 45, 17)
    npl, ntl = synth.head_logits_from_maps(t["np_bin"], t["nt"], 6)
    proc = DetectionCellPostProcessor(nr_types=6, magnification=40)
    np_prob = torch.softmax(torch.from_numpy(npl)[None].cuda(), 1)
    nt_prob = torch.softmax(torch.from_numpy(ntl)[None].cuda(), 1)
    labels, dicts = proc.post_process_batch(np_prob, torch.from_numpy(t["hv"])[None].cuda(), nt_prob)
    pm = np.concatenate([t["nt"][..., None], t["np_bin"][..., None], t["hv"].transpose(1, 2, 0)], -1).astype(np.float64)
    olab, odict = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
    assert np.array_equal(labels[0].cpu().numpy(), olab)
    glab, gdict = proc.post_process_cell_segmentation(pm)  # reference-signature entry
    assert np.array_equal(glab, olab)
    for d in (dicts[0], gdict):
        assert sorted(d) == sorted(odict)
        for k, o in odict.items():
            g = d[k]
            assert np.array_equal(g["bbox"], o["bbox"]) and np.array_equal(g["centroid"], o["centroid"])
            assert np.array_equal(g["contour"], o["contour"]) and g["type"] == o["type"] and g["type_prob"] == o["type_prob"]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "postproc_*.npz"))))
def test_against_reference_golden(path):
    g = np.load(path)
    size, n, seed, mag = (int(v) for v in g["params"])
    t = synth.synthetic_nuclei(size, n, seed, noise=float(g["noise"]))
    labels, rows, dbg = _run_gpu([t], mag)
    assert np.array_equal(dbg["blb"][0], g["blb"]) and np.array_equal(dbg["marker"][0], g["marker"])
    assert np.array_equal(dbg["dist"][0], g["dist"])
    assert np.array_equal(labels[0], g["labels"])
    d = DetectionCellPostProcessor.rows_to_dict(labels[0], rows[0])
    assert np.array_equal(np.array(sorted(d), np.int32), g["ids"])
    for i, k in enumerate(g["ids"]):
        o = d[np.int32(k)]
        assert np.array_equal(o["bbox"], g["bbox"][i]) and np.array_equal(o["centroid"], g["centroid"][i])
        assert o["type"] == g["type"][i] and o["type_prob"] == g["type_prob"][i]


def test_unknown_magnification_raises():
    with pytest.raises(NotImplementedError, match="Unknown magnification"):
        DetectionCellPostProcessor(nr_types=6, magnification=10)
