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


def _device_contours(lab, type_map=None):
    """cvb_contours on a hand-made label map, with the table built by the oracle (same row layout)."""
    import ctypes as C
    import cv2
    from cellvit_b200 import _lib as L
    from cellvit_b200.post_proc_cellvit import MAX_PTS
    H, W = lab.shape
    rows = po.instance_table(lab, type_map, 0)
    n = len(rows)
    max_rows = max(n, 1)
    labels = torch.from_numpy(lab.astype(np.int32))[None].cuda()
    table = torch.from_numpy(np.frombuffer(rows.tobytes(), np.uint8).copy()).cuda().view(1, max_rows if n else 0, 88) if n else torch.zeros(1, 1, 88, dtype=torch.uint8).cuda()
    counts = torch.tensor([n], dtype=torch.int32).cuda()
    pts = torch.full((1, max_rows, MAX_PTS, 2), -7, dtype=torch.int16).cuda()
    npts = torch.full((1, max_rows), -9, dtype=torch.int32).cuda()
    need = C.c_size_t()
    L.check(L.lib().cvb_contours_workspace_bytes(1, H, W, C.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8).cuda()
    L.check(L.lib().cvb_contours(L.ptr(labels), L.ptr(table), L.ptr(counts), 1, H, W, max_rows, MAX_PTS, L.ptr(pts), L.ptr(npts), L.ptr(ws),
                                 C.c_size_t(ws.numel()), L.stream_ptr()), "cvb_contours")
    torch.cuda.synchronize()
    ref = {}
    for r in rows:
        crop = (lab[r["rmin"]:r["rmax"], r["cmin"]:r["cmax"]] == r["id"]).astype(np.uint8)
        c = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0][0].reshape(-1, 2).astype(np.int32)
        ref[int(r["id"])] = c + np.array([r["cmin"], r["rmin"]], np.int32)
    return rows, pts[0].cpu().numpy().astype(np.int32), npts[0].cpu().numpy(), ref


def test_device_contours_match_cv2_and_flag_what_they_cannot_do():
    from scipy import ndimage
    rng = np.random.default_rng(3)
    H = W = 160
    lab = np.zeros((H, W), np.int32)
    lab[0, 0] = 0
    nid = 1  # id 1 is the dropped "first unique value" only when there is no background; here background exists
    expect_flag = set()
    # random single-component blobs with holes, lines, single pixels on a 16-px grid
    for gy in range(0, H - 16, 16):
        for gx in range(0, W - 16, 16):
            h, w = rng.integers(1, 14, 2)
            m = rng.random((h, w)) < rng.uniform(0.4, 0.95)
            cc, nc = ndimage.label(m, structure=np.ones((3, 3)))
            if nc == 0:
                continue
            keep = cc == 1 + int(np.argmax(ndimage.sum(m, cc, range(1, nc + 1))))
            if rng.random() < 0.15 and nc > 1:      # sometimes keep two components under one id -> must be flagged
                keep = cc > 0
                if ndimage.label(keep, structure=np.ones((3, 3)))[1] > 1:
                    expect_flag.add(nid)
            lab[gy + 1:gy + 1 + h, gx + 1:gx + 1 + w][keep] = nid
            nid += 1
    # one contour longer than MAX_PTS: a comb
    comb = np.zeros((12, 150), bool); comb[0] = True; comb[:, ::2] = True
    lab2 = np.zeros((H, W), np.int32); lab2[20:32, 5:155][comb] = 1; lab2[60:63, 60:64] = 2
    for L_, flags in ((lab, expect_flag), (lab2, {1})):
        rows, pts, npts, ref = _device_contours(L_)
        assert len(rows) > 1
        for i, r in enumerate(rows):
            rid = int(r["id"])
            if rid in flags:
                assert npts[i] == -1, rid
            else:
                assert npts[i] == len(ref[rid]), (rid, npts[i], len(ref[rid]))
                assert np.array_equal(pts[i, :npts[i]], ref[rid]), rid


def test_flood_paths_thin_and_huge_blobs():
    """Blobs that do not fit the shared-memory fast path of the flood kernel: thin diagonal bands (few pixels, bounding box
    far larger than the staged region -> global-memory path of the SMALL launch), a chain of touching nuclei of ~6,000
    pixels (> LARGE heap -> global heap), next to ordinary nuclei (fast path). All must equal the oracle bit for bit."""
    size = 512
    t = synth.synthetic_nuclei(size, 120, seed=21)
    yy, xx = np.mgrid[0:size, 0:size]
    np_bin, hv, nt = t["np_bin"].copy(), t["hv"].copy(), t["nt"].copy()
    rng = np.random.default_rng(5)
    # two diagonal bands, 4 px thick, 180 px long: ~700 px each, bounding box 180 x 184
    for (y0, x0, s) in ((20, 30, 1), (300, 480, -1)):
        band = np.zeros((size, size), bool)
        for k in range(180):
            band[y0 + k, x0 + s * k: x0 + s * k + 4] = True
        np_bin[band] = 1
        nt[band] = 2
        hv[0][band] = ((xx - xx[band].mean()) / 90.0)[band]
        hv[1][band] = ((yy - yy[band].mean()) / 90.0)[band]
    # a chain of 14 touching discs of radius 12 -> one blob of ~6,000 px with 14 markers
    for k in range(14):
        cy, cx = 460, 40 + 22 * k
        disc = (yy - cy) ** 2 + (xx - cx) ** 2 <= 12 ** 2
        np_bin[disc] = 1
        nt[disc] = 3
        hv[0][disc] = ((xx - cx) / 12.0)[disc]
        hv[1][disc] = ((yy - cy) / 12.0)[disc]
    hv = (hv + rng.normal(0, 0.01, hv.shape)).astype(np.float32)
    tile = {"np_bin": np_bin.astype(np.uint8), "hv": hv, "nt": nt}
    labels, rows = _check_against_oracle([tile], 40)
    assert len(rows[0]) > 100


def test_full_size_batch_bit_exact():
    """BASELINE configuration of the post-processing: a batch of four 1024 x 1024 tiles with 700 nuclei each (the bench
    workload) -- every stage output, label map and instance table bit-exact against the oracle."""
    tiles = [synth.synthetic_nuclei(1024, 700, seed=s) for s in range(4)]
    labels, rows = _check_against_oracle(tiles, 40)
    assert all(len(r) > 600 for r in rows)
