"""Oracle vs the committed golden fixtures (generated from the unmodified reference by
tests/golden/make_golden.py). Runs anywhere -- no /root/reference, no GPU."""
import glob
import os

import numpy as np
import pytest
import torch

from cellvit_b200 import synth, weights
from oracle import forward_oracle, postproc_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "postproc_*.npz"))))
def test_postproc_oracle_vs_golden(path):
    g = np.load(path)
    size, n, seed, mag = (int(v) for v in g["params"])
    d = synth.synthetic_nuclei(size, n, seed, noise=float(g["noise"]))
    lab, inter = po.proc_np_hv(d["np_bin"], d["hv"], mag, want_intermediates=True)
    assert np.array_equal(inter["blb"], g["blb"])
    assert np.array_equal(inter["marker"], g["marker"])
    assert np.array_equal(inter["dist"], g["dist"])
    assert np.array_equal(lab, g["labels"])
    rows = po.instance_table(lab, d["nt"], 6)
    dct = po.rows_to_dict(lab, rows)
    assert np.array_equal(np.array(sorted(dct), np.int32), g["ids"])
    for i, k in enumerate(g["ids"]):
        o = dct[np.int32(k)]
        assert np.array_equal(o["bbox"], g["bbox"][i])
        assert np.array_equal(o["centroid"], g["centroid"][i])
        assert o["type"] == g["type"][i] and o["type_prob"] == g["type_prob"][i]
        assert len(o["contour"]) == g["contour_len"][i]


@pytest.mark.parametrize("arch", ["ViT256", "SAM-B"])
def test_forward_oracle_vs_golden(arch):
    g = np.load(os.path.join(GOLD, f"forward_{arch}.npz"))
    size, seed, wseed = (int(v) for v in g["params"])
    sd = weights.synth_state_dict(arch, 6, 19, seed=wseed)
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=seed))
    o = forward_oracle.cellvit_forward(sd, x, arch, retrieve_tokens=True)
    for k in ("tissue_types", "nuclei_binary_map", "hv_map", "nuclei_type_map", "tokens"):
        assert np.abs(o[k].numpy() - g[k]).max() <= 5e-6, k


@pytest.mark.parametrize("tag", ["ViT256_48x80", "ViT256_shared", "SAM-B_shared_80"])
def test_forward_oracle_vs_golden_shapes_and_shared_variants(tag):
    """Fixtures made by the reference's own modules (tests/golden/make_golden.py): a non-square tile (cellvit.py:170-175) and the
    ``*Shared`` variants (cellvit_shared.py), incl. the regression head split."""
    g = np.load(os.path.join(GOLD, f"forward_{tag}.npz"))
    H, W, seed, wseed, shared, regression = (int(v) for v in g["params"])
    arch = str(g["arch"])
    sd = weights.synth_state_dict(arch, 6, 19, seed=wseed, regression_loss=bool(regression), shared=bool(shared))
    x = torch.from_numpy(synth.synthetic_tiles(1, (H, W), seed=seed))
    o = forward_oracle.cellvit_forward(sd, x, arch, retrieve_tokens=True, regression_loss=bool(regression))
    keys = ["tissue_types", "nuclei_binary_map", "hv_map", "nuclei_type_map", "tokens"] + (["regression_map"] if regression else [])
    assert sorted(k for k in g.files if k not in ("params", "arch")) == sorted(keys)
    for k in keys:
        assert o[k].shape == g[k].shape and np.abs(o[k].numpy() - g[k]).max() <= 5e-6, k


def test_watershed_tie_break_is_immaterial_on_nuclei_tiles():
    """The one unpinned choice of the P7 restatement -- how exact (value, age) ties between marker seeds are ordered -- does
    not change a single label on synthetic-nuclei tiles: an independent pure-Python flood gives the C oracle's labels under
    the oracle's (value, age, index) order, under the opposite index order, and under the array-heap mechanics of the library
    the reference calls (tools/ws_tie_sensitivity.py). On artificial flat plateaus the choice does matter; that case is the
    documented residual."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import ws_tie_sensitivity as ws
    from cellvit_b200 import synth
    t = synth.synthetic_nuclei(512, 170, seed=2)
    labels, st = po.proc_np_hv(t["np_bin"], t["hv"], 40, want_intermediates=True)
    seeds = st["dist"][st["marker"] > 0]
    assert len(seeds) - len(np.unique(seeds)) > 20          # exact seed ties do occur on this tile
    for mode in ("total", "heap", "reverse"):
        assert np.array_equal(ws.flood(st["dist"], st["marker"], st["blb"], mode), labels), mode
    # flat plateau with multi-pixel seeds: the restatement's order is still reproduced exactly, the other orders differ
    rng = np.random.default_rng(0)
    dist = np.zeros((48, 48))
    marker = np.zeros((48, 48), np.int32)
    for k in range(12):
        y, x = rng.integers(0, 46, 2)
        marker[y:y + 2, x:x + 2] = k + 1
    mask = np.ones((48, 48), np.uint8)
    want = po.watershed(dist, markers=marker, mask=mask)
    assert np.array_equal(ws.flood(dist, marker, mask, "total"), want)
    assert (ws.flood(dist, marker, mask, "reverse") != want).any()


def test_rank_transformed_flood_equals_the_oracle_flood():
    """The GPU flood's fast path (csrc/postproc.cu, RankHeap) replaces a blob's fp64 `dist` values by their rank among the blob's
    pixels and orders the heap by ONE integer rank << 22 | age << 12 | cell. Restated here in Python -- per 4-connected blob of
    the mask: bounding box + 1-pixel apron, ranks = number of strictly smaller values (ties keep equal ranks), boundary seeds
    only, labels at push time, neighbours up / left / right / down -- it must give exactly the labels of the oracle's flood
    with (value, age, index) ordering, including on a tile with deliberately tied values."""
    import heapq
    from scipy import ndimage
    from cellvit_b200 import synth
    from oracle import postproc_oracle as po

    def rank_flood(dist, marker, mask):
        out = marker.copy()
        lab, n = ndimage.label(mask)   # 4-connectivity
        for b in range(1, n + 1):
            ys, xs = np.nonzero(lab == b)
            y0, x0 = ys.min(), xs.min()
            rh, rw = ys.max() - y0 + 3, xs.max() - x0 + 3
            vals = dist[ys, xs]
            assert len(vals) <= 1023 and rh * rw <= 4096          # the fast path's limits (FL_MAXN, FL_RC)
            order = np.sort(vals)
            ranks = np.searchsorted(order, vals, side="left")   # number of strictly smaller values
            cell = (ys - y0 + 1) * rw + (xs - x0 + 1)
            slab = np.full(rh * rw, -1, np.int64)
            srank = np.zeros(rh * rw, np.int64)
            slab[cell] = out[ys, xs]
            srank[cell] = ranks
            heap = []
            for c in np.nonzero(slab > 0)[0]:
                if (slab[[c - rw, c - 1, c + 1, c + rw]] == 0).any():
                    heap.append((int(srank[c]) << 22) | int(c))
            heapq.heapify(heap)
            age = 0
            while heap:
                c = heapq.heappop(heap) & 0xFFF
                for q in (c - rw, c - 1, c + 1, c + rw):
                    if slab[q] == 0:
                        age += 1
                        slab[q] = slab[c]
                        heapq.heappush(heap, (int(srank[q]) << 22) | (age << 12) | int(q))
            assert age < 1024
            out[ys, xs] = slab[cell]
        return out

    for seed, quantise in ((3, False), (4, True)):
        t = synth.synthetic_nuclei(256, 45, seed)
        inter = po.proc_np_hv(t["np_bin"], t["hv"], 40, want_intermediates=True)
        labels, st = inter if isinstance(inter, tuple) else (inter, None)
        blb, dist, marker = st["blb"], st["dist"].copy(), st["marker"]
        if quantise:
            dist = np.round(dist * 8) / 8          # many exact ties: age and index decide
        want = po.watershed(dist, marker, blb)
        small = np.zeros_like(blb)
        lab, n = ndimage.label(blb)
        for b in range(1, n + 1):                   # keep the blobs the fast path takes (the others use the fp64 layout)
            ys, xs = np.nonzero(lab == b)
            if len(ys) <= 1023 and (ys.max() - ys.min() + 3) * (xs.max() - xs.min() + 3) <= 4096:
                small[ys, xs] = 1
        assert small.sum() > 0.5 * blb.sum()
        got = rank_flood(dist, marker * small, small)
        assert np.array_equal(got[small > 0], want[small > 0])


def test_canvas_decoder_restatement_equals_direct_decoder():
    """CPU restatement of the canvas scheme of csrc/model.cu (tile sizes that are not native to the decoder tilings): embed the skip
    tokens / the stem output top-left into a larger zero canvas, re-zero the margin after EVERY decoder layer, crop the head
    outputs -- and the result equals the decoder evaluated directly on the real size (oracle/forward_oracle.py), because inside
    the real region every 3x3 convolution then sees exactly the zero padding of the real border. Without the re-zeroing it does
    not (folded-BN shift + ReLU and ConvT bias make the margin non-zero), which the test also shows."""
    import torch
    import torch.nn.functional as F
    from cellvit_b200 import synth, weights
    from oracle import forward_oracle as fo
    torch.manual_seed(0)
    arch, (H, W), (Hc, Wc) = "ViT256", (48, 80), (64, 96)
    h, w, hc, wc = H // 16, W // 16, Hc // 16, Wc // 16
    sd = weights.synth_state_dict(arch, 6, 19, seed=3)
    x = torch.from_numpy(synth.synthetic_tiles(1, (H, W), seed=5))
    ref = fo.cellvit_forward(sd, x, arch)
    _, z = fo.vit256_encoder(sd, x, fo.VIT256_CFG)

    def run(mask_margins):
        def embed(t, s):   # [B,C,s*h,s*w] -> zero canvas [B,C,s*hc,s*wc]
            return F.pad(t, (0, s * (wc - w), 0, s * (hc - h)))

        def mask(t):       # zero everything outside the real region of a layer at scale s = t.height / hc
            if not mask_margins:
                return t
            s = t.shape[-2] // hc
            m = torch.zeros_like(t)
            m[..., : s * h, : s * w] = 1
            return t * m

        conv = lambda t, p: mask(fo._conv_block(t, sd, p))
        convt = lambda t, p: mask(F.conv_transpose2d(t, sd[p + ".weight"], sd[p + ".bias"], stride=2))
        def deconv(t, p):   # Deconv2DBlock = ConvT -> Conv3x3 -> BN -> ReLU, margin re-zeroed after the ConvT as well
            u = mask(F.conv_transpose2d(t, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], stride=2))
            return mask(F.relu(fo._bn(F.conv2d(u, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"], padding=1), sd, p + ".block.2")))

        z1, z2, z3, z4 = (embed(t, 1) for t in z)
        s0 = conv(embed(fo._conv_block(x, sd, "decoder0.0"), 16), "decoder0.1")   # the stem pads at the real border itself
        s1 = z1
        for i in range(3):
            s1 = deconv(s1, f"decoder1.{i}")
        s2 = deconv(deconv(z2, "decoder2.0"), "decoder2.1")
        s3 = deconv(z3, "decoder3.0")
        out = {}
        for key, name in (("nuclei_binary_map", "nuclei_binary_map_decoder"), ("hv_map", "hv_map_decoder"),
                          ("nuclei_type_map", "nuclei_type_maps_decoder")):
            b = convt(z4, f"{name}.bottleneck_upsampler")
            b = torch.cat([s3, b], 1)
            for i in range(3):
                b = conv(b, f"{name}.decoder3_upsampler.{i}")
            b = convt(b, f"{name}.decoder3_upsampler.3")
            for lvl, s in (("decoder2_upsampler", s2), ("decoder1_upsampler", s1)):
                b = torch.cat([s, b], 1)
                for i in range(2):
                    b = conv(b, f"{name}.{lvl}.{i}")
                b = convt(b, f"{name}.{lvl}.2")
            b = torch.cat([s0, b], 1)
            for i in range(2):
                b = conv(b, f"{name}.decoder0_header.{i}")
            out[key] = F.conv2d(b, sd[f"{name}.decoder0_header.2.weight"], sd[f"{name}.decoder0_header.2.bias"])[..., :H, :W]
        return out

    with torch.no_grad():
        good, bad = run(True), run(False)
    for k in ("nuclei_binary_map", "hv_map", "nuclei_type_map"):
        assert good[k].shape == ref[k].shape
        assert (good[k] - ref[k]).abs().max().item() <= 1e-5, k
    assert max((bad[k] - ref[k]).abs().max().item() for k in bad) > 1e-3     # the margin DOES leak without the re-zeroing


def test_product_rows_to_dict_is_the_oracles_on_the_host():
    """``DetectionCellPostProcessor.rows_to_dict`` (host side of the product: instance table (+ device contours) -> the
    reference's per-tile dict, post_proc_cellvit.py:96-151) against the oracle's dict builder on the same oracle table: the cv2
    contour path, the device-contour path (contours handed in as padded point rows), instances below 3 points dropped, key and
    value types as the reference has them (np.int32 keys, int64 bbox, float64 centroid, int32 contour)."""
    from cellvit_b200.post_proc_cellvit import MAX_PTS, ROW_DTYPE, DetectionCellPostProcessor
    d = synth.synthetic_nuclei(256, 45, seed=11)
    lab = po.proc_np_hv(d["np_bin"], d["hv"], 40)
    orows = po.instance_table(lab, d["nt"], 6)
    want = po.rows_to_dict(lab, orows)
    rows = np.zeros(len(orows), ROW_DTYPE)
    for f in ROW_DTYPE.names:
        rows[f] = orows[f]

    def same(got):
        assert list(got) == list(want) and all(type(k) is np.int32 for k in got)
        for k, w in want.items():
            g = got[k]
            for f in ("bbox", "centroid", "contour"):
                assert g[f].dtype == w[f].dtype and np.array_equal(g[f], w[f]), (k, f)
            assert g["type"] == w["type"] and g["type_prob"] == w["type_prob"]

    same(DetectionCellPostProcessor.rows_to_dict(lab, rows))                                  # contours traced by cv2 on the host
    pts = np.zeros((len(rows), MAX_PTS, 2), np.int16)
    npts = np.zeros(len(rows), np.int32)                                                      # 0 points: dropped (as the < 3 rule says)
    for i, r in enumerate(rows):
        c = want.get(np.int32(r["id"]))
        if c is not None:
            npts[i] = len(c["contour"])
            pts[i, :npts[i]] = c["contour"]
    kept = []
    same(DetectionCellPostProcessor.rows_to_dict(None, rows, True, pts, npts, kept))          # contours delivered by the device
    assert [int(rows[i]["id"]) for i in kept] == [int(k) for k in want]
    flagged = npts.copy()
    flagged[::3] = -1                                                                         # device could not trace: cv2 fallback
    same(DetectionCellPostProcessor.rows_to_dict(lab, rows, True, pts, flagged))
