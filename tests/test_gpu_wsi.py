"""GPU checks of the WSI-level rows: cell-token pooling, polygon overlap kernel and
process_wsi end to end on a synthetic preprocessed slide."""
import json

import numpy as np
import pytest
import torch

from cellvit_b200 import synth, weights
from cellvit_b200 import wsi_merge as wm
from cellvit_b200.wsi_datamodel import load_cell_graph

pytestmark = pytest.mark.gpu


def _reference_cell_token(tokens_cpu, idx, bbox, patch=16):
    """cell_detection.py:397-409, verbatim semantics (floor / ceil / uint8 cast, mean over the window)."""
    bb = bbox / patch
    bb[0, :] = np.floor(bb[0, :])
    bb[1, :] = np.ceil(bb[1, :])
    bb = bb.astype(np.uint8)
    t = tokens_cpu[idx, :, bb[0, 0]:bb[1, 0], bb[0, 1]:bb[1, 1]]
    return torch.mean(t.reshape(t.shape[0], -1).T, dim=0)


def test_cell_tokens_match_reference_expression():
    import ctypes as C
    from cellvit_b200 import _lib as L
    from cellvit_b200.post_proc_cellvit import ROW_DTYPE
    rng = np.random.default_rng(0)
    B, D, th, tw, max_rows = 2, 384, 64, 64, 512
    tokens = torch.randn(B, D, th, tw, generator=torch.Generator().manual_seed(1))
    rows = np.zeros((B, max_rows), ROW_DTYPE)
    counts = np.array([300, 17], np.int32)
    for b in range(B):
        for i in range(counts[b]):
            r0, c0 = rng.integers(0, 1000, 2)
            r1, c1 = min(1024, r0 + rng.integers(1, 70)), min(1024, c0 + rng.integers(1, 70))
            rows[b, i]["rmin"], rows[b, i]["cmin"], rows[b, i]["rmax"], rows[b, i]["cmax"] = r0, c0, r1, c1
    rows[0, 0]["rmin"], rows[0, 0]["cmin"], rows[0, 0]["rmax"], rows[0, 0]["cmax"] = 1008, 1008, 1024, 1024   # last token
    rows[0, 1]["rmin"], rows[0, 1]["cmin"], rows[0, 1]["rmax"], rows[0, 1]["cmax"] = 0, 0, 1024, 1024         # whole tile
    d_tab = torch.from_numpy(rows.view(np.uint8).reshape(B, max_rows, 88)).cuda()
    d_cnt = torch.from_numpy(counts).cuda()
    out = torch.full((B, max_rows, D), float("nan"), device="cuda")
    L.check(L.lib().cvb_cell_tokens(L.ptr(tokens.cuda()), L.ptr(d_tab), L.ptr(d_cnt), B, D, th, tw, 16, max_rows, L.ptr(out), L.stream_ptr()),
            "cvb_cell_tokens")
    out = out.cpu()
    for b in range(B):
        for i in range(counts[b]):
            r = rows[b, i]
            want = _reference_cell_token(tokens, b, np.array([[r["rmin"], r["cmin"]], [r["rmax"], r["cmax"]]], dtype=np.float64))
            assert torch.allclose(out[b, i], want, rtol=1e-5, atol=1e-6), (b, i)
        assert torch.isnan(out[b, counts[b]:]).all()   # rows beyond the count are untouched


def test_polygon_overlap_kernel_matches_host_algorithm():
    import cv2
    rng = np.random.default_rng(3)
    contours = []
    for k in range(300):
        img = np.zeros((64, 64), np.uint8)
        for _ in range(int(rng.integers(1, 4))):
            cv2.ellipse(img, (int(rng.integers(20, 44)), int(rng.integers(20, 44))), (int(rng.integers(4, 14)), int(rng.integers(3, 10))),
                        float(rng.integers(0, 180)), 0, 360, 1, -1)
        c = cv2.findContours(img, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0][0].reshape(-1, 2).astype(np.float64)
        contours.append(c + np.array([100000 + (k % 20) * 13, 50000 + (k // 20) * 13]))   # a grid of overlapping cells
    contours.append(np.array([[0, 0], [5, 0]], np.float64))                                # degenerate: 2 points
    big = np.stack([np.cos(np.linspace(0, 2 * np.pi, 200, endpoint=False)), np.sin(np.linspace(0, 2 * np.pi, 200, endpoint=False))], 1)
    contours.append(big * 40 + np.array([100100, 50100]))                                  # > 128 points: host fallback
    boxes = np.array([[c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()] for c in contours])
    pairs = wm.envelope_pairs(boxes)
    assert len(pairs) > 500
    area, inter = wm.overlap_areas(contours, pairs, torch.device("cuda", 0))
    for k in range(len(contours)):
        assert area[k] == pytest.approx(wm.polygon_area(contours[k]), rel=1e-12, abs=1e-9)
    idx = rng.choice(len(pairs), min(400, len(pairs)), replace=False)
    for k in idx:
        want = wm.polygon_intersection_area(contours[pairs[k, 0]], contours[pairs[k, 1]])
        assert inter[k] == pytest.approx(want, rel=1e-9, abs=1e-7), (k, inter[k], want)
    assert (inter >= 0).all()


def _make_slide(root, grid, tile, overlap, nuclei_fn):
    """Synthetic preprocessed slide: grid x grid tiles of one big synthetic-nuclei canvas, cut with `overlap`."""
    import yaml
    from PIL import Image
    (root / "patches").mkdir(parents=True)
    (root / "metadata").mkdir()
    yaml.safe_dump({"magnification": 40, "base_magnification": 40, "downsampling": 1, "patch_size": tile, "patch_overlap": overlap,
                    "label_map": {"background": 0}}, open(root / "metadata.yaml", "w"))
    entries = []
    rng = np.random.default_rng(5)
    for r in range(grid):
        for c in range(grid):
            name = f"s_{r}_{c}.png"
            Image.fromarray(rng.integers(0, 256, (tile, tile, 3), dtype=np.uint8)).save(root / "patches" / name)
            yaml.safe_dump({"row": r, "col": c}, open(root / "metadata" / f"s_{r}_{c}.yaml", "w"))
            entries.append({name: {"row": r, "col": c, "metadata_path": f"metadata/s_{r}_{c}.yaml"}})
    json.dump(entries, open(root / "patch_metadata.json", "w"))


def test_process_wsi_end_to_end(tmp_path):
    """2 x 2 tiles of 1024 px with 64 px overlap. Head maps are cut from ONE synthetic-nuclei canvas, so nuclei in the
    overlap bands are detected by two tiles and the clean-up must reduce them to one."""
    from cellvit_b200.cell_detection import CellSegmentationInference
    from cellvit_b200.wsi_datamodel import WSI
    tile, ov, grid = 1024, 64, 2
    root = tmp_path / "slide"
    _make_slide(root, grid, tile, ov, None)
    canvas = synth.synthetic_nuclei(2048, 1400, seed=11)
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)

    def origin(row, col):   # top-left canvas pixel (y, x) of tile (row, col); note x follows the ROW in the reference
        return int(col * tile - (col + 0.5) * ov) + ov, int(row * tile - (row + 0.5) * ov) + ov

    def override(metadata):
        maps = []
        for m in metadata:
            y0, x0 = origin(m["row"], m["col"])
            sl = (slice(y0, y0 + tile), slice(x0, x0 + tile))
            lg = synth.head_logits_from_maps(canvas["np_bin"][sl], canvas["nt"][sl], 6)
            maps.append((lg[0], lg[1], canvas["hv"][:, sl[0], sl[1]]))
        return {"nuclei_binary_map": torch.from_numpy(np.stack([m[0] for m in maps])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([m[1] for m in maps])).cuda(),
                "hv_map": torch.from_numpy(np.ascontiguousarray(np.stack([m[2] for m in maps]))).cuda()}

    wsi = WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
    out = inf.process_wsi(wsi, subdir_name="run1", patch_size=tile, overlap=ov, batch_size=2, geojson=True, num_workers=0,
                          head_override=override)
    outdir = root / "cell_detection" / "run1"
    for f in ("cells.json", "cell_detection.json", "cells.geojson", "cell_detection.geojson", "cells.pt"):
        assert (outdir / f).exists(), f
    cells = json.load(open(outdir / "cells.json"))
    assert cells["processed_patches"] == ["0_0", "0_1", "1_0", "1_1"] and len(cells["cells"]) == len(out["cells"]) > 1000
    graph = load_cell_graph(outdir / "cells.pt")
    n = len(cells["cells"])
    assert tuple(graph.x.shape) == (n, 384) and tuple(graph.positions.shape) == (n, 2) and len(graph.contours) == n
    assert torch.isfinite(graph.x).all()
    # every kept record is self-consistent with the reference's formulas
    for c in cells["cells"][:200]:
        row, col = c["patch_coordinates"]
        off = np.array([int(row * tile - (row + 0.5) * ov), int(col * tile - (col + 0.5) * ov)])
        assert c["offset_global"] == off.tolist()
        local = np.array(c["bbox"]) - off
        assert c["cell_status"] == wm.get_cell_position_marging(local, 1024, 64)
        assert c["edge_position"] == bool(local.max() == 1024 or local.min() == 0)
    # duplicates: nuclei well inside the overlap band were seen twice before the clean-up and once after it
    cent = np.array([c["centroid"] for c in cells["cells"]])
    d = np.abs(cent[:, None, :] - cent[None, :, :]).max(-1) + np.eye(n) * 1e9
    assert (d.min(1) < 1.5).sum() == 0, "near-identical centroids survived the overlap clean-up"
    status = np.array([c["cell_status"] for c in cells["cells"]])
    assert (status == 0).sum() > 0.7 * n and (status != 0).sum() > 20


def test_process_wsi_matches_reference_golden(tmp_path):
    """The whole GPU path of process_wsi (device post-processing, contour tracing, duplicate removal with the polygon kernel,
    export) against ``tests/golden/wsi_2x2_cells.json.gz`` -- the cells.json the REFERENCE's own process_wsi wrote for the same
    synthetic slide and head maps (tools/make_wsi_golden.py, oracle/wsi_fixture.py): same cells, same order, same values."""
    import gzip
    import pathlib
    from cellvit_b200.cell_detection import CellSegmentationInference
    from cellvit_b200.wsi_datamodel import WSI
    from oracle import wsi_fixture as wf
    golden = json.loads(gzip.open(pathlib.Path(__file__).parent / "golden" / "wsi_2x2_cells.json.gz").read())
    root = tmp_path / "slide"
    wf.make_slide(root)
    canvas = wf.make_canvas()
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)

    def override(metadata):
        maps = [wf.tile_maps(canvas, m["row"] * wf.GRID + m["col"]) for m in metadata]
        lg = [synth.head_logits_from_maps(m[0], m[1], 6) for m in maps]
        return {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])).cuda(),
                "hv_map": torch.from_numpy(np.ascontiguousarray(np.stack([m[2] for m in maps]))).cuda()}

    wsi = WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
    inf.process_wsi(wsi, subdir_name="g", patch_size=wf.TILE, overlap=wf.OV, batch_size=2, geojson=False, num_workers=0,
                    head_override=override)
    cells = json.load(open(root / "cell_detection" / "g" / "cells.json"))
    assert cells["processed_patches"] == golden["processed_patches"]
    assert len(cells["cells"]) == len(golden["cells"])
    for k, (a, b) in enumerate(zip(cells["cells"], golden["cells"])):
        assert a == b, (k, a, b)


def test_tilecells_equal_instance_dicts():
    """The flat-array view of a tile's cells (TileCells, used by process_wsi) holds exactly what the per-tile instance
    dicts hold: same instances in the same order, bbox / centroid / type / contour identical."""
    from cellvit_b200.cell_detection import CellSegmentationInference
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)
    B, size = 2, 256
    nuc = [synth.synthetic_nuclei(size, 45 + 5 * i, seed=90 + i) for i in range(B)]
    lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
    override = {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])).cuda(),
                "hv_map": torch.from_numpy(np.stack([n["hv"] for n in nuc])).cuda()}
    tiles = torch.from_numpy(synth.synthetic_tiles(B, size, seed=9)).pin_memory()
    (_, dicts, toks_d), = list(inf._pipeline([(tiles, None)], 40, head_override=override, with_tokens=True))
    (_, raws, toks_r), = list(inf._pipeline([(tiles, None)], 40, head_override=override, with_tokens=True, raw=True))
    for b in range(B):
        tc, d = raws[b], dicts[b]
        assert len(tc) == len(d) > 20
        rows = tc.rows[tc.valid]
        offs = np.concatenate([[0], np.cumsum(tc.lens)])
        for n, (inst_id, cell) in enumerate(d.items()):
            assert rows["id"][n] == inst_id and rows["type"][n] == cell["type"] and rows["type_prob"][n] == cell["type_prob"]
            assert np.array_equal(cell["bbox"], [[rows["rmin"][n], rows["cmin"][n]], [rows["rmax"][n], rows["cmax"][n]]])
            assert np.array_equal(cell["centroid"], [rows["cx"][n], rows["cy"][n]])
            assert np.array_equal(cell["contour"], tc.points[offs[n]:offs[n + 1]])
        assert np.array_equal(toks_r[b][tc.valid], toks_d[b])


def test_uint8_tiles_are_normalised_on_the_device_bit_identically():
    """Raw uint8 tiles (InferenceTransform(as_uint8=True)) through the pipeline give the same network input as the host
    ToTensor + Normalize: the forward outputs are bit-identical."""
    from cellvit_b200.cell_detection import CellSegmentationInference
    from cellvit_b200.wsi_datamodel import InferenceTransform
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)
    rng = np.random.default_rng(2)
    imgs = [rng.integers(0, 256, (256, 256, 3), dtype=np.uint8) for _ in range(2)]
    xf = torch.stack([InferenceTransform()(im) for im in imgs]).pin_memory()
    xu = torch.stack([InferenceTransform(as_uint8=True)(im) for im in imgs]).pin_memory()
    assert xu.dtype == torch.uint8 and tuple(xu.shape) == (2, 3, 256, 256)
    seen = {}

    def grab(tag):
        def hook(payload):
            seen[tag] = inf.model.graph_slot((2, 3, 256, 256), True, 0, torch.device("cuda", 0), argmax_maps=True)[1].clone()
            return None
        return hook

    list(inf._pipeline([(xf, None)], 40, head_override=grab("f")))
    list(inf._pipeline([(xu, None)], 40, head_override=grab("u")))
    assert torch.equal(seen["f"], seen["u"])
    assert torch.equal(seen["f"].cpu(), xf)
