"""process_wsi host logic against the REFERENCE's own process_wsi, run live in this container (skipped where
/root/reference is absent, i.e. on the GPU box).

Both sides read the same synthetic preprocessed slide from disk. The reference side is the unmodified
``CellSegmentationInference.process_wsi`` (cell_detection.py:244-483) driven on the CPU with a stand-in network that
returns head maps cut from one synthetic-nuclei canvas (so nuclei in the overlap bands are seen by two tiles); its
post-processing, per-cell records, duplicate removal and export are all the reference's code (shapely / skimage stubbed as
described in oracle/ref_shim.py). This repo's side is the real ``process_wsi`` body with the device stage (``_pipeline``)
replaced by a generator that yields the oracle's per-tile cells in the flat ``TileCells`` layout the device produces."""
import json
import types

import numpy as np
import pytest
import torch

from oracle import ref_shim, wsi_fixture as wf
from oracle.wsi_fixture import D, GRID, NUCLEI_TYPES, OV, TILE

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _run_ours(root, canvas):
    from cellvit_b200 import wsi_merge as wm
    from cellvit_b200.cell_detection import CellSegmentationInference
    from cellvit_b200.post_proc_cellvit import ROW_DTYPE, TileCells
    from cellvit_b200.wsi_datamodel import WSI
    from oracle import postproc_oracle as po

    def tile_cells(idx):
        np_bin, nt, hv = wf.tile_maps(canvas, idx)
        pm = np.concatenate([nt[..., None], np_bin[..., None], hv.transpose(1, 2, 0)], -1).astype(np.float64)
        _, inst = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
        n = len(inst)
        rows = np.zeros(n, ROW_DTYPE)
        cap = max(len(c["contour"]) for c in inst.values())
        pts, npts = np.zeros((n, cap, 2), np.int16), np.zeros(n, np.int32)
        tokens = wf.tile_tokens(idx)
        pooled = np.zeros((n, D), np.float32)
        for k, (iid, c) in enumerate(inst.items()):
            rows[k]["id"], rows[k]["type"], rows[k]["type_prob"] = iid, c["type"], c["type_prob"]
            (rows[k]["rmin"], rows[k]["cmin"]), (rows[k]["rmax"], rows[k]["cmax"]) = c["bbox"]
            rows[k]["cx"], rows[k]["cy"] = c["centroid"]
            npts[k] = len(c["contour"])
            pts[k, :npts[k]] = c["contour"]
            r0, c0, r1, c1 = rows[k]["rmin"] // 16, rows[k]["cmin"] // 16, -(-rows[k]["rmax"] // 16), -(-rows[k]["cmax"] // 16)
            pooled[k] = tokens[:, r0:r1, c0:c1].reshape(D, -1).T.mean(0).numpy()    # what cvb_cell_tokens computes (test_gpu_wsi)
        return TileCells(None, rows, pts, npts), pooled

    def pipeline(loader, magnification, head_override, with_tokens=False, raw=False, **kw):
        assert magnification == 40 and with_tokens and raw
        for patches, metadata in loader:
            assert patches.dtype == torch.uint8          # raw tiles travel; normalisation happens on the device
            out = [tile_cells(int(p[0, 0, 0])) for p in patches]
            yield metadata, [o[0] for o in out], [o[1] for o in out]

    def host_overlap(contours, pairs):
        area = np.array([wm.polygon_area(c) for c in contours])
        inter = np.array([wm.polygon_intersection_area(contours[i], contours[j]) for i, j in pairs]) if len(pairs) else np.zeros(0)
        return area, inter

    inf = object.__new__(CellSegmentationInference)
    inf.mean, inf.std, inf.device = (0.5, 0.5, 0.5), (0.5, 0.5, 0.5), "cpu"
    inf.run_conf = {"dataset_config": {"nuclei_types": dict(NUCLEI_TYPES)}}
    inf.model = types.SimpleNamespace(embed_dim=D)
    inf._pipeline = pipeline
    inf.post_process_edge_cells = lambda cell_list: wm.CellPostProcessor(cell_list, None, None, overlap_fn=host_overlap).post_process_cells()
    wsi = WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
    inf.process_wsi(wsi, subdir_name="ours", patch_size=TILE, overlap=OV, batch_size=2, geojson=True, num_workers=0)
    return root / "cell_detection" / "ours"


def _strip_ids(geo):
    for g in geo:
        g.pop("id", None)
    return geo


def test_process_wsi_files_match_reference(tmp_path):
    root = tmp_path / "slide"
    wf.make_slide(root)
    canvas = wf.make_canvas()
    ref_dir = wf.run_reference(root, canvas)
    our_dir = _run_ours(root, canvas)
    for name in ("cells.json", "cell_detection.json"):
        ref, ours = json.load(open(ref_dir / name)), json.load(open(our_dir / name))
        assert sorted(ref) == sorted(ours) == ["cells", "processed_patches", "type_map", "wsi_metadata"]
        assert ref["processed_patches"] == ours["processed_patches"] and ref["type_map"] == ours["type_map"]
        assert ref["wsi_metadata"] == ours["wsi_metadata"]
        assert len(ref["cells"]) == len(ours["cells"]) > 500
        for a, b in zip(ref["cells"], ours["cells"]):
            assert a == b
    cells = json.load(open(ref_dir / "cells.json"))["cells"]
    import gzip
    import pathlib
    golden = json.loads(gzip.open(pathlib.Path(__file__).parent / "golden" / "wsi_2x2_cells.json.gz").read())
    assert golden["cells"] == cells, "tests/golden/wsi_2x2_cells.json.gz is stale: rerun tools/make_wsi_golden.py"
    assert sum(c["cell_status"] != 0 for c in cells) > 30 and sum(c["edge_position"] for c in cells) >= 1
    for name in ("cells.geojson", "cell_detection.geojson"):
        assert _strip_ids(json.load(open(ref_dir / name))) == _strip_ids(json.load(open(our_dir / name)))
    ref_cd = ref_shim.import_reference_cell_detection()   # torch.load needs the reference's graph dataclass importable
    g_ref = torch.load(ref_dir / "cells.pt", weights_only=False)
    g_our = torch.load(our_dir / "cells.pt", weights_only=False)
    assert torch.equal(g_ref.positions, g_our.positions) and len(g_ref.contours) == len(g_our.contours)
    assert all(torch.equal(a, b) for a, b in zip(g_ref.contours, g_our.contours))
    assert g_ref.x.shape == g_our.x.shape and torch.allclose(g_ref.x, g_our.x, rtol=1e-5, atol=1e-6)
    assert g_ref.metadata == g_our.metadata
