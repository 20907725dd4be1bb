"""Drive the real ``CellSegmentationInference.process_wsi`` body on the CPU: the device stage (``_pipeline``) is replaced by a
generator that yields the oracle's per-tile cells in the flat ``TileCells`` layout the device produces, and the polygon
overlaps of the duplicate removal are evaluated with the host restatement of the kernel. Test harness (not a test)."""
import types

import numpy as np
import torch

from oracle import wsi_fixture as wf
from oracle.wsi_fixture import D, NUCLEI_TYPES, OV, TILE


def make_host_inference(canvas, cache=None):
    """CellSegmentationInference whose device stage yields oracle cells (``cache``: tile index -> cells, reused across calls)."""
    from cellvit_b200 import wsi_merge as wm
    from cellvit_b200.cell_detection import CellSegmentationInference
    from cellvit_b200.post_proc_cellvit import ROW_DTYPE, TileCells
    from cellvit_b200.wsi_datamodel import WSI
    from oracle import postproc_oracle as po

    mag = [40]

    def tile_cells(idx):
        if cache is not None and idx in cache:
            return cache[idx]
        out = _tile_cells(idx)
        if cache is not None:
            cache[idx] = out
        return out

    def _tile_cells(idx):
        np_bin, nt, hv = wf.tile_maps(canvas, idx)
        pm = np.concatenate([nt[..., None], np_bin[..., None], hv.transpose(1, 2, 0)], -1).astype(np.float64)
        _, inst = po.DetectionCellPostProcessor(6, mag[0]).post_process_cell_segmentation(pm)
        n = len(inst)
        rows = np.zeros(n, ROW_DTYPE)
        cap = max([len(c["contour"]) for c in inst.values()], default=1)
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
        assert with_tokens and raw
        mag[0] = magnification
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
    return inf


def run_host_process_wsi(root, canvas, subdir="ours", shard=None):
    from cellvit_b200.wsi_datamodel import WSI
    inf = make_host_inference(canvas)
    wsi = WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
    out = inf.process_wsi(wsi, subdir_name=subdir, patch_size=TILE, overlap=OV, batch_size=2, geojson=True, num_workers=0, shard=shard)
    return root / "cell_detection" / subdir, out
