"""Pins the CPU oracle (oracle/) against the UNMODIFIED reference imported from /root/reference and against
cv2 / scipy directly. Runs only where the reference tree exists (this container); the GPU box uses the
golden fixtures in tests/golden instead (tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from cellvit_b200 import synth, weights
from oracle import forward_oracle, postproc_oracle as po, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    captured = {}

    def spy_watershed(image, markers=None, mask=None):
        captured["dist"], captured["marker"], captured["mask"] = np.array(image), np.array(markers), np.array(mask)
        return po.watershed(image, markers=markers, mask=mask)

    cellvit, post = ref_shim.import_reference(spy_watershed)
    return cellvit, post, captured


def _pred_map(d):
    return np.concatenate([d["nt"][..., None], d["np_bin"][..., None], d["hv"].transpose(1, 2, 0)], -1).astype(np.float64)


@pytest.mark.parametrize("size,n,seed,mag,noise,crop", [(256, 40, 0, 40, 0.0, None), (256, 60, 1, 20, 0.0, None), (512, 170, 2, 40, 0.02, None),
                                                        (320, 80, 3, 40, 0.05, None), (400, 120, 4, 40, 0.02, (272, 400)),
                                                        (400, 120, 5, 20, 0.0, (399, 173))])
def test_postproc_stages_match_reference(ref, size, n, seed, mag, noise, crop):
    _, post, cap = ref
    d = synth.synthetic_nuclei(size, n, seed, noise=noise)
    if crop:   # non-square (and odd-sized) maps: nuclei cut by the new border included
        d = {k: np.ascontiguousarray(v[..., :crop[0], :crop[1]]) for k, v in d.items()}
    proc = post.DetectionCellPostProcessor(nr_types=6, magnification=mag, gt=False)
    ref_lab, ref_dict = proc.post_process_cell_segmentation(_pred_map(d))
    lab, inter = po.proc_np_hv(d["np_bin"], d["hv"], mag, want_intermediates=True)
    # P1-P6 are computed by the reference through real cv2/scipy; P7 goes through the oracle flood in both.
    assert np.array_equal(inter["blb"].astype(bool), cap["mask"].astype(bool))
    assert np.array_equal(inter["marker"], cap["marker"])
    assert np.array_equal(inter["dist"], cap["dist"]), np.abs(inter["dist"] - cap["dist"]).max()
    assert np.array_equal(lab, ref_lab)
    # P8/P9
    olab, odict = po.DetectionCellPostProcessor(6, mag).post_process_cell_segmentation(_pred_map(d))
    assert np.array_equal(olab, ref_lab)
    assert sorted(odict) == sorted(ref_dict)
    for k, r in ref_dict.items():
        o = odict[k]
        assert np.array_equal(o["bbox"], r["bbox"])
        assert np.array_equal(o["centroid"], r["centroid"]), (o["centroid"], r["centroid"])
        assert np.array_equal(o["contour"], r["contour"])
        assert o["type"] == r["type"] and o["type_prob"] == r["type_prob"]


def test_postproc_degenerate_tiles(ref):
    _, post, _ = ref
    proc = post.DetectionCellPostProcessor(nr_types=6, magnification=40)
    H = 96
    empty = np.zeros((H, H, 4))
    lab, dct = proc.post_process_cell_segmentation(empty)
    olab, odct = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(empty)
    assert np.array_equal(lab, olab) and dct == {} and odct == {}
    # all-foreground tile: exercises the np.unique(...)[1:] quirk (post_proc_cellvit.py:95)
    full = np.zeros((H, H, 4)); full[..., 1] = 1
    yy, xx = np.mgrid[0:H, 0:H]
    full[..., 2] = (xx - H / 2) / H; full[..., 3] = (yy - H / 2) / H; full[..., 0] = 2
    lab, dct = proc.post_process_cell_segmentation(full)
    olab, odct = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(full)
    assert np.array_equal(lab, olab) and sorted(dct) == sorted(odct)


# non-square and non-native tile sizes too (cellvit.py:170-175 takes any multiple of 16): the GPU shape tests (tests/test_gpu_shapes.py)
# compare the engine with the oracle at such sizes, this pins the oracle itself there -- bicubic position-embedding resize on a
# non-square grid (vits_histo.py:377-402), windows larger than the token grid and grids that need padding (image_encoder.py:263-288)
@pytest.mark.parametrize("arch,size", [("ViT256", 64), ("SAM-B", 64), ("ViT256", (48, 80)), ("ViT256", (112, 32)), ("SAM-B", 80), ("SAM-B", 240)])
def test_forward_oracle_matches_reference_modules(ref, arch, size):
    cellvit, _, _ = ref
    torch.manual_seed(0)
    if arch == "ViT256":
        m = cellvit.CellViT256(None, 6, 19)
    else:
        m = cellvit.CellViTSAM(None, 6, 19, arch)
    m.eval()
    sd = weights.synth_state_dict(arch, 6, 19, seed=3)
    m.load_state_dict(sd, strict=True)  # also pins cellvit_b200.weights.state_spec against the reference
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=5))
    with torch.no_grad():
        r = m(x, retrieve_tokens=True)
    o = forward_oracle.cellvit_forward({k: v for k, v in sd.items()}, x, arch, retrieve_tokens=True)
    for k in ("tissue_types", "nuclei_binary_map", "hv_map", "nuclei_type_map", "tokens"):
        assert r[k].shape == o[k].shape
        assert (r[k] - o[k]).abs().max().item() <= 2e-6, k


@pytest.mark.parametrize("arch,size,regression", [("ViT256", 64, False), ("SAM-B", 64, True)])
def test_forward_oracle_matches_reference_shared_modules(ref, arch, size, regression):
    """The ``*Shared`` variants (cellvit_shared.py): one decoder trunk + a 1x1 head per output. Pins ``weights.state_spec(shared=True)``
    (strict load into the reference modules, same key ORDER) and the oracle's shared path against the reference forward."""
    import importlib
    shared = importlib.import_module("models.segmentation.cell_segmentation.cellvit_shared")
    torch.manual_seed(0)
    m = shared.CellViT256Shared(None, 6, 19, regression_loss=regression) if arch == "ViT256" else \
        shared.CellViTSAMShared(None, 6, 19, arch, regression_loss=regression)
    m.eval()
    sd = weights.synth_state_dict(arch, 6, 19, seed=3, regression_loss=regression, shared=True)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=5))
    with torch.no_grad():
        r = m(x, retrieve_tokens=True)
    o = forward_oracle.cellvit_forward(dict(sd), x, arch, retrieve_tokens=True, regression_loss=regression)
    assert sorted(r) == sorted(o)
    for k in r:
        assert r[k].shape == o[k].shape
        assert (r[k] - o[k]).abs().max().item() <= 2e-6, k


def _wsi_cell_list(ref_cd, grid=2, tile=1024, ov=64, seed=17):
    """Cells of a synthetic slide (grid x grid tiles cut from one synthetic-nuclei canvas, so nuclei in the overlap bands
    appear in two tiles), as the dict list process_wsi hands to the duplicate removal -- built with the REFERENCE's own
    per-cell code (cell_detection.py:343-395) on oracle post-processing output."""
    import numpy as np
    from cellvit_b200 import synth
    from oracle import postproc_oracle as po
    side = grid * (tile - ov) + 2 * ov
    canvas = synth.synthetic_nuclei(side, int(500 * (side / 1024.0) ** 2), seed=seed)
    cells = []
    for row in range(grid):
        for col in range(grid):
            y0 = int(col * tile - (col + 0.5) * ov) + ov
            x0 = int(row * tile - (row + 0.5) * ov) + ov
            sl = (slice(y0, y0 + tile), slice(x0, x0 + tile))
            pm = np.concatenate([canvas["nt"][sl][..., None], canvas["np_bin"][sl][..., None],
                                 canvas["hv"][:, sl[0], sl[1]].transpose(1, 2, 0)], -1).astype(np.float64)
            _, inst = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
            x_global = int(row * tile - (row + 0.5) * ov)
            y_global = int(col * tile - (col + 0.5) * ov)
            for cell in inst.values():
                if cell["type"] == 0:
                    continue
                offset_global = np.array([x_global, y_global])
                d = {"bbox": (cell["bbox"] + offset_global).tolist(), "centroid": (cell["centroid"] + np.flip(offset_global)).tolist(),
                     "contour": (cell["contour"] + np.flip(offset_global)).tolist(), "type_prob": cell["type_prob"], "type": cell["type"],
                     "patch_coordinates": [row, col], "cell_status": ref_cd.get_cell_position_marging(cell["bbox"], 1024, 64),
                     "offset_global": offset_global.tolist()}
                if np.max(cell["bbox"]) == 1024 or np.min(cell["bbox"]) == 0:
                    position = ref_cd.get_cell_position(cell["bbox"], 1024)
                    d["edge_position"] = True
                    d["edge_information"] = {"position": position, "edge_patches": ref_cd.get_edge_patch(position, row, col)}
                else:
                    d["edge_position"] = False
                cells.append(d)
    return cells


@pytest.mark.parametrize("grid,seed", [(2, 17), (3, 5)])
def test_cell_post_processor_matches_reference_control_flow(grid, seed):
    """The reference's own CellPostProcessor (cell_detection.py:600-767, unmodified, shapely stubbed with the repo's polygon
    geometry) and cellvit_b200.wsi_merge.CellPostProcessor keep exactly the same cells of a synthetic overlapping slide."""
    import copy
    import logging
    from cellvit_b200 import wsi_merge as wm
    ref_cd = ref_shim.import_reference_cell_detection()
    cells = _wsi_cell_list(ref_cd, grid=grid, seed=seed)
    assert sum(c["cell_status"] != 0 for c in cells) > 50 and sum(c["edge_position"] for c in cells) > 5
    ref_keep = list(ref_cd.CellPostProcessor(copy.deepcopy(cells), logging.getLogger("ref")).post_process_cells().index.values)

    def host_overlap(contours, pairs):
        import numpy as np
        area = np.array([wm.polygon_area(c) for c in contours])
        inter = np.array([wm.polygon_intersection_area(contours[i], contours[j]) for i, j in pairs]) if len(pairs) else np.zeros(0)
        return area, inter

    mine = wm.CellPostProcessor(cells, overlap_fn=host_overlap).post_process_cells()
    assert len(ref_keep) < len(cells)            # duplicates were removed
    assert mine == [int(i) for i in ref_keep]


def test_check_wsi_matches_reference():
    """check_wsi (cell_detection.py:1008-1039): same accept / RuntimeError(message) decisions as the reference's function."""
    import types
    from cellvit_b200.cell_detection import check_wsi
    ref_cd = ref_shim.import_reference_cell_detection()
    base = {"magnification": 40, "base_magnification": 40, "downsampling": 1, "patch_size": 1024, "patch_overlap": 64}
    cases = [({}, 40.0), ({"magnification": None}, 40.0), ({"magnification": None, "downsampling": 2}, 40.0), ({"magnification": 20}, 40.0),
             ({"magnification": 20}, 20.0), ({"patch_size": 512}, 40.0), ({"patch_size": 1000}, 40.0), ({"patch_size": 2048}, 40.0),
             ({"patch_overlap": 0}, 40.0), ({"magnification": "40"}, 40.0)]

    def outcome(fn, meta, mag):
        try:
            fn(types.SimpleNamespace(metadata=dict(meta)), mag)
            return None
        except RuntimeError as e:
            return str(e)

    seen = set()
    for delta, mag in cases:
        meta = dict(base, **delta)
        got, want = outcome(check_wsi, meta, mag), outcome(ref_cd.check_wsi, meta, mag)
        assert got == want, (meta, mag, got, want)
        seen.add(want)
    assert len(seen) == 5     # accepted + the four distinct refusals


def test_generate_instance_nuclei_map_matches_reference(ref):
    """F8 (cellvit.py:385-414): the table-lookup + scatter formulation against the reference's per-instance loop, including an
    instance that is in the label map but not in the dict and a dict entry whose id is absent from the map."""
    import types
    from cellvit_b200.cellvit import CellViT
    ref_cellvit = ref[0]
    rng = np.random.default_rng(4)
    B, H, W, C = 2, 96, 80, 6
    labels = np.zeros((B, H, W), np.int64)
    dicts = []
    for b in range(B):
        ids = [3, 4, 9, 17, 40]
        for k, i in enumerate(ids):
            labels[b, 10 + 15 * k: 20 + 15 * k, 5 + 7 * b: 40 + 7 * b] = i
        d = {np.int32(i): {"type": int(rng.integers(0, C))} for i in ids if i != 9}     # 9: traced contour too short -> not listed
        d[np.int32(77)] = {"type": 2}                                                     # listed but not present
        dicts.append(d)
    lab_t = torch.tensor(labels, dtype=torch.float32)          # calculate_instance_map returns float32 label maps
    stand_in = types.SimpleNamespace(num_nuclei_classes=C)
    want = ref_cellvit.CellViT.generate_instance_nuclei_map(stand_in, lab_t, dicts)
    got = CellViT.generate_instance_nuclei_map(stand_in, lab_t, dicts)
    assert got.dtype == want.dtype and got.shape == want.shape == (B, C, H, W)
    assert torch.equal(got, want) and got.sum() > 0
