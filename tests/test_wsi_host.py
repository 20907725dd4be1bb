"""CPU checks of the WSI-level host logic: position codes against the reference's own
functions, polygon overlap against analytic and rasterised areas, envelope pairing, the duplicate-removal loop and the
on-disk WSI layout."""
import ast
import json
import os

import numpy as np
import pytest
import torch

from cellvit_b200 import wsi_merge as wm
from cellvit_b200.wsi_datamodel import load_cell_graph

REF = "/root/reference/cell_segmentation/inference/cell_detection.py"


def _free_port() -> str:
    """A TCP port nobody is listening on right now (the rendezvous of the world-size-2 runs must not collide with other jobs)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return str(s.getsockname()[1])


def _reference_functions():
    """The three pure position helpers of the reference, compiled from its source in place (nothing is copied)."""
    tree = ast.parse(open(REF).read())
    wanted = {"get_cell_position", "get_cell_position_marging", "get_edge_patch"}
    mod = ast.Module([n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted], [])
    ns = {"np": np, "List": list}
    exec(compile(mod, REF, "exec"), ns)
    return ns


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_position_codes_match_reference():
    ref = _reference_functions()
    rng = np.random.default_rng(0)
    specials = [0, 1, 63, 64, 65, 959, 960, 961, 1023, 1024]
    boxes = []
    for _ in range(4000):
        r0, c0 = (int(rng.choice(specials)) if rng.random() < 0.5 else int(rng.integers(0, 1000)) for _ in range(2))
        r1 = int(rng.choice(specials)) if rng.random() < 0.3 else r0 + int(rng.integers(1, 60))
        c1 = int(rng.choice(specials)) if rng.random() < 0.3 else c0 + int(rng.integers(1, 60))
        r1, c1 = max(r1, r0 + 1), max(c1, c0 + 1)
        boxes.append([[r0, c0], [min(r1, 1024), min(c1, 1024)]])
    boxes = np.array(boxes)
    status = wm.cell_status_batch(boxes, 1024, 64)
    for b, s in zip(boxes, status):
        assert ref["get_cell_position_marging"](b, 1024, 64) == s == wm.get_cell_position_marging(b, 1024, 64)
        pos = ref["get_cell_position"](b, 1024)
        assert pos == wm.get_cell_position(b, 1024)
        assert ref["get_edge_patch"](pos, 5, 7) == wm.get_edge_patch(pos, 5, 7)


def _raster_area(polys, ss=8):
    """Even-odd area of the intersection of polygons by supersampled point sampling (independent of the slab method)."""
    allp = np.concatenate(polys)
    x0, y0, x1, y1 = allp[:, 0].min(), allp[:, 1].min(), allp[:, 0].max(), allp[:, 1].max()
    xs = x0 + (np.arange(int((x1 - x0) * ss)) + 0.5) / ss
    ys = y0 + (np.arange(int((y1 - y0) * ss)) + 0.5) / ss
    X, Y = np.meshgrid(xs, ys)
    inside = np.ones(X.shape, bool)
    for p in polys:
        q = np.roll(p, -1, axis=0)
        ins = np.zeros(X.shape, bool)
        for (ax, ay), (bx, by) in zip(p, q):
            if ay == by:
                continue
            cond = (ay > Y) != (by > Y)
            xi = ax + (Y - ay) * (bx - ax) / (by - ay)
            ins ^= cond & (X < xi)
        inside &= ins
    return inside.sum() / ss ** 2


def test_polygon_intersection_area_exact_cases():
    sq = lambda x, y, s: np.array([[x, y], [x + s, y], [x + s, y + s], [x, y + s]], float)
    assert wm.polygon_area(sq(3, 4, 10)) == 100.0
    assert wm.polygon_intersection_area(sq(0, 0, 10), sq(5, 5, 10)) == pytest.approx(25.0, abs=1e-12)
    assert wm.polygon_intersection_area(sq(0, 0, 10), sq(10, 0, 10)) == 0.0           # touching edge
    assert wm.polygon_intersection_area(sq(0, 0, 10), sq(20, 20, 5)) == 0.0
    assert wm.polygon_intersection_area(sq(0, 0, 10), sq(2, 2, 3)) == pytest.approx(9.0, abs=1e-12)  # containment
    tri = np.array([[0, 0], [10, 0], [0, 10]], float)
    assert wm.polygon_intersection_area(tri, sq(0, 0, 5)) == pytest.approx(25.0, abs=1e-12)
    assert wm.polygon_intersection_area(tri, sq(0, 0, 10)[::-1]) == pytest.approx(50.0, abs=1e-12)  # orientation-free
    # non-convex: a U shape against a bar crossing both prongs, far from the origin (global WSI coordinates)
    u = np.array([[0, 0], [9, 0], [9, 9], [6, 9], [6, 3], [3, 3], [3, 9], [0, 9]], float) + 150000
    bar = np.array([[-1, 5], [10, 5], [10, 7], [-1, 7]], float) + 150000
    assert wm.polygon_intersection_area(u, bar) == pytest.approx(12.0, abs=1e-9)


def test_polygon_intersection_area_random_contours_vs_raster():
    import cv2
    rng = np.random.default_rng(1)
    for _ in range(12):
        polys = []
        for k in range(2):
            img = np.zeros((64, 64), np.uint8)
            for _ in range(3):
                cv2.ellipse(img, (int(rng.integers(20, 44)), int(rng.integers(20, 44))), (int(rng.integers(5, 14)), int(rng.integers(4, 10))),
                            float(rng.integers(0, 180)), 0, 360, 1, -1)
            c = cv2.findContours(img, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0][0].reshape(-1, 2).astype(float)
            polys.append(c + rng.integers(0, 5, 2))
        got = wm.polygon_intersection_area(polys[0], polys[1])
        want = _raster_area(polys)
        assert abs(got - want) <= 0.02 * max(want, 1.0) + 2.0, (got, want)
        assert wm.polygon_area(polys[0]) == pytest.approx(_raster_area(polys[:1]), rel=0.03, abs=2.0)


def test_envelope_pairs_matches_brute_force():
    rng = np.random.default_rng(2)
    lo = rng.uniform(0, 3000, (400, 2))
    boxes = np.concatenate([lo, lo + rng.uniform(1, 90, (400, 2))], 1)
    boxes[10] = [100, 100, 128, 128]; boxes[11] = [128, 128, 140, 140]   # touching corners count
    got = {tuple(p) for p in wm.envelope_pairs(boxes).tolist()}
    want = set()
    for i in range(400):
        for j in range(i + 1, 400):
            a, b = boxes[i], boxes[j]
            if a[0] <= b[2] and b[0] <= a[2] and a[1] <= b[3] and b[1] <= a[3]:
                want.add((i, j))
    assert got == want and (10, 11) in got


def _host_overlap(contours, pairs):
    area = np.array([wm.polygon_area(c) for c in contours])
    inter = np.array([wm.polygon_intersection_area(contours[i], contours[j]) for i, j in pairs]) if len(pairs) else np.zeros(0)
    return area, inter


def _cell(contour, status, patch, edge=None):
    c = {"contour": contour, "cell_status": status, "patch_coordinates": list(patch), "edge_position": edge is not None}
    if edge is not None:
        c["edge_information"] = {"position": None, "edge_patches": [list(e) for e in edge]}
    return c


def test_cell_post_processor_removes_duplicates():
    sq = lambda x, y, s: [[x, y], [x + s, y], [x + s, y + s], [x, y + s]]
    cells = [
        _cell(sq(500, 500, 20), 0, (0, 0)),                          # 0 mid cell: always kept
        _cell(sq(1000, 300, 20), 4, (0, 0)),                         # 1 margin cell of tile (0,0) ...
        _cell(sq(1002, 301, 24), 8, (0, 1)),                         # 2 ... the same nucleus seen by tile (0,1), larger
        _cell(sq(1010, 600, 14), 4, (0, 0), edge=[(0, 1)]),          # 3 touches the border, neighbour exists -> dropped
        _cell(sq(300, 1010, 14), 6, (0, 0), edge=[(1, 0)]),          # 4 touches the border, neighbour (1,0) absent -> kept
        _cell(sq(1000, 800, 20), 4, (0, 0)),                         # 5 margin cell barely touching 6 (< 1 % overlap)
        _cell(sq(1019.9, 800, 20), 8, (0, 1)),                       # 6
        _cell(sq(2000, 50, 30), 2, (0, 1)),                          # 7 chain a-b-c: 7 overlaps 8, 8 overlaps 9
        _cell(sq(2020, 50, 40), 2, (0, 1)),                          # 8
        _cell(sq(2050, 50, 30), 2, (0, 1)),                          # 9
    ]
    proc = wm.CellPostProcessor(cells, overlap_fn=_host_overlap)
    assert proc._clean_edge_cells() == [1, 2, 4, 5, 6, 7, 8, 9]
    keep = proc.post_process_cells()
    # 1 is replaced by its larger twin 2; 5/6 overlap by 0.1*20/400 = 0.5 % and both stay. Chain, as the reference's loop
    # runs it (:722-765): round 1 visits 7 -> its partner 8 is selected and consumed, 9 stays; round 2 visits 8 -> its
    # partner 9 is selected (the query itself is never a candidate, even when it is the larger one); round 3 is clean
    assert keep == [0, 2, 4, 5, 6, 9]


def test_wsi_layout_and_dataset(tmp_path):
    import yaml
    from PIL import Image
    from cellvit_b200.wsi_datamodel import InferenceTransform, PatchedWSIInference, WSI
    root = tmp_path / "slideA"
    (root / "patches").mkdir(parents=True)
    (root / "metadata").mkdir()
    meta = {"magnification": 40, "base_magnification": 40, "downsampling": 1, "patch_size": 32, "patch_overlap": 4,
            "label_map": {"background": 0, "tumor": 1}}
    yaml.safe_dump(meta, open(root / "metadata.yaml", "w"))
    rng = np.random.default_rng(0)
    entries, imgs = [], {}
    for r in range(2):
        for c in range(2):
            name = f"slideA_{r}_{c}.png"
            imgs[name] = rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)
            Image.fromarray(imgs[name]).save(root / "patches" / name)
            yaml.safe_dump({"row": r, "col": c, "background_ratio": 0.1}, open(root / "metadata" / f"slideA_{r}_{c}.yaml", "w"))
            entries.append({name: {"row": r, "col": c, "metadata_path": f"./metadata/slideA_{r}_{c}.yaml"}})
    json.dump(entries, open(root / "patch_metadata.json", "w"))
    wsi = WSI(name="slideA", patient="p", slide_path=root, patched_slide_path=root)
    assert wsi.get_number_patches() == 4 and wsi.metadata["label_map_inverse"] == {0: "background", 1: "tumor"}
    ds = PatchedWSIInference(wsi, InferenceTransform())
    x, m = ds[3]
    assert m["row"] == 1 and m["col"] == 1 and m["name"] == "slideA_1_1.png"
    want = (torch.from_numpy(imgs["slideA_1_1.png"]).permute(2, 0, 1).float() / 255 - 0.5) / 0.5
    assert torch.equal(x, want)
    xu = InferenceTransform(as_uint8=True)(Image.fromarray(imgs["slideA_1_1.png"]))
    assert xu.dtype == torch.uint8 and torch.equal(xu, torch.from_numpy(imgs["slideA_1_1.png"]).permute(2, 0, 1))
    assert torch.equal((xu.float().div(255.0) - 0.5) / 0.5, x)   # what the pipeline does with raw tiles on the device
    xb, mb = ds.collate_batch([ds[0], ds[1]])
    assert tuple(xb.shape) == (2, 3, 32, 32) and [q["col"] for q in mb] == [0, 1]


def test_convert_geojson_layout():
    from cellvit_b200.cell_detection import CellSegmentationInference
    cells = [{"type": 2, "contour": [[0, 0], [4, 0], [4, 4]], "centroid": [2.5, 1.5]},
             {"type": 1, "contour": [[9, 9], [12, 9], [12, 12]], "centroid": [11.0, 10.0]},
             {"type": 2, "contour": [[20, 20], [24, 20], [24, 24]], "centroid": [22.5, 21.5]}]
    seg = CellSegmentationInference.convert_geojson(cells, True)
    assert [f["properties"]["classification"]["name"] for f in seg] == ["Neoplastic", "Inflammatory"]
    assert seg[1]["geometry"]["type"] == "MultiPolygon" and len(seg[1]["geometry"]["coordinates"]) == 2
    assert seg[1]["geometry"]["coordinates"][0][0] == [[0, 0], [4, 0], [4, 4], [0, 0]]
    det = CellSegmentationInference.convert_geojson(cells, False)
    assert det[0]["geometry"] == {"type": "MultiPoint", "coordinates": [[11.0, 10.0]]}
    assert det[1]["properties"]["classification"]["color"] == [34, 221, 77]


_GLOO_WSI = r"""
import gzip, json, pathlib, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from oracle import wsi_fixture as wf
from wsi_host_harness import run_host_process_wsi
from cellvit_b200.wsi_datamodel import load_cell_graph
dist.init_process_group("gloo")
rank = dist.get_rank()
root = pathlib.Path(sys.argv[2])
out_dir, out = run_host_process_wsi(root, wf.make_canvas(), subdir="sharded")     # shard taken from torch.distributed
dist.barrier()
if rank == 0:
    golden = json.loads(gzip.open(sys.argv[1] + "/tests/golden/wsi_2x2_cells.json.gz").read())
    cells = json.load(open(out_dir / "cells.json"))
    assert cells["processed_patches"] == golden["processed_patches"] and cells["cells"] == golden["cells"]
    assert out["cells"] == golden["cells"]
    graph = load_cell_graph(out_dir / "cells.pt")
    assert graph.x.shape[0] == len(golden["cells"]) == len(graph.contours)
    assert graph.positions.tolist() == [[float(torch.tensor(v, dtype=torch.float32)) for v in c["centroid"]] for c in golden["cells"]]
else:
    assert out is None
dist.destroy_process_group()
print("ok", rank)
"""


def test_process_wsi_sharded_over_two_ranks_equals_reference_golden(tmp_path):
    """world_size 2 (gloo): each rank processes every second tile of the synthetic slide, rank 0 gathers the records, removes
    the cross-tile duplicates and writes the files -- identical to the reference's single-process output (the golden)."""
    import subprocess
    import sys
    from oracle import wsi_fixture as wf
    root = tmp_path / "slide"
    wf.make_slide(root)
    script = tmp_path / "gloo_wsi.py"
    script.write_text(_GLOO_WSI)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", _free_port(), str(script), repo, str(root)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_process_wsi_on_a_slide_without_cells(tmp_path):
    """No nucleus anywhere: every export file is still written, with empty cell lists and an empty graph (the reference
    raises in ``torch.stack([])`` here, cell_detection.py:462; an empty result is the useful behaviour)."""
    from oracle import wsi_fixture as wf
    from wsi_host_harness import run_host_process_wsi
    root = tmp_path / "slide"
    wf.make_slide(root)
    side = wf.GRID * (wf.TILE - wf.OV) + 2 * wf.OV
    canvas = {"np_bin": np.zeros((side, side), np.float32), "nt": np.zeros((side, side), np.float32),
              "hv": np.zeros((2, side, side), np.float32)}
    out_dir, out = run_host_process_wsi(root, canvas, subdir="empty")
    assert out["cells"] == [] and out["processed_patches"] == ["0_0", "0_1", "1_0", "1_1"]
    for name in ("cells.json", "cell_detection.json", "cells.geojson", "cell_detection.geojson"):
        data = json.load(open(out_dir / name))
        assert (data["cells"] if isinstance(data, dict) else data) == []
    graph = load_cell_graph(out_dir / "cells.pt")
    assert graph.x.shape[0] == 0 and graph.positions.shape[0] == 0 and list(graph.contours) == []


def test_inference_transform_is_bit_identical_to_torchvision():
    """``InferenceTransform`` == ``T.Compose([T.ToTensor(), T.Normalize(mean, std)])`` (cell_detection.py:214-227) bit for
    bit, for RGB, RGBA and grayscale PIL inputs and for non-default statistics."""
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    from cellvit_b200.wsi_datamodel import InferenceTransform
    rng = np.random.default_rng(7)
    for mean, std in [((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))]:
        want_fn = T.Compose([T.ToTensor(), T.Normalize(mean=mean, std=std)])
        got_fn = InferenceTransform(mean, std)
        img = Image.fromarray(rng.integers(0, 256, (64, 48, 3), dtype=np.uint8))
        assert torch.equal(got_fn(img), want_fn(img))
        rgba = Image.fromarray(rng.integers(0, 256, (32, 32, 4), dtype=np.uint8), mode="RGBA")
        assert torch.equal(got_fn(rgba), want_fn(rgba.convert("RGB")))       # the alpha plane is dropped, as PNG tiles are RGB
    # every uint8 value through both: the float path is ToTensor's x / 255 followed by (x - mean) / std
    ramp = Image.fromarray(np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(3, 2))
    assert torch.equal(InferenceTransform()(ramp), T.Compose([T.ToTensor(), T.Normalize((0.5,) * 3, (0.5,) * 3)])(ramp))


def _star_polygon(rng, cx, cy, n, r_lo, r_hi):
    ang = (np.arange(n) + rng.uniform(0.25, 0.75, n)) * (2 * np.pi / n)   # every angular gap < pi: star-shaped, hence simple
    rad = rng.uniform(r_lo, r_hi, n)
    return np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1)


def test_polygon_intersection_area_invariants():
    """Size-independent properties of the overlap predicate on random simple (star-shaped) polygons: symmetry, bounded by
    both areas, self-intersection = area, translation invariance, orientation independence, disjoint -> 0."""
    rng = np.random.default_rng(11)
    for trial in range(60):
        a = _star_polygon(rng, 50 + rng.uniform(-5, 5), 50 + rng.uniform(-5, 5), int(rng.integers(4, 40)), 4, 20)
        b = _star_polygon(rng, 50 + rng.uniform(-15, 15), 50 + rng.uniform(-15, 15), int(rng.integers(4, 40)), 4, 20)
        if trial % 2:   # integer vertices, as contours have
            a, b = np.round(a), np.round(b)
            if len(np.unique(a, axis=0)) < 3 or len(np.unique(b, axis=0)) < 3:
                continue
        aa, ab = wm.polygon_area(a), wm.polygon_area(b)
        i_ab, i_ba = wm.polygon_intersection_area(a, b), wm.polygon_intersection_area(b, a)
        tol = 1e-7 * max(aa, ab, 1.0)
        assert abs(i_ab - i_ba) <= tol
        assert -tol <= i_ab <= min(aa, ab) + tol
        if trial % 2 == 0:      # rounding can make a star polygon self-touching; the even-odd self-overlap is only its area when simple
            assert abs(wm.polygon_intersection_area(a, a) - aa) <= tol
        shift = np.array([1000.0, -250.0])
        assert abs(wm.polygon_intersection_area(a + shift, b + shift) - i_ab) <= 1e-6 * max(aa, ab, 1.0)
        assert abs(wm.polygon_intersection_area(a[::-1], b) - i_ab) <= tol
        assert wm.polygon_intersection_area(a, b + np.array([500.0, 0.0])) == 0.0


def test_process_wsi_reports_cell_counts_per_type(tmp_path):
    """``last_stats`` = what the reference logs at the end of process_wsi (cell_detection.py:470-478): cells before / after the
    clean-up and the count per nucleus type name; consistent with the written cells.json."""
    from collections import Counter
    from oracle import wsi_fixture as wf
    from wsi_host_harness import make_host_inference
    from cellvit_b200.wsi_datamodel import WSI
    root = tmp_path / "slide"
    wf.make_slide(root)
    inf = make_host_inference(wf.make_canvas())
    wsi = WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
    out = inf.process_wsi(wsi, subdir_name="s", patch_size=wf.TILE, overlap=wf.OV, batch_size=2, num_workers=0)
    cells = json.load(open(root / "cell_detection" / "s" / "cells.json"))["cells"]
    names = {v: k for k, v in wf.NUCLEI_TYPES.items()}
    assert inf.last_stats["cells"] == len(cells) == len(out.columns) < inf.last_stats["cells_before_cleaning"]
    assert inf.last_stats["per_type"] == dict(Counter(names[c["type"]] for c in cells))
    assert list(inf.last_stats["per_type"].values()) == sorted(inf.last_stats["per_type"].values(), reverse=True)


def test_contour_list_is_a_flat_view_of_the_contours():
    """``_ContourList`` (what the duplicate removal builds from the columnar store) behaves like the list of per-cell arrays and
    flattens to the same (offsets, points) the device call takes."""
    rng = np.random.default_rng(0)
    lens = [3, 7, 4, 12]
    parts = [rng.integers(0, 100, (k, 2)).astype(np.float64) for k in lens]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cl = wm._ContourList(np.concatenate(parts), off)
    assert len(cl) == 4 and all(np.array_equal(a, b) for a, b in zip(cl, parts))
    assert np.array_equal(cl[-1], parts[-1]) and np.array_equal(cl[np.int32(1)], parts[1])
    o1, p1 = wm.flatten_contours(cl)
    o2, p2 = wm.flatten_contours(parts)
    assert o1.dtype == o2.dtype == np.int32 and np.array_equal(o1, o2) and np.array_equal(p1, p2) and p1.flags.c_contiguous
    o3, p3 = wm.flatten_contours([])
    assert o3.tolist() == [0] and p3.shape == (0, 2)
