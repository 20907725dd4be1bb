"""The columnar cell store (cellvit_b200/wsi_records.py) and the native streaming export (csrc/wsi_export.cu, host only):
the files are byte-identical to ``json.dumps`` of the reference's per-cell dicts (cell_detection.py:352-409, 438-475,
538-597), compact and indented, and ``cells.pt`` names the reference's dataclass."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from cellvit_b200 import wsi_merge as wm
from cellvit_b200.wsi_records import CellColumns, LazyCellsJson, write_cells_json, write_geojson

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _random_columns(n, seed, special_floats=False):
    rng = np.random.default_rng(seed)
    lens = rng.integers(3, 40, n)
    off = np.concatenate([[0], np.cumsum(lens)])
    patch = rng.integers(0, 12, (n, 2))
    position = np.zeros((n, 4), np.int8)
    codes = list(wm._EDGE_PATCHES) + [(1, 0, 1, 0), (1, 1, 1, 1), (0, 1, 0, 1)]      # incl. codes without neighbour tiles -> null
    edge = rng.random(n) < 0.3
    for i in np.nonzero(edge)[0]:
        position[i] = codes[rng.integers(len(codes))]
    cent = rng.random((n, 2)) * rng.choice([1e-7, 1e-3, 1.0, 1e3, 1e6, 1e17], (n, 2))
    prob = rng.random(n).astype(np.float32).astype(np.float64)
    if special_floats:
        vals = [0.0, -0.0, 1.0, -1.0, 1e16, 1e15, 9999999999999998.0, 1e-4, 1e-5, 0.00012345, 123456.789, 5e-324, 1.7976931348623157e308,
                0.1, 1 / 3, 2.5e-5, 1e22, 1.5e300, 100.0, 1234567890123456.0, 12345678901234567.0, float("nan"), float("inf"), float("-inf"),
                0.30000000000000004, 4.35, 1e21, 123e-7]
        cent[:len(vals) // 2, 0] = vals[:len(vals) // 2]
        cent[:len(vals) - len(vals) // 2, 1] = vals[len(vals) // 2:]
    return CellColumns(bbox=rng.integers(-50, 100000, (n, 2, 2)), centroid=cent, type_prob=prob, type=rng.integers(1, 6, n), patch=patch,
                       status=rng.integers(0, 9, n), offset=rng.integers(-32, 90000, (n, 2)), edge=edge, position=position,
                       contour_pts=rng.integers(-10, 100000, (off[-1], 2)), contour_off=off,
                       tokens=rng.standard_normal((n, 8)).astype(np.float32))


HEADER = {"wsi_metadata": {"magnification": 40, "label_map": {"background": 0}, "path": "a/ä b", "mpp": 0.25},
          "processed_patches": ["0_0", "0_1"], "type_map": {"Background": 0, "Neoplastic": 1}}


@pytest.mark.parametrize("indent", [None, 2, 0])
@pytest.mark.parametrize("n,special", [(0, False), (1, False), (257, True)])
def test_cells_json_is_byte_identical_to_json_dumps(tmp_path, indent, n, special):
    cols = _random_columns(n, seed=n, special_floats=special)
    write_cells_json(cols, tmp_path / "cells.json", HEADER, detection=False, indent=indent)
    want = json.dumps({**HEADER, "cells": cols.to_dicts()}, indent=indent)
    assert open(tmp_path / "cells.json").read() == want
    write_cells_json(cols, tmp_path / "det.json", HEADER, detection=True, indent=indent)
    assert open(tmp_path / "det.json").read() == json.dumps({**HEADER, "cells": cols.detection_dicts()}, indent=indent)


def test_float_repr_matches_python_on_a_sweep(tmp_path):
    """Shortest round-trip digits and Python's fixed / exponent switch (1e-4 <= |x| < 1e16) over 20,000 doubles of every magnitude."""
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 63, 20000, dtype=np.int64)
    vals = bits.view(np.float64)
    vals = vals[np.isfinite(vals)]
    vals = np.concatenate([vals, -vals[:2000], 10.0 ** np.arange(-30, 31), np.arange(0, 3000) / 8.0, rng.random(2000).astype(np.float32).astype(np.float64)])
    n = len(vals) // 2
    cols = _random_columns(n, seed=1)
    cols.centroid[:] = vals[:2 * n].reshape(n, 2)
    write_geojson(cols, tmp_path / "p.geojson", False, {t: f"t{t}" for t in range(8)}, {t: [t, t, t] for t in range(8)})
    got = json.load(open(tmp_path / "p.geojson"))
    text = open(tmp_path / "p.geojson").read()
    for t, feat in zip(sorted(set(cols.type.tolist())), got):
        want = cols.centroid[cols.type == t].tolist()
        assert feat["geometry"]["coordinates"] == want                                 # round trip
        assert json.dumps(want) in text                                                # and the same characters


@pytest.mark.parametrize("indent", [None, 2])
def test_geojson_matches_convert_geojson(tmp_path, indent):
    from cellvit_b200.cell_detection import COLOR_DICT, TYPE_NUCLEI_DICT, CellSegmentationInference
    cols = _random_columns(300, seed=9)
    cells = cols.to_dicts()
    for polygons, name in ((True, "cells.geojson"), (False, "cell_detection.geojson")):
        write_geojson(cols, tmp_path / name, polygons, TYPE_NUCLEI_DICT, COLOR_DICT, indent=indent)
        text = open(tmp_path / name).read()
        got = json.loads(text)
        want = CellSegmentationInference.convert_geojson(cells, polygons)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            w["id"] = g["id"]                                                          # uuid4 per feature
        assert got == want
        assert text == json.dumps(want, indent=indent)
    empty = CellColumns.empty(8)
    write_geojson(empty, tmp_path / "e.geojson", True, TYPE_NUCLEI_DICT, COLOR_DICT)
    assert json.load(open(tmp_path / "e.geojson")) == []


def test_columns_take_concat_and_lazy_dict():
    a, b = _random_columns(40, seed=1), _random_columns(25, seed=2)
    both = CellColumns.concat([a, None, b], token_dim=8)
    assert len(both) == 65 and both.to_dicts() == a.to_dicts() + b.to_dicts()
    idx = [64, 3, 3, 40, 0]
    sub = both.take(idx)
    assert sub.to_dicts() == [both.to_dicts()[i] for i in idx]
    assert np.array_equal(sub.tokens, both.tokens[idx])
    lazy = LazyCellsJson(HEADER, sub)
    assert list(lazy) == ["wsi_metadata", "processed_patches", "type_map", "cells"] and len(lazy) == 4
    assert lazy["cells"] == sub.to_dicts() and lazy["type_map"] == HEADER["type_map"] and dict(lazy)["cells"] is lazy["cells"]
    assert len(CellColumns.concat([], token_dim=8)) == 0 and CellColumns.concat([None]).to_dicts() == []


def test_post_processor_on_columns_equals_post_processor_on_dicts():
    """The duplicate removal reads the columns directly; same kept cells as on the reference-style dict list."""
    rng = np.random.default_rng(5)
    parts = []
    for row in range(3):
        for col in range(3):
            n = 60
            c = _random_columns(n, seed=10 * row + col)
            # plausible geometry: small convex-ish contours near random centres in the tile's global frame, so that overlaps occur
            centres = rng.integers(0, 960 * 3, (n, 2))
            pts = []
            for i in range(n):
                k = c.contour_off[i + 1] - c.contour_off[i]
                ang = np.sort(rng.random(k) * 2 * math.pi)
                r = rng.integers(6, 30)
                pts.append(np.stack([centres[i, 0] + r * np.cos(ang), centres[i, 1] + r * np.sin(ang)], 1).astype(np.int64))
            c.contour_pts[:] = np.concatenate(pts)
            c.patch[:] = (row, col)
            # only border codes that have neighbour tiles (others make the reference raise)
            bad = [i for i in range(n) if c.edge[i] and tuple(c.position[i]) not in wm._EDGE_PATCHES]
            c.edge[bad] = 0
            parts.append(c)
    cols = CellColumns.concat(parts)

    def host_overlap(contours, pairs):
        area = np.array([wm.polygon_area(c) for c in contours])
        inter = np.array([wm.polygon_intersection_area(contours[i], contours[j]) for i, j in pairs]) if len(pairs) else np.zeros(0)
        return area, inter

    keep_cols = wm.CellPostProcessor(cols, overlap_fn=host_overlap).post_process_cells()
    keep_dicts = wm.CellPostProcessor(cols.to_dicts(), overlap_fn=host_overlap).post_process_cells()
    assert keep_cols == keep_dicts and 0 < len(keep_cols) < len(cols)


_LOAD_WITH_REFERENCE = r"""
import sys, torch
sys.path.insert(0, sys.argv[2])
g = torch.load(sys.argv[1], weights_only=False)           # plain torch.load, cellvit_b200 NOT importable here
assert "cellvit_b200" not in sys.modules
assert type(g).__module__ == "cell_segmentation.datasets.cell_graph_datamodel", type(g).__module__
assert isinstance(g.contours, list) and len(g.contours) == g.x.shape[0] == g.positions.shape[0] == 5
assert [tuple(c.shape) for c in g.contours] == [(3, 2), (4, 2), (5, 2), (3, 2), (6, 2)] and g.metadata == {"k": 1}
print("ok")
"""


def test_cells_pt_names_the_reference_dataclass(tmp_path):
    from cellvit_b200.wsi_datamodel import CellGraphDataWSI, SplitTensorList, load_cell_graph, save_cell_graph
    lens = [3, 4, 5, 3, 6]
    g = CellGraphDataWSI(x=torch.randn(5, 8), positions=torch.rand(5, 2), contours=SplitTensorList(torch.rand(sum(lens), 2), lens),
                         metadata={"k": 1})
    save_cell_graph(g, tmp_path / "cells.pt")
    back = load_cell_graph(tmp_path / "cells.pt")
    assert torch.equal(back.x, g.x) and torch.equal(back.positions, g.positions) and back.metadata == g.metadata
    assert all(torch.equal(a, b) for a, b in zip(back.contours, g.contours)) and len(back.contours) == 5
    import pickletools
    import zipfile
    with zipfile.ZipFile(tmp_path / "cells.pt") as z:
        data = z.read([n for n in z.namelist() if n.endswith("data.pkl")][0])
    globals_ = [arg for op, arg, _ in pickletools.genops(data) if op.name == "GLOBAL"]
    assert "cell_segmentation.datasets.cell_graph_datamodel CellGraphDataWSI" in globals_
    assert not any("cellvit_b200" in a for a in globals_)
    if os.path.isdir("/root/reference/cell_segmentation"):   # a reference-side consumer: plain torch.load in a fresh interpreter
        script = tmp_path / "load.py"
        script.write_text(_LOAD_WITH_REFERENCE)
        r = subprocess.run([sys.executable, str(script), str(tmp_path / "cells.pt"), "/root/reference"], capture_output=True, text=True,
                           cwd=str(tmp_path), timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_export_rejects_malformed_columns(tmp_path):
    """The native writer validates what it is about to index: cell indices, contour offsets, the output path."""
    from cellvit_b200 import _lib as L
    cols = _random_columns(10, seed=4)
    with pytest.raises(L.CvbError, match="out of range"):
        cols.export_json(tmp_path / "a.json", [("", "", np.array([0, 10]), 0, 1)])
    with pytest.raises(L.CvbError, match="cannot open"):
        cols.export_json(tmp_path / "no_such_dir" / "a.json", [("", "", None, 0, 1)])
    cols.contour_off[3] = cols.contour_off[4] + 5
    with pytest.raises(L.CvbError, match="non-decreasing"):
        cols.export_json(tmp_path / "a.json", [("", "", None, 0, 1)])
