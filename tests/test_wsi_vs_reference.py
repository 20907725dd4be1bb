"""process_wsi host logic against the REFERENCE's own process_wsi, run live in this container (skipped where
/root/reference is absent, i.e. on the GPU box).

Both sides read the same synthetic preprocessed slide from disk. The reference side is the unmodified
``CellSegmentationInference.process_wsi`` (cell_detection.py:244-483) driven on the CPU with a stand-in network that
returns head maps cut from one synthetic-nuclei canvas (so nuclei in the overlap bands are seen by two tiles); its
post-processing, per-cell records, duplicate removal and export are all the reference's code (shapely / skimage stubbed as
described in oracle/ref_shim.py). This repo's side is the real ``process_wsi`` body with the device stage (``_pipeline``)
replaced by a generator that yields the oracle's per-tile cells in the flat ``TileCells`` layout the device produces."""
import json
import numpy as np
import pytest
import torch

from oracle import ref_shim, wsi_fixture as wf
from wsi_host_harness import run_host_process_wsi

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _strip_ids(geo):
    for g in geo:
        g.pop("id", None)
    return geo


@pytest.mark.parametrize("magnification,downsampling", [(40, 1), (20, 2)])
def test_process_wsi_files_match_reference(tmp_path, magnification, downsampling):
    root = tmp_path / "slide"
    wf.make_slide(root, magnification, downsampling)
    canvas = wf.make_canvas()
    ref_dir = wf.run_reference(root, canvas)
    our_dir, _ = run_host_process_wsi(root, canvas)
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
    if (magnification, downsampling) == (40, 1):
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

