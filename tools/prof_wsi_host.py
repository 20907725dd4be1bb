"""CPU profile of process_wsi's host side (per-cell records, duplicate removal bookkeeping, export): the device stage is
replaced by cached oracle cells (tests/wsi_host_harness.py), so this runs without a GPU. ``python tools/prof_wsi_host.py [G]``
builds a G x G-tile slide (default 10) and prints ``last_timings``."""
import pathlib
import sys
import tempfile
import time
import types

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle import wsi_fixture as wf  # noqa: E402
import wsi_host_harness as hh  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    import json, yaml
    from PIL import Image
    from cellvit_b200 import wsi_merge as wm
    from cellvit_b200.cell_detection import CellSegmentationInference
    from cellvit_b200.wsi_datamodel import WSI
    with tempfile.TemporaryDirectory() as tmp:
        root = pathlib.Path(tmp) / "slide"
        (root / "patches").mkdir(parents=True); (root / "metadata").mkdir()
        yaml.safe_dump({"magnification": 40, "base_magnification": 40, "downsampling": 1, "patch_size": 1024, "patch_overlap": 64,
                        "label_map": {"background": 0}}, open(root / "metadata.yaml", "w"))
        entries = []
        tiny = np.zeros((16, 16, 3), np.uint8)
        for r in range(G):
            for c in range(G):
                name = f"s_{r}_{c}.png"
                tiny[0, 0, 0] = (r * G + c) % 4
                Image.fromarray(tiny).save(root / "patches" / name)
                yaml.safe_dump({"row": r, "col": c}, open(root / "metadata" / f"s_{r}_{c}.yaml", "w"))
                entries.append({name: {"row": r, "col": c, "metadata_path": f"metadata/s_{r}_{c}.yaml"}})
        json.dump(entries, open(root / "patch_metadata.json", "w"))
        canvas = wf.make_canvas()
        cache = {}
        inf = hh.make_host_inference(canvas, cache)
        wsi = WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
        for rep in range(2):
            t0 = time.perf_counter()
            out = inf.process_wsi(wsi, subdir_name="p", patch_size=1024, overlap=64, batch_size=4, geojson=True, num_workers=0)
            print(f"rep {rep}: {time.perf_counter() - t0:.2f} s, cells {len(out['cells'])}", {k: round(v, 3) for k, v in inf.last_timings.items()})
        if len(sys.argv) > 2:
            import cProfile, pstats
            pr = cProfile.Profile(); pr.enable()
            inf.process_wsi(wsi, subdir_name="p", patch_size=1024, overlap=64, batch_size=4, geojson=True, num_workers=0)
            pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
