"""Generate tests/golden/wsi_2x2_cells.json.gz: the ``cells.json`` the REFERENCE's own process_wsi (unmodified, imported from
/root/reference, see oracle/ref_shim.py for the stubbed third-party modules) writes for the synthetic 2 x 2-tile slide of
oracle/wsi_fixture.py. Run in the build container: ``python tools/make_wsi_golden.py``."""
import gzip
import json
import pathlib
import sys
import tempfile

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from oracle import wsi_fixture as wf  # noqa: E402


def main():
    out = pathlib.Path(__file__).resolve().parents[1] / "tests" / "golden" / "wsi_2x2_cells.json.gz"
    with tempfile.TemporaryDirectory() as tmp:
        root = pathlib.Path(tmp) / "slide"
        wf.make_slide(root)
        ref_dir = wf.run_reference(root, wf.make_canvas())
        cells = json.load(open(ref_dir / "cells.json"))
    payload = {"processed_patches": cells["processed_patches"], "type_map": cells["type_map"], "cells": cells["cells"],
               "fixture": {"tile": wf.TILE, "overlap": wf.OV, "grid": wf.GRID, "canvas_seed": wf.CANVAS_SEED,
                           "canvas_nuclei": wf.CANVAS_NUCLEI}}
    with gzip.GzipFile(out, "wb", compresslevel=9, mtime=0) as f:
        f.write(json.dumps(payload, separators=(",", ":")).encode())
    print(out, len(cells["cells"]), "cells", out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
