# -*- coding: utf-8 -*-
"""Synthetic preprocessed slide + the reference's own ``process_wsi`` run on it -- TEST INFRASTRUCTURE (fixtures for
tests/test_wsi_vs_reference.py, tests/test_gpu_wsi.py and tools/make_wsi_golden.py; nothing in cellvit_b200/ imports it).

``make_slide`` / ``tile_maps`` / ``tile_tokens`` are self-contained (they also run on the GPU box); ``run_reference`` imports the
unmodified reference from /root/reference (oracle/ref_shim.py) and only works in the build container."""
import json
import logging
import types

import numpy as np
import torch

from cellvit_b200 import synth
from oracle import ref_shim

TILE, OV, GRID, D = 1024, 64, 2, 96
CANVAS_SEED, CANVAS_NUCLEI = 23, 900
NUCLEI_TYPES = {"Background": 0, "Neoplastic": 1, "Inflammatory": 2, "Connective": 3, "Dead": 4, "Epithelial": 5}


def make_canvas():
    side = GRID * (TILE - OV) + 2 * OV
    return synth.synthetic_nuclei(side, CANVAS_NUCLEI, seed=CANVAS_SEED)

def make_slide(root, magnification=40, downsampling=1):
    """grid x grid PNG tiles whose red channel at pixel (0, 0) holds the tile index (the stand-in networks key on it)."""
    import yaml
    from PIL import Image
    (root / "patches").mkdir(parents=True)
    (root / "metadata").mkdir()
    yaml.safe_dump({"magnification": magnification, "base_magnification": 40, "downsampling": downsampling, "patch_size": TILE, "patch_overlap": OV,
                    "label_map": {"background": 0}}, open(root / "metadata.yaml", "w"))
    entries, rng = [], np.random.default_rng(5)
    for r in range(GRID):
        for c in range(GRID):
            name = f"s_{r}_{c}.png"
            img = rng.integers(0, 256, (TILE, TILE, 3), dtype=np.uint8)
            img[0, 0, 0] = r * GRID + c
            Image.fromarray(img).save(root / "patches" / name)
            yaml.safe_dump({"row": r, "col": c}, open(root / "metadata" / f"s_{r}_{c}.yaml", "w"))
            entries.append({name: {"row": r, "col": c, "metadata_path": f"metadata/s_{r}_{c}.yaml"}})
    json.dump(entries, open(root / "patch_metadata.json", "w"))


def tile_maps(canvas, idx):
    row, col = divmod(idx, GRID)
    y0 = int(col * TILE - (col + 0.5) * OV) + OV     # x follows the tile ROW in the reference (cell_detection.py:343-350)
    x0 = int(row * TILE - (row + 0.5) * OV) + OV
    sl = (slice(y0, y0 + TILE), slice(x0, x0 + TILE))
    return canvas["np_bin"][sl], canvas["nt"][sl], canvas["hv"][:, sl[0], sl[1]]


def tile_tokens(idx):
    return torch.randn(D, TILE // 16, TILE // 16, generator=torch.Generator().manual_seed(100 + idx))


def run_reference(root, canvas):
    ref_cd = ref_shim.import_reference_cell_detection()
    ref_cellvit, _ = ref_shim.import_reference()
    import torchvision.transforms as T

    class StandInNet:
        patch_size, num_nuclei_classes = 16, 6

        def forward(self, patches, retrieve_tokens=True):
            idxs = [int(round((float(p[0, 0, 0]) * 0.5 + 0.5) * 255)) for p in patches]
            maps = [tile_maps(canvas, i) for i in idxs]
            lg = [synth.head_logits_from_maps(m[0], m[1], 6) for m in maps]
            return {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])),
                    "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])),
                    "hv_map": torch.from_numpy(np.ascontiguousarray(np.stack([m[2] for m in maps]))),
                    "tokens": torch.stack([tile_tokens(i) for i in idxs]), "tissue_types": torch.zeros(len(idxs), 19)}

    net = StandInNet()
    net.calculate_instance_map = types.MethodType(ref_cellvit.CellViT.calculate_instance_map, net)
    inf = object.__new__(ref_cd.CellSegmentationInference)
    inf.logger, inf.device, inf.mixed_precision, inf.model = logging.getLogger("ref"), "cpu", False, net
    inf.run_conf = {"dataset_config": {"nuclei_types": dict(NUCLEI_TYPES)}}
    inf.inference_transforms = T.Compose([T.ToTensor(), T.Normalize(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))])
    wsi = ref_cd.WSI(name="slide", patient="p", slide_path=root, patched_slide_path=root)
    inf.process_wsi(wsi, subdir_name="ref", patch_size=TILE, overlap=OV, batch_size=2, geojson=True)
    return root / "cell_detection" / "ref"
