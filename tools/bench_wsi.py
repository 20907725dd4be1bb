"""process_wsi end to end on a synthetic preprocessed slide (G x G PNG tiles of 1024 px, 64 px overlap): tiles/s of the
whole reference entry point -- PNG decode (DataLoader workers), H2D, SAM-H forward, post-processing, contours, cell tokens,
D2H, per-cell records, duplicate removal with the GPU polygon-overlap kernel, JSON / graph export. Head maps are cut from one
synthetic-nuclei canvas (random-init heads are constant), so nuclei in the overlap bands are seen twice and merged.

``python tools/bench_wsi.py [G] [uint8]`` on one GPU, or under ``torchrun --nproc-per-node N`` (one process per GPU): the
slide is sharded over the ranks by process_wsi itself (tiles r, r + N, ...; records gathered on rank 0, which de-duplicates and
exports), rank 0 builds the slide folder and prints the line."""
import json, os, sys, tempfile, time
import numpy as np, torch, yaml
sys.path.insert(0, ".")
from PIL import Image
from cellvit_b200 import synth
from cellvit_b200.cell_detection import CellSegmentationInference
from cellvit_b200.cellvit import CellViTSAM
from cellvit_b200.wsi_datamodel import WSI

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
U8 = bool(int(sys.argv[2])) if len(sys.argv) > 2 else True    # raw uint8 tiles, normalised on the device (the default)
tile, ov, B = 1024, 64, 4
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    box = [tempfile.mkdtemp(prefix="wsi_") if rank == 0 else None]
    dist.broadcast_object_list(box, 0)
    root = box[0]
else:
    root = tempfile.mkdtemp(prefix="wsi_")
if rank == 0:
    os.makedirs(f"{root}/patches"); os.makedirs(f"{root}/metadata")
    yaml.safe_dump({"magnification": 40, "base_magnification": 40, "downsampling": 1, "patch_size": tile, "patch_overlap": ov,
                    "label_map": {"background": 0}}, open(f"{root}/metadata.yaml", "w"))
    rng = np.random.default_rng(0)
    entries = []
    for r in range(G):
        for c in range(G):
            name = f"s_{r}_{c}.png"
            img = (rng.integers(0, 256, (tile // 8, tile // 8, 3), dtype=np.uint8)).repeat(8, 0).repeat(8, 1)  # blocky: PNG-compressible
            Image.fromarray(img).save(f"{root}/patches/{name}")
            yaml.safe_dump({"row": r, "col": c}, open(f"{root}/metadata/s_{r}_{c}.yaml", "w"))
            entries.append({name: {"row": r, "col": c, "metadata_path": f"metadata/s_{r}_{c}.yaml"}})
    json.dump(entries, open(f"{root}/patch_metadata.json", "w"))
if world > 1:
    dist.barrier()
side = G * (tile - ov) + 2 * ov
canvas = synth.synthetic_nuclei(side, int(700 * (side / 1024.0) ** 2), seed=3)
torch.manual_seed(0)
inf = CellSegmentationInference.from_model(CellViTSAM(None, 6, 19, "SAM-H"), local)   # same seed on every rank = same weights

# head maps of every tile, resident on the device before the timed run (the override is a bench hook, not product work)
maps = {}
for r in range(G):
    for c in range(G):
        if (r * G + c) % world != rank:
            continue
        y0 = int(c * tile - (c + 0.5) * ov) + ov
        x0 = int(r * tile - (r + 0.5) * ov) + ov
        sl = (slice(y0, y0 + tile), slice(x0, x0 + tile))
        lg = synth.head_logits_from_maps(canvas["np_bin"][sl], canvas["nt"][sl], 6)
        maps[(r, c)] = (torch.from_numpy(lg[0]).cuda(), torch.from_numpy(lg[1]).cuda(),
                        torch.from_numpy(np.ascontiguousarray(canvas["hv"][:, sl[0], sl[1]])).cuda())

def override(metadata):
    m = [maps[(md["row"], md["col"])] for md in metadata]
    return {"nuclei_binary_map": torch.stack([a[0] for a in m]), "nuclei_type_map": torch.stack([a[1] for a in m]),
            "hv_map": torch.stack([a[2] for a in m])}

wsi = WSI(name="s", patient="p", slide_path=root, patched_slide_path=root)
inf.process_wsi(wsi, subdir_name="warm", batch_size=B, num_workers=8, head_override=override, uint8_tiles=U8)   # warm-up (graph capture, workers)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
out = inf.process_wsi(wsi, subdir_name="run", batch_size=B, geojson=True, num_workers=8, head_override=override, uint8_tiles=U8)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"tiles": G * G, "n_gpus": world, "uint8_tiles": U8, "cells": len(out.columns), "seconds": dt, "tiles_per_s": G * G / dt,
                      "phases_s": inf.last_timings, "note": "process_wsi incl. PNG decode, head-override host prep, dedup, JSON/GeoJSON/graph export"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
