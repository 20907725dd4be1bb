# -*- coding: utf-8 -*-
"""Inference entry points mirroring ``cell_segmentation/inference/cell_detection.py`` for the tile hot path.

Mirrored: ``CellSegmentationInference.__init__`` / checkpoint loading (cell_detection.py:93-242) and
``get_cell_predictions_with_tokens`` (:485-514). New for the B200 deployment: ``process_tiles`` -- the per-batch hot
loop of ``process_wsi`` (:306-323) over an in-memory tile stream, sharded round-robin over ``torch.distributed``
ranks with one NCCL weight broadcast and an optional all-gather of the per-tile instance tables.
WSI file IO, edge-cell merging and JSON/GeoJSON export (:244-304, :424-483, :516-902) are "next" rows.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, List, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .cellvit import CellViT, CellViT256, CellViTSAM


def unflatten_dict(d: dict, sep: str = ".") -> dict:
    """utils/tools.py:176 -- checkpoints store the run configuration flattened."""
    out: dict = {}
    for key, value in d.items():
        cur = out
        parts = key.split(sep)
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = value
    return out


def shard_indices(n_tiles: int, rank: int, world_size: int) -> List[int]:
    """Tiles rank r processes: r, r + world, ... (independent units, no data-path collective)."""
    return list(range(rank, n_tiles, world_size))


def model_from_checkpoint(checkpoint: dict) -> Union[CellViT, CellViT256, CellViTSAM]:
    """cell_detection.py:131-211: ``{"arch", "config" (flattened), "model_state_dict"}`` -> eval-mode model."""
    run_conf = unflatten_dict(checkpoint["config"], ".")
    arch = checkpoint["arch"]
    implemented = ["CellViT", "CellViT256", "CellViTSAM"]
    if arch not in implemented:
        raise NotImplementedError(f"Unknown model type. Please select one of {implemented}")
    data, mconf = run_conf["data"], run_conf.get("model", {})
    if arch == "CellViT":
        model = CellViT(num_nuclei_classes=data["num_nuclei_classes"], num_tissue_classes=data["num_tissue_classes"],
                        embed_dim=mconf["embed_dim"], input_channels=mconf.get("input_channels", 3), depth=mconf["depth"],
                        num_heads=mconf["num_heads"], extract_layers=mconf["extract_layers"],
                        regression_loss=mconf.get("regression_loss", False))
    elif arch == "CellViT256":
        model = CellViT256(model256_path=None, num_nuclei_classes=data["num_nuclei_classes"],
                           num_tissue_classes=data["num_tissue_classes"], regression_loss=mconf.get("regression_loss", False))
    else:
        model = CellViTSAM(model_path=None, num_nuclei_classes=data["num_nuclei_classes"],
                           num_tissue_classes=data["num_tissue_classes"], vit_structure=mconf["backbone"],
                           regression_loss=mconf.get("regression_loss", False))
    model.load_state_dict(checkpoint["model_state_dict"])
    model.eval()
    return model, run_conf


def broadcast_weights(model: torch.nn.Module, src: int = 0) -> None:
    """Collective C1: one flat broadcast of every floating-point parameter / buffer from ``src``."""
    import torch.distributed as dist
    tensors = [p.data for p in model.parameters()] + [b for b in model.buffers() if b.dtype.is_floating_point]
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src)
    o = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[o:o + n].view_as(t))
        o += n


class CellSegmentationInference:
    def __init__(self, model_path: Union[Path, str, dict], gpu: int, enforce_mixed_precision: bool = False) -> None:
        """cell_detection.py:93-115. ``model_path`` may also be an already loaded checkpoint dict."""
        self.device = f"cuda:{gpu}"
        ckpt = model_path if isinstance(model_path, dict) else torch.load(Path(model_path), map_location="cpu")
        self.model, self.run_conf = model_from_checkpoint(ckpt)
        self.model.to(self.device)
        tr = self.run_conf.get("transformations", {})
        norm = tr.get("normalize", {}) if isinstance(tr, dict) else {}
        self.mean = tuple(norm.get("mean", (0.5, 0.5, 0.5)))
        self.std = tuple(norm.get("std", (0.5, 0.5, 0.5)))
        # the tile engine always computes fp16 operands / fp32 accumulate (the reference's AMP mode, :314-318)
        self.mixed_precision = True if enforce_mixed_precision else self.run_conf.get("training", {}).get("mixed_precision", False)

    @classmethod
    def from_model(cls, model: CellViT, gpu: int) -> "CellSegmentationInference":
        """Wrap an already constructed (and loaded) model instead of reading a checkpoint file."""
        self = cls.__new__(cls)
        self.device = f"cuda:{gpu}"
        self.model, self.run_conf = model.eval().to(self.device), {}
        self.mean = self.std = (0.5, 0.5, 0.5)
        self.mixed_precision = True
        return self

    def get_cell_predictions_with_tokens(self, predictions: dict, magnification: int = 40) -> Tuple[List[dict], torch.Tensor]:
        """cell_detection.py:485-514."""
        predictions["nuclei_binary_map"] = F.softmax(predictions["nuclei_binary_map"], dim=1)
        predictions["nuclei_type_map"] = F.softmax(predictions["nuclei_type_map"], dim=1)
        _, instance_types = self.model.calculate_instance_map(predictions, magnification=magnification)
        tokens = predictions["tokens"].to("cpu")
        return instance_types, tokens

    def normalise(self, tiles_u8: np.ndarray) -> torch.Tensor:
        """ToTensor + Normalize (cell_detection.py:214-227): uint8 [B,H,W,3] -> float32 [B,3,H,W]."""
        x = torch.from_numpy(np.ascontiguousarray(tiles_u8)).to(self.device).permute(0, 3, 1, 2).float() / 255.0
        mean = torch.tensor(self.mean, device=self.device).view(1, 3, 1, 1)
        std = torch.tensor(self.std, device=self.device).view(1, 3, 1, 1)
        return (x - mean) / std

    @torch.no_grad()
    def process_tiles(self, batches: Iterable[torch.Tensor], magnification: int = 40, head_override: dict = None,
                      host_threads: int = 0) -> List[List[dict]]:
        """Hot loop of process_wsi (:306-323) over already normalised batches [B,3,H,W] (pinned host or device
        tensors). Per batch: H2D, forward, softmax (:500-505), device post-processing, D2H of label maps and instance
        tables; the host part (contours + dict building, post_proc_cellvit.py:96-151) of batch k overlaps the device
        work of batch k+1. ``head_override`` replaces head maps before post-processing (bench/test hook: random-init
        networks emit constant maps). Returns one list of per-tile instance dicts per batch."""
        from concurrent.futures import ThreadPoolExecutor
        from .post_proc_cellvit import DetectionCellPostProcessor
        proc = DetectionCellPostProcessor(nr_types=self.model.num_nuclei_classes, magnification=magnification, gt=False)
        results, pending = [], None
        # host_threads > 0 spreads the per-tile dict building over a thread pool; with CPython's GIL this only pays
        # when cv2.findContours dominates, so the default (0) keeps it on the calling thread.
        with ThreadPoolExecutor(max_workers=max(1, host_threads)) as pool_:
            pool = pool_ if host_threads > 0 else None
            for k, patches in enumerate(batches):
                patches = patches.to(self.device, non_blocking=True)
                predictions = self.model.forward(patches, retrieve_tokens=True)
                if head_override:
                    predictions.update(head_override)
                np_map = F.softmax(predictions["nuclei_binary_map"], dim=1)
                nt_map = F.softmax(predictions["nuclei_type_map"], dim=1)
                proc.launch_float(np_map, predictions["hv_map"], nt_map, slot=k & 1)
                if pending is not None:
                    results.append(proc.collect(pending, pool)[1])
                pending = k & 1
            if pending is not None:
                results.append(proc.collect(pending, pool)[1])
        return results

    def process_wsi(self, *args, **kwargs):
        raise NotImplementedError("WSI ingest/export is outside the tile hot path (SURVEY.md section 8f, rows N3/N4); "
                                  "feed tiles through process_tiles")
