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
        self._streams = None

    @classmethod
    def from_model(cls, model: CellViT, gpu: int) -> "CellSegmentationInference":
        """Wrap an already constructed (and loaded) model instead of reading a checkpoint file."""
        self = cls.__new__(cls)
        self.device = f"cuda:{gpu}"
        self.model, self.run_conf = model.eval().to(self.device), {}
        self.mean = self.std = (0.5, 0.5, 0.5)
        self.mixed_precision = True
        self._streams = None
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
        tensors). Three streams keep the device busy: the H2D copy of batch k+1 (copy stream) and the softmax
        (:500-505) + device post-processing + D2H of batch k (post stream) run beside the forward of batch k+1 (the
        caller's stream); the host part (dict building, post_proc_cellvit.py:96-151) of batch k overlaps them too.
        ``head_override`` replaces head maps before post-processing (bench/test hook: random-init networks emit
        constant maps). Returns one list of per-tile instance dicts per batch."""
        from concurrent.futures import ThreadPoolExecutor
        from .post_proc_cellvit import DetectionCellPostProcessor
        proc = DetectionCellPostProcessor(nr_types=self.model.num_nuclei_classes, magnification=magnification, gt=False)
        dev = torch.device(self.device)
        results, pending = [], None
        with torch.cuda.device(dev), ThreadPoolExecutor(max_workers=max(1, host_threads)) as pool_:
            # host_threads > 0 spreads the per-tile dict building over a thread pool; with CPython's GIL this only
            # pays when cv2.findContours dominates, so the default (0) keeps it on the calling thread.
            pool = pool_ if host_threads > 0 else None
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_streams", None) is None:
                self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_in, s_post = self._streams
            in_buf, consumed, keep = [None, None], [None, None], [None, None]

            def stage(k, patches):
                """H2D of batch k into device slot k & 1 on the copy stream; returns (device tensor, ready event)."""
                if patches.is_cuda:
                    return patches, None
                slot = k & 1
                if in_buf[slot] is None or in_buf[slot].shape != patches.shape or in_buf[slot].dtype != patches.dtype:
                    in_buf[slot] = torch.empty(patches.shape, dtype=patches.dtype, device=dev)
                    consumed[slot] = torch.cuda.Event()
                    consumed[slot].record(main)     # the fresh block may still be in use by earlier work on `main`
                s_in.wait_event(consumed[slot])     # the forward that last read this slot has finished
                with torch.cuda.stream(s_in):
                    in_buf[slot].copy_(patches, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(s_in)
                return in_buf[slot], ready

            it = iter(batches)
            nxt = next(it, None)
            staged = stage(0, nxt) if nxt is not None else None
            k = 0
            while staged is not None:
                x, ready = staged
                slot = k & 1
                if ready is not None:
                    main.wait_event(ready)
                predictions = self.model.forward(x, retrieve_tokens=True)
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                consumed[slot] = fwd_done
                nxt = next(it, None)
                staged = stage(k + 1, nxt) if nxt is not None else None  # overlaps this forward
                if head_override:
                    predictions.update(head_override)
                s_post.wait_event(fwd_done)
                with torch.cuda.stream(s_post):
                    np_map = F.softmax(predictions["nuclei_binary_map"], dim=1)
                    nt_map = F.softmax(predictions["nuclei_type_map"], dim=1)
                    proc.launch_float(np_map, predictions["hv_map"], nt_map, slot=slot)
                keep[slot] = (predictions, np_map, nt_map)  # alive until the D2H event of this batch has completed
                if pending is not None:
                    results.append(proc.collect(pending, pool)[1])
                    keep[pending] = None
                pending = slot
                k += 1
            if pending is not None:
                results.append(proc.collect(pending, pool)[1])
                keep[pending] = None
            main.wait_stream(s_post)
        return results

    def process_wsi(self, *args, **kwargs):
        raise NotImplementedError("WSI ingest/export is outside the tile hot path (SURVEY.md section 8f, rows N3/N4); "
                                  "feed tiles through process_tiles")
