# -*- coding: utf-8 -*-
"""Inference entry points mirroring ``cell_segmentation/inference/cell_detection.py`` for the tile hot path.

Mirrored: ``CellSegmentationInference.__init__`` / checkpoint loading (cell_detection.py:93-242) and
``get_cell_predictions_with_tokens`` (:485-514). New for the B200 deployment: ``process_tiles`` -- the per-batch hot
loop of ``process_wsi`` (:306-323) over an in-memory tile stream, sharded round-robin over ``torch.distributed``
ranks with one NCCL weight broadcast and an optional all-gather of the per-tile instance tables.
``process_wsi`` (:244-483) ingests a preprocessed WSI folder (wsi_datamodel.py), pools the cell tokens on the device,
removes cells detected twice by overlapping tiles (wsi_merge.py) and writes the reference's JSON / GeoJSON / graph files
(SURVEY 8f rows N2-N4).
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable, List, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F

from .cellvit import CellViT, CellViT256, CellViTSAM


# cell_detection.py:75-90 (QuPath colours / names of the PanNuke classes)
COLOR_DICT = {1: [255, 0, 0], 2: [34, 221, 77], 3: [35, 92, 236], 4: [254, 255, 0], 5: [255, 159, 68]}
TYPE_NUCLEI_DICT = {1: "Neoplastic", 2: "Inflammatory", 3: "Connective", 4: "Dead", 5: "Epithelial"}
DEFAULT_NUCLEI_TYPES = {"Background": 0, "Neoplastic": 1, "Inflammatory": 2, "Connective": 3, "Dead": 4, "Epithelial": 5}


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


def model_from_checkpoint(checkpoint: dict):
    """cell_detection.py:131-211: ``{"arch", "config" (flattened), "model_state_dict"}`` -> eval-mode model. All six
    architectures of the reference's ``__get_model``. (There, a ``CellViT256Shared`` checkpoint dies with ``UnboundLocalError``
    because of a misspelt name test at :192; here it loads.)"""
    from .cellvit import CellViT256Shared, CellViTSAMShared, CellViTShared
    run_conf = unflatten_dict(checkpoint["config"], ".")
    arch = checkpoint["arch"]
    implemented = ["CellViT", "CellViTShared", "CellViT256", "CellViT256Shared", "CellViTSAM", "CellViTSAMShared"]
    if arch not in implemented:
        raise NotImplementedError(f"Unknown model type. Please select one of {implemented}")
    data, mconf = run_conf["data"], run_conf.get("model", {})
    if arch in ("CellViT", "CellViTShared"):
        cls = CellViT if arch == "CellViT" else CellViTShared
        model = cls(num_nuclei_classes=data["num_nuclei_classes"], num_tissue_classes=data["num_tissue_classes"],
                    embed_dim=mconf["embed_dim"], input_channels=mconf.get("input_channels", 3), depth=mconf["depth"],
                    num_heads=mconf["num_heads"], extract_layers=mconf["extract_layers"],
                    regression_loss=mconf.get("regression_loss", False))
    elif arch in ("CellViT256", "CellViT256Shared"):
        cls = CellViT256 if arch == "CellViT256" else CellViT256Shared
        model = cls(model256_path=None, num_nuclei_classes=data["num_nuclei_classes"],
                    num_tissue_classes=data["num_tissue_classes"], regression_loss=mconf.get("regression_loss", False))
    else:
        cls = CellViTSAM if arch == "CellViTSAM" else CellViTSAMShared
        model = cls(model_path=None, num_nuclei_classes=data["num_nuclei_classes"],
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
    if hasattr(model, "invalidate_packed"):   # the copies above go through ``.data`` and leave Tensor._version unchanged
        model.invalidate_packed()


class TableGather:
    """Collective C2 in chunks: the per-tile instance tables of several steps are staged in one device buffer and exchanged
    with ONE all-gather per ``chunk`` steps (and once more at the end of the stream) instead of two collectives per step --
    a collective's CTAs spin on SMs until every peer has arrived, and the persistent tile engine needs all of them.

    Layout of one staged step: ``[B, 4 + rows * row_bytes]`` uint8 -- per tile its int32 instance count, then the first
    ``rows`` table rows. Every rank must call ``add`` / ``flush`` the same number of times."""

    def __init__(self, world: int, B: int, rows: int, row_bytes: int, chunk: int, device):
        self.world, self.B, self.rows, self.row_bytes, self.chunk = world, B, rows, row_bytes, max(1, int(chunk))
        self.staged = torch.zeros(self.chunk, B, 4 + rows * row_bytes, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros(world, self.chunk, B, 4 + rows * row_bytes, dtype=torch.uint8, device=device)
        self.n = 0
        self.n_gathered = 0
        self.collectives = 0

    def add(self, counts: torch.Tensor, table: torch.Tensor) -> None:
        """``counts`` int32 [B], ``table`` uint8 [B, max_rows, row_bytes] (device); copies are enqueued on the current stream."""
        slot = self.staged[self.n]
        slot[:, :4].copy_(counts.view(torch.uint8).view(self.B, 4))
        slot[:, 4:].copy_(table[:, :self.rows].reshape(self.B, -1))
        self.n += 1
        if self.n == self.chunk:
            self.flush()

    def flush(self) -> None:
        if self.n == 0:
            return
        import torch.distributed as dist
        dist.all_gather_into_tensor(self.gathered.view(self.world * self.chunk, self.B, -1), self.staged)
        self.n_gathered, self.n = self.n, 0
        self.collectives += 1

    def tables(self, rank: int, step: int):
        """(counts int32 [B], rows uint8 [B, rows, row_bytes]) of ``rank`` for staged step ``step`` of the last exchange."""
        assert 0 <= step < self.n_gathered
        g = self.gathered[rank, step]
        return g[:, :4].contiguous().view(torch.int32).view(self.B), g[:, 4:].view(self.B, self.rows, self.row_bytes)


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
    def _pipeline(self, items, magnification: int = 40, head_override=None, with_tokens: bool = False, host_threads: int = 0,
                  use_graphs: bool = True, raw: bool = False):
        """Generator over ``items`` = iterable of (batch [B,3,H,W] pinned-host or device tensor, payload); yields
        ``(payload, dicts, cell_tokens)`` per batch, in order, where ``dicts`` are the per-tile instance dicts and
        ``cell_tokens`` (``with_tokens``) one float32 array [n_cells, D] per tile, rows aligned with the dict order.

        Three streams keep the device busy: the H2D copy of batch k+1 (copy stream) and the
        device post-processing (+ cell-token pooling) + D2H of batch k (post stream) run beside the forward of batch
        k+1 (the caller's stream); the host part (dict building, post_proc_cellvit.py:96-151) of batch k overlaps
        them too. ``head_override`` (dict, or callable(payload) -> dict) replaces head maps before post-processing
        (bench/test hook: random-init networks emit constant maps). ``use_graphs``: the forward of each pipeline slot
        is replayed from a CUDA graph (``CellViT.graph_slot``) whose static input buffer the H2D copy writes directly; a
        slot's static outputs have been consumed (its batch collected) before the slot is replayed. ``raw``: yield
        ``TileCells`` (flat arrays) instead of per-tile dicts."""
        from concurrent.futures import ThreadPoolExecutor
        from .post_proc_cellvit import DetectionCellPostProcessor
        proc = DetectionCellPostProcessor(nr_types=self.model.num_nuclei_classes, magnification=magnification, gt=False)
        dev = torch.device(self.device)
        with torch.cuda.device(dev), ThreadPoolExecutor(max_workers=max(1, host_threads)) as pool_:
            # host_threads > 0 spreads the per-tile dict building over a thread pool; with CPython's GIL this only
            # pays when cv2.findContours dominates, so the default (0) keeps it on the calling thread.
            pool = pool_ if host_threads > 0 else None
            main = torch.cuda.current_stream(dev)
            if getattr(self, "_streams", None) is None:
                self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            s_in, s_post = self._streams
            in_buf, consumed, keep = [None, None], [None, None], [None, None]
            mean_dev = torch.tensor(self.mean, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
            std_dev = torch.tensor(self.std, dtype=torch.float32, device=dev).view(1, 3, 1, 1)
            c255_dev = torch.tensor(255.0, dtype=torch.float32, device=dev)

            def stage(k, patches):
                """Input of batch k into the device buffer of slot k & 1 (H2D on the copy stream); returns
                (device tensor, ready event, graph slot or None)."""
                slot = k & 1
                gs = None
                if use_graphs:
                    gs = self.model.graph_slot(tuple(patches.shape), True, slot, dev, argmax_maps=True)
                    buf = gs[1]
                    if consumed[slot] is None:
                        consumed[slot] = torch.cuda.Event()
                        consumed[slot].record(main)
                elif patches.is_cuda and patches.dtype != torch.uint8:
                    return patches, None, None   # already normalised, produced on `main`: the forward reads it in stream order
                else:
                    if in_buf[slot] is None or in_buf[slot].shape != patches.shape:
                        in_buf[slot] = torch.empty(patches.shape, dtype=torch.float32, device=dev)
                        consumed[slot] = torch.cuda.Event()
                        consumed[slot].record(main)  # the fresh block may still be in use by earlier work on `main`
                    buf = in_buf[slot]
                s_in.wait_event(consumed[slot])      # the forward that last read this slot has finished
                if patches.is_cuda:
                    # a device batch may have been produced lazily on the caller's stream (e.g. a generator that normalises
                    # tiles on the GPU): its producer kernels are queued on `main`, so the copy stream has to wait for them
                    produced = torch.cuda.Event()
                    produced.record(main)
                    s_in.wait_event(produced)
                    patches.record_stream(s_in)
                with torch.cuda.stream(s_in):
                    if patches.dtype == torch.uint8:
                        # raw tiles: ToTensor + Normalize (:214-227) on the device, same operations in the same order
                        u8 = patches.to(dev, non_blocking=True)
                        # (division by a 0-dim TENSOR: torch's CUDA div-by-Python-scalar multiplies by the reciprocal instead)
                        buf.copy_((u8.to(torch.float32) / c255_dev - mean_dev) / std_dev)
                    else:
                        buf.copy_(patches, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(s_in)
                return buf, ready, gs

            def finish(slot):
                payload = keep[slot][0]
                res = proc.collect(slot, pool, with_tokens=with_tokens, raw=raw)
                dicts, toks = res[1], (res[2] if with_tokens else None)
                keep[slot] = None
                return payload, dicts, toks

            it = iter(items)
            nxt = next(it, None)
            staged = (stage(0, nxt[0]), nxt[1]) if nxt is not None else None
            k, pending = 0, None
            while staged is not None:
                (x, ready, gs), payload = staged
                slot = k & 1
                if ready is not None:
                    main.wait_event(ready)
                if gs is not None:
                    gs[0].replay()
                    predictions = dict(gs[2])
                else:
                    predictions = self.model.forward(x, retrieve_tokens=True, argmax_maps=True)
                ov = head_override(payload) if callable(head_override) else head_override
                if ov:  # (before the event: an override may enqueue device work of its own on this stream)
                    predictions.update(ov)
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                consumed[slot] = fwd_done
                nxt = next(it, None)
                staged = (stage(k + 1, nxt[0]), nxt[1]) if nxt is not None else None  # overlaps this forward
                if pending is not None and proc.needs_realloc(x.shape[0], x.shape[2], x.shape[3], dev):
                    yield finish(pending)  # a larger batch / other tile size re-creates the workspace: collect first
                    pending = None
                s_post.wait_event(fwd_done)
                with torch.cuda.stream(s_post):
                    # The reference soft-maxes NP / NT first (:500-505) and then only takes their arg-max (cellvit.py:369-375);
                    # softmax is monotonic, so the arg-max of the logits is the same map -- 270 MB of traffic per batch saved.
                    # (Callers that want the probabilities use get_cell_predictions_with_tokens, which keeps the softmax.)
                    # K12 fusion: the head epilogue already wrote the uint8 arg-max planes -> no pass over the 8 logit channels.
                    # An override of the float maps (test / bench hook) that brings no planes of its own takes the float entry.
                    tok = predictions["tokens"] if with_tokens else None
                    if ov and ("nuclei_binary_map" in ov or "nuclei_type_map" in ov) and "nuclei_binary_argmax" not in ov:
                        proc.launch_float(predictions["nuclei_binary_map"], predictions["hv_map"], predictions["nuclei_type_map"], slot=slot,
                                          tokens=tok, patch_size=self.model.patch_size)
                    else:
                        proc.launch_argmax(predictions["nuclei_binary_argmax"], predictions["hv_map"], predictions["nuclei_type_argmax"],
                                           slot=slot, tokens=tok, patch_size=self.model.patch_size)
                keep[slot] = (payload, predictions)  # alive until the D2H event of this batch has completed
                if pending is not None:
                    yield finish(pending)
                pending = slot
                k += 1
            if pending is not None:
                yield finish(pending)
            main.wait_stream(s_post)

    def process_tiles(self, batches: Iterable[torch.Tensor], magnification: int = 40, head_override: dict = None,
                      host_threads: int = 0, use_graphs: bool = True) -> List[List[dict]]:
        """Hot loop of process_wsi (:306-323) over already normalised batches [B,3,H,W] (pinned host or device
        tensors), see ``_pipeline``. Returns one list of per-tile instance dicts per batch."""
        return [d for _, d, _ in self._pipeline(((b, None) for b in batches), magnification, head_override, False, host_threads, use_graphs)]

    # ------------------------------------------------------------------ WSI level
    def process_wsi(self, wsi, subdir_name: str = None, patch_size: int = 1024, overlap: int = 64, batch_size: int = 8,
                    geojson: bool = False, num_workers: int = None, head_override=None, json_indent=None,
                    uint8_tiles: bool = True, shard: Tuple[int, int] = None):
        """cell_detection.py:244-483 -- all tiles of one preprocessed WSI -> ``cells.json``, ``cell_detection.json``
        (+ ``.geojson``) and ``cells.pt`` under ``<patched_slide_path>/cell_detection[/subdir_name]``.

        Same per-cell records as the reference (global bbox / centroid / contour, ``cell_status``, edge information,
        mean cell token). The forward, post-processing, contour tracing and token pooling run on the GPU through
        ``_pipeline``; the duplicate removal of overlapping tiles uses the GPU polygon-overlap kernel (wsi_merge.py).
        The cells of the slide are kept as columns (``wsi_records.CellColumns``), never as per-cell Python objects: the
        per-tile record arithmetic is vectorised, the duplicate removal reads the columns, and the JSON / GeoJSON files are
        streamed from them by the native writer ``cvb_export_json`` (byte-identical to ``json.dumps`` of the reference's
        record dicts). Returns the ``cells.json`` dictionary as a read-only mapping whose ``"cells"`` list of per-cell dicts
        is only built when it is read (``.columns`` holds the arrays). ``head_override`` is the bench/test hook of
        ``_pipeline``. ``json_indent``: the reference writes ``indent=2`` through ujson; ``None`` (default) writes the same
        content compact (a third of the bytes), ``json_indent=2`` gives the reference's layout.
        ``uint8_tiles``: ship raw uint8 tiles from the DataLoader workers and normalise on the device (bit-identical).

        Multi-GPU (one process per GPU): with ``torch.distributed`` initialised -- or an explicit ``shard=(rank, world_size)``
        -- rank r processes tiles r, r + world, ... of the slide (no collective on the tile path), the per-tile cell records
        are gathered on rank 0 and merged in dataset order, and rank 0 alone runs the cross-tile duplicate removal and
        writes the files: the output is identical to the single-process run. The other ranks return ``None``."""
        import gc
        # The per-cell records are millions of small acyclic lists / dicts: every container allocation counts towards the
        # cyclic collector's thresholds and its full passes re-traverse all of them (measured: 4x on the record loop, 2x on
        # the export). Nothing here creates reference cycles, so the collector is paused for the duration of the call.
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._process_wsi(wsi, subdir_name, patch_size, overlap, batch_size, geojson, num_workers, head_override,
                                     json_indent, uint8_tiles, shard)
        finally:
            if gc_was_enabled:
                gc.enable()

    def _process_wsi(self, wsi, subdir_name, patch_size, overlap, batch_size, geojson, num_workers, head_override, json_indent,
                     uint8_tiles, shard):
        import os
        from torch.utils.data import DataLoader
        from .wsi_datamodel import CellGraphDataWSI, InferenceTransform, PatchedWSIInference, SplitTensorList, save_cell_graph

        dataset = PatchedWSIInference(wsi, transform=InferenceTransform(self.mean, self.std, as_uint8=uint8_tiles))
        if num_workers is None:
            num_workers = int(np.clip(int(3 / 4 * (os.cpu_count() or 16)), 1, 2 * batch_size))
        if shard is None:
            import torch.distributed as dist
            shard = (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)
        rank, world = shard
        my_tiles = shard_indices(len(dataset), rank, world)
        loader = DataLoader(dataset if world == 1 else torch.utils.data.Subset(dataset, my_tiles), batch_size=batch_size,
                            num_workers=num_workers, shuffle=False, collate_fn=dataset.collate_batch, pin_memory=True)
        nuclei_types = self.run_conf.get("dataset_config", {}).get("nuclei_types", DEFAULT_NUCLEI_TYPES)
        background = nuclei_types.get("Background", 0)
        outdir = Path(wsi.patched_slide_path) / "cell_detection"
        if subdir_name is not None:
            outdir = outdir / subdir_name
        outdir.mkdir(exist_ok=True, parents=True)

        import time
        from .wsi_records import CellColumns, LazyCellsJson, write_cells_json, write_geojson
        t_start = time.perf_counter()
        bundles = []   # one per tile, in this rank's tile order: (patch id, CellColumns of the tile or None)
        scale, psize = wsi.metadata["downsampling"], wsi.metadata["patch_size"]
        t_records = 0.0
        for metadata, tiles, toks in self._pipeline(loader, wsi.metadata["magnification"], head_override, with_tokens=True, raw=True):
            t_rec0 = time.perf_counter()
            for meta, tc, tok in zip(metadata, tiles, toks):
                row, col = meta["row"], meta["col"]
                x_global = int(row * psize * scale - (row + 0.5) * overlap)      # :343-350 (x follows the tile ROW)
                y_global = int(col * psize * scale - (col + 0.5) * overlap)
                # the per-cell records of the tile (:352-409) as columns: vectorised arithmetic on the instance table, no
                # per-cell Python objects (constants 1024 / 64 as in the reference, :372-378)
                bundles.append((f"{row}_{col}", CellColumns.from_tile(tc, tok, row, col, np.array([x_global, y_global]), background, 1024, 64)))
            t_records += time.perf_counter() - t_rec0

        if world > 1:   # gather the per-tile columns on rank 0 and restore the dataset order
            import torch.distributed as dist
            gathered = [None] * world if rank == 0 else None
            if str(self.device).startswith("cuda"):   # NCCL stages the pickled arrays through the current device
                with torch.cuda.device(torch.device(self.device)):
                    dist.gather_object(bundles, gathered, dst=0)
            else:
                dist.gather_object(bundles, gathered, dst=0)
            if rank != 0:
                self.last_timings = {"tiles": time.perf_counter() - t_start, "of_which_cell_records": t_records}
                return None
            by_tile = {t: b for r in range(world) for t, b in zip(shard_indices(len(dataset), r, world), gathered[r])}
            bundles = [by_tile[t] for t in range(len(dataset))]
        processed_patches = [b[0] for b in bundles]
        cols = CellColumns.concat([b[1] for b in bundles], token_dim=self.model.embed_dim)
        t_tiles = time.perf_counter()
        keep_idx = self.post_process_edge_cells(cols)
        t_dedup = time.perf_counter()
        kept = cols.take(keep_idx)
        graph = CellGraphDataWSI(x=torch.from_numpy(kept.tokens), positions=torch.from_numpy(kept.centroid).to(torch.float32),
                                 contours=SplitTensorList(torch.from_numpy(kept.contour_pts).to(torch.float32), np.diff(kept.contour_off).tolist()),
                                 metadata={"wsi_metadata": wsi.metadata, "nuclei_types": nuclei_types})

        # cells.json / cell_detection.json (:438-463) and the GeoJSON pair (:443-447, 457-461) streamed from the columns
        header = {"wsi_metadata": wsi.metadata, "processed_patches": processed_patches, "type_map": nuclei_types}
        write_cells_json(kept, outdir / "cells.json", header, detection=False, indent=json_indent)
        write_cells_json(kept, outdir / "cell_detection.json", header, detection=True, indent=json_indent)
        if geojson:
            write_geojson(kept, outdir / "cells.geojson", True, TYPE_NUCLEI_DICT, COLOR_DICT, indent=json_indent)
            write_geojson(kept, outdir / "cell_detection.geojson", False, TYPE_NUCLEI_DICT, COLOR_DICT, indent=json_indent)
        save_cell_graph(graph, outdir / "cells.pt")
        t_end = time.perf_counter()
        # where the wall clock went (seconds): tile stream (decode + GPU + per-cell records), duplicate removal, export
        self.last_timings = {"tiles": t_tiles - t_start, "of_which_cell_records": t_records, "dedup": t_dedup - t_tiles,
                             "export": t_end - t_dedup}
        # the per-type cell counts the reference logs at the end of process_wsi (:470-478: value_counts("type") with the type names)
        inverse = {v: k for k, v in nuclei_types.items()}
        ids, counts = np.unique(kept.type, return_counts=True)
        self.last_stats = {"cells_before_cleaning": len(cols), "cells": len(kept),
                           "per_type": {inverse.get(int(t), int(t)): int(c) for t, c in sorted(zip(ids, counts), key=lambda tc: -tc[1])}}
        return LazyCellsJson(header, kept)

    def post_process_edge_cells(self, cell_list) -> List[int]:
        """cell_detection.py:516-536 -- indices of the cells to keep after the overlap clean-up. ``cell_list``: the reference's
        list of per-cell dicts, or the columnar store of the slide (``wsi_records.CellColumns``)."""
        from .wsi_merge import CellPostProcessor
        return CellPostProcessor(cell_list, None, torch.device(self.device)).post_process_cells()

    @staticmethod
    def convert_geojson(cell_list: List[dict], polygons: bool = False) -> List[dict]:
        """cell_detection.py:538-597 -- one MultiPolygon (segmentation) or MultiPoint (detection) feature per type."""
        import uuid
        features = []
        for cell_type in sorted({c["type"] for c in cell_list}):
            cells = [c for c in cell_list if c["type"] == cell_type]
            if polygons:
                coords = [[list(c["contour"]) + [c["contour"][0]]] for c in cells]  # closed rings
            else:
                coords = [c["centroid"] for c in cells]
            features.append({
                "type": "Feature", "id": str(uuid.uuid4()),
                "geometry": {"type": "MultiPolygon" if polygons else "MultiPoint", "coordinates": coords},
                "properties": {"objectType": "annotation",
                               "classification": {"name": TYPE_NUCLEI_DICT[cell_type], "color": COLOR_DICT[cell_type]}},
            })
        return features


def check_wsi(wsi, magnification: float = 40.0) -> None:
    """cell_detection.py:1008-1039 -- refuse a preprocessed slide whose tiling does not fit the network: ``RuntimeError`` if the
    tile magnification (``metadata["magnification"]``, else ``base_magnification / downsampling``) differs from
    ``magnification``, or the tiles are not 1024 px with a 64 px overlap. Same messages as the reference."""
    meta = wsi.metadata
    if meta["magnification"] is not None:
        tile_magnification = float(meta["magnification"])
    else:
        tile_magnification = float(float(meta["base_magnification"]) / meta["downsampling"])
    if tile_magnification != magnification:
        raise RuntimeError("The magnification is not matching to the network input magnification.")
    if int(meta["patch_size"]) % 256 != 0:
        raise RuntimeError("The patch-size must be devisible by 256.")
    if meta["patch_size"] != 1024:
        raise RuntimeError("The patch-size must be 1024.")
    if meta["patch_overlap"] != 64:
        raise RuntimeError("The patch-overlap must be 64")
