# -*- coding: utf-8 -*-
"""Drop-in mirror of the reference post-processor (cell_segmentation/utils/post_proc_cellvit.py:33-153) whose
stages P1-P9 run on the GPU through ``cvb_postproc`` / ``cvb_postproc_maps`` (csrc/postproc.cu).

Instance contours (``cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0]``, post_proc_cellvit.py:106-125)
are traced on the device as well (``cvb_contours``); the host only assembles the
per-tile dicts, and calls cv2 itself for the rare instance the device flags (several 8-connected components, or
more than MAX_PTS points). There is no CPU fallback for the map stages: without libcellvit_b200.so or a CUDA device
the calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np
import torch

from . import _lib as L

ROW_DTYPE = np.dtype([("id", "<i4"), ("rmin", "<i4"), ("cmin", "<i4"), ("rmax", "<i4"), ("cmax", "<i4"),
                      ("area", "<i4"), ("type", "<i4"), ("type_prob_f", "<f4"), ("cx", "<f8"), ("cy", "<f8"),
                      ("type_prob", "<f8"), ("hist", "<i4", (8,))])
assert ROW_DTYPE.itemsize == 88
MAX_PTS = 128       # contour points kept per instance on the device (longer contours fall back to cv2)
ROWS_COPIED = 2048  # table / contour rows copied to the host eagerly per tile (more instances: second copy)


def magnification_params(magnification, gt: bool = False):
    """post_proc_cellvit.py:55-65."""
    if magnification == 40:
        object_size, k_size = 10, 21
    elif magnification == 20:
        object_size, k_size = 3, 11
    else:
        raise NotImplementedError("Unknown magnification")
    if gt:
        object_size, k_size = 100, 21
    return object_size, k_size


class _Workspace:
    """Device scratch + output buffers for one (B, H, W); reused across calls. Outputs (label maps, instance tables,
    contours) exist once per pipeline slot, so batch k's results stay intact on the device while batch k+1 runs."""

    N_SLOTS = 2

    def __init__(self, B, H, W, device, max_rows):
        need = C.c_size_t()
        L.check(L.lib().cvb_postproc_workspace_bytes(B, H, W, C.byref(need)), "cvb_postproc_workspace_bytes")
        self.key = (B, H, W, str(device), max_rows)
        self.ws = torch.empty(need.value, dtype=torch.uint8, device=device)
        L.check(L.lib().cvb_contours_workspace_bytes(B, H, W, C.byref(need)), "cvb_contours_workspace_bytes")
        self.cws = torch.empty(need.value, dtype=torch.uint8, device=device)
        self.dev = [dict(labels=torch.empty(B, H, W, dtype=torch.int32, device=device),
                         table=torch.empty(B, max_rows, ROW_DTYPE.itemsize, dtype=torch.uint8, device=device),
                         counts=torch.empty(B, dtype=torch.int32, device=device),
                         pts=torch.empty(B, max_rows, MAX_PTS, 2, dtype=torch.int16, device=device),
                         npts=torch.empty(B, max_rows, dtype=torch.int32, device=device)) for _ in range(self.N_SLOTS)]
        # the synchronous entry points use slot 0
        d0 = self.dev[0]
        self.labels, self.table, self.counts, self.pts, self.npts = d0["labels"], d0["table"], d0["counts"], d0["pts"], d0["npts"]
        # pinned host slots so that the host can finish batch k while the device runs batch k+1
        self.host = [dict(labels=torch.empty(B, H, W, dtype=torch.int32).pin_memory(),
                          table=torch.empty(B, max_rows, ROW_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
                          pts=torch.empty(B, max_rows, MAX_PTS, 2, dtype=torch.int16).pin_memory(),
                          npts=torch.empty(B, max_rows, dtype=torch.int32).pin_memory(),
                          counts=torch.empty(B, dtype=torch.int32).pin_memory(), event=None) for _ in range(self.N_SLOTS)]

    def launch_contours(self, B, H, W, max_rows, slot=0, connected=True):
        """``connected``: the slot's label map was written by cvb_postproc, whose instances are 4-connected by construction
        (markers are connected components, the flood labels 4-neighbours only): the 8-connected multi-component check of
        cvb_contours is skipped (workspace = NULL)."""
        d = self.dev[slot]
        L.check(L.lib().cvb_contours(L.ptr(d["labels"]), L.ptr(d["table"]), L.ptr(d["counts"]), B, H, W, max_rows, MAX_PTS,
                                     L.ptr(d["pts"]), L.ptr(d["npts"]), None if connected else L.ptr(self.cws),
                                     C.c_size_t(0 if connected else self.cws.numel()), L.stream_ptr()),
                "cvb_contours")

    def copy_to_host(self, slot, n_rows, B=None):
        """Enqueue the D2H copies of one batch on the current stream. Every copy is a CONTIGUOUS block: a strided
        slice such as ``table[:, :n_rows]`` would make torch stage it through a temporary and block the host until
        the device has drained (which serialises the host dict building with the next batch's device work)."""
        h, d = self.host[slot], self.dev[slot]
        B = d["labels"].shape[0] if B is None else B
        h["B"] = B
        h["labels"][:B].copy_(d["labels"][:B], non_blocking=True)
        h["counts"][:B].copy_(d["counts"][:B], non_blocking=True)
        for b in range(B):
            h["table"][b, :n_rows].copy_(d["table"][b, :n_rows], non_blocking=True)
            h["pts"][b, :n_rows].copy_(d["pts"][b, :n_rows], non_blocking=True)
            h["npts"][b, :n_rows].copy_(d["npts"][b, :n_rows], non_blocking=True)
        h["rows_copied"] = n_rows
        h["stream"] = torch.cuda.current_stream()
        h["event"] = torch.cuda.Event()
        h["event"].record()
        return h


class TileCells:
    """The cells of one tile as flat arrays (no per-cell Python objects): table ``rows`` (ROW_DTYPE), and the contours of
    the rows in ``valid`` (>= 3 points, post_proc_cellvit.py:110-113) concatenated in ``points`` [sum(lens), 2] (x, y)
    with ``lens``. Instances the device could not trace (several 8-connected components, > MAX_PTS points) are resolved
    with cv2 on the label map, exactly as ``rows_to_dict`` does."""

    def __init__(self, labels, rows, pts, npts, with_types=True):
        self.rows, self.with_types = rows, with_types
        npts = np.asarray(npts)
        flagged = np.nonzero(npts < 0)[0]
        if len(flagged) == 0:
            self.valid = np.nonzero(npts >= 3)[0]
            self.lens = npts[self.valid].astype(np.int64)
            mask = np.arange(pts.shape[1])[None, :] < self.lens[:, None]
            self.points = pts[self.valid][mask].astype(np.int32)
        else:  # rare: per-instance path
            import cv2
            valid, lens, chunks = [], [], []
            for i in range(len(rows)):
                if npts[i] >= 3:
                    c = pts[i, :npts[i]].astype(np.int32)
                elif npts[i] >= 0:
                    continue
                else:
                    r = rows[i]
                    crop = (labels[r["rmin"]:r["rmax"], r["cmin"]:r["cmax"]] == r["id"]).astype(np.uint8)
                    c = np.squeeze(cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)[0][0].astype("int32"))
                    if c.ndim != 2 or c.shape[0] < 3:
                        continue
                    c = c + np.array([r["cmin"], r["rmin"]], dtype=np.int32)
                valid.append(i); lens.append(len(c)); chunks.append(c)
            self.valid = np.array(valid, dtype=np.int64)
            self.lens = np.array(lens, dtype=np.int64)
            self.points = np.concatenate(chunks) if chunks else np.zeros((0, 2), np.int32)

    def __len__(self):
        return len(self.valid)


class DetectionCellPostProcessor:
    def __init__(self, nr_types: int = None, magnification=40, gt: bool = False, max_rows: int = 8192) -> None:
        self.nr_types = nr_types
        self.magnification = magnification
        self.gt = gt
        self.object_size, self.k_size = magnification_params(magnification, gt)
        self.max_rows = max_rows
        self._wsp = None

    # ------------------------------------------------------------------ device entry points
    def needs_realloc(self, B, H, W, device) -> bool:
        """True if a batch of this shape does not fit the current workspace (a smaller batch of the same tile size does)."""
        w = self._wsp
        return w is None or w.key[1:] != (H, W, str(device), self.max_rows) or B > w.key[0]

    def _workspace(self, B, H, W, device) -> _Workspace:
        """Workspace for up to B tiles of H x W; a smaller batch (the ragged tail of a tile stream) reuses the buffers of a
        larger one -- reallocating would drop the pinned results of a batch that has not been collected yet."""
        if self.needs_realloc(B, H, W, device):
            self._wsp = _Workspace(B, H, W, device, self.max_rows)
        return self._wsp

    def run_maps(self, np_bin: torch.Tensor, hv: torch.Tensor, type_map: torch.Tensor = None, debug: bool = False):
        """np_bin uint8 [B,H,W], hv float32 [B,2,H,W], type_map int32 [B,H,W] (all CUDA). Returns
        (labels int32 [B,H,W] device, rows list of structured arrays, debug dict)."""
        if not np_bin.is_cuda:
            raise RuntimeError("cellvit_b200 has no CPU path: post-processing inputs must be CUDA tensors")
        B, H, W = np_bin.shape
        np_bin, hv = np_bin.contiguous(), hv.contiguous().float()
        type_map = None if type_map is None else type_map.contiguous().to(torch.int32)
        with torch.cuda.device(np_bin.device):
            w = self._workspace(B, H, W, np_bin.device)
            dbg = {}
            if debug:
                dbg = {"blb": torch.empty(B, H, W, dtype=torch.uint8, device=np_bin.device),
                       "dist": torch.empty(B, H, W, dtype=torch.float64, device=np_bin.device),
                       "marker": torch.empty(B, H, W, dtype=torch.int32, device=np_bin.device)}
            L.check(L.lib().cvb_postproc_maps(L.ptr(np_bin), L.ptr(hv), L.ptr(type_map), B, H, W, int(self.nr_types or 0),
                                              self.object_size, self.k_size, L.ptr(w.labels), L.ptr(w.table), L.ptr(w.counts),
                                              self.max_rows, L.ptr(dbg.get("blb")), L.ptr(dbg.get("dist")), L.ptr(dbg.get("marker")),
                                              L.ptr(w.ws), C.c_size_t(w.ws.numel()), L.stream_ptr()), "cvb_postproc_maps")
            return w.labels[:B], self._rows(w, B), dbg

    def run_float(self, np_map: torch.Tensor, hv: torch.Tensor, nt_map: torch.Tensor = None):
        """np_map [B,2,H,W], hv [B,2,H,W], nt_map [B,C,H,W] float32 CUDA (probabilities or logits; argmax on device)."""
        if not np_map.is_cuda:
            raise RuntimeError("cellvit_b200 has no CPU path: post-processing inputs must be CUDA tensors")
        if self.gt:
            raise NotImplementedError("gt=True is served by run_maps (object_size=100)")
        B, _, H, W = np_map.shape
        np_map, hv = np_map.contiguous().float(), hv.contiguous().float()
        nt_map = None if nt_map is None else nt_map.contiguous().float()
        with torch.cuda.device(np_map.device):
            w = self._workspace(B, H, W, np_map.device)
            L.check(L.lib().cvb_postproc(L.ptr(np_map), L.ptr(hv), L.ptr(nt_map), B, H, W, 0 if nt_map is None else nt_map.shape[1],
                                         int(self.magnification), L.ptr(w.labels), L.ptr(w.table), L.ptr(w.counts), self.max_rows,
                                         L.ptr(w.ws), C.c_size_t(w.ws.numel()), L.stream_ptr()), "cvb_postproc")
            return w.labels[:B], self._rows(w, B)

    # ------------------------------------------------------------------ asynchronous (pipelined) use
    def launch_float(self, np_map: torch.Tensor, hv: torch.Tensor, nt_map: torch.Tensor, slot: int, table_rows: int = ROWS_COPIED,
                     tokens: torch.Tensor = None, patch_size: int = 16):
        """Enqueue cvb_postproc and the D2H copies of its results into pinned host slot ``slot`` (0/1) on the current
        stream; returns immediately. ``collect(slot)`` waits for that slot and builds the per-tile dicts. With
        ``tokens`` [B,D,h,w] the per-cell mean tokens (cell_detection.py:397-409) are pooled on the device too."""
        B, _, H, W = np_map.shape
        with torch.cuda.device(np_map.device):
            w = self._workspace(B, H, W, np_map.device)
            d = w.dev[slot]
            nt = nt_map if self.nr_types is not None else None
            L.check(L.lib().cvb_postproc(L.ptr(np_map), L.ptr(hv), L.ptr(nt), B, H, W, 0 if nt is None else nt.shape[1],
                                         int(self.magnification), L.ptr(d["labels"]), L.ptr(d["table"]), L.ptr(d["counts"]),
                                         self.max_rows, L.ptr(w.ws), C.c_size_t(w.ws.numel()), L.stream_ptr()), "cvb_postproc")
            w.launch_contours(B, H, W, self.max_rows, slot)
            if tokens is not None:
                tokens = tokens.contiguous().float()
                D, th, tw = tokens.shape[1:]
                if d.get("cell_tokens") is None or d["cell_tokens"].shape[-1] != D:
                    d["cell_tokens"] = torch.empty(B, self.max_rows, D, dtype=torch.float32, device=np_map.device)
                L.check(L.lib().cvb_cell_tokens(L.ptr(tokens), L.ptr(d["table"]), L.ptr(d["counts"]), B, D, th, tw, int(patch_size),
                                                self.max_rows, L.ptr(d["cell_tokens"]), L.stream_ptr()), "cvb_cell_tokens")
            w.copy_to_host(slot, min(table_rows, self.max_rows), B)

    def launch_argmax(self, np_argmax: torch.Tensor, hv: torch.Tensor, nt_argmax: torch.Tensor, slot: int, table_rows: int = ROWS_COPIED,
                      tokens: torch.Tensor = None, patch_size: int = 16):
        """``launch_float`` for the arg-max planes of ``CellViT.forward(..., argmax_maps=True)`` (uint8 [B,H,W], CUDA): the
        planes are consumed in place by ``cvb_postproc_argmax`` -- no preparation pass over the 8 logit channels."""
        B, H, W = np_argmax.shape
        assert np_argmax.dtype == torch.uint8 and np_argmax.is_contiguous() and hv.is_contiguous()
        with torch.cuda.device(np_argmax.device):
            w = self._workspace(B, H, W, np_argmax.device)
            d = w.dev[slot]
            nt = nt_argmax if self.nr_types is not None else None
            assert nt is None or (nt.dtype == torch.uint8 and nt.is_contiguous())
            L.check(L.lib().cvb_postproc_argmax(L.ptr(np_argmax), L.ptr(hv), L.ptr(nt), B, H, W, 0 if nt is None else int(self.nr_types),
                                                int(self.magnification), L.ptr(d["labels"]), L.ptr(d["table"]), L.ptr(d["counts"]),
                                                self.max_rows, L.ptr(w.ws), C.c_size_t(w.ws.numel()), L.stream_ptr()), "cvb_postproc_argmax")
            w.launch_contours(B, H, W, self.max_rows, slot)
            if tokens is not None:
                tokens = tokens.contiguous().float()
                D, th, tw = tokens.shape[1:]
                if d.get("cell_tokens") is None or d["cell_tokens"].shape[-1] != D:
                    d["cell_tokens"] = torch.empty(B, self.max_rows, D, dtype=torch.float32, device=np_argmax.device)
                L.check(L.lib().cvb_cell_tokens(L.ptr(tokens), L.ptr(d["table"]), L.ptr(d["counts"]), B, D, th, tw, int(patch_size),
                                                self.max_rows, L.ptr(d["cell_tokens"]), L.stream_ptr()), "cvb_cell_tokens")
            w.copy_to_host(slot, min(table_rows, self.max_rows), B)

    def collect(self, slot: int, pool=None, with_tokens: bool = False, raw: bool = False):
        """Wait for host slot ``slot`` and assemble the per-tile instance dicts (reference layout). Returns
        (label maps, dicts) -- or (label maps, dicts, cell tokens per tile) when ``with_tokens``. ``raw``: instead of dicts,
        one ``TileCells`` per tile (instance table + device contours as flat arrays) for callers that process a tile's
        cells vectorised (process_wsi); the token rows then align with the table rows."""
        w = self._wsp
        h = w.host[slot]
        h["event"].synchronize()
        nb = h.get("B", h["counts"].shape[0])
        counts = h["counts"].numpy()[:nb]
        if (counts > self.max_rows).any():
            raise L.CvbError(f"instance table overflow: {int(counts.max())} rows needed, max_rows={self.max_rows}")
        d = w.dev[slot]
        if (counts > h["rows_copied"]).any():  # rare: more instances than the eagerly copied prefix
            with torch.cuda.stream(h["stream"]):  # the slot's device buffers are intact until its next launch
                h["table"].copy_(d["table"]); h["pts"].copy_(d["pts"]); h["npts"].copy_(d["npts"])
                torch.cuda.current_stream().synchronize()
        lab, tab, pts, npts = h["labels"].numpy()[:nb], h["table"].numpy(), h["pts"].numpy(), h["npts"].numpy()
        with_types = self.nr_types is not None
        kept = [[] for _ in range(len(counts))]

        def one(b):
            n = int(counts[b])
            rows = np.frombuffer(tab[b, :n].tobytes(), dtype=ROW_DTYPE)
            if raw:
                kept[b] = None
                return TileCells(lab[b], rows, pts[b, :n], npts[b, :n], with_types)
            return self.rows_to_dict(lab[b], rows, with_types, pts[b, :n], npts[b, :n], kept[b])

        dicts = [one(b) for b in range(len(counts))] if pool is None else list(pool.map(one, range(len(counts))))
        if not with_tokens:
            return lab, dicts
        toks = []
        with torch.cuda.stream(h["stream"]):
            for b in range(len(counts)):
                t = d["cell_tokens"][b, :int(counts[b])].cpu().numpy()  # exact row count, known only now
                if kept[b] is None:
                    toks.append(t)
                else:
                    toks.append(t[kept[b]] if len(kept[b]) else np.zeros((0, t.shape[-1]), np.float32))
        return lab, dicts, toks

    def _rows(self, w: _Workspace, B: int) -> List[np.ndarray]:
        counts = w.counts[:B].cpu().numpy()  # synchronises the stream
        if (counts > self.max_rows).any():
            raise L.CvbError(f"instance table overflow: {int(counts.max())} rows needed, max_rows={self.max_rows}")
        n_max = int(counts.max()) if counts.size else 0
        tab = w.table[:B, :max(n_max, 1)].cpu().numpy()
        return [np.frombuffer(tab[b].tobytes(), dtype=ROW_DTYPE, count=int(counts[b])) for b in range(len(counts))]

    # ------------------------------------------------------------------ host glue (contours, dict format)
    @staticmethod
    def rows_to_dict(labels: np.ndarray, rows: np.ndarray, with_types: bool = True, pts: np.ndarray = None,
                     npts: np.ndarray = None, kept: list = None) -> dict:
        """Instance table (+ device contours) -> the reference's per-tile dict (post_proc_cellvit.py:96-151).
        Without device contours, or for instances the device flagged (npts < 0), the contour comes from cv2.
        ``kept`` (optional list) receives the table row index of every instance that made it into the dict."""
        out = {}
        ids, tp, ty = rows["id"].astype(np.int32), rows["type_prob"].tolist(), rows["type"].tolist()
        # bbox / centroid arrays of all instances at once; the dict entries are rows of them (no per-instance array construction)
        bbox_all = np.stack([np.stack([rows["rmin"], rows["cmin"]], 1), np.stack([rows["rmax"], rows["cmax"]], 1)], 1).astype(np.int64)
        cent_all = np.stack([rows["cx"], rows["cy"]], 1)
        nn = npts.tolist() if npts is not None else None
        if pts is not None:
            # the device's padded point rows -> ONE compact int32 array of all contour points of the tile (a private copy of the
            # pinned buffer); the dict entries below are slices of it
            lens = np.clip(np.asarray(npts), 0, None)
            flat = pts[np.arange(pts.shape[1])[None, :] < lens[:, None]].astype(np.int32)
            offs = np.concatenate([[0], np.cumsum(lens)]).tolist()
        for i, inst_id in enumerate(ids):
            if nn is not None and nn[i] >= 0:
                if nn[i] < 3:
                    continue  # "< 3 points dont make a contour" (post_proc_cellvit.py:110-113)
                contour = flat[offs[i]:offs[i + 1]]
            else:
                import cv2
                (rmin, cmin), (rmax, cmax) = bbox_all[i].tolist()
                crop = (labels[rmin:rmax, cmin:cmax] == inst_id).astype(np.uint8)
                cnts = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
                contour = np.squeeze(cnts[0][0].astype("int32"))
                if contour.shape[0] < 3 or contour.ndim != 2:
                    continue
                contour[:, 0] += cmin
                contour[:, 1] += rmin
            if kept is not None:
                kept.append(i)
            out[inst_id] = {
                "bbox": bbox_all[i],
                "centroid": cent_all[i],
                "contour": contour,
                "type_prob": tp[i] if with_types else None,
                "type": ty[i] if with_types else None,
            }
        return out

    def post_process_batch(self, np_map: torch.Tensor, hv_map: torch.Tensor, nt_map: torch.Tensor) -> Tuple[torch.Tensor, List[dict]]:
        """Batched device path used by ``CellViT.calculate_instance_map`` (cellvit.py:360-381)."""
        if not np_map.is_cuda:
            raise RuntimeError("cellvit_b200 has no CPU path: post-processing inputs must be CUDA tensors")
        self.launch_float(np_map.contiguous().float(), hv_map.contiguous().float(),
                          None if nt_map is None else nt_map.contiguous().float(), slot=0)
        _, dicts = self.collect(0)
        return self._wsp.dev[0]["labels"][:np_map.shape[0]], dicts

    def post_process_cell_segmentation(self, pred_map: np.ndarray) -> Tuple[np.ndarray, dict]:
        """Reference signature (post_proc_cellvit.py:67-153): pred_map [H,W,4] = (type, np, h, v) or [H,W,3]."""
        pred_map = np.asarray(pred_map)
        if self.nr_types is not None:
            pred_type = torch.from_numpy(np.ascontiguousarray(pred_map[..., 0].astype(np.int32)))[None].cuda()
            pred_inst = pred_map[..., 1:]
        else:
            pred_type, pred_inst = None, pred_map
        pred = np.array(pred_inst, dtype=np.float32)
        np_bin = torch.from_numpy(np.ascontiguousarray((pred[..., 0] >= 0.5).astype(np.uint8)))[None].cuda()
        hv = torch.from_numpy(np.ascontiguousarray(pred[..., 1:3].transpose(2, 0, 1)))[None].cuda()
        labels, rows, _ = self.run_maps(np_bin, hv, pred_type)
        w = self._wsp
        H, W = np_bin.shape[-2:]
        w.launch_contours(1, H, W, self.max_rows)
        n = len(rows[0])
        lab = labels[0].cpu().numpy()
        return lab.copy(), self.rows_to_dict(lab, rows[0], self.nr_types is not None, w.pts[0, :max(n, 1)].cpu().numpy()[:n],
                                             w.npts[0, :max(n, 1)].cpu().numpy()[:n])
