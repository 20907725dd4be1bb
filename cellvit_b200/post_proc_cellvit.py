# -*- coding: utf-8 -*-
"""Drop-in mirror of the reference post-processor (cell_segmentation/utils/post_proc_cellvit.py:33-153) whose
stages P1-P9 run on the GPU through ``cvb_postproc`` / ``cvb_postproc_maps`` (csrc/postproc.cu).

Only the contour of each instance (``cv2.findContours`` on the bbox crop, post_proc_cellvit.py:106-125) is still
computed on the host from the device label map -- SURVEY.md section 8f row N1. There is no CPU fallback for the
map stages: without libcellvit_b200.so or a CUDA device the calls raise.
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
    """Device scratch + output buffers for one (B, H, W); reused across calls."""

    def __init__(self, B, H, W, device, max_rows):
        need = C.c_size_t()
        L.check(L.lib().cvb_postproc_workspace_bytes(B, H, W, C.byref(need)), "cvb_postproc_workspace_bytes")
        self.key = (B, H, W, str(device), max_rows)
        self.ws = torch.empty(need.value, dtype=torch.uint8, device=device)
        self.labels = torch.empty(B, H, W, dtype=torch.int32, device=device)
        self.table = torch.empty(B, max_rows, ROW_DTYPE.itemsize, dtype=torch.uint8, device=device)
        self.counts = torch.empty(B, dtype=torch.int32, device=device)


class DetectionCellPostProcessor:
    def __init__(self, nr_types: int = None, magnification=40, gt: bool = False, max_rows: int = 8192) -> None:
        self.nr_types = nr_types
        self.magnification = magnification
        self.gt = gt
        self.object_size, self.k_size = magnification_params(magnification, gt)
        self.max_rows = max_rows
        self._wsp = None

    # ------------------------------------------------------------------ device entry points
    def _workspace(self, B, H, W, device) -> _Workspace:
        key = (B, H, W, str(device), self.max_rows)
        if self._wsp is None or self._wsp.key != key:
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
            return w.labels, self._rows(w), dbg

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
            return w.labels, self._rows(w)

    def _rows(self, w: _Workspace) -> List[np.ndarray]:
        counts = w.counts.cpu().numpy()  # synchronises the stream
        if (counts > self.max_rows).any():
            raise L.CvbError(f"instance table overflow: {int(counts.max())} rows needed, max_rows={self.max_rows}")
        n_max = int(counts.max()) if counts.size else 0
        tab = w.table[:, :max(n_max, 1)].cpu().numpy()
        return [np.frombuffer(tab[b].tobytes(), dtype=ROW_DTYPE, count=int(counts[b])) for b in range(len(counts))]

    # ------------------------------------------------------------------ host glue (contours, dict format)
    @staticmethod
    def rows_to_dict(labels: np.ndarray, rows: np.ndarray, with_types: bool = True) -> dict:
        """Instance table -> the reference's per-tile dict (post_proc_cellvit.py:96-151)."""
        import cv2
        out = {}
        for r in rows:
            rmin, cmin, rmax, cmax = int(r["rmin"]), int(r["cmin"]), int(r["rmax"]), int(r["cmax"])
            crop = (labels[rmin:rmax, cmin:cmax] == r["id"]).astype(np.uint8)
            cnts = cv2.findContours(crop, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
            contour = np.squeeze(cnts[0][0].astype("int32"))
            if contour.shape[0] < 3 or contour.ndim != 2:
                continue
            contour[:, 0] += cmin
            contour[:, 1] += rmin
            out[np.int32(r["id"])] = {
                "bbox": np.array([[rmin, cmin], [rmax, cmax]]),
                "centroid": np.array([r["cx"], r["cy"]]),
                "contour": contour,
                "type_prob": float(r["type_prob"]) if with_types else None,
                "type": int(r["type"]) if with_types else None,
            }
        return out

    def post_process_batch(self, np_map: torch.Tensor, hv_map: torch.Tensor, nt_map: torch.Tensor) -> Tuple[torch.Tensor, List[dict]]:
        """Batched device path used by ``CellViT.calculate_instance_map`` (cellvit.py:360-381)."""
        labels, rows = self.run_float(np_map, hv_map, nt_map if self.nr_types is not None else None)
        lab_host = labels.cpu().numpy()
        dicts = [self.rows_to_dict(lab_host[b], rows[b], self.nr_types is not None) for b in range(lab_host.shape[0])]
        return labels, dicts

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
        lab = labels[0].cpu().numpy()
        return lab, self.rows_to_dict(lab, rows[0], self.nr_types is not None)
