# -*- coding: utf-8 -*-
"""ctypes front-end of oracle/postproc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may import this module.
It mirrors the call surface of the reference post-processor
(cell_segmentation/utils/post_proc_cellvit.py:33-153) so parity tests read like reference usage.

Parity status: stages P1-P6, P8, P9 are pinned against cv2 4.13 / scipy 1.18 and the reference's own
code imported from /root/reference (tests/test_oracle_vs_reference.py, golden fixtures under
tests/golden/). Stage P7 (skimage.segmentation.watershed, scikit-image==0.19.3) is PARITY UNPINNED:
the library is neither installed nor vendored; the restatement follows its published algorithm with
ties resolved by the total order (value, age, pixel index).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class InstRow(C.Structure):
    _fields_ = [("id", C.c_int32), ("rmin", C.c_int32), ("cmin", C.c_int32), ("rmax", C.c_int32),
                ("cmax", C.c_int32), ("area", C.c_int32), ("type", C.c_int32), ("type_prob_f", C.c_float),
                ("cx", C.c_double), ("cy", C.c_double), ("type_prob", C.c_double), ("hist", C.c_int32 * 8)]


ROW_DTYPE = np.dtype([("id", "<i4"), ("rmin", "<i4"), ("cmin", "<i4"), ("rmax", "<i4"), ("cmax", "<i4"),
                      ("area", "<i4"), ("type", "<i4"), ("type_prob_f", "<f4"), ("cx", "<f8"), ("cy", "<f8"),
                      ("type_prob", "<f8"), ("hist", "<i4", (8,))])


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcvoracle.so")
    src = os.path.join(_HERE, "postproc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcvoracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.cvo_proc_np_hv.restype = C.c_int
        L.cvo_proc_np_hv.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cvo_watershed.restype = None
        L.cvo_watershed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cvo_instance_table.restype = C.c_int
        L.cvo_instance_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        assert L.cvo_row_bytes() == ROW_DTYPE.itemsize == C.sizeof(InstRow)
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def magnification_params(magnification: int, gt: bool = False):
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


def watershed(image, markers=None, mask=None):
    """Drop-in for ``skimage.segmentation.watershed(image, markers=, mask=)`` (connectivity 1)."""
    image = np.ascontiguousarray(image, np.float64)
    H, W = image.shape
    mask_ = np.ones((H, W), np.uint8) if mask is None else np.ascontiguousarray(mask != 0, np.uint8)
    markers = np.ascontiguousarray(markers, np.int32)
    out = np.empty((H, W), np.int32)
    lib().cvo_watershed(_p(image), _p(markers), _p(mask_), H, W, _p(out))
    return out


def proc_np_hv(np_bin, hv, magnification=40, gt=False, want_intermediates=False):
    """Stages P1-P7. np_bin [H,W] (0/1), hv [2,H,W] float32 -> int32 labels [H,W]."""
    object_size, ksize = magnification_params(magnification, gt)
    np_bin = np.ascontiguousarray(np_bin, np.uint8)
    hv = np.ascontiguousarray(hv, np.float32)
    H, W = np_bin.shape
    labels = np.empty((H, W), np.int32)
    blb = np.empty((H, W), np.uint8) if want_intermediates else None
    dist = np.empty((H, W), np.float64) if want_intermediates else None
    marker = np.empty((H, W), np.int32) if want_intermediates else None
    rc = lib().cvo_proc_np_hv(_p(np_bin), _p(hv), H, W, object_size, ksize, _p(labels), _p(blb), _p(dist), _p(marker))
    if rc != 0:
        raise RuntimeError(f"cvo_proc_np_hv failed: {rc}")
    if want_intermediates:
        return labels, {"blb": blb, "dist": dist, "marker": marker}
    return labels


def instance_table(labels, type_map, nr_types):
    labels = np.ascontiguousarray(labels, np.int32)
    H, W = labels.shape
    tm = None if type_map is None else np.ascontiguousarray(type_map, np.int32)
    cap = 4096
    while True:
        rows = np.zeros(cap, ROW_DTYPE)
        n = lib().cvo_instance_table(_p(labels), _p(tm), H, W, int(nr_types or 0), _p(rows), cap)
        if n >= 0:
            return rows[:n]
        cap = -n


def rows_to_dict(labels, rows, with_contours=True):
    """Instance table -> the reference's per-tile dict (post_proc_cellvit.py:96-151)."""
    import cv2
    out = {}
    for r in rows:
        rmin, cmin, rmax, cmax = int(r["rmin"]), int(r["cmin"]), int(r["rmax"]), int(r["cmax"])
        contour = None
        if with_contours:
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
            "type_prob": float(r["type_prob"]),
            "type": int(r["type"]),
        }
    return out


class DetectionCellPostProcessor:
    """Oracle twin of the reference class (post_proc_cellvit.py:33-153)."""

    def __init__(self, nr_types=None, magnification=40, gt=False):
        self.nr_types, self.magnification, self.gt = nr_types, magnification, gt
        self.object_size, self.k_size = magnification_params(magnification, gt)

    def post_process_cell_segmentation(self, pred_map):
        pred_map = np.asarray(pred_map)
        if self.nr_types is not None:
            pred_type = pred_map[..., 0].astype(np.int32)
            pred_inst = pred_map[..., 1:]
        else:
            pred_type, pred_inst = None, pred_map
        pred = np.array(pred_inst, dtype=np.float32)
        np_bin = (pred[..., 0] >= 0.5).astype(np.uint8)
        hv = np.ascontiguousarray(pred[..., 1:3].transpose(2, 0, 1))
        labels = proc_np_hv(np_bin, hv, self.magnification, self.gt)
        rows = instance_table(labels, pred_type, self.nr_types)
        return labels, rows_to_dict(labels, rows)
