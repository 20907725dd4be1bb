# -*- coding: utf-8 -*-
"""Columnar store of the cells of one slide, and the native export of the WSI result files.

The reference builds one Python dict per cell (``cell_segmentation/inference/cell_detection.py:352-409``), keeps them in
lists, filters the lists and hands them to ``ujson.dump`` (``:438-475``). On a slide with 10^5 cells those dicts and their
JSON encoding are most of the wall clock once the network runs on a B200. Here a slide's cells live in a handful of numpy
arrays (``CellColumns``): per-tile arithmetic is vectorised, the duplicate removal reads the columns it needs, the files are
streamed from the columns by ``cvb_export_json`` (csrc/wsi_export.cu, byte-identical to ``json.dumps``), and the
reference's list of per-cell dicts is only materialised when a caller asks for it (``to_dicts``).
"""
from __future__ import annotations

import ctypes as C
import json
from collections.abc import Mapping
from typing import List, Optional

import numpy as np

from .wsi_merge import _EDGE_PATCHES, cell_status_batch

JSON_CELLS, JSON_DETECTION, JSON_POLYGONS, JSON_POINTS = 0, 1, 2, 3


class _CColumns(C.Structure):
    _fields_ = [("n", C.c_longlong), ("bbox", C.c_void_p), ("centroid", C.c_void_p), ("contour_pts", C.c_void_p),
                ("contour_off", C.c_void_p), ("type_prob", C.c_void_p), ("type", C.c_void_p), ("patch", C.c_void_p),
                ("status", C.c_void_p), ("offset", C.c_void_p), ("edge", C.c_void_p), ("position", C.c_void_p)]


class _CSection(C.Structure):
    _fields_ = [("head", C.c_char_p), ("tail", C.c_char_p), ("idx", C.c_void_p), ("n_idx", C.c_longlong), ("kind", C.c_int),
                ("depth", C.c_int)]


_FIELDS = (("bbox", np.int64, (2, 2)), ("centroid", np.float64, (2,)), ("type_prob", np.float64, ()), ("type", np.int64, ()),
           ("patch", np.int64, (2,)), ("status", np.int64, ()), ("offset", np.int64, (2,)), ("edge", np.uint8, ()),
           ("position", np.int8, (4,)))


class CellColumns:
    """Cells of a slide (or of one tile) as columns; row order = the reference's ``cell_dict_wsi`` order.

    bbox int64 [n,2,2], centroid f64 [n,2], type_prob f64 [n], type int64 [n], patch int64 [n,2] (tile row, col), status int64
    [n] (``cell_status``), offset int64 [n,2] (``offset_global``), edge uint8 [n] (``edge_position``), position int8 [n,4]
    ([top, right, down, left], meaningful where ``edge``), contour_pts int64 [P,2] + contour_off int64 [n+1], tokens f32 [n,D]
    (mean cell tokens, may be None)."""

    def __init__(self, **cols):
        for name, dtype, tail in _FIELDS:
            a = np.ascontiguousarray(cols[name], dtype=dtype)
            assert a.shape[1:] == tail, (name, a.shape)
            setattr(self, name, a)
        self.contour_pts = np.ascontiguousarray(cols["contour_pts"], dtype=np.int64).reshape(-1, 2)
        self.contour_off = np.ascontiguousarray(cols["contour_off"], dtype=np.int64)
        self.tokens = cols.get("tokens")
        n = len(self.type)
        assert all(len(getattr(self, f)) == n for f, _, _ in _FIELDS) and len(self.contour_off) == n + 1
        assert n == 0 or self.contour_off[-1] == len(self.contour_pts)

    def __len__(self) -> int:
        return len(self.type)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def empty(token_dim: Optional[int] = None) -> "CellColumns":
        z = {name: np.zeros((0,) + tail, dtype) for name, dtype, tail in _FIELDS}
        tok = None if token_dim is None else np.zeros((0, token_dim), np.float32)
        return CellColumns(contour_pts=np.zeros((0, 2), np.int64), contour_off=np.zeros(1, np.int64), tokens=tok, **z)

    @staticmethod
    def from_tile(tc, tokens, row: int, col: int, offset_global: np.ndarray, background: int, patch_size: int = 1024,
                  margin: int = 64) -> Optional["CellColumns"]:
        """The records of one tile (cell_detection.py:343-409, vectorised over the tile's instance table ``tc`` =
        ``post_proc_cellvit.TileCells``); None when the tile holds no (non-background) cell."""
        rows = tc.rows[tc.valid]
        sel = rows["type"] != background
        if not sel.any():
            return None
        rows = rows[sel]
        n = len(rows)
        offset_global = np.asarray(offset_global, dtype=np.int64)
        flip = offset_global[::-1]
        bbox_local = np.stack([np.stack([rows["rmin"], rows["cmin"]], 1), np.stack([rows["rmax"], rows["cmax"]], 1)], 1).astype(np.int64)
        flat = bbox_local.reshape(n, -1)
        edge = (flat.max(1) == patch_size) | (flat.min(1) == 0)                              # :376-378
        # get_cell_position (:789-819): [top, right, down, left]
        position = np.stack([bbox_local[:, 0, 0] == 0, bbox_local[:, 1, 1] == patch_size, bbox_local[:, 1, 0] == patch_size,
                             bbox_local[:, 0, 1] == 0], 1).astype(np.int8)
        lens = tc.lens[sel]
        off = np.zeros(n + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        pts = tc.points[np.repeat(sel, tc.lens)].astype(np.int64) + flip
        tok = None if tokens is None else np.ascontiguousarray(tokens[tc.valid[sel]])
        return CellColumns(bbox=bbox_local + offset_global, centroid=np.stack([rows["cx"], rows["cy"]], 1) + flip,
                           type_prob=rows["type_prob"], type=rows["type"], patch=np.tile(np.array([row, col], np.int64), (n, 1)),
                           status=cell_status_batch(bbox_local, patch_size, margin), offset=np.tile(offset_global, (n, 1)),
                           edge=edge, position=position, contour_pts=pts, contour_off=off, tokens=tok)

    @staticmethod
    def concat(parts: List["CellColumns"], token_dim: Optional[int] = None) -> "CellColumns":
        parts = [p for p in parts if p is not None and len(p)]
        if not parts:
            return CellColumns.empty(token_dim)
        cols = {name: np.concatenate([getattr(p, name) for p in parts]) for name, _, _ in _FIELDS}
        lens = np.concatenate([np.diff(p.contour_off) for p in parts])
        off = np.zeros(len(lens) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        tok = np.concatenate([p.tokens for p in parts]) if all(p.tokens is not None for p in parts) else None
        return CellColumns(contour_pts=np.concatenate([p.contour_pts for p in parts]), contour_off=off, tokens=tok, **cols)

    def take(self, idx) -> "CellColumns":
        """The cells ``idx`` (in that order) as a new store."""
        idx = np.asarray(idx, dtype=np.int64)
        lens = np.diff(self.contour_off)[idx]
        off = np.zeros(len(idx) + 1, np.int64)
        np.cumsum(lens, out=off[1:])
        gather = np.repeat(self.contour_off[:-1][idx] - off[:-1], lens) + np.arange(int(off[-1]), dtype=np.int64)
        cols = {name: getattr(self, name)[idx] for name, _, _ in _FIELDS}
        tok = None if self.tokens is None else self.tokens[idx]
        return CellColumns(contour_pts=self.contour_pts[gather], contour_off=off, tokens=tok, **cols)

    # ------------------------------------------------------------------ views the duplicate removal reads
    def contour(self, i: int) -> np.ndarray:
        return self.contour_pts[self.contour_off[i]:self.contour_off[i + 1]]

    def first_edge_patch(self, i: int):
        """``edge_information["edge_patches"][0]`` of edge cell i (cell_detection.py:656); TypeError for border codes without
        neighbour tiles, where the reference subscripts ``None``."""
        offs = _EDGE_PATCHES.get(tuple(int(p) for p in self.position[i]))
        if offs is None:
            raise TypeError("'NoneType' object is not subscriptable")
        return (int(self.patch[i, 0]) + offs[0][0], int(self.patch[i, 1]) + offs[0][1])

    # ------------------------------------------------------------------ the reference's per-cell dicts
    def to_dicts(self) -> List[dict]:
        """``cell_dict_wsi`` of the reference: one dict per cell (cell_detection.py:352-395), same keys in the same order."""
        bbox, cent = self.bbox.tolist(), self.centroid.tolist()
        pts, off = self.contour_pts.tolist(), self.contour_off.tolist()
        probs, types, patch = self.type_prob.tolist(), self.type.tolist(), self.patch.tolist()
        status, offset, edge, position = self.status.tolist(), self.offset.tolist(), self.edge.tolist(), self.position.tolist()
        out = []
        for n in range(len(types)):
            d = {"bbox": bbox[n], "centroid": cent[n], "contour": pts[off[n]:off[n + 1]], "type_prob": probs[n], "type": types[n],
                 "patch_coordinates": patch[n], "cell_status": status[n], "offset_global": offset[n]}
            if edge[n]:
                offs = _EDGE_PATCHES.get(tuple(position[n]))
                d["edge_position"] = True
                d["edge_information"] = {"position": position[n],
                                         "edge_patches": None if offs is None else [[patch[n][0] + a, patch[n][1] + b] for a, b in offs]}
            else:
                d["edge_position"] = False
            out.append(d)
        return out

    def detection_dicts(self) -> List[dict]:
        """``cell_dict_detection`` (cell_detection.py:410-416)."""
        return [{"bbox": b, "centroid": c, "type": t} for b, c, t in zip(self.bbox.tolist(), self.centroid.tolist(), self.type.tolist())]

    # ------------------------------------------------------------------ native export
    def _c_struct(self) -> _CColumns:
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        return _CColumns(len(self), p(self.bbox), p(self.centroid), p(self.contour_pts), p(self.contour_off), p(self.type_prob),
                         p(self.type), p(self.patch), p(self.status), p(self.offset), p(self.edge), p(self.position))

    def export_json(self, path, sections, indent: Optional[int] = None) -> None:
        """``sections`` = [(head str, tail str, idx or None (= all cells), kind, depth)]: writes head + array + tail per section."""
        from . import _lib as L
        keep, arr = [], (_CSection * len(sections))()
        for k, (head, tail, idx, kind, depth) in enumerate(sections):
            idx = np.arange(len(self), dtype=np.int64) if idx is None else np.ascontiguousarray(idx, dtype=np.int64)
            keep.append(idx)
            arr[k] = _CSection(head.encode("ascii"), tail.encode("ascii"), idx.ctypes.data_as(C.c_void_p), len(idx), kind, depth)
        cs = self._c_struct()
        L.check(L.lib().cvb_export_json(str(path).encode(), C.byref(cs), arr, len(sections), -1 if indent is None else int(indent)),
                "cvb_export_json")


_MARK = "@@cvb-array@@"


def _split(obj, indent):
    """json.dumps of ``obj`` (which holds the marker string once) split around the marker: (head, tail)."""
    s = json.dumps(obj, indent=indent)
    head, tail = s.split('"' + _MARK + '"')
    return head, tail


def write_cells_json(cols: CellColumns, path, header: dict, detection: bool = False, indent: Optional[int] = None) -> None:
    """``cells.json`` / ``cell_detection.json`` (cell_detection.py:438-463): ``header`` holds wsi_metadata, processed_patches and
    type_map; the ``cells`` array is streamed from the columns."""
    head, tail = _split({**header, "cells": _MARK}, indent)
    cols.export_json(path, [(head, tail, None, JSON_DETECTION if detection else JSON_CELLS, 1)], indent)


def write_geojson(cols: CellColumns, path, polygons: bool, type_names: dict, colors: dict, indent: Optional[int] = None) -> None:
    """cell_detection.py:538-597 ``convert_geojson`` + dump: one MultiPolygon (segmentation) or MultiPoint (detection) feature per
    cell type, coordinates streamed from the columns."""
    import uuid
    types = sorted(set(cols.type.tolist()))
    if not types:
        with open(path, "w") as f:
            f.write(json.dumps([], indent=indent))
        return
    feats = [{"type": "Feature", "id": str(uuid.uuid4()),
              "geometry": {"type": "MultiPolygon" if polygons else "MultiPoint", "coordinates": _MARK},
              "properties": {"objectType": "annotation", "classification": {"name": type_names[t], "color": colors[t]}}} for t in types]
    # render the list once with one marker per feature, then cut it at the markers
    pieces = json.dumps(feats, indent=indent).split('"' + _MARK + '"')
    sections = []
    for k, t in enumerate(types):
        tail = pieces[k + 1] if k == len(types) - 1 else ""
        sections.append((pieces[k], tail, np.nonzero(cols.type == t)[0], JSON_POLYGONS if polygons else JSON_POINTS, 3))
    cols.export_json(path, sections, indent)


class LazyCellsJson(Mapping):
    """What ``process_wsi`` returns: the ``cells.json`` dictionary, whose ``cells`` list (one dict per cell, the reference's
    records) is only built when it is read -- the files are written from the columns (``.columns``)."""

    def __init__(self, header: dict, columns: CellColumns):
        self._header, self.columns, self._cells = dict(header), columns, None

    def __getitem__(self, key):
        if key == "cells":
            if self._cells is None:
                self._cells = self.columns.to_dicts()
            return self._cells
        return self._header[key]

    def __iter__(self):
        yield from self._header
        yield "cells"

    def __len__(self):
        return len(self._header) + 1
