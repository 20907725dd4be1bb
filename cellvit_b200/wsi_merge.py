# -*- coding: utf-8 -*-
"""WSI-level cell bookkeeping of ``cell_segmentation/inference/cell_detection.py``:
per-cell position codes (:771-902) and the removal of cells that overlapping tiles detected twice
(``CellPostProcessor`` :600-767).

The reference leans on pandas ``iterrows`` and shapely (``Polygon``, ``STRtree``); shapely is not installed here, and
the pairwise overlap areas are the expensive part, so they are computed on the GPU (``cvb_polygon_overlap``,
csrc/wsi_merge.cu) for the envelope-intersecting pairs found by a grid hash on the host. The greedy, order-dependent
selection loop of ``_remove_overlap`` is kept as is (it is cheap once the overlap graph exists).
PARITY NOTE: intersection areas are exact for simple polygons (even-odd rule); GEOS's ``buffer(0)`` repair of
self-touching contours (:700-713) is not reproduced, and the STRtree hit order (which only breaks ties between equally
large candidates) is replaced by ascending cell index.
"""
from __future__ import annotations

import ctypes as C
from collections import deque
from typing import Dict, List, Sequence, Tuple

import numpy as np

# ------------------------------------------------------------------------------------------------ position codes


def get_cell_position(bbox: np.ndarray, patch_size: int = 1024) -> List[int]:
    """cell_detection.py:789-819 -- [top, right, down, left] flags of a cell touching the tile border."""
    return [int(bbox[0, 0] == 0), int(bbox[1, 1] == patch_size), int(bbox[1, 0] == patch_size), int(bbox[0, 1] == 0)]


def get_cell_position_marging(bbox: np.ndarray, patch_size: int = 1024, margin: int = 64) -> int:
    """cell_detection.py:822-874 -- 0 = mid, 1..8 clockwise from top-left for cells inside the overlap margin."""
    return int(cell_status_batch(np.asarray(bbox)[None], patch_size, margin)[0])


def cell_status_batch(bbox: np.ndarray, patch_size: int = 1024, margin: int = 64) -> np.ndarray:
    """Vectorised ``get_cell_position_marging`` over bboxes [n,2,2] = [[rmin,cmin],[rmax,cmax]]."""
    bbox = np.asarray(bbox)
    rmin, cmin, rmax, cmax = bbox[:, 0, 0], bbox[:, 0, 1], bbox[:, 1, 0], bbox[:, 1, 1]
    hi = patch_size - margin
    in_margin = (bbox.reshape(len(bbox), -1).max(1) > hi) | (bbox.reshape(len(bbox), -1).min(1) < margin)
    top, left, right, bottom = rmin < margin, cmin < margin, cmax > hi, rmax > hi
    status = np.select(
        [top & left, top & right, top, right & bottom, right, bottom & left, bottom, left],
        [1, 3, 2, 5, 4, 7, 6, 8], default=0)
    return np.where(in_margin, status, 0).astype(np.int64)


_EDGE_PATCHES = {
    (1, 0, 0, 0): ((-1, 0),),
    (1, 1, 0, 0): ((-1, 0), (-1, 1), (0, 1)),
    (0, 1, 0, 0): ((0, 1),),
    (0, 1, 1, 0): ((0, 1), (1, 1), (1, 0)),
    (0, 0, 1, 0): ((1, 0),),
    (0, 0, 1, 1): ((1, 0), (1, -1), (0, -1)),
    (0, 0, 0, 1): ((0, -1),),
    (1, 0, 0, 1): ((0, -1), (-1, -1), (-1, 0)),
}


def get_edge_patch(position: Sequence[int], row: int, col: int):
    """cell_detection.py:877-902 -- neighbour tiles a border-touching cell continues into (None for other codes)."""
    offs = _EDGE_PATCHES.get(tuple(int(p) for p in position))
    return None if offs is None else [[row + dr, col + dc] for dr, dc in offs]


# ------------------------------------------------------------------------------------------------ polygon overlap


def polygon_area(p: np.ndarray) -> float:
    p = np.asarray(p, dtype=np.float64)
    if len(p) < 3:
        return 0.0
    x, y = p[:, 0] - p[0, 0], p[:, 1] - p[0, 1]
    return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(np.roll(x, -1), y)))


def polygon_intersection_area(a: np.ndarray, b: np.ndarray) -> float:
    """Host version of the device algorithm (csrc/wsi_merge.cu) -- used for the pairs the kernel flags and by tests."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if len(a) < 3 or len(b) < 3:
        return 0.0
    o = a[0].copy()
    a, b = a - o, b - o
    a2, b2 = np.roll(a, -1, axis=0), np.roll(b, -1, axis=0)
    xs = [a[:, 0], b[:, 0]]
    r, s = a2 - a, b2 - b
    den = r[:, None, 0] * s[None, :, 1] - r[:, None, 1] * s[None, :, 0]
    qp = b[None, :, :] - a[:, None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (qp[..., 0] * s[None, :, 1] - qp[..., 1] * s[None, :, 0]) / den
        u = (qp[..., 0] * r[:, None, 1] - qp[..., 1] * r[:, None, 0]) / den
    hit = (den != 0) & (t >= 0) & (t <= 1) & (u >= 0) & (u <= 1)
    ii, _ = np.nonzero(hit)
    xs.append(a[ii, 0] + t[hit] * r[ii, 0])
    xs = np.unique(np.concatenate(xs))
    area = 0.0

    def chords(p, p2, xm):
        m = (p[:, 0] < xm) != (p2[:, 0] < xm)
        return np.sort(p[m, 1] + (xm - p[m, 0]) * (p2[m, 1] - p[m, 1]) / (p2[m, 0] - p[m, 0]))

    for x0, x1 in zip(xs[:-1], xs[1:]):
        xm = 0.5 * (x0 + x1)
        if not (x0 < xm < x1):
            continue
        ya, yb = chords(a, a2, xm), chords(b, b2, xm)
        i = j = 0
        ln = 0.0
        while i + 1 < len(ya) and j + 1 < len(yb):
            lo, hi = max(ya[i], yb[j]), min(ya[i + 1], yb[j + 1])
            if hi > lo:
                ln += hi - lo
            if ya[i + 1] < yb[j + 1]:
                i += 2
            else:
                j += 2
        area += ln * (x1 - x0)
    return float(area)


def envelope_pairs(boxes: np.ndarray, cell: float = 128.0) -> np.ndarray:
    """All index pairs (i < j) whose axis-aligned envelopes [xmin,ymin,xmax,ymax] intersect (touching counts, as
    for shapely's STRtree.query) -- grid hash + exact filter. Returns int32 [n_pairs, 2] sorted lexicographically."""
    n = len(boxes)
    if n < 2:
        return np.zeros((0, 2), np.int32)
    b = np.asarray(boxes, dtype=np.float64)
    g0 = np.floor(b[:, :2] / cell).astype(np.int64)
    g1 = np.floor(b[:, 2:] / cell).astype(np.int64)
    span = g1 - g0 + 1
    keys, ids = [], []
    for dx in range(int(span[:, 0].max())):
        for dy in range(int(span[:, 1].max())):
            m = (span[:, 0] > dx) & (span[:, 1] > dy)
            if m.any():
                keys.append(((g0[m, 0] + dx) << 32) ^ ((g0[m, 1] + dy) & 0xFFFFFFFF))
                ids.append(np.nonzero(m)[0])
    keys, ids = np.concatenate(keys), np.concatenate(ids)
    order = np.lexsort((ids, keys))
    keys, ids = keys[order], ids[order]
    # all (i < j) pairs inside each run of equal keys, without a Python loop over the runs: element k pairs with element k + d of
    # the sorted arrays for d = 1, 2, ... while both lie in the same run (runs are short: a grid cell holds a handful of boxes)
    out = []
    d = 1
    while d < len(keys):
        same = keys[d:] == keys[:-d]
        if not same.any():
            break
        out.append(np.stack([ids[:-d][same], ids[d:][same]], 1))     # ids ascend inside a run, so column 0 < column 1
        d += 1
    if not out:
        return np.zeros((0, 2), np.int32)
    p = np.unique(np.concatenate(out), axis=0)
    bi, bj = b[p[:, 0]], b[p[:, 1]]
    keep = (bi[:, 0] <= bj[:, 2]) & (bj[:, 0] <= bi[:, 2]) & (bi[:, 1] <= bj[:, 3]) & (bj[:, 1] <= bi[:, 3])
    return p[keep].astype(np.int32)


class _ContourList:
    """Read-only sequence of contours [k_i, 2] stored as ONE array + offsets (what the columnar cell store holds): element access
    is a slice, and ``overlap_areas`` ships ``flat`` / ``offsets`` to the device as they are instead of re-concatenating."""

    def __init__(self, flat: np.ndarray, offsets: np.ndarray):
        self.flat, self.offsets = flat, offsets

    def __len__(self):
        return len(self.offsets) - 1

    def __getitem__(self, i):
        if i < 0:
            i += len(self)
        return self.flat[self.offsets[i]:self.offsets[i + 1]]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def flatten_contours(contours) -> Tuple[np.ndarray, np.ndarray]:
    """(offsets int32 [n+1], points fp64 [sum k_i, 2]) of a list of contours or of a ``_ContourList`` (which is already flat)."""
    if isinstance(contours, _ContourList):
        return contours.offsets.astype(np.int32), np.ascontiguousarray(contours.flat, dtype=np.float64).reshape(-1, 2)
    off = np.zeros(len(contours) + 1, np.int32)
    off[1:] = np.cumsum([len(c) for c in contours])
    pts = np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1, 2) for c in contours]) if off[-1] else np.zeros((0, 2))
    return off, pts


def overlap_areas(contours: List[np.ndarray], pairs: np.ndarray, device=None) -> Tuple[np.ndarray, np.ndarray]:
    """(polygon areas [n], intersection areas [n_pairs]) through cvb_polygon_overlap on ``device``."""
    import torch
    from . import _lib as L
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    n = len(contours)
    if n == 0:
        return np.zeros(0), np.zeros(0)
    off, pts = flatten_contours(contours)
    with torch.cuda.device(device):
        d_pts = torch.from_numpy(np.ascontiguousarray(pts)).to(device)
        d_off = torch.from_numpy(off).to(device)
        d_pairs = torch.from_numpy(np.ascontiguousarray(pairs, dtype=np.int32)).to(device)
        d_area = torch.empty(n, dtype=torch.float64, device=device)
        d_inter = torch.empty(max(len(pairs), 1), dtype=torch.float64, device=device)
        L.check(L.lib().cvb_polygon_overlap(L.ptr(d_pts), L.ptr(d_off), n, L.ptr(d_pairs), len(pairs), L.ptr(d_area), L.ptr(d_inter),
                                            L.stream_ptr()), "cvb_polygon_overlap")
        area, inter = d_area.cpu().numpy(), d_inter.cpu().numpy()[:len(pairs)]
    for k in np.nonzero(inter < 0)[0]:  # contours beyond the kernel's limits
        inter[k] = polygon_intersection_area(contours[pairs[k, 0]], contours[pairs[k, 1]])
    return area, inter


# ------------------------------------------------------------------------------------------------ CellPostProcessor


class CellPostProcessor:
    """cell_detection.py:600-767. ``cell_list`` entries need: contour, cell_status, edge_position, edge_information
    (for edge cells), patch_coordinates [row, col] -- or ``cell_list`` is the columnar store of the slide
    (``wsi_records.CellColumns``), which holds the same information without per-cell objects.
    ``post_process_cells`` returns the sorted indices to keep."""

    def __init__(self, cell_list, logger=None, device=None, overlap_fn=None) -> None:
        self.logger = logger
        self.device = device
        self.overlap_fn = overlap_fn or overlap_areas
        self.cols = cell_list if hasattr(cell_list, "contour_off") else None
        self.cells = None if self.cols is not None else cell_list
        if self.cols is not None:
            status = self.cols.status
        else:
            status = np.array([c["cell_status"] for c in cell_list], dtype=np.int64) if cell_list else np.zeros(0, np.int64)
        self.mid_idx = np.nonzero(status == 0)[0]
        self.margin_idx = np.nonzero(status != 0)[0]

    def _log(self, msg):
        if self.logger is not None:
            self.logger.info(msg)

    def post_process_cells(self) -> List[int]:
        self._log("Finding edge-cells for merging")
        cleaned = self._clean_edge_cells()
        self._log("Removal of cells detected multiple times")
        cleaned = self._remove_overlap(cleaned)
        return sorted(set(self.mid_idx.tolist()) | set(cleaned))

    def _clean_edge_cells(self) -> List[int]:
        """:640-672 -- margin cells that do not touch the border, plus border cells whose (first) neighbour tile has no
        margin cell at all (i.e. was not processed / is empty there)."""
        if self.cols is not None:
            c = self.cols
            existing = set(map(tuple, c.patch[self.margin_idx].tolist()))
            edge = c.edge[self.margin_idx] != 0
            keep = self.margin_idx[~edge].tolist()
            keep += [int(i) for i in self.margin_idx[edge] if c.first_edge_patch(int(i)) not in existing]
            return sorted(keep)
        existing = {tuple(self.cells[i]["patch_coordinates"]) for i in self.margin_idx}
        keep = []
        for i in self.margin_idx:
            c = self.cells[i]
            if not c["edge_position"]:
                keep.append(int(i))
            elif tuple(c["edge_information"]["edge_patches"][0]) not in existing:
                keep.append(int(i))
        return sorted(keep)

    def _remove_overlap(self, cleaned: List[int]) -> List[int]:
        """:674-767 -- up to 20 rounds of: walk the cells in index order; a cell with overlapping partners (intersection
        > 1 % of either area) is replaced by the largest of those partners, partners are consumed."""
        if len(cleaned) < 2:
            return list(cleaned)
        if self.cols is not None:
            # contours and their envelopes straight from the columns: one gather + four segmented reductions
            idx = np.asarray(cleaned, dtype=np.int64)
            off = self.cols.contour_off
            lens = (off[1:] - off[:-1])[idx]
            seg = np.zeros(len(idx) + 1, np.int64)
            np.cumsum(lens, out=seg[1:])
            flat = self.cols.contour_pts[np.repeat(off[:-1][idx] - seg[:-1], lens) + np.arange(int(seg[-1]), dtype=np.int64)].astype(np.float64)
            contours = _ContourList(flat, seg)
            boxes = np.tile(np.array([0.0, 0.0, -1.0, -1.0]), (len(idx), 1))
            ne = lens > 0
            if ne.any():
                st = seg[:-1][ne]   # (reduceat needs non-empty segments: a start index repeated for an empty one would read its neighbour)
                boxes[ne] = np.stack([np.minimum.reduceat(flat[:, 0], st), np.minimum.reduceat(flat[:, 1], st),
                                      np.maximum.reduceat(flat[:, 0], st), np.maximum.reduceat(flat[:, 1], st)], 1)
        else:
            contours = [np.asarray(self.cells[i]["contour"], dtype=np.float64).reshape(-1, 2) for i in cleaned]
            boxes = np.array([[c[:, 0].min(), c[:, 1].min(), c[:, 0].max(), c[:, 1].max()] if len(c) else [0, 0, -1, -1] for c in contours])
        pairs = envelope_pairs(boxes)
        area, inter = self.overlap_fn(contours, pairs, self.device) if self.overlap_fn is overlap_areas else self.overlap_fn(contours, pairs)
        with np.errstate(divide="ignore", invalid="ignore"):
            strong = (inter / area[pairs[:, 0]] > 0.01) | (inter / area[pairs[:, 1]] > 0.01) if len(pairs) else np.zeros(0, bool)
        nbrs: Dict[int, List[int]] = {}
        for (i, j) in pairs[strong]:
            nbrs.setdefault(int(i), []).append(int(j))
            nbrs.setdefault(int(j), []).append(int(i))
        # Cells without any strong partner take no part in the loop below: they are never consumed (only partners are) and always
        # keep themselves, in every round -- so the rounds only walk the cells that have partners (a few per cent of a slide).
        isolated = np.ones(len(cleaned), bool)
        isolated[pairs[strong].ravel()] = False
        alive = sorted(nbrs)               # local ids, ascending == ascending cell index
        for iteration in range(20):
            alive_set = set(alive)
            merged, iterated, overlaps = deque(), set(), 0
            for q in alive:
                if q in iterated:
                    continue
                sub = [p for p in sorted(nbrs.get(q, ())) if p in alive_set and p not in iterated]
                if sub:
                    overlaps += len(sub)
                    iterated.update(sub)
                    merged.append(sub[int(np.argmax([area[p] for p in sub]))])
                else:
                    merged.append(q)
                iterated.add(q)
            self._log(f"Iteration {iteration}: Found overlap of # cells: {overlaps}")
            if overlaps == 0:
                self._log("Found all overlapping cells")
                break
            alive = sorted(set(merged))
        keep = isolated
        keep[alive] = True
        return [cleaned[k] for k in np.nonzero(keep)[0]]
