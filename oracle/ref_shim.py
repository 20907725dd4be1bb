# -*- coding: utf-8 -*-
"""Import the UNMODIFIED reference from /root/reference in this container -- TEST INFRASTRUCTURE.

The reference needs ``skimage`` (not installed, no network). Only two names are touched on the
tile-inference path: ``skimage.segmentation.watershed`` (post_proc_cellvit.py:20,247) and
``skimage.draw.polygon`` (tools.py:21, unused on the path). They are provided through ``sys.modules``
stubs; ``watershed`` is bound to the oracle restatement (PARITY UNPINNED for that stage, see
oracle/postproc_oracle.py). /root/reference does not exist on the GPU box: nothing under ``-m gpu``,
``smoke()`` or bench.py calls this module.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("CELLVIT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "segmentation"))


def install_stubs(watershed_fn=None):
    if "skimage" in sys.modules and not getattr(sys.modules["skimage"], "_cvb_stub", False):
        return
    if watershed_fn is None:
        from oracle.postproc_oracle import watershed as watershed_fn
    sk = types.ModuleType("skimage"); sk._cvb_stub = True
    seg = types.ModuleType("skimage.segmentation"); seg.watershed = watershed_fn
    draw = types.ModuleType("skimage.draw")
    draw.polygon = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("skimage.draw stub"))
    sk.segmentation, sk.draw = seg, draw
    sys.modules.update({"skimage": sk, "skimage.segmentation": seg, "skimage.draw": draw})


def import_reference(watershed_fn=None):
    """Returns (cellvit_module, post_proc_module) of the reference."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    install_stubs(watershed_fn)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    saved_warn = warnings.warn  # the reference monkey-patches warnings.warn at import (post_proc_cellvit.py:30)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib
        cellvit = importlib.import_module("models.segmentation.cell_segmentation.cellvit")
        post = importlib.import_module("cell_segmentation.utils.post_proc_cellvit")
    warnings.warn = saved_warn
    if watershed_fn is not None:
        # the module binds `watershed` by name at import time (post_proc_cellvit.py:20): rebind it, so that a module that was
        # already imported with another flood (another test module ran first) calls the one asked for here
        post.watershed = watershed_fn
    return cellvit, post


def import_reference_cell_detection():
    """The reference's ``cell_segmentation/inference/cell_detection.py`` (unmodified), for pinning the WSI-level control flow
    (``CellPostProcessor``, position helpers). Its third-party imports that are not installed here are stubbed:

    * ``ujson`` -> ``json``; ``pandarallel`` -> ``DataFrame.parallel_apply = DataFrame.apply``;
    * ``shapely`` (Polygon / STRtree, ``requirements.txt:27``) -> a minimal stand-in whose geometry is the repo's own exact
      polygon-intersection restatement (cellvit_b200/wsi_merge.py) and whose ``STRtree.query`` returns envelope hits in
      insertion order. So a comparison through this shim pins the reference's *control flow* (edge-cell cleaning, greedy
      merge loop, index bookkeeping) -- the geometric predicate itself stays "parity unpinned" against real GEOS."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import json
    import numpy as np
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from cellvit_b200 import wsi_merge as wm

    uj = types.ModuleType("ujson")
    uj.dump, uj.dumps, uj.load, uj.loads = json.dump, json.dumps, json.load, json.loads
    pp = types.ModuleType("pandarallel")

    class _Pandarallel:
        @staticmethod
        def initialize(**kwargs):
            import pandas as pd
            pd.DataFrame.parallel_apply = pd.DataFrame.apply

    pp.pandarallel = _Pandarallel

    class _Area:
        def __init__(self, a):
            self.area = a

    class Polygon:
        def __init__(self, shell):
            self.pts = np.asarray(shell.pts if isinstance(shell, Polygon) else list(shell), dtype=np.float64).reshape(-1, 2)
            self.uid = None

        is_valid = True

        @property
        def area(self):
            return wm.polygon_area(self.pts)

        @property
        def bounds(self):
            return (self.pts[:, 0].min(), self.pts[:, 1].min(), self.pts[:, 0].max(), self.pts[:, 1].max())

        def intersection(self, other):
            return _Area(wm.polygon_intersection_area(self.pts, other.pts))

        def buffer(self, d):
            return self

    class MultiPolygon(list):
        pass

    class STRtree:
        def __init__(self, geoms):
            self.geoms = list(geoms)
            self.boxes = np.array([g.bounds for g in self.geoms], dtype=np.float64).reshape(-1, 4)

        def query(self, geom):
            b = geom.bounds
            hit = (self.boxes[:, 0] <= b[2]) & (b[0] <= self.boxes[:, 2]) & (self.boxes[:, 1] <= b[3]) & (b[1] <= self.boxes[:, 3])
            return [self.geoms[i] for i in np.nonzero(hit)[0]]

    class ShapelyDeprecationWarning(Warning):
        pass

    sh, st = types.ModuleType("shapely"), types.ModuleType("shapely.strtree")
    er, ge = types.ModuleType("shapely.errors"), types.ModuleType("shapely.geometry")
    st.STRtree, er.ShapelyDeprecationWarning, ge.Polygon, ge.MultiPolygon = STRtree, ShapelyDeprecationWarning, Polygon, MultiPolygon
    sh.strtree, sh.errors, sh.geometry, sh._cvb_stub = st, er, ge, True
    for name, mod in {"ujson": uj, "pandarallel": pp, "shapely": sh, "shapely.strtree": st, "shapely.errors": er,
                      "shapely.geometry": ge}.items():
        sys.modules.setdefault(name, mod)
    saved_warn = warnings.warn
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import importlib
        mod = importlib.import_module("cell_segmentation.inference.cell_detection")
    warnings.warn = saved_warn
    return mod
