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
    return cellvit, post
