# -*- coding: utf-8 -*-
"""``pickle_module`` for ``torch.save`` / ``torch.load`` of ``cells.pt``.

The reference stores a ``cell_segmentation.datasets.cell_graph_datamodel.CellGraphDataWSI`` instance
(``cell_segmentation/inference/cell_detection.py:462-468``), and its downstream tooling ``torch.load``s exactly that. This
package's dataclass has the same fields but lives in ``cellvit_b200.wsi_datamodel``; pickled by its own name, the file could only
be read where this package is importable. The pickler below names the REFERENCE's class in the stream (a ``GLOBAL`` record is
just a module path and a name), so a reference-side consumer gets the reference's own dataclass back; the unpickler resolves
that name to the reference's class when it is importable and to this package's otherwise (``wsi_datamodel.load_cell_graph``).
"""
import pickle
from pickle import *  # noqa: F401,F403 -- torch expects the surface of the pickle module (load, dump, HIGHEST_PROTOCOL, ...)

REF_MODULE, REF_NAME = "cell_segmentation.datasets.cell_graph_datamodel", "CellGraphDataWSI"


class Pickler(pickle._Pickler):  # the pure-Python pickler: save_global can be overridden
    def save_global(self, obj, name=None):
        from .wsi_datamodel import CellGraphDataWSI
        if obj is CellGraphDataWSI:
            if self.proto >= 4:
                self.save(REF_MODULE)
                self.save(REF_NAME)
                self.write(pickle.STACK_GLOBAL)
            else:
                self.write(pickle.GLOBAL + REF_MODULE.encode("ascii") + b"\n" + REF_NAME.encode("ascii") + b"\n")
            self.memoize(obj)
            return
        super().save_global(obj, name)


class Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) == (REF_MODULE, REF_NAME):
            try:
                return super().find_class(module, name)      # the reference is importable: its own dataclass
            except (ImportError, AttributeError):
                from .wsi_datamodel import CellGraphDataWSI
                return CellGraphDataWSI
        return super().find_class(module, name)
