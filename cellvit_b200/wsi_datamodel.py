# -*- coding: utf-8 -*-
"""Tile ingest for WSI inference: the on-disk layout the reference's preprocessing
writes and its inference reads.

Mirrors the parts of ``datamodel/wsi_datamodel.py:21-138`` (``WSI``), ``preprocessing/encoding/datasets/
patched_wsi_inference.py:15-87`` (``PatchedWSIInference``) and ``datamodel/graph_datamodel.py`` /
``cell_segmentation/datasets/cell_graph_datamodel.py`` (graph containers) that ``process_wsi`` touches:

    <patched_slide_path>/metadata.yaml          magnification, base_magnification, downsampling, patch_size,
                                                patch_overlap, label_map, ...
    <patched_slide_path>/patch_metadata.json    [{"<wsi>_<row>_<col>.png": {"row", "col", "metadata_path", ...}}, ...]
    <patched_slide_path>/patches/<name>.png     RGB tiles
    <patched_slide_path>/<metadata_path>        per-patch yaml (row, col, ...)
"""
from __future__ import annotations

import json
import logging
from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, List, Tuple, Union

import numpy as np
import torch
import yaml
from torch.utils.data import Dataset


@dataclass
class WSI:
    """One (already patched) whole-slide image; same fields as the reference dataclass."""

    name: str
    patient: str
    slide_path: Union[str, Path]
    patched_slide_path: Union[str, Path] = None
    embedding_name: Union[str, Path] = None
    label: Union[str, int, float, np.ndarray] = None
    logger: logging.Logger = None

    metadata: dict = field(init=False, repr=False)
    all_patch_metadata: dict = field(init=False, repr=False)
    patches_list: List = field(init=False, repr=False)
    patch_transform: Callable = field(init=False, repr=False)

    def __post_init__(self):
        self.slide_path = Path(self.slide_path).resolve()
        if self.patched_slide_path is not None:
            self.patched_slide_path = Path(self.patched_slide_path).resolve()
            with open(self.patched_slide_path / "metadata.yaml", "r") as f:
                self.metadata = yaml.safe_load(f)
            self.metadata["label_map_inverse"] = {v: k for k, v in self.metadata["label_map"].items()}
            with open(self.patched_slide_path / "patch_metadata.json", "r") as f:
                entries = json.load(f)
            self.patches_list = [str(next(iter(e))) for e in entries]
            self.all_patch_metadata = {str(next(iter(e))): e[next(iter(e))] for e in entries}
            self.patch_transform = None
        if self.logger is not None:
            self.logger.debug(repr(self))

    def load_patch_metadata(self, patch_name: str) -> dict:
        """Per-patch yaml (path relative to the patched slide folder) + ``name``."""
        with open(self.patched_slide_path / self.all_patch_metadata[patch_name]["metadata_path"], "r") as f:
            meta = yaml.safe_load(f)
        meta["name"] = patch_name
        return meta

    def set_patch_transform(self, transform: Callable) -> None:
        self.patch_transform = transform

    def process_patch_image(self, patch_name: str, transform: Callable = None) -> Tuple[torch.Tensor, dict]:
        from PIL import Image
        patch = Image.open(self.patched_slide_path / "patches" / patch_name)
        if transform:
            patch = transform(patch)
        return patch, self.load_patch_metadata(patch_name)

    def get_number_patches(self) -> int:
        return int(len(self.patches_list))


class PatchedWSIInference(Dataset):
    """One item per tile of one WSI: (transformed tile, patch metadata dict)."""

    def __init__(self, wsi_object: WSI, transform: Callable) -> None:
        assert isinstance(wsi_object, WSI), "Must be a WSI-object"
        assert wsi_object.patched_slide_path is not None, "Please provide a WSI that already has been patched into slices"
        self.transform = transform
        self.wsi_object = wsi_object

    def __getitem__(self, idx: int):
        return self.wsi_object.process_patch_image(patch_name=self.wsi_object.patches_list[idx], transform=self.transform)

    def __len__(self) -> int:
        return int(self.wsi_object.get_number_patches())

    @staticmethod
    def collate_batch(batch: List[Tuple]) -> Tuple[torch.Tensor, List[dict]]:
        patches, metadata = zip(*batch)
        return torch.stack(patches), list(metadata)


class InferenceTransform:
    """``T.Compose([T.ToTensor(), T.Normalize(mean, std)])`` (cell_detection.py:214-227) without torchvision:
    PIL / uint8 HWC -> float32 CHW in [0,1] -> (x - mean) / std. RGBA inputs keep their first three channels."""

    def __init__(self, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5), as_uint8: bool = False):
        """``as_uint8``: return the raw uint8 CHW tile instead; the tile pipeline then applies the same two operations on
        the device (bit-identical), which cuts the worker -> main -> pinned -> device traffic per tile from 12.6 MB to 3.1 MB."""
        self.mean = torch.tensor(mean, dtype=torch.float32).view(3, 1, 1)
        self.std = torch.tensor(std, dtype=torch.float32).view(3, 1, 1)
        self.as_uint8 = as_uint8

    def __call__(self, img) -> torch.Tensor:
        a = np.asarray(img)
        if a.ndim == 2:
            a = np.repeat(a[..., None], 3, axis=2)
        a = np.array(a[..., :3])  # writable copy
        x = torch.from_numpy(a).permute(2, 0, 1)
        if self.as_uint8:
            return x.contiguous()
        x = x.to(torch.float32).div(255.0)
        return (x - self.mean) / self.std


class _TorchSplit:
    """Pickles as the call ``torch.split(flat, lens)`` (only torch / builtins are needed to load it)."""

    def __init__(self, flat: torch.Tensor, lens: List[int]):
        self.flat, self.lens = flat, lens

    def __reduce__(self):
        return torch.split, (self.flat, self.lens)


class SplitTensorList(list):
    """``List[Tensor]`` whose elements are consecutive row blocks of ONE tensor (the contours of all cells of a slide).
    In memory it is an ordinary list of views; ``torch.save`` writes it as ``list(torch.split(flat, lens))`` -- one storage
    and one length list instead of one pickled tensor object per cell (70 us each: 4 s on a slide with 60,000 cells) --
    and ``torch.load`` gives back a plain ``list`` of tensors with the same values, as the reference stores
    (cell_detection.py:462-468)."""

    def __init__(self, flat: torch.Tensor, lens: List[int]):
        super().__init__(flat.split_with_sizes(lens) if len(lens) else [])
        self._flat, self._lens = flat, list(lens)

    def __reduce__(self):
        return list, (_TorchSplit(self._flat, self._lens),)


@dataclass
class GraphDataWSI:
    x: torch.Tensor
    positions: torch.Tensor
    metadata: dict


@dataclass
class CellGraphDataWSI(GraphDataWSI):
    contours: List[torch.Tensor]


def save_cell_graph(graph: CellGraphDataWSI, path) -> None:
    """``torch.save(graph, cells.pt)`` (cell_detection.py:462-468) with the class recorded under the REFERENCE's module path
    (``cell_segmentation.datasets.cell_graph_datamodel.CellGraphDataWSI``): a reference-side ``torch.load`` returns the
    reference's own dataclass, with the fields x / positions / contours / metadata it expects."""
    from . import _graph_pickle
    torch.save(graph, path, pickle_module=_graph_pickle)


def load_cell_graph(path, map_location="cpu"):
    """Reads a ``cells.pt`` written by this package or by the reference: the reference's dataclass where it is importable, this
    package's ``CellGraphDataWSI`` (same fields) otherwise."""
    from . import _graph_pickle
    return torch.load(path, map_location=map_location, pickle_module=_graph_pickle, weights_only=False)
