# -*- coding: utf-8 -*-
"""Drop-in mirrors of the reference model classes for the tile-inference path.

``CellViT`` / ``CellViT256`` / ``CellViTSAM`` keep the reference constructor signatures, public attributes,
``state_dict`` keys (so reference checkpoints ``load_state_dict(strict=True)``), ``forward`` contract and the
``calculate_instance_map`` / ``generate_instance_nuclei_map`` helpers
(models/segmentation/cell_segmentation/cellvit.py:26-151, 153-210, 332-414, 428-494, 496-667).

The forward itself runs entirely in libcellvit_b200.so (``cvb_forward``): there is no PyTorch/CPU fallback --
calling it without the CUDA library or on CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import packing, weights


class ModelDesc(C.Structure):
    _fields_ = [("sam", C.c_int), ("embed_dim", C.c_int), ("depth", C.c_int), ("num_heads", C.c_int),
                ("window_size", C.c_int), ("n_global", C.c_int), ("global_idx", C.c_int * 8), ("extract", C.c_int * 4),
                ("n_np_out", C.c_int), ("n_nt", C.c_int), ("n_tissue", C.c_int), ("skip11", C.c_int), ("skip12", C.c_int),
                ("bott_pad", C.c_int), ("shared_decoder", C.c_int)]


class _Node(nn.Module):
    """Anonymous container so that dotted reference keys map onto a module tree."""


def _build_tree(root: nn.Module, spec, seed: int):
    for key, (shape, kind) in spec.items():
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, _Node())
            mod = mod._modules[p]
        t = weights.synth_tensor(key, shape, kind, seed)
        if kind in ("bn_mean", "bn_var", "bn_nbt"):
            mod.register_buffer(parts[-1], t)
        else:
            mod.register_parameter(parts[-1], nn.Parameter(t))


class CellViT(nn.Module):
    """cellvit.py:26-151 -- ViT encoder + three U-Net style decoder branches (NP, HV, NT) + tissue logits."""

    def __init__(self, num_nuclei_classes: int, num_tissue_classes: int, embed_dim: int, input_channels: int,
                 depth: int, num_heads: int, extract_layers: List, mlp_ratio: float = 4, qkv_bias: bool = True,
                 drop_rate: float = 0, attn_drop_rate: float = 0, drop_path_rate: float = 0,
                 regression_loss: bool = False, _arch: str = "ViT256", _shared: bool = False):
        super().__init__()
        assert len(extract_layers) == 4, "Please provide 4 layers for skip connections"
        if input_channels != 3 or mlp_ratio != 4 or not qkv_bias:
            raise NotImplementedError("cellvit_b200 supports input_channels=3, mlp_ratio=4, qkv_bias=True")
        self.patch_size = 16
        self.num_tissue_classes = num_tissue_classes
        self.num_nuclei_classes = num_nuclei_classes
        self.embed_dim = embed_dim
        self.input_channels = input_channels
        self.depth = depth
        self.num_heads = num_heads
        self.mlp_ratio = mlp_ratio
        self.qkv_bias = qkv_bias
        self.extract_layers = extract_layers
        self.drop_rate = drop_rate
        self.attn_drop_rate = attn_drop_rate
        self.drop_path_rate = drop_path_rate
        self.regression_loss = regression_loss
        self.skip_dim_11, self.skip_dim_12, self.bottleneck_dim = weights.decoder_dims(embed_dim)
        self.branches_output = {"nuclei_binary_map": 2 + (2 if regression_loss else 0), "hv_map": 2,
                                "nuclei_type_maps": num_nuclei_classes}
        self._arch = _arch
        self._shared = bool(_shared)   # the *Shared variants (cellvit_shared.py): one decoder trunk, three 1x1 heads
        self._sam = _arch != "ViT256"
        self._global_idx = tuple(weights.SAM_CFG[_arch]["global_idx"]) if self._sam else ()
        spec = weights.state_spec(_arch, num_nuclei_classes, num_tissue_classes, regression_loss,
                                  embed_dim=embed_dim, depth=depth, num_heads=num_heads, shared=self._shared)
        _build_tree(self, spec, seed=int(torch.initial_seed() & 0xFFFF))
        self._handle = None
        self._packed = {}        # name -> device tensor (kept alive for the C side)
        self._packed_key = None  # (device, pack epoch, hash of the per-tensor (address, version) pairs)
        self._pack_epoch = 0
        self._size_key = None
        self._ws = None
        self._capture_stream = None
        self._graphs = {}        # forward_graphed: (shape, tokens, slot, device) -> (graph, static input, static outputs)

    # ------------------------------------------------------------------ engine plumbing
    def _cfg(self):
        return dict(sam=self._sam, embed_dim=self.embed_dim, depth=self.depth, num_heads=self.num_heads, window=14,
                    global_idx=self._global_idx, skip11=self.skip_dim_11, skip12=self.skip_dim_12, bott=self.bottleneck_dim,
                    shared=self._shared)

    def _ensure_handle(self):
        if self._handle is None:
            d = ModelDesc(sam=int(self._sam), embed_dim=self.embed_dim, depth=self.depth, num_heads=self.num_heads,
                          window_size=14 if self._sam else 0, n_global=len(self._global_idx),
                          n_np_out=self.branches_output["nuclei_binary_map"], n_nt=self.num_nuclei_classes,
                          n_tissue=self.num_tissue_classes, skip11=self.skip_dim_11, skip12=self.skip_dim_12,
                          bott_pad=packing.pad64(self.bottleneck_dim), shared_decoder=int(self._shared))
            for i, g in enumerate(self._global_idx):
                d.global_idx[i] = g
            for i, e in enumerate(self.extract_layers):
                d.extract[i] = int(e)
            h = C.c_void_p()
            L.check(L.lib().cvb_model_create(C.byref(d), C.byref(h)), "cvb_model_create")
            self._handle = h
        return self._handle

    def _register(self, tensors):
        lib = L.lib()
        for name, t in tensors.items():
            self._packed[name] = t
            L.check(lib.cvb_model_set_param(self._handle, name.encode(), C.c_void_p(t.data_ptr())), "cvb_model_set_param")

    def _ensure_packed(self, device, h, w):
        self._ensure_handle()
        # one (address, version) pair per tensor: sums / xors of them can collide, and an in-place write through ``.data``
        # does not bump the version at all -- callers that do that (``broadcast_weights``) call ``invalidate_packed()``
        key = (str(device), self._pack_epoch,
               hash(tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))))
        if key != self._packed_key:  # weights were (re)loaded, updated in place or moved: repack once
            sd_dev = {k: v.to(device) for k, v in self.state_dict().items()}
            self._register(packing.pack_static(sd_dev, self._cfg()))
            self._packed_key, self._size_key = key, None
            self._graphs = {}  # captured graphs hold the old tensors' addresses
        if self._size_key != (h, w):
            sd_dev = {k: v.to(device) for k, v in self.state_dict().items()
                      if k.startswith("encoder.pos_embed") or "rel_pos" in k or "cls_token" in k}
            self._register(packing.pack_for_size(sd_dev, self._cfg(), h, w))
            self._size_key = (h, w)
            self._graphs = {}

    def invalidate_packed(self) -> None:
        """Force a repack of the kernel-layout weights (and drop captured CUDA graphs) at the next forward: call after
        writing parameters through ``.data`` or any other route that leaves ``Tensor._version`` unchanged."""
        self._pack_epoch += 1

    def _load_from_state_dict(self, *args, **kwargs):
        self._pack_epoch += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def set_engine_option(self, name: str, value: int) -> None:
        """Per-model engine option (``cvb_model_set_option``), e.g. ``("attention_tc", 0)`` to run the mma.sync attention
        kernels instead of the tcgen05 ones. Captured CUDA graphs are dropped."""
        L.check(L.lib().cvb_model_set_option(self._ensure_handle(), name.encode(), int(value)), "cvb_model_set_option")
        self._graphs = {}

    def __del__(self):
        try:
            if self._handle is not None:
                L.lib().cvb_model_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ reference API
    def forward(self, x: torch.Tensor, retrieve_tokens: bool = False, argmax_maps: bool = False) -> dict:
        """cellvit.py:153-210 / :586-644. Raw logits, fp32, on x.device. ``argmax_maps`` (not in the reference): also return
        ``nuclei_binary_argmax`` / ``nuclei_type_argmax`` uint8 [B,H,W], the arg-max planes the post-processing consumes,
        written by the fused head epilogue (``cvb_forward_argmax``)."""
        assert x.shape[-2] % self.patch_size == 0, "Input images must be divisible by the patch size"
        assert x.shape[-1] % self.patch_size == 0, "Input images must be divisible by the patch size"
        if not x.is_cuda:
            raise RuntimeError("cellvit_b200 has no CPU path: move the input to a CUDA device")
        B, Cin, H, W = x.shape
        assert Cin == 3
        x = x.contiguous().float()
        h, w = H // 16, W // 16
        with torch.cuda.device(x.device):
            self._ensure_packed(x.device, h, w)
            lib = L.lib()
            need = C.c_size_t()
            L.check(lib.cvb_model_workspace_bytes(self._handle, B, H, W, C.byref(need)), "cvb_model_workspace_bytes")
            if self._ws is None or self._ws.numel() < need.value or self._ws.device != x.device:
                self._ws = torch.empty(need.value, dtype=torch.uint8, device=x.device)
                self._graphs = {}  # captured graphs hold the old workspace address
            n_np = self.branches_output["nuclei_binary_map"]
            o_np = torch.empty(B, n_np, H, W, device=x.device)
            o_hv = torch.empty(B, 2, H, W, device=x.device)
            o_nt = torch.empty(B, self.num_nuclei_classes, H, W, device=x.device)
            o_ti = torch.empty(B, self.num_tissue_classes, device=x.device)
            o_tok = torch.empty(B, self.embed_dim, h, w, device=x.device) if retrieve_tokens else None
            if argmax_maps:
                a_np = torch.empty(B, H, W, dtype=torch.uint8, device=x.device)
                a_nt = torch.empty(B, H, W, dtype=torch.uint8, device=x.device)
                L.check(lib.cvb_forward_argmax(self._handle, L.ptr(x), B, H, W, L.ptr(o_np), L.ptr(o_hv), L.ptr(o_nt), L.ptr(o_ti),
                                               L.ptr(o_tok), L.ptr(a_np), L.ptr(a_nt), L.ptr(self._ws), C.c_size_t(self._ws.numel()),
                                               L.stream_ptr()), "cvb_forward_argmax")
            else:
                L.check(lib.cvb_forward(self._handle, L.ptr(x), B, H, W, L.ptr(o_np), L.ptr(o_hv), L.ptr(o_nt), L.ptr(o_ti),
                                        L.ptr(o_tok), L.ptr(self._ws), C.c_size_t(self._ws.numel()), L.stream_ptr()), "cvb_forward")
        out = {"tissue_types": o_ti}
        if argmax_maps:
            out["nuclei_binary_argmax"], out["nuclei_type_argmax"] = a_np, a_nt
        if self.regression_loss:
            out["nuclei_binary_map"], out["regression_map"] = o_np[:, :2], o_np[:, 2:]
        else:
            out["nuclei_binary_map"] = o_np
        out["hv_map"] = o_hv
        out["nuclei_type_map"] = o_nt
        if retrieve_tokens:
            out["tokens"] = o_tok
        return out

    def forward_graphed(self, x: torch.Tensor, retrieve_tokens: bool = True, slot: int = 0, argmax_maps: bool = False) -> dict:
        """``forward`` replayed from a CUDA graph (one per input shape and ``slot``): the ~290 launches of a SAM-H forward
        become one graph launch, which removes the 2 us of idle time at every kernel boundary. The returned tensors are
        STATIC: the next call with the same ``slot`` overwrites them (the tile pipeline alternates two slots and has
        consumed a slot's outputs before it is replayed). ``x`` may be a host tensor (pinned: asynchronous H2D straight
        into the graph's input buffer)."""
        assert x.shape[-2] % self.patch_size == 0, "Input images must be divisible by the patch size"
        assert x.shape[-1] % self.patch_size == 0, "Input images must be divisible by the patch size"
        dev = x.device if x.is_cuda else next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("cellvit_b200 has no CPU path: move the model to a CUDA device")
        graph, static_x, out = self.graph_slot(tuple(x.shape), retrieve_tokens, slot, dev, argmax_maps)
        static_x.copy_(x, non_blocking=True)
        graph.replay()
        return out

    def graph_slot(self, shape, retrieve_tokens: bool, slot: int, dev, argmax_maps: bool = False):
        """(CUDA graph, static input tensor, static output dict) of ``forward`` for one input shape and pipeline slot;
        captured on first use. Repacking the weights (new checkpoint, other tile size) drops all captured graphs."""
        B, _, H, W = shape
        with torch.cuda.device(dev):
            self._ensure_packed(dev, H // 16, W // 16)
            key = (tuple(shape), bool(retrieve_tokens), int(slot), str(dev), bool(argmax_maps))
            g = self._graphs.get(key)
            if g is None:
                static_x = torch.zeros(tuple(shape), dtype=torch.float32, device=dev)
                with torch.no_grad():
                    self.forward(static_x, retrieve_tokens, argmax_maps)  # eager warm-up: sizes the workspace, sets kernel attributes
                    torch.cuda.synchronize(dev)
                    graph = torch.cuda.CUDAGraph()
                    # Captured on a HIGH-priority stream: kernel nodes keep the priority of the stream they were captured on,
                    # so the forward's CTAs are placed before those of the post-processing stream (default = lowest
                    # priority) whenever both have work pending -- the post-processing fills what the forward leaves free
                    if self._capture_stream is None or self._capture_stream.device != dev:
                        self._capture_stream = torch.cuda.Stream(dev, priority=-1)
                    with torch.cuda.graph(graph, stream=self._capture_stream):
                        out = self.forward(static_x, retrieve_tokens, argmax_maps)
                g = self._graphs[key] = (graph, static_x, out)
        return g

    def calculate_instance_map(self, predictions: OrderedDict, magnification=40) -> Tuple[torch.Tensor, List[dict]]:
        """cellvit.py:332-383. ``predictions`` hold post-softmax NP/NT maps [B,C,H,W] and the HV map (only the argmax
        of NP/NT is used). Returns (float32 CPU tensor [B,H,W], list of per-tile instance dicts)."""
        from .post_proc_cellvit import DetectionCellPostProcessor
        proc = DetectionCellPostProcessor(nr_types=self.num_nuclei_classes, magnification=magnification, gt=False)
        labels, dicts = proc.post_process_batch(predictions["nuclei_binary_map"], predictions["hv_map"],
                                                predictions["nuclei_type_map"])
        return torch.Tensor(labels.cpu().numpy()).type(torch.float32), dicts

    def generate_instance_nuclei_map(self, instance_maps: torch.Tensor, type_preds: List[dict]) -> torch.Tensor:
        """cellvit.py:385-414 -- [B, num_nuclei_classes, H, W] float32: plane ``type`` holds the instance id on the pixels of
        every instance listed in ``type_preds[b]``; instances without a dict entry stay 0. One id -> type table lookup and
        one scatter per tile instead of the reference's full-image compare per instance (same result)."""
        batch_size, hh, ww = instance_maps.shape
        out = torch.zeros((batch_size, self.num_nuclei_classes, hh, ww))
        for i in range(batch_size):
            lab = instance_maps[i].to("cpu", torch.int64)
            type_of = torch.full((int(lab.max()) + 2,), -1, dtype=torch.int64)
            for nuclei, spec in type_preds[i].items():
                if 0 <= int(nuclei) < len(type_of):
                    type_of[int(nuclei)] = int(spec["type"])
            plane = type_of[lab.clamp(min=0)]
            plane[lab < 0] = -1
            listed = plane >= 0
            out[i].scatter_(0, plane.clamp(min=0).unsqueeze(0), (lab * listed).to(torch.float32).unsqueeze(0))
        return out

    def freeze_encoder(self):
        """cellvit.py:416-420 (the tissue head stays trainable)."""
        for name, p in self._modules["encoder"].named_parameters():
            if name.split(".")[0] != "head":
                p.requires_grad = False

    def unfreeze_encoder(self):
        for p in self._modules["encoder"].parameters():
            p.requires_grad = True


class CellViT256(CellViT):
    """cellvit.py:428-494 -- HIPT ViT-S/16 backbone (embed 384, depth 12, 6 heads, skips after 3/6/9/12)."""

    def __init__(self, model256_path, num_nuclei_classes: int, num_tissue_classes: int, drop_rate: float = 0,
                 attn_drop_rate: float = 0, drop_path_rate: float = 0, regression_loss: bool = False):
        self.patch_size = 16
        self.model256_path = model256_path
        super().__init__(num_nuclei_classes=num_nuclei_classes, num_tissue_classes=num_tissue_classes, embed_dim=384,
                         input_channels=3, depth=12, num_heads=6, extract_layers=[3, 6, 9, 12], mlp_ratio=4, qkv_bias=True,
                         drop_rate=drop_rate, attn_drop_rate=attn_drop_rate, drop_path_rate=drop_path_rate,
                         regression_loss=regression_loss, _arch="ViT256")

    def load_pretrained_encoder(self, model256_path: str):
        """cellvit.py:483-494: HIPT checkpoint, key 'teacher', strip 'module.' / 'backbone.' prefixes."""
        state_dict = torch.load(str(model256_path), map_location="cpu")["teacher"]
        state_dict = {k.replace("module.", ""): v for k, v in state_dict.items()}
        state_dict = {k.replace("backbone.", ""): v for k, v in state_dict.items()}
        own = self._modules["encoder"].state_dict()
        msg = self._modules["encoder"].load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=False)
        print(f"Loading checkpoint: {msg}")


class CellViTSAM(CellViT):
    """cellvit.py:496-667 -- SAM ViTDet backbone (SAM-B / SAM-L / SAM-H)."""

    def __init__(self, model_path, num_nuclei_classes: int, num_tissue_classes: int, vit_structure, drop_rate: float = 0,
                 regression_loss: bool = False):
        if vit_structure.upper() not in weights.SAM_CFG:
            raise NotImplementedError("Unknown ViT-SAM backbone structure")
        cfg = weights.SAM_CFG[vit_structure.upper()]
        self.model_path = model_path
        super().__init__(num_nuclei_classes=num_nuclei_classes, num_tissue_classes=num_tissue_classes,
                         embed_dim=cfg["embed_dim"], input_channels=3, depth=cfg["depth"], num_heads=cfg["num_heads"],
                         extract_layers=list(cfg["extract"]), mlp_ratio=4, qkv_bias=True, drop_rate=drop_rate,
                         attn_drop_rate=0, drop_path_rate=0, regression_loss=regression_loss, _arch=vit_structure.upper())
        self.prompt_embed_dim = 256
        self.encoder_global_attn_indexes = list(cfg["global_idx"])

    def load_pretrained_encoder(self, model_path):
        """cellvit.py:574-584: SAM image-encoder weights, strict=False."""
        state_dict = torch.load(str(model_path), map_location="cpu")
        own = self._modules["encoder"].state_dict()
        msg = self._modules["encoder"].load_state_dict({k: v for k, v in state_dict.items() if k in own}, strict=False)
        print(f"Loading checkpoint: {msg}")


class CellViTShared(CellViT):
    """cellvit_shared.py:23-331 -- CellViT with ONE shared upsampling trunk (``decoder.*``) and a 1x1 convolution per output
    (``nuclei_binary_map_decoder``, ``hv_map_decoder``, ``nuclei_type_maps_decoder``) on its 64-channel feature map."""

    def __init__(self, num_nuclei_classes: int, num_tissue_classes: int, embed_dim: int, input_channels: int, depth: int,
                 num_heads: int, extract_layers: List, mlp_ratio: float = 4, qkv_bias: bool = True, drop_rate: float = 0,
                 attn_drop_rate: float = 0, drop_path_rate: float = 0, regression_loss: bool = False, _arch: str = "ViT256"):
        super().__init__(num_nuclei_classes, num_tissue_classes, embed_dim, input_channels, depth, num_heads, extract_layers,
                         mlp_ratio, qkv_bias, drop_rate, attn_drop_rate, drop_path_rate, regression_loss, _arch=_arch, _shared=True)


class CellViT256Shared(CellViTShared):
    """cellvit_shared.py:333-393."""

    def __init__(self, model256_path, num_nuclei_classes: int, num_tissue_classes: int, drop_rate: float = 0,
                 attn_drop_rate: float = 0, drop_path_rate: float = 0, regression_loss: bool = False):
        self.model256_path = model256_path
        super().__init__(num_nuclei_classes, num_tissue_classes, 384, 3, 12, 6, [3, 6, 9, 12], 4, True, drop_rate, attn_drop_rate,
                         drop_path_rate, regression_loss, _arch="ViT256")

    load_pretrained_encoder = CellViT256.load_pretrained_encoder


class CellViTSAMShared(CellViTShared):
    """cellvit_shared.py:396-560."""

    def __init__(self, model_path, num_nuclei_classes: int, num_tissue_classes: int, vit_structure, drop_rate: float = 0,
                 regression_loss: bool = False):
        if vit_structure.upper() not in weights.SAM_CFG:
            raise NotImplementedError("Unknown ViT-SAM backbone structure")
        cfg = weights.SAM_CFG[vit_structure.upper()]
        self.model_path = model_path
        super().__init__(num_nuclei_classes, num_tissue_classes, cfg["embed_dim"], 3, cfg["depth"], cfg["num_heads"],
                         list(cfg["extract"]), 4, True, drop_rate, 0, 0, regression_loss, _arch=vit_structure.upper())
        self.prompt_embed_dim = 256
        self.encoder_global_attn_indexes = list(cfg["global_idx"])

    load_pretrained_encoder = CellViTSAM.load_pretrained_encoder
