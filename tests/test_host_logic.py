"""CPU-side checks: the C-ABI library exports every declared symbol, the module mirrors keep the reference
state_dict surface and error behaviour, packing shapes, sharding logic (incl. a world-size-2 gloo run)."""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> str:
    """A TCP port nobody is listening on right now (the rendezvous of the world-size-2 runs must not collide with other jobs)."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return str(s.getsockname()[1])


@pytest.fixture(scope="module")
def lib_path():
    from cellvit_b200 import build
    return build.build()


def test_c_abi_exports_every_declared_symbol(lib_path):
    hdr = open(os.path.join(ROOT, "include", "cellvit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(cvb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 14
    L = ctypes.CDLL(lib_path)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.cvb_version() == 200
    # argument validation that needs no GPU
    L.cvb_last_error.restype = ctypes.c_char_p
    need = ctypes.c_size_t()
    assert L.cvb_postproc_workspace_bytes(2, 1024, 1024, ctypes.byref(need)) == 0 and need.value > 2 * 1024 * 1024 * 44
    assert L.cvb_postproc_workspace_bytes(0, 1024, 1024, ctypes.byref(need)) == -1
    assert L.cvb_model_create(None, None) == -1 and b"null" in L.cvb_last_error()


def test_every_exported_symbol_is_declared_in_a_header(lib_path):
    """No undeclared exports: the dynamic symbol table of libcellvit_b200.so holds exactly the cvb_* functions that
    include/cellvit_b200.h (the boundary) and include/cellvit_b200_debug.h (test / profiling hooks) declare; both compile as C99."""
    def declared(name):
        h = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", name)).read(), flags=re.S)
        return set(re.findall(r"\b(cvb_[a-z0-9_]+)\s*\(", h))
    api, dbg = declared("cellvit_b200.h"), declared("cellvit_b200_debug.h")
    assert not (api & dbg)
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.split() and l.split()[-1].startswith("cvb_")}
    assert exported == api | dbg, (sorted(exported - api - dbg), sorted((api | dbg) - exported))
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "cellvit_b200_debug.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_model_desc_validation_and_shape_errors(lib_path):
    from cellvit_b200.cellvit import CellViT256, ModelDesc
    L = ctypes.CDLL(lib_path)
    L.cvb_last_error.restype = ctypes.c_char_p
    m = CellViT256(None, 6, 19)
    h = m._ensure_handle()
    need = ctypes.c_size_t()
    assert L.cvb_model_workspace_bytes(h, 1, 256, 256, ctypes.byref(need)) == 0 and need.value > 0
    assert L.cvb_model_workspace_bytes(h, 1, 250, 250, ctypes.byref(need)) == -2     # not divisible by 16
    assert b"divisible by the patch size" in L.cvb_last_error()
    small = need.value
    assert L.cvb_model_workspace_bytes(h, 1, 320, 320, ctypes.byref(need)) == 0      # any multiple of 16: decoder on a 512 canvas
    canvas = need.value
    assert L.cvb_model_workspace_bytes(h, 1, 512, 512, ctypes.byref(need)) == 0 and small < need.value < canvas
    assert L.cvb_model_workspace_bytes(h, 1, 208, 1024, ctypes.byref(need)) == 0     # non-square (ViT-S only)
    assert L.cvb_model_workspace_bytes(h, 1, 1040, 256, ctypes.byref(need)) == -2    # beyond the 1024-pixel edge
    assert b"1024" in L.cvb_last_error()
    from cellvit_b200.cellvit import CellViTSAM
    ms = CellViTSAM(None, 6, 19, "SAM-B")
    hs = ms._ensure_handle()
    assert L.cvb_model_workspace_bytes(hs, 1, 400, 400, ctypes.byref(need)) == 0
    assert L.cvb_model_workspace_bytes(hs, 1, 256, 512, ctypes.byref(need)) == -2    # SAM: square token grids only (utils.py:222-224)
    assert b"square" in L.cvb_last_error()
    bad = ModelDesc(sam=0, embed_dim=100, depth=1, num_heads=3)
    out = ctypes.c_void_p()
    assert L.cvb_model_create(ctypes.byref(bad), ctypes.byref(out)) != 0


@pytest.mark.parametrize("arch", ["ViT256", "SAM-B", "SAM-H"])
def test_state_dict_surface_matches_reference_spec(arch):
    from cellvit_b200 import weights
    from cellvit_b200.cellvit import CellViT256, CellViTSAM
    m = CellViT256(None, 6, 19) if arch == "ViT256" else CellViTSAM(None, 6, 19, arch)
    sd = weights.synth_state_dict(arch, 6, 19, seed=1)
    own = m.state_dict()
    assert list(own.keys()) == list(sd.keys())
    assert all(tuple(own[k].shape) == tuple(sd[k].shape) for k in sd)
    m.load_state_dict(sd, strict=True)
    assert m.patch_size == 16 and m.num_nuclei_classes == 6 and m.embed_dim == (384 if arch == "ViT256" else weights.SAM_CFG[arch]["embed_dim"])
    n_params = sum(p.numel() for p in m.parameters())
    if arch == "ViT256":
        assert n_params == 46_750_349   # logs_paper/PanNuke/CellViTHV/ViT256/Best-Setting/Fold-1/logs.log:718
    if arch == "SAM-H":
        assert n_params == 699_741_149  # logs_paper/PanNuke/CellViTHV/SAM-H/Fold-1/logs.log:957


def test_forward_contract_errors_without_gpu():
    from cellvit_b200.cellvit import CellViT256, CellViTSAM
    m = CellViT256(None, 6, 19)
    with pytest.raises(AssertionError, match="divisible by the patch size"):
        m(torch.zeros(1, 3, 250, 256))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 256, 256))
    with pytest.raises(NotImplementedError, match="Unknown ViT-SAM backbone structure"):
        CellViTSAM(None, 6, 19, "SAM-X")
    from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor
    with pytest.raises(NotImplementedError, match="Unknown magnification"):
        DetectionCellPostProcessor(6, magnification=30)
    assert (DetectionCellPostProcessor(6, 40).object_size, DetectionCellPostProcessor(6, 20).k_size) == (10, 11)
    assert (DetectionCellPostProcessor(6, 40, gt=True).object_size, DetectionCellPostProcessor(6, 40, gt=True).k_size) == (100, 21)


def test_packing_layouts():
    from cellvit_b200 import packing, weights
    from cellvit_b200.cellvit import CellViT256
    sd = weights.synth_state_dict("ViT256", 6, 19, seed=2)
    cfg = CellViT256(None, 6, 19)._cfg()
    P = packing.pack_static(sd, cfg)
    P.update(packing.pack_for_size(sd, cfg, 16, 16))
    # conv weight: k = tap * Cpad + c, concat split (312 -> 320 twice)
    w = sd["hv_map_decoder.decoder3_upsampler.0.block.0.weight"]
    pk = P["hv.d3.0.w"].float().view(320, 3, 3, 640)
    assert torch.equal(pk[:312, :, :, :312], w[:, :312].permute(0, 2, 3, 1).half().float())
    assert torch.equal(pk[:312, :, :, 320:632], w[:, 312:].permute(0, 2, 3, 1).half().float())
    assert pk[312:].abs().sum() == 0 and pk[:, :, :, 312:320].abs().sum() == 0 and pk[:, :, :, 632:].abs().sum() == 0
    # folded BN: relu(scale * conv + shift) == relu(bn(conv + bias))
    p = "decoder2.1.block"
    s, sh = P["decoder2.1.conv.scale"], P["decoder2.1.conv.shift"]
    y = torch.randn(256)
    bn = (y + sd[p + ".1.bias"] - sd[p + ".2.running_mean"]) / torch.sqrt(sd[p + ".2.running_var"] + 1e-5) * sd[p + ".2.weight"] + sd[p + ".2.bias"]
    assert torch.allclose(y * s + sh, bn, atol=1e-5)
    # ConvTranspose: rows ordered (dy, dx, co)
    wt = sd["decoder1.0.block.0.weight"]
    pt = P["decoder1.0.ct.w"].float().view(2, 2, 256, 384)
    assert torch.equal(pt[1, 0], wt[:, :, 1, 0].t().half().float())
    assert P["pos"].shape == (257, 384) and P["clspos"].shape == (384,)


def test_packing_sam_rel_tables():
    """SAM window blocks: rel-pos tables [27, hd] plus their concatenation (rows 0.. = rel_h, rows 32.. = rel_w) that
    the tcgen05 window attention uses as one K-major operand; global blocks get tables resized to the token grid."""
    from cellvit_b200 import packing, weights
    from cellvit_b200.cellvit import CellViTSAM
    sd = weights.synth_state_dict("SAM-B", 6, 19, seed=2)
    cfg = CellViTSAM(None, 6, 19, "SAM-B")._cfg()
    P = packing.pack_static(sd, cfg)
    win = [i for i in range(cfg["depth"]) if i not in cfg["global_idx"]][0]
    rh, rw, cat = P[f"b{win}.relh"], P[f"b{win}.relw"], P[f"b{win}.relcat"]
    assert tuple(rh.shape) == (27, 64) and tuple(cat.shape) == (64, 64) and cat.dtype == torch.float16
    assert torch.equal(cat[:27], rh) and torch.equal(cat[32:59], rw)
    assert cat[27:32].abs().sum() == 0 and cat[59:].abs().sum() == 0
    assert torch.equal(rh.float(), sd[f"encoder.blocks.{win}.attn.rel_pos_h"].half().float())
    G = packing.pack_for_size(sd, cfg, 16, 16)
    g = cfg["global_idx"][0]
    assert tuple(G[f"b{g}.relh"].shape) == (31, 64) and f"b{g}.relcat" not in P


def test_shard_indices_cover_all_tiles_once():
    from cellvit_b200.cell_detection import shard_indices, unflatten_dict
    for n, world in [(10, 1), (10, 3), (7, 8), (10000, 8)]:
        got = sorted(i for r in range(world) for i in shard_indices(n, r, world))
        assert got == list(range(n))
    assert unflatten_dict({"data.num_nuclei_classes": 6, "model.backbone": "SAM-H"}) == {"data": {"num_nuclei_classes": 6}, "model": {"backbone": "SAM-H"}}


_GLOO = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from cellvit_b200.cell_detection import broadcast_weights, shard_indices
from cellvit_b200.cellvit import CellViT256
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)                    # different weights per rank before the broadcast
m = CellViT256(None, 6, 19)
with torch.no_grad():
    for p in m.parameters():
        p.add_(float(rank))
broadcast_weights(m, 0)
chk = torch.cat([p.detach().reshape(-1)[:4] for p in m.parameters()]).double().sum()
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
assert all(torch.equal(allc[0], c) for c in allc), allc
# tile sharding + all-gather of per-tile "instance counts" gives the same table on every rank, in tile order
n_tiles = 11
mine = shard_indices(n_tiles, rank, world)
counts = torch.zeros(n_tiles, dtype=torch.int64)
for i in mine:
    counts[i] = 100 + i
dist.all_reduce(counts)
assert counts.tolist() == [100 + i for i in range(n_tiles)]
# chunked exchange of instance tables: 5 steps with a chunk of 4 -> one collective after step 4 and one at the flush
from cellvit_b200.cell_detection import TableGather
B, rows, rb = 3, 8, 88
g = TableGather(world, B, rows, rb, 4, "cpu")
def fake(r, step):
    gen = torch.Generator().manual_seed(1000 * r + step)
    return (torch.randint(0, 9, (B,), generator=gen, dtype=torch.int32),
            torch.randint(0, 256, (B, 16, rb), generator=gen, dtype=torch.uint8))
for step in range(5):
    g.add(*fake(rank, step))
    if step == 3:
        assert g.collectives == 1 and g.n_gathered == 4
        for r in range(world):
            for k in range(4):
                c, t = g.tables(r, k)
                wc, wt = fake(r, k)
                assert torch.equal(c, wc) and torch.equal(t, wt[:, :rows])
g.flush()
assert g.collectives == 2 and g.n_gathered == 1
for r in range(world):
    c, t = g.tables(r, 0)
    wc, wt = fake(r, 4)
    assert torch.equal(c, wc) and torch.equal(t, wt[:, :rows])
g.flush()
assert g.collectives == 2
dist.destroy_process_group()
print("ok", rank)
'''


def test_world_size_2_gloo_broadcast_and_sharding(tmp_path):
    script = tmp_path / "gloo_job.py"
    script.write_text(_GLOO)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", _free_port(), str(script), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_public_header_is_plain_c(tmp_path):
    """include/cellvit_b200.h is the drop-in boundary: it must compile as strict C99 (no C++-isms, no CUDA or torch headers) and
    a C translation unit calling through it must link against the shared library."""
    from cellvit_b200 import build
    lib = build.build()
    src = tmp_path / "boundary.c"
    src.write_text('#include "cellvit_b200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    size_t need = 0;\n'
                   '    if (cvb_postproc_workspace_bytes(0, 1024, 1024, &need) != CVB_EARG) return 2;   /* no GPU needed */\n'
                   '    printf("%d %s\\n", cvb_version(), cvb_last_error());\n'
                   '    /* the WSI export entry point from plain C: two cells, one on the tile border */\n'
                   '    {\n'
                   '        int64_t bbox[8] = {0, 5, 9, 12, 40, 41, 50, 52}, pts[12] = {5, 0, 12, 0, 12, 9, 41, 40, 52, 40, 52, 50}, off[3] = {0, 3, 6};\n'
                   '        double cen[4] = {8.5, 4.25, 46.0, 45.0}, prob[2] = {0.75, 1.0};\n'
                   '        int64_t type[2] = {1, 2}, patch[4] = {0, 1, 0, 1}, status[2] = {2, 0}, offs[4] = {-32, 928, -32, 928};\n'
                   '        uint8_t edge[2] = {1, 0};\n'
                   '        int8_t pos[8] = {1, 0, 0, 0, 0, 0, 0, 0};\n'
                   '        cvb_cell_columns c = {2, bbox, cen, pts, off, prob, type, patch, status, offs, edge, pos};\n'
                   '        int64_t idx[2] = {0, 1};\n'
                   '        cvb_json_section s = {"{\\"cells\\": ", "}", idx, 2, CVB_JSON_CELLS, 1};\n'
                   '        if (cvb_export_json("cells.json", &c, &s, 1, -1) != CVB_OK) return 3;\n'
                   '    }\n'
                   '    return 0;\n}\n')
    exe = tmp_path / "boundary"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        lib, f"-Wl,-rpath,{os.path.dirname(lib)}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, cwd=str(tmp_path))
    assert run.returncode == 0 and run.stdout.split()[0] == "200", (run.returncode, run.stdout, run.stderr)
    import json
    cells = json.load(open(tmp_path / "cells.json"))["cells"]
    assert cells[0] == {"bbox": [[0, 5], [9, 12]], "centroid": [8.5, 4.25], "contour": [[5, 0], [12, 0], [12, 9]], "type_prob": 0.75, "type": 1,
                        "patch_coordinates": [0, 1], "cell_status": 2, "offset_global": [-32, 928], "edge_position": True,
                        "edge_information": {"position": [1, 0, 0, 0], "edge_patches": [[-1, 1]]}}
    assert cells[1]["edge_position"] is False and "edge_information" not in cells[1] and cells[1]["centroid"] == [46.0, 45.0]


def test_window_tc_operand_layout_reproduces_the_reference_attention():
    """The tcgen05 window-attention kernel (csrc/window_tc.cu) computes S = [Q | Gsel] [K | Sel]^T, where a softmax thread
    builds its Gsel half from the G = Q Rcat^T row with a register barrel shifter. This restates that data flow in torch on
    the CPU -- in[i] = G[26 - i] / scale packed in pairs, shift s = 13 - qpos applied as 8 / 4 / 2 / 1 stages on packed
    registers, QG row = [Q[64:80] | Gsel_h (14) 0 0 | Gsel_w (14) 0 0 | unused], Sel[k] = ones at kh(k) and 16 + kw(k), keys
    padded to 208, the key split 96 / 112 of the two threads of a row with a shared maximum, row sums from a ones operand --
    and checks the result against attention with the decomposed relative position bias (image_encoder.py:354-392)."""
    torch.manual_seed(0)
    g, hd, S, SP = 14, 80, 196, 208
    scale = hd ** -0.5
    q, k, v = torch.randn(S, hd), torch.randn(S, hd), torch.randn(S, hd)
    Rh, Rw = torch.randn(2 * g - 1, hd) * 0.2, torch.randn(2 * g - 1, hd) * 0.2
    # reference: softmax(scale q k^T + rel_h + rel_w) v, rel from the UNSCALED q
    idx = (torch.arange(g)[:, None] - torch.arange(g)[None, :]) + g - 1
    rq = q.view(g, g, hd)
    rel_h = torch.einsum("hwc,hkc->hwk", rq, Rh[idx])
    rel_w = torch.einsum("hwc,wkc->hwk", rq, Rw[idx])
    attn = ((q * scale) @ k.T).view(g, g, g, g) + rel_h[..., :, None] + rel_w[..., None, :]
    want = attn.view(S, S).softmax(-1) @ v
    relcat = torch.zeros(64, hd); relcat[:27] = Rh; relcat[32:59] = Rw
    G = q @ relcat.T                                                 # the G MMA: [S, 64]

    def barrel14(p, s):
        """p: 15 'registers' of (lo, hi) pairs; returns 7 registers holding in[s .. s + 13] (the kernel's barrel14)."""
        t1 = [p[min(i + 4, 14)] if s & 8 else p[i] for i in range(11)]
        t2 = [t1[i + 2] if s & 4 else t1[i] for i in range(9)]
        t3 = [t2[i + 1] if s & 2 else t2[i] for i in range(8)]
        return [(t3[i][1], t3[i + 1][0]) if s & 1 else t3[i] for i in range(7)]

    qg = torch.zeros(S, 64)
    qg[:, :16] = q[:, 64:80]
    for t in range(S):
        qh, qw = divmod(t, g)
        for half, qpos in ((0, qh), (1, qw)):
            gv = (G[t, 32 * half:32 * half + 32] / scale).tolist()
            p = [(gv[26 - 2 * i], gv[25 - 2 * i]) for i in range(13)] + [(gv[0], 0.0), (0.0, 0.0)]
            o7 = barrel14(p, 13 - qpos)
            vals = [x for pair in o7 for x in pair] + [0.0, 0.0]
            qg[t, 16 + 16 * half:32 + 16 * half] = torch.tensor(vals)
    sel = torch.zeros(SP, 64)
    for key in range(S):
        kh, kw = divmod(key, g)
        sel[key, kh] = 1.0
        sel[key, 16 + kw] = 1.0
    kp = torch.zeros(SP, hd); kp[:S] = k
    vp = torch.zeros(SP, hd); vp[:S] = v
    ones = torch.zeros(SP); ones[:S] = 1.0
    # the kernel's seven k-steps: Q[0:64] K[0:64], Q[64:80] (QG cols 0..15) K[64:80], Gsel (QG cols 16..47) Sel (cols 0..31)
    s_acc = q[:, :64] @ kp[:, :64].T + qg[:, :16] @ kp[:, 64:80].T + qg[:, 16:48] @ sel[:, :32].T
    sl2 = scale * math.log2(math.e)
    m = torch.maximum(s_acc[:, :96].max(1, keepdim=True).values, s_acc[:, 96:S].max(1, keepdim=True).values)
    p = torch.exp2((s_acc - m) * sl2)
    p[:, S:] = 0.0                                                   # keys 196..207: weight 0
    got = (p @ vp) / (p @ ones)[:, None]
    assert torch.allclose(got, want, rtol=1e-4, atol=1e-5), (got - want).abs().max()


@pytest.mark.parametrize("arch", ["ViT256", "SAM-B"])
def test_shared_decoder_variants_surface(arch, lib_path):
    """The ``*Shared`` variants (cellvit_shared.py): reference constructor signatures, the reference's state_dict keys in its
    order (weights.state_spec(shared=True) is pinned against the reference modules in tests/test_oracle_vs_reference.py), a
    third of the decoder workspace, and model_from_checkpoint accepts the arch names of the reference's __get_model."""
    from cellvit_b200 import weights
    from cellvit_b200.cell_detection import model_from_checkpoint
    from cellvit_b200.cellvit import CellViT256, CellViT256Shared, CellViTSAM, CellViTSAMShared, CellViTShared
    m = CellViT256Shared(None, 6, 19) if arch == "ViT256" else CellViTSAMShared(None, 6, 19, arch, regression_loss=True)
    assert isinstance(m, CellViTShared)
    sd = weights.synth_state_dict(arch, 6, 19, seed=1, regression_loss=arch != "ViT256", shared=True)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    assert m.branches_output == {"nuclei_binary_map": 2 if arch == "ViT256" else 4, "hv_map": 2, "nuclei_type_maps": 6}
    assert any(k.startswith("decoder.decoder3_upsampler.") for k in sd) and "nuclei_binary_map_decoder.weight" in sd
    ckpt = {"arch": "CellViT256Shared" if arch == "ViT256" else "CellViTSAMShared",
            "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": arch if arch != "ViT256" else "default",
                       "model.regression_loss": arch != "ViT256"},
            "model_state_dict": sd}
    loaded, _ = model_from_checkpoint(ckpt)
    assert type(loaded).__name__ == ckpt["arch"] and not loaded.training
    full = CellViT256(None, 6, 19) if arch == "ViT256" else CellViTSAM(None, 6, 19, arch)
    L = ctypes.CDLL(lib_path)
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    assert L.cvb_model_workspace_bytes(m._ensure_handle(), 2, 256, 256, ctypes.byref(a)) == 0
    assert L.cvb_model_workspace_bytes(full._ensure_handle(), 2, 256, 256, ctypes.byref(b)) == 0
    assert 0 < a.value <= b.value
    with pytest.raises(NotImplementedError):
        model_from_checkpoint(dict(ckpt, arch="CellViTUnknown"))


@pytest.mark.parametrize("B,h,w,ws", [(2, 64, 64, 14), (1, 16, 16, 14), (2, 25, 25, 14), (1, 13, 13, 14), (1, 28, 28, 14), (3, 5, 9, 4)])
def test_window_partition_index_maps_restated(B, h, w, ws):
    """CPU restatement of the index arithmetic the windowed blocks rely on (csrc/tc_gemm.cu map_out_row, csrc/ops_misc.cu
    window_pad_fill_kernel, csrc/window_tc.cu un-partitioned store) against torch's window_partition (image_encoder.py:263-288):
    the raster -> window-order row map (TC_ROW_TO_WINDOW) hits exactly the rows that hold real tokens, the padding-token
    enumeration of the fill kernel hits exactly the other rows, and the window -> raster map (TC_ROW_WINDOW / un_g) inverts it."""
    import torch.nn.functional as F
    g = max((h + ws - 1) // ws, (w + ws - 1) // ws)
    P = g * ws
    tok = torch.arange(1, B * h * w + 1, dtype=torch.float32).view(B, h, w, 1)            # token ids in raster order, 0 = padding
    padded = F.pad(tok, (0, 0, 0, P - w, 0, P - h))
    win = padded.view(B, g, ws, g, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1).long().numpy()   # window_partition, flattened rows

    def to_window(m):   # TC_ROW_TO_WINDOW
        b, rem = divmod(m, h * w)
        y, x = divmod(rem, w)
        wy, wx = y // ws, x // ws
        return b * P * P + (wy * g + wx) * ws * ws + (y - wy * ws) * ws + (x - wx * ws)

    def to_raster(r):   # TC_ROW_WINDOW and the un-partitioned store of window_tc_kernel: -1 = padding token
        b, rem = divmod(r, P * P)
        wi, t = divmod(rem, ws * ws)
        y, x = (wi // g) * ws + t // ws, (wi % g) * ws + t % ws
        return (b * h + y) * w + x if (y < h and x < w) else -1

    real_rows = [to_window(m) for m in range(B * h * w)]
    assert [int(win[r]) for r in real_rows] == list(range(1, B * h * w + 1))
    right, per_img = h * (P - w), h * (P - w) + (P - h) * P
    pad_rows = []
    for r in range(B * per_img):                              # window_pad_fill_kernel's enumeration
        b, i = divmod(r, per_img)
        if i < right:
            y, x = i // (P - w), w + i % (P - w)
        else:
            y, x = h + (i - right) // P, (i - right) % P
        wy, wx = y // ws, x // ws
        pad_rows.append(b * P * P + (wy * g + wx) * ws * ws + (y - wy * ws) * ws + (x - wx * ws))
    assert len(set(pad_rows)) == len(pad_rows) and sorted(pad_rows) == [r for r in range(len(win)) if win[r] == 0]
    assert sorted(real_rows + pad_rows) == list(range(len(win)))
    assert [to_raster(r) for r in range(len(win))] == [int(v) - 1 for v in win]
