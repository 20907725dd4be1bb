"""Any tile size divisible by the patch size, as the reference forward accepts (cellvit.py:170-175, 603-608): the encoder
runs on the real token grid, the decoder on a zero-extended canvas (csrc/model.cu), the outputs are cropped. Checked
against the fp32 oracle (oracle/forward_oracle.py) on the same seeded inputs; tolerance 1e-3 abs on the head maps."""
import ctypes as C

import pytest
import torch

from cellvit_b200 import _lib as L
from cellvit_b200 import synth, weights

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _model(arch):
    from cellvit_b200.cellvit import CellViT256, CellViTSAM
    sd = weights.synth_state_dict(arch, 6, 19, seed=3)
    m = CellViT256(None, 6, 19) if arch == "ViT256" else CellViTSAM(None, 6, 19, arch)
    m.load_state_dict(sd, strict=True)
    return sd, m.cuda().eval()


def _oracle_on_device(sd, x, arch):
    """fp32 oracle evaluated with plain torch ops on the GPU (TF32 off), one tile at a time."""
    from oracle import forward_oracle
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    refs = [forward_oracle.cellvit_forward(sd_dev, x[b:b + 1].cuda(), arch, retrieve_tokens=True) for b in range(x.shape[0])]
    return {k: torch.cat([r[k] for r in refs]).cpu() for k in refs[0]}


def _check(out, ref):
    errs = {k: (out[k].cpu() - ref[k]).abs().max().item() for k in ref}
    print(errs)
    for k in ("nuclei_binary_map", "hv_map", "nuclei_type_map"):
        assert out[k].shape == ref[k].shape and out[k].dtype == torch.float32
        assert errs[k] <= TOL, errs
    assert errs["tissue_types"] <= 5e-3, errs
    assert out["tokens"].shape == ref["tokens"].shape
    assert errs["tokens"] <= 2e-2 * max(1.0, ref["tokens"].abs().max().item()), errs


def test_copy_planes_and_zero_margin():
    g = torch.Generator(device="cuda").manual_seed(1)
    lib = L.lib()
    for dt, nbytes, per in ((torch.float32, 4, 1), (torch.uint8, 1, 1), (torch.float16, 16, 8)):
        src = (torch.rand(3, 5, 7 * per, device="cuda", generator=g) * 200).to(dt)
        big = torch.full((3, 8, 12 * per), 7, device="cuda", dtype=dt)
        L.check(lib.cvb_op_copy_planes(L.ptr(src), 5, 7, L.ptr(big), 8, 12, C.c_longlong(3), nbytes, L.stream_ptr()), "embed")
        want = torch.zeros_like(big)
        want[:, :5, :7 * per] = src
        assert torch.equal(big, want)
        small = torch.full((3, 4, 3 * per), 7, device="cuda", dtype=dt)
        L.check(lib.cvb_op_copy_planes(L.ptr(big), 8, 12, L.ptr(small), 4, 3, C.c_longlong(3), nbytes, L.stream_ptr()), "crop")
        assert torch.equal(small, src[:, :4, :3 * per])
    for (H, W, vH, vW) in ((8, 12, 5, 7), (8, 12, 8, 7), (8, 12, 5, 12), (8, 12, 8, 12), (6, 4, 1, 1)):
        buf = torch.rand(2, H, W, 24, device="cuda", generator=g).half() + 1
        want = buf.clone()
        want[:, vH:] = 0
        want[:, :, vW:] = 0
        L.check(lib.cvb_op_zero_margin(L.ptr(buf), C.c_longlong(2), H, W, 48, vH, vW, L.stream_ptr()), "margin")
        assert torch.equal(buf, want), (H, W, vH, vW)


# (H, W): 272 x 400 -> 17 x 25 tokens on a 32 x 32 canvas; 208 x 1024 -> 13 x 64 on a NON-SQUARE 16 x 64 canvas (one native
# dimension); 1024 x 640 -> 64 x 40 on 64 x 64; 96 x 128 -> 48 tokens (fewer rows than one GEMM tile, mma.sync attention)
@pytest.mark.parametrize("H,W,B", [(272, 400, 2), (208, 1024, 1), (1024, 640, 1), (96, 128, 2), (512, 256, 1)])
def test_vit256_any_tile_size_matches_oracle(H, W, B):
    sd, m = _model("ViT256")
    x = torch.from_numpy(synth.synthetic_tiles(B, (H, W), seed=5))
    ref = _oracle_on_device(sd, x, "ViT256")
    with torch.no_grad():
        out = m(x.cuda(), retrieve_tokens=True, argmax_maps=True)
    torch.cuda.synchronize()
    _check(out, ref)
    # the arg-max planes of the fused head are cropped with the logits
    assert out["nuclei_type_argmax"].shape == (B, H, W)
    assert torch.equal(out["nuclei_type_argmax"].long(), out["nuclei_type_map"].argmax(1))
    assert torch.equal(out["nuclei_binary_argmax"].long(), out["nuclei_binary_map"][:, :2].argmax(1))


# 208 -> 13 tokens (smaller than one 14 x 14 window), 304 -> 19 (two windows per edge, canvas 32), 400 -> 25, 640 -> 40 (canvas 64)
@pytest.mark.parametrize("arch,size", [("SAM-B", 208), ("SAM-B", 304), ("SAM-B", 400), ("SAM-H", 640)])
def test_sam_any_square_tile_size_matches_oracle(arch, size):
    sd, m = _model(arch)
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=5))
    ref = _oracle_on_device(sd, x, arch)
    with torch.no_grad():
        out = m(x.cuda(), retrieve_tokens=True)
    torch.cuda.synchronize()
    _check(out, ref)


def test_square_canvas_option_gives_the_same_maps():
    sd, m = _model("ViT256")
    x = torch.from_numpy(synth.synthetic_tiles(1, (208, 528), seed=8)).cuda()
    with torch.no_grad():
        a = m(x)
        m.set_engine_option("square_canvas", 1)
        b = m(x)
    torch.cuda.synchronize()
    for k in ("nuclei_binary_map", "hv_map", "nuclei_type_map"):
        # same arithmetic per output pixel up to the accumulation order of a different tiling
        assert (a[k] - b[k]).abs().max().item() <= 1e-4, k


def test_native_sizes_do_not_take_the_canvas_path():
    """256 / 512 / 1024 tiles launch exactly the kernels they launched before the canvas path existed (no copies, no masks)."""
    sd, m = _model("ViT256")
    lib = L.lib()
    lib.cvb_launch_count.restype = C.c_longlong
    counts = {}
    for size in ((256, 256), (256, 512), (272, 256)):
        x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=2)).cuda()
        with torch.no_grad():
            m(x)
            torch.cuda.synchronize()
            lib.cvb_launch_count(1)
            m(x)
            torch.cuda.synchronize()
        counts[size] = lib.cvb_launch_count(1)
    assert counts[(256, 256)] == counts[(256, 512)], counts      # both native: same launch sequence
    assert counts[(272, 256)] > counts[(256, 256)], counts       # canvas: embeds, masks, crops on top


def test_shape_errors_follow_the_reference():
    sd, m = _model("SAM-B")
    with pytest.raises(AssertionError):                          # cellvit.py:603-608
        m(torch.zeros(1, 3, 250, 256, device="cuda"))
    with pytest.raises(RuntimeError, match="square"):            # utils.py:222-224 cannot broadcast a non-square grid
        m(torch.zeros(1, 3, 256, 512, device="cuda"))
    sd, m = _model("ViT256")
    with pytest.raises(RuntimeError, match="1024"):
        m(torch.zeros(1, 3, 1040, 256, device="cuda"))


def test_process_tiles_on_a_non_native_tile_size_matches_oracle():
    """272 x 400 tiles through the product pipeline (CUDA-graph replay of the canvas forward, device post-processing, contours,
    cell tokens) against the oracle post-processing of the same head maps."""
    import numpy as np
    from cellvit_b200.cell_detection import CellSegmentationInference
    from oracle import postproc_oracle as po
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)
    B, H, W = 2, 272, 400
    nuc = [synth.synthetic_nuclei(400, 80 + 10 * i, seed=20 + i) for i in range(B)]
    nuc = [{k: np.ascontiguousarray(v[..., :H, :W]) for k, v in n.items() if k in ("np_bin", "nt", "hv")} for n in nuc]
    lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
    override = {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])).cuda(),
                "hv_map": torch.from_numpy(np.stack([n["hv"] for n in nuc])).cuda()}
    tiles = torch.from_numpy(synth.synthetic_tiles(B, (H, W), seed=9)).pin_memory()
    res = [(d, t) for _, d, t in inf._pipeline(((tiles, k) for k in range(3)), 40, head_override=override, with_tokens=True)]
    assert len(res) == 3
    for dicts, toks in res:
        for b in range(B):
            pm = np.concatenate([nuc[b]["nt"][..., None], nuc[b]["np_bin"][..., None], nuc[b]["hv"].transpose(1, 2, 0)], -1).astype(np.float64)
            _, odict = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
            assert sorted(dicts[b]) == sorted(odict) and len(odict) > 5
            for k, ov in odict.items():
                gv = dicts[b][k]
                assert np.array_equal(gv["bbox"], ov["bbox"]) and np.array_equal(gv["centroid"], ov["centroid"])
                assert np.array_equal(gv["contour"], ov["contour"]) and gv["type"] == ov["type"] and gv["type_prob"] == ov["type_prob"]
            assert toks[b].shape == (len(odict), 384)


@pytest.mark.parametrize("arch,size", [("SAM-B", 256), ("SAM-H", 400), ("SAM-B", 1024)])
def test_window_pad_skip_gives_the_same_forward(arch, size):
    """The QKV GEMM of the windowed blocks runs over the real tokens only (rows scattered into window order by the epilogue,
    padding rows = bias): same maps as running it over the zero-padded window-partitioned rows (image_encoder.py:180-184)."""
    sd, m = _model(arch)
    x = torch.from_numpy(synth.synthetic_tiles(2, size, seed=4)).cuda()
    with torch.no_grad():
        a = m(x, retrieve_tokens=True)
        m.set_engine_option("window_pad_skip", 0)
        b = m(x, retrieve_tokens=True)
    torch.cuda.synchronize()
    for k in ("nuclei_binary_map", "hv_map", "nuclei_type_map", "tissue_types", "tokens"):
        d = (a[k] - b[k]).abs().max().item()
        print(k, "max abs difference", d, "bit-equal", torch.equal(a[k], b[k]))
        assert d <= 1e-5 * max(1.0, b[k].abs().max().item()), (k, d)


@pytest.mark.parametrize("arch,size,B,regression", [("ViT256", 256, 2, False), ("SAM-B", 256, 1, True), ("ViT256", (272, 400), 1, True)])
def test_shared_decoder_variants_match_oracle(arch, size, B, regression):
    """CellViT256Shared / CellViTSAMShared (cellvit_shared.py:147-231): one upsampling trunk, three 1x1 heads -- against the fp32
    oracle (pinned to the reference's Shared modules in tests/test_oracle_vs_reference.py), native and canvas tile sizes."""
    from cellvit_b200.cellvit import CellViT256Shared, CellViTSAMShared
    sd = weights.synth_state_dict(arch, 6, 19, seed=3, regression_loss=regression, shared=True)
    m = CellViT256Shared(None, 6, 19, regression_loss=regression) if arch == "ViT256" else CellViTSAMShared(None, 6, 19, arch, regression_loss=regression)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    x = torch.from_numpy(synth.synthetic_tiles(B, size, seed=5))
    from oracle import forward_oracle
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    refs = [forward_oracle.cellvit_forward(sd_dev, x[b:b + 1].cuda(), arch, retrieve_tokens=True, regression_loss=regression) for b in range(B)]
    ref = {k: torch.cat([r[k] for r in refs]).cpu() for k in refs[0]}
    with torch.no_grad():
        out = m(x.cuda(), retrieve_tokens=True, argmax_maps=True)
    torch.cuda.synchronize()
    assert ("regression_map" in out) == regression
    if regression:
        assert (out["regression_map"].cpu() - ref["regression_map"]).abs().max().item() <= TOL
    _check(out, {k: v for k, v in ref.items() if k != "regression_map"})
    assert torch.equal(out["nuclei_type_argmax"].long(), out["nuclei_type_map"].argmax(1))
    assert torch.equal(out["nuclei_binary_argmax"].long(), out["nuclei_binary_map"][:, :2].argmax(1))
