"""Parity cases the first round left open ("parity hardening"):
 (a) head maps at trained-checkpoint magnitude (|logit| ~ 10): post-softmax probabilities within 1e-3 of the fp32 oracle;
 (b) the benchmarked configuration itself (CellViT-SAM-H, B = 4 tiles of 1024^2) against the fp32 oracle;
 (c) post-processing fed with the MODEL'S OWN head outputs (non-degenerate maps) against the oracle's post-processing of
     the same maps;
 (d) a lazily produced device-tensor tile stream and raw uint8 device tiles through the pipeline."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from cellvit_b200 import synth, weights
from oracle import forward_oracle, postproc_oracle as po

pytestmark = pytest.mark.gpu
HEADS = ("nuclei_binary_map", "hv_map", "nuclei_type_maps")


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _model(arch, sd):
    from cellvit_b200.cellvit import CellViT256, CellViTSAM
    m = CellViT256(None, 6, 19) if arch == "ViT256" else CellViTSAM(None, 6, 19, arch)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


def _scaled_heads(arch, seed, x, target=10.0, smooth_upsampling=False):
    """Synthetic weights whose 1x1 heads are rescaled so that every head map spans roughly +-target (random-init logits are
    <= 0.15, which hides relative errors: SURVEY.md section 7.3-2). The factor comes from the fp32 oracle's own head outputs."""
    sd = weights.synth_state_dict(arch, 6, 19, seed=seed)
    if smooth_upsampling:   # random 2x2 transposed convolutions emit a pixel-scale checkerboard: make them nearest-neighbour-like
        for k, v in sd.items():
            if v.ndim == 4 and tuple(v.shape[-2:]) == (2, 2):
                sd[k] = v[:, :, :1, :1].expand_as(v).contiguous()
    ref = forward_oracle.cellvit_forward(sd, x, arch, retrieve_tokens=False)
    for branch, key in zip(HEADS, ("nuclei_binary_map", "hv_map", "nuclei_type_map")):
        w, b = f"{branch}_decoder.decoder0_header.2.weight", f"{branch}_decoder.decoder0_header.2.bias"
        r = ref[key]
        centred = r - r.mean(dim=(0, 2, 3), keepdim=True)
        f = target / max(centred.abs().max().item(), 1e-6)
        # logits' = f * (logits - mean): scale the weights, shift the bias so that the per-channel mean is zero
        sd[b] = (sd[b] - r.mean(dim=(0, 2, 3))) * f
        sd[w] = sd[w] * f
    return sd


@pytest.mark.parametrize("arch,size", [("ViT256", 256), ("SAM-B", 256)])
def test_forward_at_trained_logit_magnitude(arch, size):
    """Random-init heads emit |logit| <= 0.15, where the 1e-3 abs bar is a ~1e-2 RELATIVE bar. With the heads rescaled to
    |logit| ~ 10 the engine's fp16 operands / fp16 inter-layer activations (fp32 accumulate) show their real accuracy:
    ~1e-3 relative on the logits (measured 8.8e-4 on ViT-256), i.e. up to ~1e-2 abs at this magnitude and a few 1e-3 on the
    post-softmax probabilities. That is the accuracy class of the reference's OWN mixed-precision inference mode
    (torch.autocast fp16, cell_detection.py:314-318): the test evaluates the fp32 oracle under autocast on the GPU and requires
    the engine to be no further from fp32 than 1.5x that, plus fixed bars: logits 2e-3 relative, probabilities 6e-3, arg-max
    identical wherever the fp32 top-2 probability gap exceeds 1e-2."""
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=11))
    sd = _scaled_heads(arch, 5, x)
    ref = forward_oracle.cellvit_forward(sd, x, arch, retrieve_tokens=False)
    assert ref["nuclei_binary_map"].abs().max() > 5 and ref["nuclei_type_map"].abs().max() > 5   # the heads really are at magnitude ~10
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.float16):
        amp = {k: v.float().cpu() for k, v in forward_oracle.cellvit_forward(sd_dev, x.cuda(), arch, retrieve_tokens=False).items()}
    m = _model(arch, sd)
    with torch.no_grad():
        out = {k: v.cpu() for k, v in m(x.cuda()).items()}
    for k in ("nuclei_binary_map", "nuclei_type_map", "hv_map"):
        scale = max(1.0, ref[k].abs().max().item())
        logit_err = (out[k] - ref[k]).abs().max().item()
        amp_err = (amp[k] - ref[k]).abs().max().item()
        print(f"{arch} {k}: max |x| {scale:.2f}  engine err {logit_err:.2e} ({logit_err / scale:.1e} rel)  reference-under-autocast err {amp_err:.2e}")
        assert logit_err <= 2e-3 * scale, (k, logit_err)
        assert logit_err <= 1.5 * amp_err + 1e-4, (k, logit_err, amp_err)
        if k != "hv_map":
            p_out, p_ref = F.softmax(out[k], 1), F.softmax(ref[k], 1)
            prob_err = (p_out - p_ref).abs().max().item()
            print(f"{arch} {k}: probability err {prob_err:.2e} (autocast: {(F.softmax(amp[k], 1) - p_ref).abs().max().item():.2e})")
            assert prob_err <= 6e-3, (k, prob_err)
            clear = (p_ref.topk(2, 1).values.diff(dim=1).abs() > 1e-2)[:, 0]   # pixels whose arg-max is not a near-tie
            assert torch.equal(out[k].argmax(1)[clear], ref[k].argmax(1)[clear])


def test_forward_sam_h_batch4_1024_matches_oracle_on_device():
    """The benchmarked configuration: CellViT-SAM-H, B = 4 tiles of 1024^2 in ONE cvb_forward call (fp32 oracle evaluated on the
    GPU tile by tile, TF32 off)."""
    arch, B = "SAM-H", 4
    sd = weights.synth_state_dict(arch, 6, 19, seed=3)
    x = torch.from_numpy(synth.synthetic_tiles(B, 1024, seed=16)).cuda()
    sd_dev = {k: v.cuda() for k, v in sd.items()}
    refs = [{k: v.cpu() for k, v in forward_oracle.cellvit_forward(sd_dev, x[b:b + 1], arch, retrieve_tokens=True).items()} for b in range(B)]
    del sd_dev
    torch.cuda.empty_cache()
    m = _model(arch, sd)
    with torch.no_grad():
        out = m(x, retrieve_tokens=True)
    torch.cuda.synchronize()
    for k in ("nuclei_binary_map", "hv_map", "nuclei_type_map"):
        err = (out[k].cpu() - torch.cat([r[k] for r in refs])).abs().max().item()
        print(k, err)
        assert err <= 1e-3, (k, err)


@pytest.mark.parametrize("seed", [12, 16, 18])
def test_postprocessing_of_the_models_own_head_maps_matches_oracle(seed):
    """The device post-processing consumes the forward's own output buffers (fused head epilogue -> prep_float_kernel hand-off) and
    must equal the oracle's post-processing of the very same maps copied to the host. Random-init networks emit degenerate maps
    (SURVEY.md 8d), so the network is made to produce nuclei-like ones: a tile with nuclei-sized blobs, nearest-neighbour-like
    transposed convolutions (no checkerboard), heads rescaled to |logit| ~ 10 with zero-mean channels. The seeds are ones for
    which the CPU oracle's own maps give ~20 instances per tile (most random HV heads leave no marker at all)."""
    from cellvit_b200.post_proc_cellvit import DetectionCellPostProcessor
    arch, size, B = "ViT256", 256, 1
    x = torch.from_numpy(synth.synthetic_tiles(B, size, seed=21)) * 0.02
    mask = torch.from_numpy(synth.synthetic_nuclei(size, 30, seed=60)["np_bin"].astype(np.float32))
    mask = F.avg_pool2d(mask[None, None], 5, 1, 2)[0, 0]
    x[0] += torch.stack([1.4 * mask - 0.6, 0.9 * mask - 0.4, -1.1 * mask + 0.5])
    sd = _scaled_heads(arch, seed, x, smooth_upsampling=True)
    m = _model(arch, sd)
    with torch.no_grad():
        out = m(x.cuda(), retrieve_tokens=True)
    proc = DetectionCellPostProcessor(nr_types=6, magnification=40)
    labels, dicts = proc.post_process_batch(out["nuclei_binary_map"], out["hv_map"], out["nuclei_type_map"])
    labels = labels.cpu().numpy()
    np_bin = out["nuclei_binary_map"].argmax(1).cpu().numpy().astype(np.uint8)
    nt = out["nuclei_type_map"].argmax(1).cpu().numpy()
    hv = out["hv_map"].cpu().numpy()
    # the product pipeline (K12 fusion): arg-max planes written by the head epilogue, consumed in place by cvb_postproc_argmax
    with torch.no_grad():
        out2 = m(x.cuda(), retrieve_tokens=True, argmax_maps=True)
    assert torch.equal(out2["nuclei_binary_argmax"], out2["nuclei_binary_map"].argmax(1).to(torch.uint8))
    assert torch.equal(out2["nuclei_type_argmax"], out2["nuclei_type_map"].argmax(1).to(torch.uint8))
    assert torch.equal(out2["nuclei_binary_map"], out["nuclei_binary_map"]) and torch.equal(out2["hv_map"], out["hv_map"])
    from cellvit_b200.cell_detection import CellSegmentationInference
    inf = CellSegmentationInference.from_model(m, 0)
    piped = [d for _, d, _ in inf._pipeline([(x.pin_memory(), None)], 40, use_graphs=True)]
    for b in range(B):
        pm = np.concatenate([nt[b][..., None], np_bin[b][..., None], hv[b].transpose(1, 2, 0)], -1).astype(np.float64)
        olab, odict = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
        print("seed", seed, "foreground", float(np_bin[b].mean()), "instances", len(odict))
        assert sorted(piped[0][b]) == sorted(odict)
        for k, ov in odict.items():
            assert np.array_equal(piped[0][b][k]["contour"], ov["contour"]) and piped[0][b][k]["type"] == ov["type"]
        assert len(odict) >= 5, "the construction no longer gives a non-degenerate instance map"
        assert np.array_equal(labels[b], olab)
        assert sorted(dicts[b]) == sorted(odict)
        for k, ov in odict.items():
            gv = dicts[b][k]
            assert np.array_equal(gv["bbox"], ov["bbox"]) and np.array_equal(gv["centroid"], ov["centroid"])
            assert np.array_equal(gv["contour"], ov["contour"]) and gv["type"] == ov["type"] and gv["type_prob"] == ov["type_prob"]


def test_pipeline_accepts_lazy_device_batches_and_uint8_device_tiles():
    """process_tiles documents device tensors as valid input. A generator that normalises tiles on the GPU enqueues its producer
    kernels on the caller's stream while the copy stream stages the next batch (the copy must wait for them); raw uint8
    device tiles must be normalised on the non-graph path too."""
    from cellvit_b200.cell_detection import CellSegmentationInference
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)
    size, B, n_batches = 256, 2, 5
    rng = np.random.default_rng(0)
    u8 = [rng.integers(0, 256, size=(B, size, size, 3), dtype=np.uint8) for _ in range(n_batches)]
    nuc = [synth.synthetic_nuclei(size, 30, seed=80 + i) for i in range(B)]
    lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
    captured = []

    def override(payload):   # keep the forward's own tokens: they depend on the input really having arrived in time
        return {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])).cuda(),
                "hv_map": torch.from_numpy(np.stack([n["hv"] for n in nuc])).cuda()}

    def tokens_of(batches, graphs):
        toks = []
        for _, _, t in inf._pipeline(batches, 40, head_override=override, with_tokens=True, use_graphs=graphs):
            toks.append(np.concatenate(t))
        return toks

    def norm(t, device):   # ToTensor + Normalize with true divisions, the operation order of the pipeline's uint8 path
        x = torch.from_numpy(t).to(device).permute(0, 3, 1, 2).to(torch.float32) / torch.tensor(255.0, device=device)
        return ((x - torch.tensor(0.5, device=device)) / torch.tensor(0.5, device=device)).contiguous()

    want = tokens_of(((norm(t, "cpu").pin_memory(), None) for t in u8), True)                     # pinned-host float batches
    for graphs in (True, False):
        lazy = tokens_of(((norm(t, "cuda"), None) for t in u8), graphs)                            # produced lazily on the GPU
        raw = tokens_of(((torch.from_numpy(t).permute(0, 3, 1, 2).contiguous().cuda(), None) for t in u8), graphs)   # uint8 device tiles
        for a, b, c in zip(want, lazy, raw):
            assert np.array_equal(a, b), "lazily produced device batch: the forward read a stale / unwritten input buffer"
            assert np.array_equal(a, c), "uint8 device tiles were not normalised"
