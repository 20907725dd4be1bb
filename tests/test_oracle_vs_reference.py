"""Pins the CPU oracle (oracle/) against the UNMODIFIED reference imported from /root/reference and against
cv2 / scipy directly. Runs only where the reference tree exists (this container); the GPU box uses the
golden fixtures in tests/golden instead (tests/test_oracle_golden.py)."""
import numpy as np
import pytest
import torch

from cellvit_b200 import synth, weights
from oracle import forward_oracle, postproc_oracle as po, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    captured = {}

    def spy_watershed(image, markers=None, mask=None):
        captured["dist"], captured["marker"], captured["mask"] = np.array(image), np.array(markers), np.array(mask)
        return po.watershed(image, markers=markers, mask=mask)

    cellvit, post = ref_shim.import_reference(spy_watershed)
    return cellvit, post, captured


def _pred_map(d):
    return np.concatenate([d["nt"][..., None], d["np_bin"][..., None], d["hv"].transpose(1, 2, 0)], -1).astype(np.float64)


@pytest.mark.parametrize("size,n,seed,mag,noise", [(256, 40, 0, 40, 0.0), (256, 60, 1, 20, 0.0), (512, 170, 2, 40, 0.02),
                                                   (320, 80, 3, 40, 0.05)])
def test_postproc_stages_match_reference(ref, size, n, seed, mag, noise):
    _, post, cap = ref
    d = synth.synthetic_nuclei(size, n, seed, noise=noise)
    proc = post.DetectionCellPostProcessor(nr_types=6, magnification=mag, gt=False)
    ref_lab, ref_dict = proc.post_process_cell_segmentation(_pred_map(d))
    lab, inter = po.proc_np_hv(d["np_bin"], d["hv"], mag, want_intermediates=True)
    # P1-P6 are computed by the reference through real cv2/scipy; P7 goes through the oracle flood in both.
    assert np.array_equal(inter["blb"].astype(bool), cap["mask"].astype(bool))
    assert np.array_equal(inter["marker"], cap["marker"])
    assert np.array_equal(inter["dist"], cap["dist"]), np.abs(inter["dist"] - cap["dist"]).max()
    assert np.array_equal(lab, ref_lab)
    # P8/P9
    olab, odict = po.DetectionCellPostProcessor(6, mag).post_process_cell_segmentation(_pred_map(d))
    assert np.array_equal(olab, ref_lab)
    assert sorted(odict) == sorted(ref_dict)
    for k, r in ref_dict.items():
        o = odict[k]
        assert np.array_equal(o["bbox"], r["bbox"])
        assert np.array_equal(o["centroid"], r["centroid"]), (o["centroid"], r["centroid"])
        assert np.array_equal(o["contour"], r["contour"])
        assert o["type"] == r["type"] and o["type_prob"] == r["type_prob"]


def test_postproc_degenerate_tiles(ref):
    _, post, _ = ref
    proc = post.DetectionCellPostProcessor(nr_types=6, magnification=40)
    H = 96
    empty = np.zeros((H, H, 4))
    lab, dct = proc.post_process_cell_segmentation(empty)
    olab, odct = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(empty)
    assert np.array_equal(lab, olab) and dct == {} and odct == {}
    # all-foreground tile: exercises the np.unique(...)[1:] quirk (post_proc_cellvit.py:95)
    full = np.zeros((H, H, 4)); full[..., 1] = 1
    yy, xx = np.mgrid[0:H, 0:H]
    full[..., 2] = (xx - H / 2) / H; full[..., 3] = (yy - H / 2) / H; full[..., 0] = 2
    lab, dct = proc.post_process_cell_segmentation(full)
    olab, odct = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(full)
    assert np.array_equal(lab, olab) and sorted(dct) == sorted(odct)


@pytest.mark.parametrize("arch,size", [("ViT256", 64), ("SAM-B", 64)])
def test_forward_oracle_matches_reference_modules(ref, arch, size):
    cellvit, _, _ = ref
    torch.manual_seed(0)
    if arch == "ViT256":
        m = cellvit.CellViT256(None, 6, 19)
    else:
        m = cellvit.CellViTSAM(None, 6, 19, arch)
    m.eval()
    sd = weights.synth_state_dict(arch, 6, 19, seed=3)
    m.load_state_dict(sd, strict=True)  # also pins cellvit_b200.weights.state_spec against the reference
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=5))
    with torch.no_grad():
        r = m(x, retrieve_tokens=True)
    o = forward_oracle.cellvit_forward({k: v for k, v in sd.items()}, x, arch, retrieve_tokens=True)
    for k in ("tissue_types", "nuclei_binary_map", "hv_map", "nuclei_type_map", "tokens"):
        assert r[k].shape == o[k].shape
        assert (r[k] - o[k]).abs().max().item() <= 2e-6, k
