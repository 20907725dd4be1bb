"""Oracle vs the committed golden fixtures (generated from the unmodified reference by
tests/golden/make_golden.py). Runs anywhere -- no /root/reference, no GPU."""
import glob
import os

import numpy as np
import pytest
import torch

from cellvit_b200 import synth, weights
from oracle import forward_oracle, postproc_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "postproc_*.npz"))))
def test_postproc_oracle_vs_golden(path):
    g = np.load(path)
    size, n, seed, mag = (int(v) for v in g["params"])
    d = synth.synthetic_nuclei(size, n, seed, noise=float(g["noise"]))
    lab, inter = po.proc_np_hv(d["np_bin"], d["hv"], mag, want_intermediates=True)
    assert np.array_equal(inter["blb"], g["blb"])
    assert np.array_equal(inter["marker"], g["marker"])
    assert np.array_equal(inter["dist"], g["dist"])
    assert np.array_equal(lab, g["labels"])
    rows = po.instance_table(lab, d["nt"], 6)
    dct = po.rows_to_dict(lab, rows)
    assert np.array_equal(np.array(sorted(dct), np.int32), g["ids"])
    for i, k in enumerate(g["ids"]):
        o = dct[np.int32(k)]
        assert np.array_equal(o["bbox"], g["bbox"][i])
        assert np.array_equal(o["centroid"], g["centroid"][i])
        assert o["type"] == g["type"][i] and o["type_prob"] == g["type_prob"][i]
        assert len(o["contour"]) == g["contour_len"][i]


@pytest.mark.parametrize("arch", ["ViT256", "SAM-B"])
def test_forward_oracle_vs_golden(arch):
    g = np.load(os.path.join(GOLD, f"forward_{arch}.npz"))
    size, seed, wseed = (int(v) for v in g["params"])
    sd = weights.synth_state_dict(arch, 6, 19, seed=wseed)
    x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=seed))
    o = forward_oracle.cellvit_forward(sd, x, arch, retrieve_tokens=True)
    for k in ("tissue_types", "nuclei_binary_map", "hv_map", "nuclei_type_map", "tokens"):
        assert np.abs(o[k].numpy() - g[k]).max() <= 5e-6, k
