"""Generates the golden fixtures in this directory by running the UNMODIFIED reference from /root/reference
(through oracle/ref_shim.py) in the dev container. The GPU box has no /root/reference; tests there use
these files. Re-run:  python tests/golden/make_golden.py

Fixtures (all derived from seeds, so only outputs are stored):
  postproc_<i>.npz : reference post-processing of cellvit_b200.synth.synthetic_nuclei(size, n, seed, noise)
                     -> labels, blb, marker, dist (captured at the skimage.watershed call site,
                     post_proc_cellvit.py:247) and the instance table fields (bbox, centroid, type, type_prob).
                     NOTE: the flood itself (P7) is the oracle restatement -- parity unpinned for that stage.
  forward_<arch>.npz : reference nn.Module forward on synthetic_tiles with weights.synth_state_dict(arch, seed=3)
  forward_<tag>.npz  : the same for a non-square tile (48 x 80) and for the *Shared modules (cellvit_shared.py), params =
                       [H, W, tile seed, weight seed, shared, regression_loss]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cellvit_b200 import synth, weights  # noqa: E402
from oracle import postproc_oracle as po, ref_shim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
POSTPROC_CASES = [(128, 14, 0, 40, 0.0), (128, 20, 1, 20, 0.0), (192, 30, 2, 40, 0.03), (256, 48, 3, 40, 0.0)]
FORWARD_CASES = [("ViT256", 64, 5), ("SAM-B", 64, 5)]
# (file tag, arch, tile size, shared-decoder variant, regression_loss): non-square / non-native tile sizes and the *Shared modules
FORWARD_CASES_EXTRA = [("ViT256_48x80", "ViT256", (48, 80), False, False), ("ViT256_shared", "ViT256", 64, True, False),
                       ("SAM-B_shared_80", "SAM-B", 80, True, True)]


def main():
    cap = {}

    def spy(image, markers=None, mask=None):
        cap["dist"], cap["marker"], cap["mask"] = np.array(image), np.array(markers), np.array(mask)
        return po.watershed(image, markers=markers, mask=mask)

    cellvit, post = ref_shim.import_reference(spy)
    for i, (size, n, seed, mag, noise) in enumerate(POSTPROC_CASES):
        d = synth.synthetic_nuclei(size, n, seed, noise=noise)
        pm = np.concatenate([d["nt"][..., None], d["np_bin"][..., None], d["hv"].transpose(1, 2, 0)], -1).astype(np.float64)
        lab, dct = post.DetectionCellPostProcessor(6, mag).post_process_cell_segmentation(pm)
        ids = np.array(sorted(dct), np.int32)
        np.savez_compressed(
            os.path.join(HERE, f"postproc_{i}.npz"), params=np.array([size, n, seed, mag], np.int64), noise=np.float64(noise),
            labels=lab.astype(np.int32), blb=cap["mask"].astype(np.uint8), marker=cap["marker"].astype(np.int32),
            dist=cap["dist"].astype(np.float64), ids=ids,
            bbox=np.array([dct[k]["bbox"] for k in ids], np.int32).reshape(-1, 2, 2),
            centroid=np.array([dct[k]["centroid"] for k in ids], np.float64).reshape(-1, 2),
            type=np.array([dct[k]["type"] for k in ids], np.int32), type_prob=np.array([dct[k]["type_prob"] for k in ids], np.float64),
            contour_len=np.array([len(dct[k]["contour"]) for k in ids], np.int32))
    for arch, size, seed in FORWARD_CASES:
        m = (cellvit.CellViT256(None, 6, 19) if arch == "ViT256" else cellvit.CellViTSAM(None, 6, 19, arch)).eval()
        m.load_state_dict(weights.synth_state_dict(arch, 6, 19, seed=3), strict=True)
        x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=seed))
        with torch.no_grad():
            r = m(x, retrieve_tokens=True)
        np.savez_compressed(os.path.join(HERE, f"forward_{arch}.npz"), params=np.array([size, seed, 3], np.int64),
                            **{k: v.numpy() for k, v in r.items()})
    import importlib
    shared_mod = importlib.import_module("models.segmentation.cell_segmentation.cellvit_shared")
    for tag, arch, size, shared, regression in FORWARD_CASES_EXTRA:
        mod = shared_mod if shared else cellvit
        cls = {("ViT256", False): "CellViT256", ("ViT256", True): "CellViT256Shared", ("SAM-B", False): "CellViTSAM", ("SAM-B", True): "CellViTSAMShared"}
        ctor = getattr(mod, cls[(arch, shared)])
        m = (ctor(None, 6, 19, regression_loss=regression) if arch == "ViT256" else ctor(None, 6, 19, arch, regression_loss=regression)).eval()
        m.load_state_dict(weights.synth_state_dict(arch, 6, 19, seed=3, regression_loss=regression, shared=shared), strict=True)
        x = torch.from_numpy(synth.synthetic_tiles(1, size, seed=5))
        with torch.no_grad():
            r = m(x, retrieve_tokens=True)
        hw = (size, size) if isinstance(size, int) else size
        np.savez_compressed(os.path.join(HERE, f"forward_{tag}.npz"), params=np.array([hw[0], hw[1], 5, 3, int(shared), int(regression)], np.int64),
                            arch=np.array(arch), **{k: v.numpy() for k, v in r.items()})


if __name__ == "__main__":
    main()
    print("golden fixtures written to", HERE)
