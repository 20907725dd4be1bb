"""End-to-end API on the GPU: CellSegmentationInference.process_tiles / calculate_instance_map produce the
reference's per-tile dict layout and match the oracle exactly when the head maps are identical."""
import numpy as np
import pytest
import torch

from cellvit_b200 import synth, weights
from oracle import postproc_oracle as po

pytestmark = pytest.mark.gpu


def _same_dict(g, o):
    assert sorted(g) == sorted(o)
    for k, ov in o.items():
        gv = g[k]
        assert np.array_equal(gv["bbox"], ov["bbox"]) and np.array_equal(gv["centroid"], ov["centroid"])
        assert np.array_equal(gv["contour"], ov["contour"]) and gv["type"] == ov["type"] and gv["type_prob"] == ov["type_prob"]


def test_process_tiles_pipeline_matches_oracle():
    from cellvit_b200.cell_detection import CellSegmentationInference
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)
    B, size = 2, 256
    nuc = [synth.synthetic_nuclei(size, 35 + 5 * i, seed=40 + i) for i in range(B)]
    lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
    override = {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg])).cuda(),
                "hv_map": torch.from_numpy(np.stack([n["hv"] for n in nuc])).cuda()}
    tiles = torch.from_numpy(synth.synthetic_tiles(B, size, seed=9)).pin_memory()
    for threads in (0, 4):
        res = inf.process_tiles([tiles, tiles, tiles], magnification=40, head_override=override, host_threads=threads)
        assert len(res) == 3
        for batch in res:
            for b in range(B):
                pm = np.concatenate([nuc[b]["nt"][..., None], nuc[b]["np_bin"][..., None], nuc[b]["hv"].transpose(1, 2, 0)], -1).astype(np.float64)
                _, odict = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
                _same_dict(batch[b], odict)
    # reference-style call sequence: forward -> get_cell_predictions_with_tokens
    with torch.no_grad():
        pred = inf.model(tiles.cuda(), retrieve_tokens=True)
    pred.update(override)
    inst, tokens = inf.get_cell_predictions_with_tokens(pred, magnification=40)
    assert tokens.device.type == "cpu" and tuple(tokens.shape) == (B, 384, 16, 16)
    assert len(inst) == B and len(inst[0]) > 10
    lab, dicts = inf.model.calculate_instance_map(pred, 40)
    assert lab.dtype == torch.float32 and lab.device.type == "cpu" and tuple(lab.shape) == (B, size, size)
    nuc_map = inf.model.generate_instance_nuclei_map(lab, dicts)
    assert tuple(nuc_map.shape) == (B, 6, size, size)


def test_pipeline_ragged_last_batch_and_graph_replay():
    """Batches of 3, 3 and 1 tiles through process_tiles with and without CUDA-graph replay: identical dicts, and the
    static graph outputs of a slot are not clobbered before its batch has been collected."""
    from cellvit_b200.cell_detection import CellSegmentationInference
    ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
            "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3)}
    inf = CellSegmentationInference(ckpt, gpu=0)
    size = 256
    nuc = [synth.synthetic_nuclei(size, 30 + 3 * i, seed=70 + i) for i in range(7)]
    lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
    batches = [(0, 3), (3, 6), (6, 7)]
    tiles = [torch.from_numpy(synth.synthetic_tiles(b - a, size, seed=a)).pin_memory() for a, b in batches]

    def override_for(k):
        a, b = batches[k]
        return {"nuclei_binary_map": torch.from_numpy(np.stack([l[0] for l in lg[a:b]])).cuda(),
                "nuclei_type_map": torch.from_numpy(np.stack([l[1] for l in lg[a:b]])).cuda(),
                "hv_map": torch.from_numpy(np.stack([n["hv"] for n in nuc[a:b]])).cuda()}

    outs = {}
    for graphs in (True, False):
        res = [d for _, d, _ in inf._pipeline(((t, k) for k, t in enumerate(tiles)), 40, head_override=override_for, use_graphs=graphs)]
        assert [len(r) for r in res] == [3, 3, 1]
        outs[graphs] = res
    flat = lambda res: [d for batch in res for d in batch]
    for i, (g, e) in enumerate(zip(flat(outs[True]), flat(outs[False]))):
        _same_dict(g, e)
        pm = np.concatenate([nuc[i]["nt"][..., None], nuc[i]["np_bin"][..., None], nuc[i]["hv"].transpose(1, 2, 0)], -1).astype(np.float64)
        _, odict = po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
        _same_dict(g, odict)
