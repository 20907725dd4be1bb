"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU, NCCL): rank r processes tiles r, r+W, ...
with weights broadcast from rank 0; the per-tile results are gathered on rank 0 and compared with rank 0 processing ALL tiles alone."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cellvit_b200 import synth, weights  # noqa: E402
from cellvit_b200.cell_detection import CellSegmentationInference, broadcast_weights, shard_indices  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# every rank starts from DIFFERENT weights: only the broadcast can make the results agree
ckpt = {"arch": "CellViT256", "config": {"data.num_nuclei_classes": 6, "data.num_tissue_classes": 19, "model.backbone": "default"},
        "model_state_dict": weights.synth_state_dict("ViT256", 6, 19, seed=3 + rank)}
inf = CellSegmentationInference(ckpt, gpu=local)
broadcast_weights(inf.model, 0)
size, n_tiles = 256, 6
nuc = [synth.synthetic_nuclei(size, 25 + 3 * i, seed=90 + i) for i in range(n_tiles)]
lg = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
tiles = synth.synthetic_tiles(n_tiles, size, seed=4)


def run(idx):
    out = {}
    batches = ((torch.from_numpy(tiles[i:i + 1]).pin_memory(), i) for i in idx)
    ov = lambda i: {"nuclei_binary_map": torch.from_numpy(lg[i][0][None]).cuda(), "nuclei_type_map": torch.from_numpy(lg[i][1][None]).cuda(),
                    "hv_map": torch.from_numpy(nuc[i]["hv"][None]).cuda()}
    for i, dicts, toks in inf._pipeline(batches, 40, head_override=ov, with_tokens=True):
        out[i] = ({int(k): {f: np.asarray(v[f]).tolist() for f in ("bbox", "centroid", "contour", "type", "type_prob")} for k, v in dicts[0].items()},
                  toks[0].tobytes())
    return out


mine = run(shard_indices(n_tiles, rank, world))
gathered = [None] * world if rank == 0 else None
dist.gather_object(mine, gathered, dst=0)
ok = True
if rank == 0:
    merged = {}
    for g in gathered:
        merged.update(g)
    alone = run(range(n_tiles))
    ok = sorted(merged) == list(range(n_tiles)) and all(pickle.dumps(merged[i]) == pickle.dumps(alone[i]) for i in range(n_tiles))
    print("NCCL_EQUAL_OK" if ok else "NCCL_EQUAL_MISMATCH", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
