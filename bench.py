#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Headline benchmark: 1024x1024 tiles/s, CellViT inference + HoVer-Net post-processing (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--arch SAM-H|ViT256] [--batch B]

Default workload = BASELINE.json configs[2] (the configuration the metric is quoted on): CellViT-SAM-H, B = 4 synthetic
1024^2 RGB tiles per GPU per step, on-GPU post-processing. `--arch ViT256` (default batch 8) is configs[1].

A step = one batch through the hot path: cvb_forward (random-init weights of the reference architecture) + cvb_postproc
+ cvb_contours. Random-init networks emit spatially constant argmax maps, so -- as the survey
prescribes -- post-processing runs on seeded synthetic-nuclei head maps (700 nuclei / tile) that are resident on the
device in place of the head outputs; the forward still runs in full on the synthetic tiles every step.

 value    : tiles/s with inputs resident in HBM (forward + device post-processing: label maps, instance tables, contours;
            post-processing of batch k overlaps the forward of batch k+1 on a second stream, as in the product pipeline).
 overlap  : forward-only and post-processing-only step times measured in the same run; loss_ms = ms_per_step - forward_only_ms
            is what the post-processing costs on top of the forward it hides behind.
 e2e      : tiles/s through the public Python API (CellSegmentationInference.process_tiles) with HOST buffers: pinned H2D of
            the tiles and D2H of label maps + tables + contours every step, host dict building included.
 roofline : tile-engine kernel (tc_kernel / conv_patch_kernel, tcgen05) -- algorithmic FLOPs / sum of its CUDA-event launch
            times; `postproc` = HBM-bound post-processing kernels against the committed ncu DRAM-byte capture.
 cpu_baseline : the oracle port (oracle/, fp32 torch CPU forward + C post-processing) on the host cores, N=1 only.
--impl reference : the same oracle port as the reference's CPU path (the reference itself is Python and needs
 /root/reference, which does not exist on the GPU box). It uses no GPU.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 1024
N_NUCLEI = 700
# algorithmic GFLOP per tile (shared skip decoders once) and the share outside the tile-engine kernel
# (attention core Q K^T / P V / rel-pos einsums: SAM-H 28 windowed blocks x 5.3 + 4 global x 87.2; ViT-S 12 blocks x 25.8)
ARCHS = {
    "SAM-H": dict(batch=4, gflop=9816.8, gflop_attn=28 * 5.3 + 4 * 87.2, name="CellViT-SAM-H"),
    "ViT256": dict(batch=8, gflop=3378.9, gflop_attn=12 * 25.8, name="CellViT-256"),
}


def metric_name(arch):
    return f"1024x1024 tiles/sec ({ARCHS[arch]['name']} inference+postproc)"


def workload_name(arch):
    # the workload both arms run (config.workload); the arms differ only in how many tiles make one step
    return (f"{ARCHS[arch]['name']} inference + HV watershed post-processing on injected synthetic-nuclei head maps ({N_NUCLEI} nuclei/tile), "
            f"synthetic {TILE}x{TILE} tiles, random-init weights")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1397.5), d.get("hbm_gbs", 6450.6), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def cpu_tile_seconds(arch: str, threads: int, tiles: int = 1):
    """Oracle port on the host: fp32 torch forward (all threads) + C post-processing (1 thread), per 1024^2 tile."""
    import torch
    from cellvit_b200 import synth, weights
    from oracle import forward_oracle, postproc_oracle as po
    torch.set_num_threads(threads)
    sd = weights.synth_state_dict(arch, 6, 19, seed=0)
    nuc = synth.synthetic_nuclei(TILE, N_NUCLEI, 0)
    t_f = t_p = 0.0
    for i in range(tiles):
        x = torch.from_numpy(synth.synthetic_tiles(1, TILE, seed=i))
        t0 = time.perf_counter()
        forward_oracle.cellvit_forward(sd, x, arch, retrieve_tokens=True)
        t1 = time.perf_counter()
        pm = np.concatenate([nuc["nt"][..., None], nuc["np_bin"][..., None], nuc["hv"].transpose(1, 2, 0)], -1).astype(np.float64)
        po.DetectionCellPostProcessor(6, 40).post_process_cell_segmentation(pm)
        t2 = time.perf_counter()
        t_f += t1 - t0
        t_p += t2 - t1
    return t_f / tiles, t_p / tiles


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    budget = 150.0
    tf, tp = cpu_tile_seconds(args.arch, cores, 1)            # also the warm-up
    per = tf + tp
    steps = max(1, min(args.steps, int(budget // per)))
    t0 = time.perf_counter()
    tf2, tp2 = cpu_tile_seconds(args.arch, cores, steps)
    el = time.perf_counter() - t0
    v = steps / el
    sample = f"{steps} tile(s) of 1 (steps capped by a {budget:.0f}s budget); forward {tf2:.2f}s + postproc {tp2:.2f}s per tile"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.arch), "value": v, "unit": "tiles/s", "n_gpus": args.gpus, "gpus_used": 0,
        "steps": steps, "warmup": 1,
        "ms_per_step": 1000.0 * el / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": workload_name(args.arch),
                                        "batch": "1 tile per step on the host cores (bounded sample; CPU port of the reference, no GPU used)"},
        "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from cellvit_b200 import _lib as L
    from cellvit_b200 import synth
    from cellvit_b200.cellvit import CellViT256, CellViTSAM
    from cellvit_b200.post_proc_cellvit import MAX_PTS, ROWS_COPIED, DetectionCellPostProcessor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    lib.cvb_launch_count.restype = C.c_longlong
    if os.environ.get("CVB_SKIP_FLOOD") == "1":   # timing experiment: how much of the overlap loss the flood kernels cause
        lib.cvb_debug_postproc_skip_flood(1)
    if os.environ.get("CVB_FLOOD_LARGE_CTAS"):
        lib.cvb_debug_flood_large_ctas(int(os.environ["CVB_FLOOD_LARGE_CTAS"]))
    if os.environ.get("CVB_POST_CTAS"):
        lib.cvb_debug_postproc_max_ctas(int(os.environ["CVB_POST_CTAS"]))
    arch = args.arch
    cfg = ARCHS[arch]
    B, K, Wm = args.batch or cfg["batch"], args.steps, args.warmup
    gflop_tc = cfg["gflop"] - cfg["gflop_attn"]

    # ---- model: every rank builds the architecture, rank 0's weights are broadcast once over NCCL (collective C1)
    torch.manual_seed(0)
    model = (CellViTSAM(None, 6, 19, arch) if arch != "ViT256" else CellViT256(None, 6, 19)).eval().to(dev)
    if world > 1:
        with torch.no_grad():
            tensors = list(model.parameters()) + [b for b in model.buffers() if b.dtype.is_floating_point]
            flat = torch.cat([p.detach().reshape(-1).float() for p in tensors])
            dist.broadcast(flat, 0)
            o = 0
            for t in tensors:
                n = t.numel(); t.copy_(flat[o:o + n].view_as(t)); o += n
            del flat
        model.invalidate_packed()
    proc = DetectionCellPostProcessor(nr_types=6, magnification=40)

    # ---- inputs: rank r owns tiles r, r+world, ... of the synthetic stream (weak scaling: B tiles per rank per step)
    tiles_host = torch.from_numpy(synth.synthetic_tiles(B, TILE, seed=1000 + rank)).pin_memory()
    x_dev = tiles_host.to(dev)
    nuc = [synth.synthetic_nuclei(TILE, N_NUCLEI, seed=rank * B + i) for i in range(B)]
    logits = [synth.head_logits_from_maps(n["np_bin"], n["nt"], 6) for n in nuc]
    np_dev = torch.from_numpy(np.stack([l[0] for l in logits])).to(dev)
    nt_dev = torch.from_numpy(np.stack([l[1] for l in logits])).to(dev)
    hv_dev = torch.from_numpy(np.stack([n["hv"] for n in nuc])).to(dev)
    # what the fused head epilogue hands to the post-processing in the product pipeline: uint8 arg-max planes (K12 fusion); here
    # derived once from the injected logits, the forward below still writes its own planes every step
    np_arg_dev = (np_dev[:, 1] > np_dev[:, 0]).to(torch.uint8).contiguous()
    nt_arg_dev = nt_dev.argmax(1).to(torch.uint8).contiguous()

    # post-processing on a default (= lowest) priority stream; the forward graph's kernels carry high priority (cellvit.py)
    s_post = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)

    gather = None
    if world > 1 and args.gather_every > 1:
        from cellvit_b200.cell_detection import TableGather
        gather = TableGather(world, B, 1024, 88, args.gather_every, dev)

    def flush_gather():
        if gather is not None:
            with torch.cuda.stream(s_post):
                gather.flush()

    def forward_once():
        with torch.no_grad():
            if args.graphs:
                model.forward_graphed(x_dev, retrieve_tokens=True, slot=0, argmax_maps=True)
            else:
                model(x_dev, retrieve_tokens=True, argmax_maps=True)

    def post_once(exchange=True):
        w = proc._workspace(B, TILE, TILE, dev)
        L.check(lib.cvb_postproc_argmax(L.ptr(np_arg_dev), L.ptr(hv_dev), L.ptr(nt_arg_dev), B, TILE, TILE, 6, 40, L.ptr(w.labels), L.ptr(w.table),
                                        L.ptr(w.counts), proc.max_rows, L.ptr(w.ws), C.c_size_t(w.ws.numel()), L.stream_ptr()), "cvb_postproc_argmax")
        w.launch_contours(B, TILE, TILE, proc.max_rows)   # per-instance contours on the device (cvb_contours)
        if not exchange:
            return
        if gather is not None:  # collective C2 staged and exchanged once per --gather-every steps
            gather.add(w.counts, w.table)
        elif world > 1:  # collective C2 per step: all-gather of the per-tile instance tables (counts, then the first 1024 rows)
            cnts = [torch.empty_like(w.counts) for _ in range(world)]
            dist.all_gather(cnts, w.counts)
            part = w.table[:, :1024].contiguous()
            tabs = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(tabs, part)

    def step_device():
        # forward of batch k on the main stream; post-processing of batch k on a second stream once that forward has
        # finished (the dependency of the real pipeline), so it overlaps the forward of batch k+1 -- the same
        # structure as CellSegmentationInference.process_tiles
        forward_once()
        if args.no_post:
            return
        fwd_done = torch.cuda.Event()
        fwd_done.record(main)
        s_post.wait_event(fwd_done)
        with torch.cuda.stream(s_post):
            post_once()

    from cellvit_b200.cell_detection import CellSegmentationInference
    inf = CellSegmentationInference.from_model(model, local)
    override = {"nuclei_binary_argmax": np_arg_dev, "hv_map": hv_dev, "nuclei_type_argmax": nt_arg_dev}  # injected synthetic nuclei

    def run_e2e(n_batches):
        # public API: pinned-host tiles in, per-tile instance dicts out (H2D, forward, device post-processing,
        # D2H of label maps + tables + contours, host dicts; host work of batch k overlaps device work of batch k+1)
        res = inf.process_tiles([tiles_host] * n_batches, magnification=40, head_override=override, use_graphs=bool(args.graphs))
        return sum(len(d) for d in res[-1])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n, flush=None):
        """n calls of fn between two events on the main stream (the post stream is joined before the second one)."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        if flush is not None:
            flush()
        main.wait_stream(s_post)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b)

    # launches of one forward (a graph replay does not pass through the launch counter)
    lib.cvb_launch_count(1)
    with torch.no_grad():
        model(x_dev, retrieve_tokens=True)
    fwd_launches = int(lib.cvb_launch_count(1))
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(Wm):
        step_device()
    flush_gather()
    main.wait_stream(s_post)
    sync_all()
    lib.cvb_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps_done = K
    if args.dynamic_queue and world > 1:
        # C4 mode: the K * world batches of the stream sit in ONE queue (a counter in the c10d store of the rendezvous); a rank
        # takes the next batch whenever its host loop is free, so a rank that falls behind (larger blobs, lower clocks) simply
        # processes fewer batches. Every rank enqueues at most 2 steps ahead of its GPU, as the product pipeline does.
        store = dist.distributed_c10d._get_default_store()
        total = K * world
        steps_done = 0
        inflight = []
        slow = float(os.environ.get("CVB_SLOW_RANK_MS", "0")) / 1e3 if rank == world - 1 else 0.0   # test hook: a straggling rank
        while int(store.add("cvb_bench_next_batch", 1)) <= total:
            if slow:
                time.sleep(slow)
            step_device()
            steps_done += 1
            ev = torch.cuda.Event()
            ev.record(main)
            inflight.append(ev)
            if len(inflight) > 2:
                inflight.pop(0).synchronize()
        # the staged-table exchange needs the same number of add() calls on every rank: pad with empty steps. The ranks agree on
        # the maximum through the STORE, not through a collective: an NCCL call here would be ordered differently against the
        # pending chunk all-gathers on a rank that ran more steps (collectives of one communicator must be issued in the same
        # order everywhere -- an all-reduce at this point deadlocked the first 8-GPU run)
        if gather is not None:
            store.set(f"cvb_bench_steps_{rank}", str(steps_done))
            n_max = max(int(store.get(f"cvb_bench_steps_{r}")) for r in range(world))
            with torch.cuda.stream(s_post):
                w = proc._workspace(B, TILE, TILE, dev)
                for _ in range(n_max - steps_done):
                    gather.add(w.counts, w.table)
    else:
        for _ in range(K):
            step_device()
    flush_gather()            # (chunked exchange: the tail of the stream is gathered inside the timed region)
    main.wait_stream(s_post)  # the timed region ends when the last batch's post-processing has finished
    e1.record()
    sync_all()
    launches = int(lib.cvb_launch_count(1)) + (K * fwd_launches if args.graphs else 0)
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    done = torch.tensor([steps_done], device=dev)
    if world > 1:
        dist.all_reduce(done)
    total_steps = int(done.item())                    # == K * world (static shards: K per rank)
    value = B * total_steps / (ms_total / 1000.0)

    # ---- what the overlap costs: the same step without post-processing, and the post-processing alone (rank-local; max over ranks)
    Ko = max(3, min(K, 10))
    ov = None
    if not args.no_post:
        sync_all()
        f_ms = timed(forward_once, Ko) / Ko

        def post_alone():
            with torch.cuda.stream(s_post):
                post_once(exchange=False)
        p_ms = timed(post_alone, Ko) / Ko
        c_ms = None
        if world > 1:
            def exch():
                with torch.cuda.stream(s_post):
                    w = proc._workspace(B, TILE, TILE, dev)
                    if gather is not None:
                        gather.add(w.counts, w.table)
                    else:
                        cnts = [torch.empty_like(w.counts) for _ in range(world)]
                        dist.all_gather(cnts, w.counts)
                        part = w.table[:, :1024].contiguous()
                        tabs = [torch.empty_like(part) for _ in range(world)]
                        dist.all_gather(tabs, part)
            sync_all()
            n_ex = max(Ko, args.gather_every * 2)
            c_ms = timed(exch, n_ex, flush_gather) / n_ex
            t3 = torch.tensor([f_ms, p_ms, c_ms], device=dev)
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
            f_ms, p_ms, c_ms = (float(v) for v in t3.tolist())
        ov = {"forward_only_ms": f_ms, "post_only_ms": p_ms, "loss_ms": ms_total / K - f_ms, "steps": Ko,
              "collective_ms_per_step": c_ms,
              "collective": None if world == 1 else ("ncclAllGather of the staged instance tables once per %d steps (TableGather)" % args.gather_every
                                                     if gather is not None else "two ncclAllGather per step (counts, then 1024 table rows)")}

    # ---- e2e through the Python API with host buffers
    e2e = None
    if not args.no_post:
        run_e2e(2)
        sync_all()
        Ke = max(2, min(K, 20))
        t0 = time.perf_counter()
        run_e2e(Ke)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        rows = min(ROWS_COPIED, proc.max_rows)
        e2e = {"value": world * B * Ke / float(te.item()), "unit": "tiles/s", "h2d_bytes_per_step": tiles_host.numel() * 4,
               # label maps + counts + per tile the eagerly copied rows of the instance table, contour points and point counts
               "d2h_bytes_per_step": B * TILE * TILE * 4 + B * 4 + B * rows * (88 + MAX_PTS * 2 * 2 + 4), "steps": Ke,
               "api": "CellSegmentationInference.process_tiles: H2D + forward + device post-processing + contours + D2H + host dicts (3 streams, 2-deep pipeline)"}

    # ---- roofline leg: per-launch CUDA-event timing of the tile-engine kernel over Kp more steps (not under a profiler)
    roof = None
    cpu_base = None
    if rank == 0:
        Kp = max(1, min(K, 3))
        tc_ms, tc_n, tc_fl = C.c_double(), C.c_int(), C.c_double()
        tot_ms, tot_n, tot_fl = 0.0, 0, 0.0
        for _ in range(Kp):
            L.check(lib.cvb_tc_profile_begin(4096), "cvb_tc_profile_begin")
            with torch.no_grad():
                model(x_dev, retrieve_tokens=True)
            L.check(lib.cvb_tc_profile_end(C.byref(tc_ms), C.byref(tc_n), C.byref(tc_fl)), "cvb_tc_profile_end")
            tot_ms += tc_ms.value; tot_n += tc_n.value; tot_fl += tc_fl.value
        peak, hbm, how = _peaks()
        achieved = gflop_tc * B * Kp / (tot_ms / 1000.0) / 1000.0  # TFLOP/s
        traffic, traffic_src = None, None
        for tname in ("r2_traffic.json", "r1_traffic.json"):   # the newest committed ncu pass of this same step (tile-engine launches)
            tpath = os.path.join(ROOT, "profiles", tname)
            if arch == "SAM-H" and B == 4 and os.path.exists(tpath):  # DRAM bytes per tile-engine launch
                tj = json.load(open(tpath))
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = f"profiles/{tname} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean over the step's {tj.get('launches')} tile-engine launches)"
                break
        roof = {"bound": "tensor", "kernel": "tc_kernel / conv_patch_kernel (tcgen05 GEMM / implicit-GEMM conv)", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak, "peak_source": how, "traffic": traffic, "traffic_unit": "bytes/launch",
                "traffic_source": traffic_src,
                "launches_per_step": tot_n // Kp, "kernel_ms_per_step": tot_ms / Kp,
                "executed_tflop_per_step": tot_fl / Kp / 1e12, "algorithmic_tflop_per_step": gflop_tc * B / 1000.0,
                "share_of_step": (tot_ms / Kp) / (ms_total / K),
                "whole_step_frac": (cfg["gflop"] * B / (ms_total / K)) / peak}  # all algorithmic FLOPs of the step / step time / peak
        # HBM-bound post-processing kernels: committed ncu captures at this shape (tools/postproc_traffic.py). The arg-max entry
        # (cvb_postproc_argmax, what this bench and the product pipeline run) first, the float entry's older capture as fallback.
        for pname in ("r2_postproc_traffic_k12.json", "r2_postproc_traffic.json"):
            ppath = os.path.join(ROOT, "profiles", pname)
            if os.path.exists(ppath):
                roof["postproc"] = json.load(open(ppath))
                break
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            tf, tp = cpu_tile_seconds(arch, cores, 1)
            cpu_base = {"value": 1.0 / (tf + tp), "unit": "tiles/s", "cores": cores, "kind": "port",
                        "sample": f"1 tile: oracle fp32 forward {tf:.2f}s ({cores} threads) + C post-processing {tp:.2f}s (1 thread)"}
            # informative second baseline: how the reference itself runs on a GPU box (cell_detection.py:306-323): eager PyTorch
            # forward under fp16 autocast on the GPU, then the per-tile post-processing on ONE host thread, strictly in sequence
            try:
                from cellvit_b200 import weights
                from oracle import forward_oracle
                del model
                torch.cuda.empty_cache()
                sd_dev = {k: v.to(dev) for k, v in weights.synth_state_dict(arch, 6, 19, seed=0).items()}
                n_t = 4
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                    forward_oracle.cellvit_forward(sd_dev, x_dev[:1], arch, retrieve_tokens=True)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for i in range(n_t):
                        forward_oracle.cellvit_forward(sd_dev, x_dev[i % B:i % B + 1], arch, retrieve_tokens=True)
                    torch.cuda.synchronize()
                    t_fwd = (time.perf_counter() - t0) / n_t
                cpu_base["gpu_eager"] = {"value": 1.0 / (t_fwd + tp), "unit": "tiles/s", "kind": "port",
                                         "sample": f"oracle forward in eager PyTorch on this GPU under fp16 autocast {t_fwd * 1e3:.0f} ms/tile "
                                                   f"+ C post-processing on one host thread {tp:.2f}s/tile, sequential (the reference's own GPU mode)"}
            except Exception as ex:   # informative only
                cpu_base["gpu_eager"] = {"unavailable": repr(ex)[:200]}
    if world > 1:
        dist.barrier()
    if rank == 0:
        par = f"tiles sharded over {world} GPU(s), NCCL weight broadcast"
        if world > 1:
            par += (", per-step all-gather of instance tables" if gather is None else f", instance tables all-gathered every {args.gather_every} steps")
        print(json.dumps({
            "metric": metric_name(arch), "value": value, "unit": "tiles/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic",
            "config": {"workload": workload_name(arch),
                       "batch": f"{B} tiles per GPU per step" + (", forward only (--no-post)" if args.no_post else ", post-processing on the GPU"),
                       "l2": "per-step working set (fp16 weights + >10 GB activations) exceeds the 126 MB L2; no explicit flush",
                       "forward_launch": "CUDA graph replay (kernel nodes at high stream priority)" if args.graphs else "eager",
                       "parallelism": par},
            "clocks": clocks, "timed_region_s": ms_total / 1000.0, "tiles_processed": B * total_steps,
            "queue": "dynamic (one batch counter in the c10d store)" if (args.dynamic_queue and world > 1) else "static round-robin shards",
            "overlap": ov, "e2e": e2e,
            "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu_base}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arch", default="SAM-H", choices=sorted(ARCHS), help="SAM-H (default; BASELINE configs[2]) or ViT256 (configs[1], batch 8)")
    ap.add_argument("--batch", type=int, default=0, help="tiles per GPU per step (default: 4 for SAM-H, 8 for ViT256)")
    ap.add_argument("--graphs", type=int, default=1, help="1 (default): the forward is replayed from a CUDA graph, as in the product pipeline; 0: eager launches")
    ap.add_argument("--gather-every", type=int, default=8, help="N>1 only: exchange the instance tables once per this many steps (TableGather, default 8); "
                    "1 = two all-gathers per step")
    ap.add_argument("--dynamic-queue", action="store_true", help="N>1 only (BASELINE configs[3]): the steps * N batches are handed out from one queue "
                    "instead of static round-robin shards; ms_per_step is then the job time / steps")
    ap.add_argument("--no-post", action="store_true", help="forward only: no post-processing, no e2e leg (quantifies what the post-processing costs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling / quick iteration runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
