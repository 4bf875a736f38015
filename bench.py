#!/usr/bin/env python
"""Benchmark of the DGCNN kNN-graph + EdgeConv stack (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = forward + backward of the three EdgeConv layers (dynamic kNN graph -> edge
MLP -> GroupNorm -> LeakyReLU -> max over k, M4:488-505) over one batch of synthetic
ABC-shaped clouds: B = 16 clouds x 10 000 points, k = 50, mode 0 (configs[1] of
BASELINE.json) per GPU; with N > 1 the batch is sharded by cloud (weak scaling: 16
clouds per GPU) and the weight gradients are all-reduced once per step over NCCL.

The timed step is one CUDA-graph replay of forward + backward (same kernels as the eager call sequence, run as
`--streams` half-batch branches on separate streams inside the graph; `--no-graph` times the eager sequence, whose
per-call breakdown is reported either way), followed by the all-reduce when N > 1.

Prints ONE JSON line (rank 0).  `value` = clouds/s with inputs resident in HBM, timed with
CUDA events, max over ranks; `e2e` = the same step driven from pinned HOST buffers through
the public API (H2D of the batch and D2H of the loss inside the timed region);
`roofline` = the dominant kernel (the feature-space kNN scan, timed alone by the library's event pair) against the
measured tensor peak, with the whole kNN call beside it (`roofline.call`);
`cpu_baseline` = the oracle (a torch-CPU restatement of the reference path) on this box's
host cores, on a bounded sample.  `--impl reference` times that CPU path alone.

Extra keys on the same line (skipped with `--no-extras`, all outside the timed region of the headline):
`gpu_reference` (the oracle's torch path on this GPU), `config3_knn_sweep`, `config4_full_step` (per-point training
step, mode 5), `config4_fixed_global_batch` (128 clouds over N GPUs), `config5_large_clouds` (4 x 100 k points),
`bf16_storage_mode`, and for N > 1 `allreduce_check` (the collective checked against an all-gather) and `multi_gpu`
(every rank's own step time, the spread between ranks, what remains for the collective).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "point clouds/sec (10k pts, k=50) fwd+bwd"
UNIT = "clouds/s"
B_PER_GPU, NPTS, KNN = 16, 10000, 50
# algorithmic work per cloud (N=1e4, k=50, mode-0 stack)
GFLOP_PER_CLOUD = 101.1
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                d = json.load(f)
            d["_source"] = "measured"
            return d
        except Exception:
            pass
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------- CPU path
def cpu_reference_clouds_per_s(steps: int, warmup: int, threads: int):
    """The oracle's DGCNNEncoderGn.edge_stack forward+backward, one 10k-point cloud per step."""
    from oracle import dgcnn_oracle as orc
    from gcanet_b200.synth import abc_like_batch
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    enc = orc.DGCNNEncoderGn(mode=0, nn_nb=KNN, input_channels=6)
    x = torch.from_numpy(abc_like_batch(1, NPTS, seed=1234))
    cot = [torch.randn(1, c, NPTS) for c in (64, 64, 128)]

    def step():
        enc.zero_grad(set_to_none=True)
        outs = enc.edge_stack(x)
        torch.autograd.backward(outs, cot)            # upstream gradients are given, as for ours
        return float(outs[2].detach().sum())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    warm = max(1, args.warmup)                 # the driver compares the line's warm-up count with the one it asked for
    cps, sec = cpu_reference_clouds_per_s(args.steps, warm, threads)
    sample = f"{args.steps} steps x 1 cloud (10k pts, k=50, mode 0) fwd+bwd, {cpu_model_name()}"
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DGCNN kNN+EdgeConv stack fwd+bwd, 10k pts, k=50, mode 0 (CPU: one cloud per step)",
                   "batch_per_step": 1, "points": NPTS, "k": KNN},
        "cpu_baseline": {"value": cps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": cps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference torch path restated in oracle/dgcnn_oracle.py (bit-identical to the reference source on "
                "the golden fixtures); /root/reference is not present on the GPU box and its model file cannot be "
                "imported (spconv etc. absent)",
    }
    print(json.dumps(line), flush=True)
    return 0



# ---------------------------------------------------------------------------- same-GPU reference + other configs
def gpu_reference_leg(dev):
    """The reference torch path (oracle restatement, bit-identical to the reference source on the golden fixtures)
    on THIS GPU: DGCNNEncoderGn.edge_stack forward+backward, mode 0, k = 50, B as large as fits (the path
    materialises [B,N,N] distances and [B,2C,N,k] edge tensors, M4:36-41, M4:120-123).  SURVEY 8(d) / BASELINE.md 3.3:
    the GPU-vs-GPU bar.  torch defaults (matmul fp32, cuDNN conv TF32 allowed), as the reference would run."""
    from oracle import dgcnn_oracle as orc
    from gcanet_b200.synth import abc_like_batch
    torch.manual_seed(0)
    enc = orc.DGCNNEncoderGn(mode=0, nn_nb=KNN, input_channels=6).to(dev)
    for B in (16, 8, 4, 2, 1):
        try:
            x = torch.from_numpy(abc_like_batch(B, NPTS, seed=1234)).to(dev)
            cot = [torch.randn(B, c, NPTS, device=dev) for c in (64, 64, 128)]

            def step():
                enc.zero_grad(set_to_none=True)
                outs = enc.edge_stack(x)
                torch.autograd.backward(outs, cot)

            step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                step()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 3
            peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
            del x, cot
            torch.cuda.empty_cache()
            return {"value": B / (ms / 1e3), "unit": UNIT, "batch": B, "ms_per_step": ms, "steps": 3,
                    "peak_memory_gb": round(peak_gb, 1),
                    "what": "oracle (restated reference torch path: matmul + topk + gather + conv2d + group_norm + max, "
                            "autograd backward) on cuda, fp32, same clouds / k / layers as `value`"}
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            continue
    return None


def _median_ms(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def extra_configs(enc, dev, G):
    """BASELINE configs[2] (kNN sweep C = 3/64/128, k = 20/50, B = 16 x 10k, the public `knn` call: ordered int64 lists)
    and configs[4] (large-cloud stress, B = 4 x 100k, k = 50: the three-layer stack forward+backward)."""
    import gcanet_b200 as gb
    from gcanet_b200.synth import abc_like_batch
    out = {}
    x = torch.from_numpy(abc_like_batch(B_PER_GPU, NPTS, seed=1234)).to(dev)
    with torch.no_grad():
        x1, x2, x3 = enc.edge_stack(x)
    sweep = {}
    for C, t in ((3, x), (64, x1.contiguous()), (128, x3.contiguous())):
        for k in (20, 50):
            ms = _median_ms(lambda: gb.knn(t, k, k))
            sweep[f"C={C},k={k}"] = {"ms": round(ms, 4), "tflops_algorithmic": round(2.0 * NPTS * NPTS * C * B_PER_GPU / ms / 1e9, 1)}
    out["config3_knn_sweep"] = {"workload": "knn(x,k,k), B=16 x 10k pts; xyz clouds (C=3), layer-1 (C=64) and layer-3 (C=128) "
                                            "activations of this encoder; nearest-first int64 lists; median of 7 calls",
                                "results": sweep}
    # the materialising API-parity entry points: pure data movement, quoted against HBM
    peaks = load_peaks()
    hbm = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    api = {}
    with torch.no_grad():
        idx20 = gb.knn(x1.contiguous(), 20, 20)
        idx3 = gb.knn(x, 20, 20)
        for name, t, idx, cch in (("get_graph_feature[C=64,k=20]", x1.contiguous(), idx20, 64), ("get_graph_feature[C=3,k=20]", x, idx3, 3)):
            ms = _median_ms(lambda: gb.get_graph_feature(t, 20, 20, idx=idx))
            wr = B_PER_GPU * NPTS * 20 * 2 * cch * 4
            api[name] = {"ms": round(ms, 4), "gb_written": round(wr / 1e9, 3), "frac_of_hbm_peak": round(wr / ms / 1e6 / hbm, 3)}
        i32 = idx20[:, :, :16].to(torch.int32).contiguous()
        feats = x1.contiguous()
        ms = _median_ms(lambda: gb.grouping_operation(feats, i32))
        wr = B_PER_GPU * 64 * NPTS * 16 * 4
        api["grouping_operation[C=64,np=10000,ns=16]"] = {"ms": round(ms, 4), "gb_written": round(wr / 1e9, 3),
                                                         "frac_of_hbm_peak": round(wr / ms / 1e6 / hbm, 3)}
        ref3 = x[:, :3].contiguous()
        ms = _median_ms(lambda: gb.knn_cuda(ref3, ref3[:, :, :2048].contiguous(), 3))
        api["KNN_CUDA[dim=3,Nr=10000,Nq=2048,k=3]"] = {"ms": round(ms, 4)}
    out["api_parity_kernels"] = {"workload": "B=16 x 10k pts; bytes = the result tensor the reference's signature requires (written once)",
                                 "results": api}
    del x1, x2, x3, idx20, idx3, i32, feats
    # bf16-storage mode of the same stack step ("bf16 on 1xB200" of configs[1]); fp32 stays the headline (parity mode)
    cot = [torch.randn(B_PER_GPU, c, NPTS, device=dev) for c in (64, 64, 128)]
    hot16 = [p for n, p in enc.named_parameters() if n.split(".")[0] in ("conv1", "conv2", "conv3", "bn1", "bn2", "bn3")]

    def step16():
        for p in hot16:
            p.grad = None
        torch.autograd.backward(enc.edge_stack(x), cot)

    enc.storage = "bf16"
    ms16 = _median_ms(step16, reps=15, warm=3)
    enc.storage = "fp32"
    ms32 = _median_ms(step16, reps=15, warm=3)
    out["bf16_storage_mode"] = {"workload": "the headline stack step with the projected operand [P|Q] of the three EdgeConv layers stored "
                                            "in bf16 (tolerance 2e-2 of the activation scale, tests/test_gpu_bf16.py); median of 15 steps",
                                "dtype": "bf16 storage of [P|Q], fp32 accumulate / statistics / gradients",
                                "ms_per_step": round(ms16, 4), "clouds_per_s": round(B_PER_GPU / (ms16 / 1e3), 1),
                                "fp32_ms_per_step_same_loop": round(ms32, 4)}
    del cot
    # config 5
    B5, N5 = 4, 100000
    x5 = torch.from_numpy(abc_like_batch(B5, N5, seed=555)).to(dev)
    cot5 = [torch.randn(B5, c, N5, device=dev) for c in (64, 64, 128)]
    hot = [p for n, p in enc.named_parameters() if n.split(".")[0] in ("conv1", "conv2", "conv3", "bn1", "bn2", "bn3")]

    def step5():
        for p in hot:
            p.grad = None
        outs = enc.edge_stack(x5)
        torch.autograd.backward(outs, cot5)

    ms5 = _median_ms(step5, reps=5, warm=2)
    knn5 = _median_ms(lambda: G.knn_graph(x5, KNN, KNN, want64=False, want32=True, ordered=False), reps=5, warm=1)
    out["config5_large_clouds"] = {"workload": "kNN+EdgeConv stack fwd+bwd, B=4 x 100k pts, k=50, mode 0; median of 5 steps",
                                   "ms_per_step": round(ms5, 3), "clouds_per_s": round(B5 / (ms5 / 1e3), 1),
                                   "points_per_s": round(B5 * N5 / (ms5 / 1e3)), "xyz_knn_ms": round(knn5, 3)}
    del x5, cot5
    torch.cuda.empty_cache()
    out["config4_full_step"] = full_step_leg(dev, B_PER_GPU)
    return out


def full_step_leg(dev, batch, steps=10):
    """BASELINE configs[3] as SURVEY 8(d) defines it for this environment: the three-layer stack + encoder tail + per-point
    heads + normals EdgeConv + offset-prediction block + NLL / offset-L1 losses, forward + backward (everything of
    forward_train before the proposal grouping, which needs spconv / softgroup.ops)."""
    import gcanet_b200 as gb
    from gcanet_b200.model import PrimitivesEmbeddingPerPoint, nll_loss, offset_l1_loss
    from gcanet_b200.synth import abc_like_batch
    torch.manual_seed(0)
    net = PrimitivesEmbeddingPerPoint(emb_size=64, num_primitives=10, mode=5, num_channels=6, nn_nb=KNN).to(dev)
    c = torch.from_numpy(abc_like_batch(batch, NPTS, seed=4321, with_normals=True)).to(dev)
    pts, nrm = c[:, :3].transpose(1, 2).contiguous(), c[:, 3:].transpose(1, 2).contiguous()
    g = torch.Generator(device="cpu").manual_seed(3)
    t_gt = torch.randint(0, 10, (batch, NPTS), generator=g).to(dev)
    i_gt = torch.randint(-1, 12, (batch, NPTS), generator=g).to(dev)
    off_gt = (torch.randn(batch, NPTS, 3, generator=g) * 0.05).to(dev)

    def step():
        net.zero_grad(set_to_none=True)
        o = net(pts, nrm)
        loss = nll_loss(o["type_per_point"], t_gt) + 10.0 * offset_l1_loss(o["pt_offsets"], i_gt, off_gt)
        loss.backward()
        return loss

    before = gb._cabi.launch_count()
    ms = _median_ms(step, reps=steps, warm=3)
    launches = (gb._cabi.launch_count() - before) // (steps + 3)
    loss = float(step().detach())
    del net
    torch.cuda.empty_cache()
    return {"workload": f"per-point GCANet step fwd+bwd, B={batch} x 10k pts, k=50, mode 5: stack + tail + heads + normals EdgeConv + "
                        "offset block + NLL / offset-L1 losses (no proposal grouping / spconv head); median of 10 steps",
            "ms_per_step": round(ms, 3), "clouds_per_s": round(batch / (ms / 1e3), 1), "library_launches_per_step": int(launches),
            "loss": round(loss, 4)}


# ---------------------------------------------------------------------------- GPU path
def run_ours(args):
    import torch.distributed as dist
    import gcanet_b200 as gb
    from gcanet_b200 import _cabi, functional as G
    from gcanet_b200.parallel import GradBucket, init_from_env
    from gcanet_b200.synth import abc_like_batch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (gcanet_b200 has no CPU path; use --impl reference for the CPU arm)")
    rank, world, local = init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _cabi.check(_cabi.lib().gcanet_check_device(), "check_device")

    torch.manual_seed(0)
    enc = gb.DGCNNEncoderGn(mode=0, nn_nb=KNN, input_channels=6).to(dev)
    hot = [p for n, p in enc.named_parameters()
           if n.split(".")[0] in ("conv1", "conv2", "conv3", "bn1", "bn2", "bn3")]
    bucket = GradBucket(hot, assume_uniform=True).attach_sinks() if world > 1 else None

    # each rank owns its own 16 clouds (weak scaling); cloud seeds are disjoint across ranks
    x_host = torch.from_numpy(abc_like_batch(B_PER_GPU, NPTS, seed=1234, first_cloud=rank * B_PER_GPU)).pin_memory()
    x_dev = x_host.to(dev)
    gen = torch.Generator(device="cpu").manual_seed(7 + rank)
    cot = [torch.randn(B_PER_GPU, c, NPTS, generator=gen).to(dev) for c in (64, 64, 128)]
    loss_host = torch.zeros(1).pin_memory()

    nbr = args.streams
    part = B_PER_GPU // nbr
    cot_parts = [[c[i * part:(i + 1) * part].contiguous() for c in cot] for i in range(nbr)]
    branch_streams = [torch.cuda.Stream() for _ in range(nbr)] if nbr > 1 else None
    if nbr > 1:
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)     # the branches share the parameters on purpose

    def step_core(x):
        # forward of the three EdgeConv layers, backward from given upstream gradients (what the
        # consumer of x1|x2|x3 hands back); the step's result is a checksum of x3
        for p in hot:
            p.grad = None
        if branch_streams is None or not torch.cuda.is_current_stream_capturing():
            outs = enc.edge_stack(x)
            torch.autograd.backward(outs, cot)
            return outs[2].detach().sum()
        # inside the CUDA graph the batch runs as `nbr` branches of B / nbr clouds, so the issue/tensor-bound kNN of one
        # part overlaps the L2-bound gather / scatter of another and the small prep kernels stop serialising the device
        # (eagerly the host cannot feed two streams fast enough: 5.34 vs 5.01 ms; as two graph branches 4.57 vs 4.73 ms)
        cur = torch.cuda.current_stream()
        outs = []
        for i, s in enumerate(branch_streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(enc.edge_stack(x[i * part:(i + 1) * part]))
        for s, o, ch in zip(branch_streams, outs, cot_parts):
            with torch.cuda.stream(s):
                torch.autograd.backward(o, ch)
        for s in branch_streams:
            cur.wait_stream(s)
        return torch.stack([o[2].detach().sum() for o in outs]).sum()

    def step(x):
        loss = step_core(x)
        if bucket is not None:
            bucket.all_reduce_mean()                    # one NCCL all-reduce of the flat gradient bucket
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """(total ms, median ms of one step): CUDA events on the launching stream, barrier + synchronize on both sides,
        max over ranks.  The total over exactly `steps` steps gives `value`; the per-step median is reported beside it."""
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        barrier()
        evs[0].record()
        for i in range(steps):
            fn()
            evs[i + 1].record()
        barrier()
        ms = evs[0].elapsed_time(evs[-1])
        med = statistics.median(evs[i].elapsed_time(evs[i + 1]) for i in range(steps))
        if world > 1:
            t = torch.tensor([ms, med], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, med = float(t[0]), float(t[1])
        return ms, med

    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    torch.cuda.synchronize()

    # --- eager pass with per-call events: the breakdown and the dominant kernel's time (not the headline)
    n_eager = min(args.steps, 20)
    G.enable_kernel_timing(True)
    l0 = _cabi.launch_count()
    ms_eager_total, _ = timed(lambda: step(x_dev), n_eager)
    launches_per_step = (_cabi.launch_count() - l0) // n_eager
    per_call = G.kernel_timings_ms()
    G.enable_kernel_timing(False)
    ms_eager = ms_eager_total / n_eager

    # --- the dominant KERNEL alone: the library brackets the scan kernel(s) of every kNN call with its own pair of CUDA
    # events on the launching stream (gcanet_knn_probe_arm / _read, include/gcanet_b200.h); a separate short eager pass,
    # because reading the probe waits for the device after each kNN call
    scan_ms = {}
    try:
        G.enable_scan_probe(True)
        for _ in range(min(args.steps, 10)):
            step_core(x_dev)                            # no collective here: a probe failure must not desynchronise ranks
        torch.cuda.synchronize()
        scan_ms = G.scan_kernel_timings_ms()
    except Exception as exc:                            # noqa: BLE001 -- the roofline then falls back to the whole call, and says so
        scan_ms = {"error": f"{type(exc).__name__}: {exc}"}
    finally:
        G.enable_scan_probe(False)

    # --- the step as ONE CUDA graph: forward + backward of the stack captured once, replayed per step (the C-ABI never
    # allocates or synchronises, so everything it enqueues is capturable); the gradient all-reduce stays outside the
    # graph, on the same stream, right after the replay.  Same kernels, same order -- the ~100 launches just stop
    # paying their host-side latency one by one.
    graph, static_x, static_loss, graph_error = None, x_dev.clone(), None, None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step_core(static_x)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step_core(static_x)
            graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:                        # noqa: BLE001 -- fall back to the eager step, and say so
            graph, graph_error = None, f"{type(exc).__name__}: {exc}"
            torch.cuda.synchronize()

    def step_resident():
        if graph is None:
            return step(x_dev)
        graph.replay()                                  # static_x holds the resident batch
        if bucket is not None:
            bucket.all_reduce_mean()
        return static_loss

    for _ in range(3):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, ms_median = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    launches = launches_per_step * args.steps
    ms_step = ms_total / args.steps
    value = world * B_PER_GPU / (ms_step / 1e3)

    # --- end to end: pinned host batch -> H2D -> step -> D2H loss, every step
    def e2e_step():
        if graph is None:
            loss = step(x_host.to(dev, non_blocking=True))
        else:
            static_x.copy_(x_host, non_blocking=True)   # H2D of this step's batch into the graph's input buffer
            graph.replay()
            if bucket is not None:
                bucket.all_reduce_mean()
            loss = static_loss
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

    for _ in range(2):
        e2e_step()
    ms_e2e_total, ms_e2e_median = timed(e2e_step, args.steps)
    ms_e2e = ms_e2e_total / args.steps
    e2e_value = world * B_PER_GPU / (ms_e2e / 1e3)
    if graph is not None:
        static_x.copy_(x_dev)

    # --- N > 1: where the multi-GPU step's extra time goes.  Every rank times its own replay WITHOUT the collective (no
    # barrier between steps); the synchronous step costs the slowest rank's time plus the all-reduce, so
    # `ms_per_step - max(per-rank)` is the collective with its stream hand-offs and `max - mean` the spread between ranks
    # (their clouds differ, so the pruned kNN scans do different amounts of work)
    multi_gpu = None
    if world > 1 and graph is not None:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        evs[0].record()
        for _ in range(args.steps):
            graph.replay()
        evs[1].record()
        torch.cuda.synchronize()
        own = torch.tensor([evs[0].elapsed_time(evs[1]) / args.steps], device=dev)
        allr = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(allr, own)
        per_rank = [round(float(t[0]), 4) for t in allr]
        multi_gpu = {"per_rank_ms_without_collective": per_rank,
                     "rank_spread_ms": round(max(per_rank) - sum(per_rank) / world, 4),
                     "collective_ms": round(max(0.0, ms_step - max(per_rank)), 4)}     # (two separate timing loops: +-0.03 ms)

    # --- N > 1: the collective itself, checked on the hardware (outside the timed regions): every rank's own gradients
    # are all-gathered and averaged with torch ops, and must equal what GradBucket.all_reduce_mean left in p.grad
    allreduce_check = None
    if world > 1:
        for p in hot:
            p.grad = None
        outs = enc.edge_stack(x_dev)
        torch.autograd.backward(outs, cot)
        mine = torch.cat([p.grad.reshape(-1) for p in hot]).clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        want = torch.stack(gathered).double().mean(0)
        bucket.all_reduce_mean()
        got = torch.cat([p.grad.reshape(-1) for p in hot]).double()
        err = ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).reshape(1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        differ = (torch.stack(gathered).std(0).max() > 0).float().reshape(1)      # ranks really hold different gradients
        allreduce_check = {"max_rel_err": float(err[0]), "ranks_differ": bool(differ[0] > 0), "floats": int(mine.numel())}

    # --- BASELINE configs[3], literal split: a FIXED global batch of 128 clouds sharded over the ranks (128 / N per GPU),
    # same stack step + gradient all-reduce; strong scaling, reported beside the weak-scaling headline
    fixed = None
    if not args.no_extras and 128 % world == 0:
        per = 128 // world
        xf = torch.from_numpy(abc_like_batch(per, NPTS, seed=1234, first_cloud=rank * per)).to(dev)
        cotf = [torch.randn(per, c, NPTS, generator=gen).to(dev) for c in (64, 64, 128)]

        def step_fixed():
            for p in hot:
                p.grad = None
            outs = enc.edge_stack(xf)
            torch.autograd.backward(outs, cotf)
            if bucket is not None:
                bucket.all_reduce_mean()

        for _ in range(3):
            step_fixed()
        ms_f, med_f = timed(step_fixed, 5)
        fixed = {"global_batch": 128, "batch_per_gpu": per, "ms_per_step": ms_f / 5, "ms_per_step_median": med_f,
                 "value": 128 / (ms_f / 5 / 1e3), "unit": UNIT, "scaling": "strong", "steps": 5}
        del xf, cotf
        torch.cuda.empty_cache()

    extras = {}
    if world == 1 and not args.no_extras:
        extras = extra_configs(enc, dev, G)

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peaks = load_peaks()
    breakdown = {k: round(sum(v) / n_eager, 4) for k, v in sorted(per_call.items())}
    # dominant kernel: the feature-space kNN scan (two launches of C=64 per step)
    tag = "knn_graph[C=64,metric=0]"
    knn_ms = per_call.get(tag, [])
    roofline = None
    # `traffic` is by definition a profiler counter (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full
    # capture of this kernel, per launch): it cannot be measured inside an un-profiled run, so it is read from the newest
    # committed capture and the file is named in `traffic_source`; null when no capture of the current kernel exists
    traffic, traffic_source, scan_traffic = None, None, None
    for name in ("r02_dominant_kernel.json", "r01_dominant_kernel.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                dk = json.load(f)
            traffic = int(dk["dram_bytes_read"]) + int(dk["dram_bytes_write"])
            traffic_source = f"profiles/{name} (ncu --set full, {dk.get('kernel', 'dominant kernel')}, per launch)"
            if "scan" in dk:                            # the scan kernel's own counters, for the kernel-level line
                scan_traffic = int(dk["scan"]["dram_bytes_read"]) + int(dk["scan"]["dram_bytes_write"])
            break
        except Exception:
            continue
    if knn_ms:
        # Per SURVEY 8(d) the figure is 2*N^2*C FLOP per cloud, B clouds per launch.  `achieved` / `frac` are the dominant
        # KERNEL's (knn_tcp_scan_kernel<64>: its launch duration measured live by the library's event pair); `call` is
        # the same work divided by the whole gcanet_knn_graph call (preparation + scan + exact re-rank), the number
        # earlier rounds reported as `frac`.
        avg_ms = sum(knn_ms) / len(knn_ms)
        flop = 2.0 * NPTS * NPTS * 64 * B_PER_GPU                 # 2*N^2*C per cloud
        peak = float(peaks["bf16_tflops_sustained"])
        call_achieved = flop / (avg_ms * 1e-3) / 1e12
        call = {"achieved": call_achieved, "frac": call_achieved / peak, "ms_per_call": avg_ms, "traffic": traffic,
                "what": "PCA / Hilbert-order preparation + scan + exact fp32 re-rank (one gcanet_knn_graph call)"}
        kern_ms = scan_ms.get(tag) if isinstance(scan_ms, dict) else None
        if kern_ms:
            k_ms = sum(kern_ms) / len(kern_ms)
            achieved = flop / (k_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor",
                        "kernel": "knn_tcp_scan_kernel<64> (tcgen05 bf16x3 distance scan with box pruning, feature-space kNN C=64), "
                                  "algorithmic 2*N^2*C FLOP per cloud x 16 clouds per launch",
                        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                        "traffic": scan_traffic, "traffic_source": traffic_source, "ms_per_launch": k_ms,
                        "ms_per_launch_median": statistics.median(kern_ms), "launches_timed": len(kern_ms),
                        "timed_by": "CUDA events recorded by the library around the kernel on the launching stream "
                                    "(gcanet_knn_probe_arm / gcanet_knn_probe_read), eager pass",
                        "launches_per_step": len(knn_ms) / n_eager, "call": call,
                        "peak_source": f"{peaks['_source']} bf16_tflops_sustained (kernel timed inside a long step)",
                        "frac_of_burst_peak": (achieved / float(peaks["bf16_tflops"])) if "bf16_tflops" in peaks else None}
        else:
            roofline = {"bound": "tensor",
                        "kernel": "feature-space kNN C=64, WHOLE CALL (scan-kernel probe unavailable: "
                                  f"{scan_ms.get('error', 'no scan bracketed') if isinstance(scan_ms, dict) else 'n/a'}): preparation + "
                                  "knn_tcp_scan_kernel + exact re-rank, algorithmic 2*N^2*C FLOP per cloud",
                        "achieved": call_achieved, "peak": peak, "unit": "TFLOP/s", "frac": call_achieved / peak,
                        "traffic": traffic, "traffic_source": traffic_source, "ms_per_launch": avg_ms,
                        "launches_per_step": len(knn_ms) / n_eager, "call": call,
                        "peak_source": f"{peaks['_source']} bf16_tflops_sustained (kernel timed inside a long step)"}
    step_tflops = GFLOP_PER_CLOUD * 1e9 * B_PER_GPU / (ms_step * 1e-3) / 1e12

    gpu_reference = None
    if world == 1 and not args.no_extras:
        gpu_reference = gpu_reference_leg(dev)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cps, sec = cpu_reference_clouds_per_s(steps=3, warmup=1, threads=threads)
        cpu_baseline = {"value": cps, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"3 clouds (1 per step, 10k pts, k=50, mode 0) fwd+bwd after 1 warm-up, "
                                  f"{sec:.2f} s/cloud, {cpu_model_name()}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "ms_per_step_median": ms_median,
        "cuda_graph": graph is not None if graph_error is None else graph_error, "ms_per_step_eager": ms_eager,
        "graph_branches": args.streams if graph is not None else 1,
        "value_from_median": world * B_PER_GPU / (ms_median / 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DGCNN kNN+EdgeConv stack fwd+bwd, B=16 x 10k pts, k=50, mode 0 (BASELINE configs[1])",
                   "batch_per_gpu": B_PER_GPU, "global_batch": world * B_PER_GPU, "points": NPTS, "k": KNN,
                   "parallelism": f"dp{world} (clouds sharded, one NCCL all-reduce of the weight gradients per step)"
                   if world > 1 else "single GPU",
                   "l2": "no flush needed: one step streams >1 GB of activations/neighbour lists, far above the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "ms_per_step_median": ms_e2e_median,
                "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "step_tflops_algorithmic": step_tflops,
        "step_frac_of_tensor_peak": step_tflops / float(peaks["bf16_tflops_sustained"]),
        "breakdown_ms_per_step": breakdown,
        "cpu_baseline": cpu_baseline,
        "gpu_reference": gpu_reference,
        "allreduce_check": allreduce_check, "multi_gpu": multi_gpu,
        "config4_fixed_global_batch": fixed,
    }
    line.update(extras)
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the ~15 s CPU leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step instead of the CUDA-graph replay")
    ap.add_argument("--streams", type=int, choices=[1, 2, 4, 8], default=2,
                    help="inside the graph the batch runs as this many branches of B / streams clouds (default 2)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the same-GPU reference leg and the config 3 / config 5 keys (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
