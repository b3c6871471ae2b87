#!/usr/bin/env python
"""Benchmark of the doubly contrastive loss hot path (BASELINE.json metric: contrastive-loss fwd+bwd
anchors/sec, and % of bf16 tensor-core peak for the similarity kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = one forward + backward of the pixel-level contrastive loss module over one synthetic
ACDC-shaped batch (label down-sampling, argmax, hard-anchor sampling with the host RNG, gather,
N x N contrast, gradient back to the dense NCHW embedding gradient).
  N = 1 : workload cfg2 (BASELINE.json configs[1]: batch 8 @ 2048x1024, 128-d, 8192 anchors)
  N > 1 : workload cfg4 (configs[3]: 65536 anchors, anchor rows sharded over the ranks, contrast
          set all-gathered with NCCL; fixed total work => "scaling": "strong")
`value` has inputs resident in HBM; `e2e` times the same step through the module with pinned
HOST inputs (H2D of feats/labels/predict inside the timed region, D2H of the loss).
`--impl reference` times the CPU restatement of the reference (oracle port: the reference is
Python and `/root/reference` does not travel to the GPU box) on the host cores.
"""
import argparse
import dataclasses
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "contrastive_loss_fwd_bwd_anchors_per_sec"
UNIT = "anchors/s"
DIM = 128


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_tflops=float(p["bf16_tflops"]), hbm_gbs=float(p["hbm_gbs"]), source="measured (burst)")
    return dict(bf16_tflops=1590.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def ncu_traffic(workload, world):
    """DRAM bytes per step of the similarity kernels (sweep P + backward) from the committed `ncu --set full`
    capture of this workload on one GPU (profiles/r01r_ncu_dram_traffic.json); None where none was taken."""
    path = os.path.join(ROOT, "profiles", "r01r_ncu_dram_traffic.json")
    if world != 1 or not os.path.exists(path):
        return None
    try:
        with open(path) as f:
            return json.load(f).get(workload, {}).get("similarity_step")
    except (OSError, ValueError):
        return None


def pick_workload(args, world):
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS
    name = args.workload
    if name == "auto":
        name = "cfg2" if world == 1 else "cfg4"
    return WORKLOADS[name]


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ---------------------------------------------------------------------------------------------
def cpu_port_step(wl, data, seed):
    from oracle import dcl_oracle as O
    crit = O.PixelContrastPort()
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    x = data["feats"].detach().clone().requires_grad_(True)
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    loss = crit(x, labels=data["labels"], predict=data["predict"])
    loss.backward()
    dt = time.perf_counter() - t0
    return dt, crit.last_plan.A * crit.last_plan.n_view, float(loss.item())


def reference_sample_workload(wl):
    """Bounded sample for the CPU arm: 2 of the workload's images, max_samples scaled alike (so n_view is
    unchanged) but capped at 4096 anchors; the reference's cost per anchor grows with the number of anchors, so
    this sample flatters the CPU number."""
    b = min(2, wl.B)
    # at most 4096 anchors per CPU step (the dense N x N autograd of the reference grows as N^2: 16384 anchors would
    # take ~30 s and several GB per step); for cfg2 this is the proportional 2048
    return dataclasses.replace(wl, name=wl.name + "_sample", B=b, max_samples=min(wl.max_samples * b // wl.B, 4096))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from doubly_contrastive_semseg_b200.synthetic import make_inputs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl_full = pick_workload(args, max(world, args.gpus))
    wl = reference_sample_workload(wl_full)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    data = make_inputs(wl, seed=1, device="cpu")
    n = 0
    for s in range(min(args.warmup, 1)):
        cpu_port_step(wl, data, 1000 + s)
    times = []
    for s in range(args.steps):
        dt, n, _ = cpu_port_step(wl, data, 2000 + s)
        times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    sample = "%d of %d images of %s per step (%d anchors), fp32 torch-CPU restatement of utils/loss.py" % (
        wl.B, wl_full.B, wl_full.name, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": total / len(times) * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl_full.name, "sample": sample, "B": wl.B, "label_hw": [wl.H, wl.W],
                   "embed_hw": [wl.h, wl.w], "anchors_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    import doubly_contrastive_semseg_b200 as pkg
    from doubly_contrastive_semseg_b200 import _lib
    from doubly_contrastive_semseg_b200 import loss as L
    from doubly_contrastive_semseg_b200.synthetic import make_inputs
    _lib.load()

    wl = pick_workload(args, world)
    if wl.B % world:
        raise SystemExit("workload batch %d is not divisible by %d ranks" % (wl.B, world))
    # every rank generates the same global batch and keeps its own images (data-parallel shard)
    data = make_inputs(wl, seed=1, device=dev)
    bl = wl.B // world
    sl = slice(rank * bl, (rank + 1) * bl)
    feats = data["feats"][sl].contiguous()
    labels = data["labels"][sl].contiguous()
    predict = data["predict"][sl].contiguous()
    del data
    torch.cuda.empty_cache()

    if world > 1:
        crit = pkg.ShardedPixelContrastLoss(device=dev)
    else:
        crit = pkg.PixelContrastLoss(device=dev)
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views

    sim_events = []
    L.set_profile_hook(lambda name, a, b: sim_events.append((name, a, b)))

    feats.requires_grad_(True)

    # Identical CPU generator state on every rank, set ONCE: every rank consumes the same stream (each replays the
    # global plan), so the states stay identical, and an untouched generator lets the library's look-ahead thread
    # regenerate mt19937 state blocks off the step's critical path.
    torch.manual_seed(1234)

    def step(seed, f=feats, lab=labels, pred=predict):
        f.grad = None
        loss = crit(f, labels=lab, predict=pred)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        step(s)
    barrier()
    sim_events.clear()
    L.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record()
        for s in range(args.steps):
            step(100 + s)
        ev1.record()
        barrier()
    launches = L.launch_count()
    ms_total = ev0.elapsed_time(ev1)
    n_global = crit.last_n_global if world > 1 else crit.last_layout.n
    n_local = crit.last_layout.n
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n_global * args.steps / (ms_total * 1e-3)
    sim_ms = sum(a.elapsed_time(b) for _, a, b in sim_events)
    sim_launches = len(sim_events)
    clocks = clk.summary()

    # ---- end to end: pinned host inputs, H2D inside the timed region, loss read back
    h_feats = feats.detach().cpu().pin_memory()
    h_labels = labels.cpu().pin_memory()
    h_predict = predict.cpu().pin_memory()
    d_feats = torch.empty_like(feats).requires_grad_(True)
    d_labels, d_predict = torch.empty_like(labels), torch.empty_like(predict)
    h2d = h_feats.numel() * 4 + h_labels.numel() * 8 + h_predict.numel() * 4

    def e2e_step(seed):
        with torch.no_grad():
            d_feats.copy_(h_feats, non_blocking=True)
            d_labels.copy_(h_labels, non_blocking=True)
            d_predict.copy_(h_predict, non_blocking=True)
        loss = step(seed, d_feats, d_labels, d_predict)
        return float(loss.item())            # D2H read of the step's result

    for s in range(min(args.warmup, 3)):
        e2e_step(s)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        e2e_step(100 + s)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_global * args.steps / (float(t.item()) * 1e-3)
    L.set_profile_hook(None)

    if rank == 0:
        peaks = load_peaks()
        flops = 6.0 * n_local * n_global * DIM * args.steps        # this rank's share of 6 N^2 D
        achieved = flops / (sim_ms * 1e-3) / 1e12 if sim_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": wl.name, "B": wl.B, "label_hw": [wl.H, wl.W], "embed_hw": [wl.h, wl.w],
                       "embed_dim": DIM, "classes": wl.K, "max_samples": wl.max_samples,
                       "max_views": wl.max_views, "anchors": n_global, "anchors_per_gpu": n_local,
                       "sampler_rng": "exact (torch CPU mt19937 replay)",
                       "l2": "inputs (%.0f MB embeddings) exceed the 126 MB L2; no explicit flush" % (feats.numel() * 4 / 1e6),
                       "parallelism": "rows sharded x%d, contrast set NCCL all-gathered" % world if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "traffic": ncu_traffic(wl.name, world),
                         "kernel": "similarity sweeps + fused backward (dcl_contrast_fwd + dcl_contrast_bwd)",
                         "algorithmic": "6*N_local*N*D flops per step", "ms_per_step": sim_ms / args.steps,
                         "timed_calls": sim_launches, "peak_source": peaks["source"]},
        }
        if world == 1 and not args.no_cpu_baseline:
            from doubly_contrastive_semseg_b200.synthetic import make_inputs as mk
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            cpu_data = mk(wl, seed=1, device="cpu")
            dt, n_cpu, _ = cpu_port_step(wl, cpu_data, 100)
            line["cpu_baseline"] = {"value": n_cpu / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "1 fwd+bwd of %s (%d anchors) on the host, fp32 torch-CPU restatement "
                                              "of utils/loss.py, %.1f s" % (wl.name, n_cpu, dt)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
