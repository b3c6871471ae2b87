#!/usr/bin/env python
"""Benchmark of the doubly contrastive loss hot path (BASELINE.json metric: contrastive-loss fwd+bwd
anchors/sec, and % of bf16 tensor-core peak for the similarity kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = one forward + backward of the pixel-level contrastive loss module over one synthetic
ACDC-shaped batch (label down-sampling, argmax, hard-anchor sampling with torch's CPU generator, gather,
N x N contrast, gradient back to the dense NCHW embedding gradient).

The headline line is the SAME workload at every N, so that the driver's 1/2/4/8 series is one scaling curve:
  cfg4 (BASELINE.json configs[3]): batch 8 @ 2048x1024, 65536 anchors; on N > 1 GPUs the anchor rows are sharded
  over the ranks (one process per GPU) and the contrast set is all-gathered with NCCL ("scaling": "strong").
At N = 1 the line also carries, under "workloads", the same measurements for
  cfg2 (configs[1]): batch 8 @ 2048x1024, 8192 anchors (pixel term), and
  cfg3 (configs[2]): the doubly contrastive step (pixel term + 32x32 image-level term) on the 2x16-crop batch,
  cfg5 (configs[4]): one GPU's share (8 images x 2 crops at 1024x512) of the full SwiftNet-RN18 training step,
and "roofline_hbm": achieved GB/s of the HBM-bound kernels (sampler, gather, scatter, global average pool).
`value` has inputs resident in HBM; `e2e` times the same step through the module with pinned HOST inputs
(H2D of feats/labels/predict inside the timed region, D2H of the loss).
`--impl reference` times the reference's own utils/loss.py (oracle/_ref/loss.py, placed there by
oracle/build_ref.py; the oracle's port when that file is absent) on the host cores.
"""
import argparse
import dataclasses
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "contrastive_loss_fwd_bwd_anchors_per_sec"
UNIT = "anchors/s"
DIM = 128
REF_BUDGET_S = 330.0          # the reference arm sizes its sample so that steps + warm-up fit in about this


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_tflops=float(p["bf16_tflops"]), hbm_gbs=float(p["hbm_gbs"]), source="measured (burst)")
    return dict(bf16_tflops=1590.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """SM clock / throttle-reason samples taken DURING the timed regions by a native NVML thread inside the library
    (dcl_clock_sampler_start / _stop, csrc/dcl_clocks.cpp; 2 ms period, started and stopped around every timed
    region).  `nvidia-smi -lms` is not used: while it polls, even at 100 ms, every step of the 0.4 ms workload takes
    0.25 ms longer (measured, profiles/r02j_*), and its start-up can stall a step for 100+ ms."""

    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    PERIOD_US = 2000

    def __init__(self, lib):
        self.lib = lib
        self.regions = []
        self.active = False
        # first use loads and initialises NVML (tens of milliseconds of driver activity): do it now, not inside the
        # first timed region
        self.start()
        time.sleep(0.02)
        self.stop()
        self.regions = []

    def start(self):
        if os.environ.get("DCL_BENCH_NO_SAMPLER"):          # experiments only
            return
        self.active = self.lib.dcl_clock_sampler_start(self.PERIOD_US) == 0

    def stop(self):
        if not self.active:
            return
        import ctypes
        out = (ctypes.c_double * 5)()
        self.lib.dcl_clock_sampler_stop(ctypes.cast(out, ctypes.c_void_p))
        self.active = False
        if out[0] > 0:
            self.regions.append([float(v) for v in out])

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
               "window": "inside the timed regions only (device-resident and end-to-end steps of every workload in this "
                         "line); native NVML thread, %d ms period" % (self.PERIOD_US // 1000)}
        if not self.regions:
            return out
        mask = 0
        for r in self.regions:
            mask |= int(r[3])
        out.update(sm_mhz=float(np.median([r[1] for r in self.regions])), sm_mhz_min_region=float(min(r[1] for r in self.regions)),
                   sm_max_mhz=float(max(r[2] for r in self.regions)), samples=int(sum(r[0] for r in self.regions)),
                   power_w_max=float(max(r[4] for r in self.regions)))
        out["reasons"] = sorted(k for k, bit in self.BITS.items() if mask & bit)
        return out


def ncu_traffic(workload, world):
    """DRAM bytes per step of the similarity kernels (sweep P + backward) from the committed `ncu --set full`
    capture of this workload on one GPU (profiles/*_ncu_dram_traffic.json, newest round first); None where none was
    taken."""
    if world != 1:
        return None
    for name in ("r02_ncu_dram_traffic.json", "r01r_ncu_dram_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            try:
                with open(path) as f:
                    v = json.load(f).get(workload, {}).get("similarity_step")
                if v is not None:
                    return v
            except (OSError, ValueError):
                pass
    return None


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own file when oracle/_ref holds it, the oracle's port otherwise
# ---------------------------------------------------------------------------------------------
def cpu_module(wl):
    """-> (PixelContrastLoss-like module on the CPU, kind)"""
    from oracle import build_ref
    ref = build_ref.load()
    if ref is not None:
        crit, kind = ref.PixelContrastLoss(device="cpu"), "reference"
    else:
        from oracle import dcl_oracle as O
        crit, kind = O.PixelContrastPort(), "port"
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    return crit, kind


def cpu_step(crit, data, seed):
    """One forward + backward on the host; returns (seconds, loss).  The reference prints on every call (loss.py:270)."""
    import contextlib
    import io
    x = data["feats"].detach().clone().requires_grad_(True)
    torch.manual_seed(seed)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        loss = crit(x, labels=data["labels"], predict=data["predict"])
    loss.backward()
    return time.perf_counter() - t0, float(loss.item())


def cpu_anchors(wl, data):
    """Rows of the contrast matrix the reference forms on this input (A * n_view, loss.py:290-291)."""
    from oracle import dcl_oracle as O
    lab = O.downsample_labels(data["labels"].numpy(), wl.h, wl.w)
    A = 0
    for b in range(lab.shape[0]):
        vals, counts = np.unique(lab[b], return_counts=True)
        A += sum(1 for v, c in zip(vals, counts) if v != 255 and c > wl.max_views)
    return A * min(wl.max_samples // max(A, 1), wl.max_views)


def reference_workload(args, world):
    """The workload the reference arm runs: the headline workload when the reference can hold it, else the largest
    the reference's dense N x N autograd can (cfg2 = the same eight images with max_samples 8192: at 65536 anchors the
    reference needs ~190 GB of N x N temporaries)."""
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS
    name = args.workload if args.workload != "auto" else "cfg4"
    wl = WORKLOADS[name]
    note = None
    if wl.max_samples > 16384:
        note = "reference OOM at %d anchors (dense N x N autograd, ~190 GB); timed on cfg2 = the same images, max_samples 8192" % wl.max_samples
        wl = dataclasses.replace(WORKLOADS["cfg2"], name=name + "->cfg2")
    return name, wl, note


def shrink(wl, factor):
    """A proportional sample of a workload: fewer images, max_samples scaled alike (n_view unchanged)."""
    b = max(1, int(wl.B * factor))
    return dataclasses.replace(wl, name=wl.name + "_sample", B=b, max_samples=max(wl.max_views, wl.max_samples * b // wl.B))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from doubly_contrastive_semseg_b200.synthetic import make_inputs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    headline, wl, note = reference_workload(args, max(world, args.gpus))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # One untimed step at full size decides: if steps + warm-up fit the budget the whole workload is timed; otherwise a
    # proportional sample is (fewer images, max_samples scaled alike; the reference's cost grows about quadratically
    # with the number of images - one full-size index_put per anchor).
    total_steps = args.steps + args.warmup
    data = make_inputs(wl, seed=1, device="cpu")
    crit, kind = cpu_module(wl)
    t_full, _ = cpu_step(crit, data, 999)
    sample = wl
    if t_full * total_steps > REF_BUDGET_S:
        f = (REF_BUDGET_S / (t_full * total_steps)) ** 0.5
        sample = shrink(wl, max(f, 1.0 / wl.B))
    data = make_inputs(sample, seed=1, device="cpu")
    crit, kind = cpu_module(sample)
    n = cpu_anchors(sample, data)
    for s in range(args.warmup):
        cpu_step(crit, data, 1000 + s)
    times = []
    for s in range(args.steps):
        dt, _ = cpu_step(crit, data, 2000 + s)
        times.append(dt)
    total = sum(times)
    value = n * len(times) / total
    what = ("%s: %d of %d images per step, max_samples %d, max_views %d (%d anchors); %s on %d host threads"
            % (wl.name, sample.B, wl.B, sample.max_samples, sample.max_views, n,
               "the reference's own utils/loss.py (oracle/_ref/loss.py)" if kind == "reference"
               else "fp32 torch-CPU restatement of utils/loss.py (oracle port)", torch.get_num_threads()))
    if note:
        what += "; " + note
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3,
        "ms_per_step_median": float(np.median(times)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": headline, "sample": what, "B": sample.B, "label_hw": [sample.H, sample.W],
                   "embed_hw": [sample.h, sample.w], "embed_dim": DIM, "classes": sample.K,
                   "max_samples": sample.max_samples, "max_views": sample.max_views, "anchors": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class Timer:
    """K steps bracketed by events on the current stream + a per-step event after every step."""

    clocks = None            # ClockSampler shared by every timed region of the process

    def __init__(self, barrier):
        self.barrier = barrier

    def run(self, step, steps, first_seed=100, before_timed=None):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        # the sampler thread starts BEFORE the barrier: its first NVML queries (the slow ones, and with several ranks
        # they queue on NVML's lock) then land in the barrier, not in the first timed step
        if Timer.clocks is not None:
            Timer.clocks.start()
        # like timeit: no cyclic garbage collection inside the timed region (a generation-2 pass is a millisecond,
        # several sub-millisecond steps); collected right before instead
        import gc
        gc_was = gc.isenabled()
        gc.collect()
        gc.disable()
        # two more untimed steps instead of a sleep: the sampler's first NVML queries land here and in the barrier, and the
        # host thread enters the timed region warm (a 5 ms sleep made the first timed step 2-3x slower on the host side)
        for s in range(2):
            step(first_seed - 2 + s)
        if before_timed is not None:
            before_timed()
        self.barrier()
        try:
            return self._run(step, steps, first_seed, marks)
        finally:
            if gc_was:
                gc.enable()
            if Timer.clocks is not None:
                Timer.clocks.stop()

    def _run(self, step, steps, first_seed, marks):
        marks[0].record()
        nomarks = bool(os.environ.get("DCL_BENCH_NO_MARKS"))          # experiments only
        for s in range(steps):
            step(first_seed + s)
            if not nomarks or s == steps - 1:
                marks[s + 1].record()
        if nomarks:
            self.barrier()
            total = marks[0].elapsed_time(marks[-1])
            return total, [total / steps] * steps
        self.barrier()
        total = marks[0].elapsed_time(marks[-1])
        per = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
        return total, per


def measure_pixel(pkg, L, wl, world, rank, dev, args, barrier, dist, with_e2e=True):
    """The pixel term of `wl` on this rank's images.  Returns a dict of measurements (identical on every rank after the
    max-reduction) and the module."""
    from doubly_contrastive_semseg_b200.synthetic import make_inputs
    data = make_inputs(wl, seed=1, device=dev)
    bl = wl.B // world
    sl = slice(rank * bl, (rank + 1) * bl)
    feats = data["feats"][sl].contiguous().requires_grad_(True)
    labels = data["labels"][sl].contiguous()
    predict = data["predict"][sl].contiguous()
    full = data if world > 1 else None
    if world == 1:
        del data
    crit = pkg.ShardedPixelContrastLoss(device=dev) if world > 1 else pkg.PixelContrastLoss(device=dev)
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    out = {}

    # ---- sharded == single GPU on the concatenated batch (the bench workload itself), before anything is timed
    if world > 1:
        ref = pkg.PixelContrastLoss(device=dev)
        ref.max_samples, ref.max_views = wl.max_samples, wl.max_views
        xf = full["feats"].clone().requires_grad_(True)
        torch.manual_seed(4321)
        loss_ref = ref(xf, labels=full["labels"], predict=full["predict"])
        loss_ref.backward()
        torch.manual_seed(4321)
        feats.grad = None
        loss = crit(feats, labels=labels, predict=predict)
        loss.backward()
        e_loss = abs(loss.item() - loss_ref.item()) / abs(loss_ref.item())
        e_grad = float((feats.grad - xf.grad[sl]).abs().max() / xf.grad.abs().max())
        same_n = int(crit.last_n_global == ref.last_layout.n)
        t = torch.tensor([e_loss, e_grad, 1.0 - same_n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_loss, e_grad, bad_n = (float(v) for v in t)
        out["sharded_parity"] = {"loss_rel": e_loss, "grad_rel": e_grad, "anchors_equal": bad_n == 0.0,
                                 "reference": "single-GPU module on the concatenated batch, same generator state",
                                 "ok": bool(e_loss <= 1e-5 and e_grad <= 2e-3 and bad_n == 0.0)}
        del ref, xf, full, loss_ref
        torch.cuda.empty_cache()

    # the roofline's clock: events around the contrast forward / backward of every timed step, recorded by the library on
    # the step's stream (DCL_BENCH_NO_HOOK: experiments without them)
    sim_on = not os.environ.get("DCL_BENCH_NO_HOOK")
    # Identical CPU generator state on every rank, set ONCE: every rank consumes the same stream (each replays the
    # global plan), so the states stay identical, and an untouched generator lets the library keep its look-ahead
    # of the mt19937 stream (host thread + device mirror) running off the step's critical path.
    torch.manual_seed(1234)

    host_log = []                     # experiments only (DCL_BENCH_DUMP_STEPS): host wall time and sections per step
    dump = bool(os.environ.get("DCL_BENCH_DUMP_STEPS"))

    def step(seed, f=feats, lab=labels, pred=predict):
        f.grad = None
        t0 = time.perf_counter()
        loss = crit(f, labels=lab, predict=pred)
        t1 = time.perf_counter()
        loss.backward()
        if dump:
            import ctypes
            t2 = time.perf_counter()
            sec = (ctypes.c_longlong * 8)()
            _lib_handle.dcl_step_timing(sec)
            py = L._DEBUG_PY_TIMES[-1] if L._DEBUG_PY_TIMES else []
            host_log.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, [v / 1e6 for v in sec[:7]] + [-1.0] + list(py)))
        return loss

    from doubly_contrastive_semseg_b200 import _lib as _lib_mod
    _lib_handle = _lib_mod.load()
    for s in range(args.warmup):
        step(s)
    barrier()

    def before_timed():
        host_log.clear()
        L.reset_launch_count()
        if sim_on:
            L.set_sim_timing(True)

    total, per = Timer(barrier).run(step, args.steps, before_timed=before_timed)
    launches = L.launch_count()
    n_global = crit.last_n_global
    n_local = crit.last_layout.n
    sim_fwd, sim_bwd, sim_steps = L.sim_times() if sim_on else (0.0, 0.0, 0)
    L.set_sim_timing(False)
    sim_ms = sim_fwd + sim_bwd
    t = torch.tensor([total, float(np.median(per))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, med = float(t[0]), float(t[1])
    if dump and rank == 0:
        import ctypes
        print("per-step ms (%s):" % wl.name, " ".join("%.3f" % v for v in per), file=sys.stderr, flush=True)
        medp = float(np.median(per))
        for i, v in enumerate(per):
            if v > 1.5 * medp and i < len(host_log):
                print("  slow step %d: %.3f ms gpu; host forward %.3f ms backward %.3f ms; sections(ms) %s"
                      % (i, v, host_log[i][0], host_log[i][1], " ".join("%.3f" % x for x in host_log[i][2])), file=sys.stderr)
        w = (ctypes.c_longlong * 2)()
        _lib_handle.dcl_host_lookahead_wait(w)
        print("  look-ahead waits: total %.3f ms, longest %.3f ms" % (w[0] / 1e6, w[1] / 1e6), file=sys.stderr, flush=True)
    out.update(value=n_global * args.steps / (total * 1e-3), ms_per_step=total / args.steps, ms_per_step_median=med,
               anchors=n_global, anchors_per_gpu=n_local, gpu_launches=launches, sim_ms_per_step=sim_ms / args.steps,
               timed_calls=2 * sim_steps, embed_mb=feats.numel() * 4 / 1e6)

    # ---- end to end: pinned host inputs, H2D inside the timed region, loss read back
    if with_e2e:
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
        total_e, _ = Timer(barrier).run(e2e_step, args.steps)
        t = torch.tensor([total_e], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["e2e"] = {"value": n_global * args.steps / (float(t.item()) * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
        del h_feats, h_labels, h_predict, d_feats, d_labels, d_predict
    del feats, labels, predict
    torch.cuda.empty_cache()
    return out, crit


def measure_doubly(pkg, L, wl, dev, args, barrier):
    """cfg3: both contrastive terms on the two-crop batch through DoublyContrastiveLoss, trainer.py:143-158 weighting."""
    import types
    from doubly_contrastive_semseg_b200.synthetic import make_inputs
    data = make_inputs(wl, seed=1, device=dev)
    feats = data["feats"].requires_grad_(True)                      # [2B,128,h,w]
    labels, predict, weather = data["labels"], data["predict"], data["weather"]
    torch.manual_seed(7)
    crit = pkg.DoublyContrastiveLoss(device=dev, opts=types.SimpleNamespace(deeplab=False))
    crit.pixel.max_samples, crit.pixel.max_views = wl.max_samples, wl.max_views
    torch.manual_seed(1234)

    def step(seed, f=feats, lab=labels, pred=predict, wt=weather):
        f.grad = None
        sup, pix = crit(f, labels=lab, predict=pred, class_labels=wt)
        total = (sup + pix) / wl.B
        total.backward()
        return total

    for s in range(args.warmup):
        step(s)
    barrier()
    total, per = Timer(barrier).run(step, args.steps, before_timed=L.reset_launch_count)
    launches = L.launch_count()
    n_pix = crit.pixel.last_layout.n
    rows = n_pix + 2 * wl.B
    out = dict(value=rows * args.steps / (total * 1e-3), ms_per_step=total / args.steps,
               ms_per_step_median=float(np.median(per)), anchors=rows, pixel_anchors=n_pix, image_rows=2 * wl.B,
               gpu_launches=launches, embed_mb=feats.numel() * 4 / 1e6)
    # end to end
    h = [t.detach().cpu().pin_memory() for t in (feats, labels, predict, weather)]
    d_feats = torch.empty_like(feats).requires_grad_(True)
    d_rest = [torch.empty_like(t) for t in (labels, predict, weather)]
    h2d = sum(t.numel() * t.element_size() for t in h)

    def e2e_step(seed):
        with torch.no_grad():
            d_feats.copy_(h[0], non_blocking=True)
            for dst, src in zip(d_rest, h[1:]):
                dst.copy_(src, non_blocking=True)
        return float(step(seed, d_feats, *d_rest).item())

    for s in range(2):
        e2e_step(s)
    total_e, _ = Timer(barrier).run(e2e_step, args.steps)
    out["e2e"] = {"value": rows * args.steps / (total_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": 4}
    del data, feats, d_feats, h, d_rest
    torch.cuda.empty_cache()
    return out


def hbm_kernels(L, _lib, dev, peaks, reps=7):
    """CUDA-event times of the HBM-bound kernels at the cfg2 shape (sampler front end, gather, scatter) and the cfg3
    shape (global average pool), each launched alone through its C entry point with L2 flushed in between (a 512 MB
    fill).  Algorithmic bytes are SURVEY 8d's: what the stage has to move at minimum."""
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
    from doubly_contrastive_semseg_b200.loss import _p, _stream
    wl = WORKLOADS["cfg2"]
    d = make_inputs(wl, seed=1, device=dev)
    B, h, w, hw = wl.B, wl.h, wl.w, wl.h * wl.w
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    crit = L.PixelContrastLoss(device=dev)
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    torch.manual_seed(5)
    out = crit.sample(d["feats"], d["labels"], d["predict"])
    pix, y_dev, n = out
    n_pad = pix.shape[0]
    lay = crit.last_layout
    req_dev = torch.from_numpy(lay.req.reshape(-1).copy()).to(dev)
    res = {}

    def timed(name, fn, nbytes, note):
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        us = float(np.median(ts)) * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        res[name] = {"us": round(us, 2), "algorithmic_bytes": int(nbytes), "GBps": round(gbs, 1),
                     "frac": round(gbs / peaks["hbm_gbs"], 4), "what": note}

    code_chunk = {}

    def f_classify():
        code_chunk["v"] = L.classify(d["labels"], d["predict"], h, w)
    timed("classify", f_classify, B * 19 * hw * 4 + B * hw * 8 + B * hw * 2,
          "k_classify: predict read + one label per output pixel + code write")
    code, chunk, counts = code_chunk["v"]
    rowof = torch.empty(B * hw, dtype=torch.int32, device=dev)
    timed("select", lambda: L.select_pixels(code, chunk, B, hw, req_dev, n_pad, rowof), n_pad * (16 + 4 + 4) + n_pad * 4096,
          "k_select (+ 4 MB clear of the pixel->row map): one 2048-pixel chunk of codes scanned per row")
    tiles_box = {}

    def f_gather():
        tiles_box["v"] = L.gather_tiles(d["feats"], pix, n_pad, rowof)
    timed("gather", f_gather, n_pad * DIM * 4 + n_pad * DIM * 2,
          "k_gather_px: N*D f32 read at stride h*w, chunk-wise in pixel order, + bf16 tile write")
    dF = torch.randn(n_pad, DIM, device=dev)
    g = torch.ones((), device=dev)
    dfe = torch.zeros_like(d["feats"])
    timed("scatter", lambda: _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n_pad, _p(g), _p(dfe), B, hw, 0, _p(rowof), _stream()),
          2 * n_pad * DIM * 4, "k_scatter_px: N*D f32 read + N*D scattered f32 writes, chunk-wise in pixel order")
    timed("zero_fill", lambda: _lib.call("dcl_zero_fill", _p(dfe), dfe.numel() * 4, 0, _stream()), dfe.numel() * 4,
          "k_zero_fill: dense gradient clear (runs on a second stream underneath the sweeps in the real step)")
    gap_g = torch.randn(B * DIM, device=dev)
    timed("dense_grad", lambda: _lib.call("dcl_dense_grad", _p(dF), _p(rowof), B, _p(g), _p(gap_g), _p(dfe), B, hw, _stream()),
          dfe.numel() * 4 + n_pad * DIM * 4, "dcl_dense_grad: pooled-gradient broadcast (k_gap_bwd) + anchor gradients added by k_scatter_px (doubly step)")
    sampler_us = res["classify"]["us"] + res["select"]["us"] + res["gather"]["us"]
    sampler_bytes = B * 19 * hw * 4 + B * hw * 8 + n_pad * DIM * 4 + n_pad * DIM * 2
    res["sampler_total"] = {"us": round(sampler_us, 2), "algorithmic_bytes": int(sampler_bytes),
                            "GBps": round(sampler_bytes / (sampler_us * 1e-6) / 1e9, 1),
                            "frac": round(sampler_bytes / (sampler_us * 1e-6) / 1e9 / peaks["hbm_gbs"], 4),
                            "what": "classify + select + gather against SURVEY 8d's sampler bytes (cfg2: 94 MB)"}
    del d, dfe, dF
    torch.cuda.empty_cache()
    # segmentation-loss neighbour at the cfg2 shapes: logits [8,19,256,512], labels / EDT weights [8,1024,2048]
    gq = torch.Generator(device=dev).manual_seed(3)
    logits = 2.0 * torch.randn(wl.B, 19, wl.h, wl.w, generator=gq, device=dev)
    target = torch.randint(0, 19, (wl.B, wl.H, wl.W), generator=gq, device=dev)
    alpha = torch.rand(wl.B, wl.H, wl.W, generator=gq, device=dev)
    cw = torch.ones(19, device=dev)
    unscaled = torch.empty_like(logits)
    loss_n = torch.empty(2, device=dev)
    fws = int(_lib.load().dcl_focal_workspace_bytes(wl.B, wl.h, wl.w))
    fbuf = torch.empty(fws, dtype=torch.uint8, device=dev)
    timed("focal", lambda: _lib.call("dcl_focal_fwd", _p(logits), _p(target), _p(alpha), _p(cw), wl.B, 19, wl.h, wl.w,
                                     wl.H, wl.W, 255, 0.5, 0, _p(unscaled), _p(loss_n), _p(fbuf), fws, _stream()),
          target.numel() * 12 + 2 * logits.numel() * 4,
          "k_focal: BoundaryAwareFocalLoss forward + gradient in one pass (labels 8 B + EDT weight 4 B per pixel, "
          "logits read, gradient written); MUFU-bound")
    del logits, target, alpha, unscaled
    wl3 = WORKLOADS["cfg3"]
    x = torch.randn(2 * wl3.B, DIM, wl3.h, wl3.w, device=dev)
    pooled = torch.empty(2 * wl3.B * DIM, device=dev)
    R, hw3 = 2 * wl3.B * DIM, wl3.h * wl3.w
    timed("gap_fwd", lambda: _lib.call("dcl_gap_fwd", _p(x), R, hw3, _p(pooled), _stream()), x.numel() * 4,
          "k_gap_fwd at the cfg3 shape [32,128,256,512]")
    timed("gap_bwd", lambda: _lib.call("dcl_gap_bwd", _p(pooled), R, hw3, _p(x), 0, _stream()), x.numel() * 4,
          "k_gap_bwd at the cfg3 shape (dense broadcast write)")
    return res


def tensor_roofline(m, n_local, n_global, steps, peaks, traffic):
    flops = 6.0 * n_local * n_global * DIM            # this rank's share of 6 N^2 D, per step
    achieved = flops / (m["sim_ms_per_step"] * 1e-3) / 1e12 if m["sim_ms_per_step"] > 0 else 0.0
    return {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / peaks["bf16_tflops"], "traffic": traffic,
            "kernel": "similarity sweeps + fused backward (dcl_contrast_fwd + dcl_contrast_bwd inside dcl_step_fwd)",
            "algorithmic": "6*N_local*N*D flops per step", "ms_per_step": m["sim_ms_per_step"],
            "timed_calls": m["timed_calls"], "peak_source": peaks["source"]}


def config_of(wl, m, world):
    return {"workload": wl.name, "B": wl.B, "label_hw": [wl.H, wl.W], "embed_hw": [wl.h, wl.w],
            "embed_dim": DIM, "classes": wl.K, "max_samples": wl.max_samples, "max_views": wl.max_views,
            "anchors": m["anchors"], "anchors_per_gpu": m.get("anchors_per_gpu", m["anchors"]),
            "sampler_rng": "exact (torch CPU mt19937; look-ahead stream mirrored on the GPU, permutations replayed there)",
            "l2": "inputs (%.0f MB embeddings per GPU) exceed the 126 MB L2; no explicit flush" % m["embed_mb"],
            "timed_region": "CUDA events on the step's stream, barrier + synchronize on both sides, Python's cyclic GC "
                            "off inside it (timeit convention)",
            "parallelism": "rows sharded x%d, contrast set NCCL all-gathered" % world if world > 1 else "single GPU"}


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
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS
    _lib.load()

    name = args.workload if args.workload != "auto" else "cfg4"
    wl = WORKLOADS[name]
    if wl.B % world:
        raise SystemExit("workload batch %d is not divisible by %d ranks" % (wl.B, world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peaks = load_peaks()
    clk = Timer.clocks = ClockSampler(_lib.load())
    if True:
        if wl.two_crop:
            m = measure_doubly(pkg, L, wl, dev, args, barrier)
        else:
            m, _ = measure_pixel(pkg, L, wl, world, rank, dev, args, barrier, dist)
        extra = {}
        if world == 1 and args.workload == "auto":
            # the other single-GPU configurations of BASELINE.json, same measurements
            extra["cfg2"], _ = measure_pixel(pkg, L, workload_named("cfg2"), 1, 0, dev, args, barrier, dist)
            extra["cfg3"] = measure_doubly(pkg, L, workload_named("cfg3"), dev, args, barrier)
    clocks = clk.summary()

    parity = m.pop("sharded_parity", None)
    if parity is not None and not parity["ok"]:
        if rank == 0:
            print(json.dumps({"error": "sharded result differs from the single-GPU module", "sharded_parity": parity}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        raise SystemExit(2)

    if rank == 0:
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "ms_per_step_median": m["ms_per_step_median"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config_of(wl, m, world), "clocks": clocks, "e2e": m["e2e"],
            "gpu_launches": m["gpu_launches"],
        }
        if not wl.two_crop:
            line["roofline"] = tensor_roofline(m, m["anchors_per_gpu"], m["anchors"], args.steps, peaks,
                                               ncu_traffic(wl.name, world))
        if parity is not None:
            line["sharded_parity"] = parity
        if extra:
            blocks = {}
            for k, e in extra.items():
                wk = WORKLOADS[k]
                b = {"value": e["value"], "unit": UNIT, "ms_per_step": e["ms_per_step"],
                     "ms_per_step_median": e["ms_per_step_median"], "config": config_of(wk, e, 1), "e2e": e["e2e"],
                     "gpu_launches": e["gpu_launches"]}
                if not wk.two_crop:
                    b["roofline"] = tensor_roofline(e, e["anchors_per_gpu"], e["anchors"], args.steps, peaks,
                                                    ncu_traffic(k, 1))
                else:
                    b["config"]["pixel_anchors"], b["config"]["image_rows"] = e["pixel_anchors"], e["image_rows"]
                    b["config"]["step"] = "(supcon + pixel) / batch_size through DoublyContrastiveLoss, one backward (trainer.py:143-158)"
                blocks[k] = b
            line["workloads"] = blocks
        if world == 1 and args.workload == "auto" and not args.no_train_step:
            # BASELINE.json configs[4] on ONE of its eight GPUs: the full SwiftNet-RN18 training step (network in
            # cuDNN under bf16 autocast, the three losses of this repository, fused Adam), 8 images x 2 crops per GPU
            from tools import train_bench
            t5 = train_bench.run(train_bench.parse(["--steps", str(min(args.steps, 10)), "--warmup", "5"]), init_dist=False)
            line.setdefault("workloads", {})["cfg5"] = {
                "value": t5["value"], "unit": t5["unit"], "ms_per_step": t5["ms_per_step"], "config": t5["config"],
                "losses": t5["losses"], "peak_mem_gb": t5["peak_mem_gb"],
                "note": "per-GPU share of configs[4] (batch 64 over 8 GPUs = 8 images per GPU); tools/train_bench.py under "
                        "torchrun runs it on N GPUs (DDP, pixel term sharded)"}
        if world == 1 and not args.no_hbm:
            hb = hbm_kernels(L, _lib, dev, peaks)
            top = "gap_fwd"
            line["roofline_hbm"] = {"bound": "hbm", "achieved": hb[top]["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": hb[top]["frac"], "kernel": "k_gap_fwd (largest HBM mover of the doubly step)",
                                    "peak_source": peaks["source"], "kernels": hb,
                                    "note": "each kernel alone, CUDA events, L2 flushed between launches; bytes = SURVEY 8d algorithmic bytes"}
        if world == 1 and not args.no_cpu_baseline:
            from doubly_contrastive_semseg_b200.synthetic import make_inputs as mk
            wc = WORKLOADS["cfg2"]
            torch.set_num_threads(os.cpu_count() or 1)
            cpu_data = mk(wc, seed=1, device="cpu")
            crit_cpu, kind = cpu_module(wc)
            n_cpu = cpu_anchors(wc, cpu_data)
            dt, _ = cpu_step(crit_cpu, cpu_data, 100)
            line["cpu_baseline"] = {"value": n_cpu / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                                    "sample": "1 fwd+bwd of cfg2 (%d anchors; the same images as %s, max_samples 8192: the "
                                              "reference cannot hold 65536 anchors) on the host, %s, %.1f s"
                                              % (n_cpu, wl.name, "the reference's own utils/loss.py (oracle/_ref)"
                                                 if kind == "reference" else "oracle port", dt)}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def workload_named(name):
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS
    return WORKLOADS[name]


def main():
    if os.environ.get("DCL_BENCH_NO_GC"):                   # experiments only
        import gc
        gc.disable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hbm", action="store_true")
    ap.add_argument("--no-train-step", action="store_true", help="skip the cfg5 block (full training step)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
