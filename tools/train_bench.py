"""BASELINE config 5 on this framework: the full SwiftNet-RN18 training step with the doubly contrastive loss and the
boundary-aware focal loss, `--batch` images per GPU (two crops each) at 1024x512, synthetic ACDC-shaped data.

    python tools/train_bench.py [--batch 8] [--steps 10] [--warmup 3] [--no-amp]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py ...

Prints one JSON line (rank 0): images/s over all ranks, ms per step (CUDA events, max over ranks), the share of the
step spent in the losses (forward + their part of the backward is not separable; the forward calls are timed), and
whether the ranks' weights are still identical after the last step."""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--no-amp", action="store_true")
    ap.add_argument("--max-samples", type=int, default=0,
                    help="PixelContrastLoss.max_samples; 0 = max(1024, 128 x global batch): the reference's 1024 gives "
                         "n_view = 0 (and the reference crashes, loss.py:345) once the global batch holds more than 1024 "
                         "(image, class) pairs, e.g. 64 images x 19 classes")
    ap.add_argument("--no-cudnn-benchmark", action="store_true")
    ap.add_argument("--nchw", action="store_true", help="contiguous NCHW activations instead of channels_last")
    ap.add_argument("--profile", action="store_true", help="CUPTI kernel table of two steps after the timed ones (stderr)")
    return ap.parse_args(argv)


def run(a, init_dist=True):
    """-> the result dict on rank 0 (None elsewhere)"""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.backends.cudnn.benchmark = not a.no_cudnn_benchmark
    if world > 1 and init_dist:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    import doubly_contrastive_semseg_b200 as pkg
    from doubly_contrastive_semseg_b200.swiftnet import fill_deterministic
    dev = torch.device("cuda", local)
    opts = types.SimpleNamespace(amp=not a.no_amp, batch_size=a.batch * world, channels_last=not a.nchw)
    step = pkg.TrainStep(opts, device=dev)
    fill_deterministic(step.net, 1)
    step.pixelcontrast_criterion.max_samples = a.max_samples if a.max_samples > 0 else max(1024, 128 * a.batch * world)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    B, H, W = a.batch, a.height, a.width
    coarse = torch.randint(0, 19, (B, H // 64, W // 64), generator=g, device=dev)
    labels = coarse.repeat_interleave(64, 1).repeat_interleave(64, 2).contiguous()
    labels[torch.rand(B, H, W, generator=g, device=dev) < 0.05] = 255
    sample = {"left": torch.rand(2 * B, 3, H, W, generator=g, device=dev) * 255.0,
              "label": labels,
              "weather": (torch.arange(B, device=dev) + rank) % 4,
              "label_distance_weight": torch.rand(B, H, W, generator=g, device=dev)}
    pristine = sample["label"].clone()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out = None
    for i in range(a.warmup + a.steps):
        if i == a.warmup:
            if world > 1:
                torch.distributed.barrier()
            torch.cuda.synchronize()
            ev[0].record()
        sample["label"].copy_(pristine)                 # the focal loss rewrites ignore -> 0 in place
        out = step(sample)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / a.steps
    t = torch.tensor([ms], device=dev)
    sync = True
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        w = torch.cat([p.detach().flatten()[:64] for p in step.net.parameters()])
        lo, hi = w.clone(), w.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        sync = bool(torch.equal(lo, hi))
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                sample["label"].copy_(pristine)
                step(sample)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70), file=sys.stderr)
    res = None
    if rank == 0:
        res = ({"workload": "cfg5", "metric": "train_step_images_per_sec", "value": B * world / (float(t) * 1e-3),
                          "unit": "images/s", "n_gpus": world, "ms_per_step": float(t), "steps": a.steps, "warmup": a.warmup,
                          "config": {"model": "SwiftNet-RN18 pyramid (random init)", "images_per_gpu": B, "crops_per_image": 2,
                                     "image_hw": [H, W], "embed_hw": [H // 4, W // 4], "criterion": step.opts.criterion, "max_samples": step.pixelcontrast_criterion.max_samples,
                                     "max_views": step.pixelcontrast_criterion.max_views,
                                     "amp_bf16": not a.no_amp, "channels_last": not a.nchw, "cudnn_benchmark": not a.no_cudnn_benchmark, "optimizer": "Adam (fused), two lr groups"},
                          "losses": {k: float(v) for k, v in out.items()}, "weights_in_sync": sync,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
    if world > 1 and init_dist:
        torch.distributed.destroy_process_group()
    return res


def main():
    res = run(parse())
    if res is not None:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
