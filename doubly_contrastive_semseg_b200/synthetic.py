"""Synthetic ACDC-shaped inputs for tests and bench: no dataset, no network.

Configurations follow BASELINE.json `configs`; every tensor is produced from a seeded generator
on the requested device so the CPU oracle and the CUDA path can be fed identical data.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class Workload:
    name: str
    B: int            # images (first crop)
    H: int            # label resolution
    W: int
    h: int            # embedding resolution (stride 4)
    w: int
    K: int            # label classes actually present
    max_samples: int
    max_views: int
    two_crop: bool = False   # image-level term: features are [2B,...]


WORKLOADS = {
    # reference defaults on the CPU-runnable case: N = 2*19*2 = 76
    "cfg1": Workload("cfg1", 2, 512, 1024, 128, 256, 19, 1024, 2),
    "cfg1_n988": Workload("cfg1_n988", 2, 512, 1024, 128, 256, 19, 1024, 32),
    # headline single-GPU case: A = 16*8 = 128 anchors x 64 views = 8192 rows
    "cfg2": Workload("cfg2", 8, 1024, 2048, 256, 512, 16, 8192, 64),
    # doubly contrastive: pixel term on the first 16 crops + 32x32 image-level term
    "cfg3": Workload("cfg3", 16, 1024, 2048, 256, 512, 16, 8192, 64, two_crop=True),
    # 65536 rows (sharded across 2/4/8 GPUs)
    "cfg4": Workload("cfg4", 8, 1024, 2048, 256, 512, 16, 65536, 512),
    # small shapes for quick tests
    "tiny": Workload("tiny", 2, 64, 128, 16, 32, 5, 1024, 8),
    "small": Workload("small", 4, 256, 512, 64, 128, 16, 2048, 32),
}


def nearest_index(out_size: int, in_size: int, device) -> torch.Tensor:
    """Legacy 'nearest' source index (float32 scale), as F.interpolate(mode='nearest')."""
    scale = torch.tensor(float(in_size), dtype=torch.float32) / torch.tensor(float(out_size), dtype=torch.float32)
    idx = torch.floor(torch.arange(out_size, dtype=torch.float32) * scale).long().clamp(max=in_size - 1)
    return idx.to(device)


def make_inputs(wl: Workload, seed: int = 0, device="cpu", block: int = 32, ignore_frac: float = 0.05,
                correct_frac: float = 0.7):
    """-> dict(feats [B or 2B,128,h,w] f32, labels [B,H,W] i64, predict [B,19,h,w] f32, weather [B,1] i64)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(1000 + seed)
    B, H, W, h, w = wl.B, wl.H, wl.W, wl.h, wl.w
    bh, bw = (H + block - 1) // block, (W + block - 1) // block
    coarse = torch.randint(0, wl.K, (B, bh, bw), generator=g, device=dev)
    labels = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].contiguous()
    labels[torch.rand((B, H, W), generator=g, device=dev) < ignore_frac] = 255
    iy, ix = nearest_index(h, H, dev), nearest_index(w, W, dev)
    lab_ds = labels[:, iy][:, :, ix]                                   # [B,h,w]
    centroid = 0.5 * torch.randn((19, 128), generator=g, device=dev)
    nb = 2 * B if wl.two_crop else B
    feats = 0.5 * torch.randn((nb, 128, h, w), generator=g, device=dev)
    cidx = lab_ds.clamp(max=18)
    add = centroid[cidx].permute(0, 3, 1, 2) * (lab_ds != 255).unsqueeze(1)
    feats[:B] += add
    if wl.two_crop:
        feats[B:] += add
    predict = torch.randn((B, 19, h, w), generator=g, device=dev)
    boost = (torch.rand((B, h, w), generator=g, device=dev) < correct_frac) & (lab_ds != 255)
    predict.scatter_add_(1, cidx.unsqueeze(1), 3.0 * boost.unsqueeze(1).float())
    weather = torch.randint(0, 4, (B, 1), generator=g, device=dev)
    return dict(feats=feats, labels=labels.long(), predict=predict, weather=weather)
