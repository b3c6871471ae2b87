"""Minimal driver for ncu on the HBM-bound kernels: one cfg2 pixel step (classify, prefix, select, gather, scatter)
and one GAP forward / backward at the cfg3 shape ([32,128,256,512], 2.15 GB each way)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg   # noqa: E402
from doubly_contrastive_semseg_b200.loss import _GapFn   # noqa: E402
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs   # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = WORKLOADS["cfg2"]
d = make_inputs(wl, seed=1, device="cuda")
crit = pkg.PixelContrastLoss(device="cuda")
crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
x = d["feats"].requires_grad_(True)
torch.manual_seed(7)
for _ in range(reps):
    x.grad = None
    loss = crit(x, labels=d["labels"], predict=d["predict"])
    loss.backward()
torch.cuda.synchronize()
print("pixel loss", float(loss))
del d, x
torch.cuda.empty_cache()
# the doubly step at the cfg3 shape: global average pool over [32,128,256,512] (2.15 GB), one-pass dense gradient
import types
wl3 = WORKLOADS["cfg3"]
d3 = make_inputs(wl3, seed=1, device="cuda")
both = pkg.DoublyContrastiveLoss(device="cuda", opts=types.SimpleNamespace(deeplab=False))
both.pixel.max_samples, both.pixel.max_views = wl3.max_samples, wl3.max_views
big = d3["feats"].requires_grad_(True)
for _ in range(reps):
    big.grad = None
    sup, pix = both(big, labels=d3["labels"], predict=d3["predict"], class_labels=d3["weather"])
    ((sup + pix) / wl3.B).backward()
torch.cuda.synchronize()
print("doubly", float(sup.detach()), float(pix.detach()))
