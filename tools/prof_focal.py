"""Minimal driver for ncu on the fused BoundaryAwareFocalLoss kernel at the cfg2 shapes (batch 8, labels 1024x2048,
logits 256x512, 19 classes): `reps` forward + backward calls."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg   # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator(device="cuda").manual_seed(3)
B, C, h, w, H, W = 8, 19, 256, 512, 1024, 2048
logits = (2.0 * torch.randn(B, C, h, w, generator=g, device="cuda")).requires_grad_(True)
target = torch.randint(0, C, (B, H, W), generator=g, device="cuda")
alpha = torch.rand(B, H, W, generator=g, device="cuda")
opts = types.SimpleNamespace(with_depth_level_loss=False, criterion="x", no_class_weights=False, no_EDT=False)
crit = pkg.BoundaryAwareFocalLoss(gamma=0.5, num_classes=C, ignore_id=255, weight=torch.ones(C, device="cuda"),
                                  device="cuda", opts=opts)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(reps):
    logits.grad = None
    if i == reps - 1:
        ev[0].record()
    loss = crit(logits, target, {"label_distance_weight": alpha})
    loss.backward()
    if i == reps - 1:
        ev[1].record()
torch.cuda.synchronize()
print("focal loss", float(loss.detach()), "last fwd+bwd %.1f us" % (ev[0].elapsed_time(ev[1]) * 1e3))
if len(sys.argv) > 2 and sys.argv[2] == "torch":
    # the same loss as plain PyTorch on the GPU (the oracle's restatement of loss.py:27-80: up-sampled logits, log-softmax,
    # gather, weights), for scale only
    from oracle import dcl_oracle as O
    port = O.BoundaryFocalPort(gamma=0.5, num_classes=C, ignore_id=255, weight=torch.ones(C, device="cuda"), device="cuda", opts=opts)
    xr = logits.detach().clone().requires_grad_(True)
    for i in range(3):
        xr.grad = None
        if i == 2:
            torch.cuda.synchronize()
            ev[0].record()
        lr_ = port(xr, target.clone(), {"label_distance_weight": alpha})
        lr_.backward()
    ev[1].record()
    torch.cuda.synchronize()
    print("plain torch on the GPU: loss", float(lr_.detach()), "fwd+bwd %.1f us, peak memory %.2f GB"
          % (ev[0].elapsed_time(ev[1]) * 1e3, torch.cuda.max_memory_allocated() / 2 ** 30))
