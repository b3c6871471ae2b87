"""torch.profiler timeline of the module step (GPU kernels, memsets, copies):
    python tools/torch_profile.py [cfg2|cfg4]           pixel term
    python tools/torch_profile.py cfg3 doubly           both terms on the two-crop batch (DoublyContrastiveLoss)"""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
d = make_inputs(wl, seed=1, device="cuda")
doubly = len(sys.argv) > 2 and sys.argv[2] == "doubly"
if doubly:
    import types
    crit = pkg.DoublyContrastiveLoss(device="cuda", opts=types.SimpleNamespace(deeplab=False))
    crit.pixel.max_samples, crit.pixel.max_views = wl.max_samples, wl.max_views
    feats = d["feats"].requires_grad_(True)
    def step():
        feats.grad = None
        sup, pix = crit(feats, labels=d["labels"], predict=d["predict"], class_labels=d["weather"])
        ((sup + pix) / wl.B).backward()
else:
    crit = pkg.PixelContrastLoss(device="cuda")
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    feats = d["feats"][: wl.B].contiguous().requires_grad_(True)
    def step():
        feats.grad = None
        loss = crit(feats, labels=d["labels"], predict=d["predict"])
        loss.backward()
torch.manual_seed(1)
for _ in range(6): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last step only
n = len(evs) // 5
last = evs[-n:]
t0 = last[0].time_range.start
for e in last:
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:90]}")
print("step span %.1f us; sum of kernel times %.1f us" % (last[-1].time_range.end - t0, sum(e.time_range.end - e.time_range.start for e in last)))
