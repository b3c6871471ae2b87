"""cProfile of the module step on the GPU box: where does the host time of a cfg2 step go?"""
import cProfile, pstats, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
d = make_inputs(wl, seed=1, device="cuda")
crit = pkg.PixelContrastLoss(device="cuda")
crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
feats = d["feats"][: wl.B].contiguous().requires_grad_(True)
def step():
    feats.grad = None
    loss = crit(feats, labels=d["labels"], predict=d["predict"])
    loss.backward()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
torch.cuda.synchronize()
print("step wall us", (time.perf_counter() - t0) / 20 * 1e6)
# host-only time per step (no final sync inside)
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host issue time per step us", (t1 - t0) / 20 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr).sort_stats("cumulative")
st.print_stats(28)
