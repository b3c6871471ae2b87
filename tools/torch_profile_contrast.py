"""torch.profiler (CUPTI) timeline of one contrast forward + backward chain on n random rows."""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
_lib.load().dcl_debug_flags(flags)
g = torch.Generator(device="cuda").manual_seed(n)
y = torch.randint(0, 16, (n,), generator=g, device="cuda").sort().values.int()
Z = torch.randn(n, 128, generator=g, device="cuda")
n_pad = (n + 127) // 128 * 128
nJ = n_pad // 128
tiles, sq = L.pack_rows(Z, n_pad)
def step():
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
    return L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        torch.cuda._sleep(600_000)
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
k = len(evs) // 4
last = evs[-k:]
t0 = last[0].time_range.end
print(f"n={n} flags={flags}")
for e in last[1:]:
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {e.name[:70]}")
print(f"span {last[-1].time_range.end - last[1].time_range.start:.1f} us")
