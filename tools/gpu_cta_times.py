"""Per-CTA entry / exit times (globaltimer, clock64) of the sweep-P, k_rows and backward kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
g = torch.Generator(device="cuda").manual_seed(n)
y = torch.randint(0, K, (n,), generator=g, device="cuda").sort().values.int()
Z = torch.randn(n, 128, generator=g, device="cuda")
n_pad = (n + 127) // 128 * 128
tiles, sq = L.pack_rows(Z, n_pad); nJ = n_pad // 128
def step():
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
    return L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
for _ in range(5): step()
buf = torch.zeros(4 * 256 * 4, dtype=torch.int64, device="cuda")
lib.dcl_debug_cta_times(buf.data_ptr())
torch.cuda._sleep(400_000)
step(); torch.cuda.synchronize()
lib.dcl_debug_cta_times(None)
t = buf.cpu().numpy().reshape(4, 256, 4)
t0 = None
for slot, name in ((0, "sweep P"), (2, "k_rows"), (1, "backward")):
    a = t[slot]; m = a[:, 0] > 0; a = a[m]
    if t0 is None: t0 = a[:, 0].min()
    dur_ns = a[:, 2] - a[:, 0]; dur_clk = a[:, 3] - a[:, 1]
    print(f"{name:9s}: {m.sum():3d} CTAs; first entry {(a[:,0].min()-t0)/1e3:7.2f} us, last entry {(a[:,0].max()-t0)/1e3:7.2f}, "
          f"first exit {(a[:,2].min()-t0)/1e3:7.2f}, last exit {(a[:,2].max()-t0)/1e3:7.2f} | per-CTA us min/med/max "
          f"{dur_ns.min()/1e3:.2f}/{np.median(dur_ns)/1e3:.2f}/{dur_ns.max()/1e3:.2f} | clk min/med/max {dur_clk.min()}/{int(np.median(dur_clk))}/{dur_clk.max()} "
          f"| MHz {np.median(dur_clk / np.maximum(dur_ns, 1)) * 1e3:.0f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", f"cta_times_{n}_{K}.npz"), t=t, y=y.cpu().numpy(), n=n)
