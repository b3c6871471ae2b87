"""Host-side (Python) cost of a PixelContrastLoss step around the C call, and of its pieces in isolation."""
import os, sys, time, ctypes
os.environ["DCL_DEBUG_PY_TIMES"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import doubly_contrastive_semseg_b200 as pkg
from doubly_contrastive_semseg_b200 import _lib, loss as L
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
wl = WORKLOADS["cfg2"]
d = make_inputs(wl, seed=1, device="cuda")
crit = pkg.PixelContrastLoss(device="cuda")
crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
x = d["feats"].requires_grad_(True)
torch.manual_seed(1234)
fw, bw = [], []
for s in range(40):
    x.grad = None
    t0 = time.perf_counter()
    loss = crit(x, labels=d["labels"], predict=d["predict"])
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    fw.append((t1 - t0) * 1e6); bw.append((t2 - t1) * 1e6)
torch.cuda.synchronize()
py = np.array(L._DEBUG_PY_TIMES[10:]) * 1e3
print("forward call us median", np.median(fw[10:]), "backward call", np.median(bw[10:]))
print("_run_step sections us (buffers+dzero, carve, rng+struct, C call):", np.median(py, axis=0))
# pieces in isolation
class _Id(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        return a.view_as(a)
    @staticmethod
    def backward(ctx, g):
        return g
def t(f, n=200):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
print("_check_inputs %.1f" % t(lambda: crit._check_inputs(x, d["labels"], d["predict"])))
print("_verify_host_rng %.1f" % t(L._verify_host_rng))
print("torch.cuda.device ctx %.1f" % t(lambda: torch.cuda.device(x.device).__enter__()))
print("_stream %.1f" % t(L._stream))
print("empty_like feats %.1f" % t(lambda: torch.empty_like(x)))
print("autograd Function.apply of a no-op %.1f" % t(lambda: _Id.apply(x)))
