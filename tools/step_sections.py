"""Where one PixelContrastLoss step spends its host time (dcl_step_timing) and its wall time, per workload.
    python tools/step_sections.py [workload] [steps]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg                                   # noqa: E402
from doubly_contrastive_semseg_b200 import _lib, loss as L                    # noqa: E402
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs   # noqa: E402

wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
d = make_inputs(wl, seed=1, device="cuda")
crit = pkg.PixelContrastLoss(device="cuda")
crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
x = d["feats"].requires_grad_(True)
torch.manual_seed(1234)
lib = _lib.load()
names = ["issue classify", "+count wait", "+host plan", "+upload queued", "+plan issued", "+sel/gather/fwd", "+bwd issued"]
for mode in ("default", "same-stream fill", "host plan"):
    L._SIDE_STREAM_FILL = mode != "same-stream fill"
    L._DEVICE_PLAN = mode != "host plan"
    rows, walls, fw, bw = [], [], [], []
    for s in range(steps):
        x.grad = None
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = crit(x, labels=d["labels"], predict=d["predict"])
        t1 = time.perf_counter()
        loss.backward()
        t2 = time.perf_counter()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        out = (ctypes.c_longlong * 8)()
        lib.dcl_step_timing(out)
        rows.append([out[i] / 1e3 for i in range(7)])
        walls.append((t3 - t0) * 1e6); fw.append((t1 - t0) * 1e6); bw.append((t2 - t1) * 1e6)
    st = (ctypes.c_longlong * 4)()
    lib.dcl_host_lookahead_stats(st)
    med = np.median(np.array(rows[3:]), axis=0)
    print("%s / %s: synced step %.0f us (forward call %.0f, backward call %.0f); device plan %d; lookahead stats %s"
          % (wl.name, mode, np.median(walls[3:]), np.median(fw[3:]), np.median(bw[3:]), crit.last_plan is not None and
             int(crit.__dict__.get("_sb")["info"][5]), list(st)))
    print("   " + "  ".join("%s %.0f" % (n, v) for n, v in zip(names, med)))
# back-to-back (no sync between steps): what the bench measures
for mode in ("default", "same-stream fill", "host plan"):
    L._SIDE_STREAM_FILL = mode != "same-stream fill"
    L._DEVICE_PLAN = mode != "host plan"
    for s in range(5):
        x.grad = None
        crit(x, labels=d["labels"], predict=d["predict"]).backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(20):
        x.grad = None
        crit(x, labels=d["labels"], predict=d["predict"]).backward()
    b.record()
    torch.cuda.synchronize()
    print("%s / %s: back-to-back %.1f us per step" % (wl.name, mode, a.elapsed_time(b) / 20 * 1e3))
