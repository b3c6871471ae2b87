"""Wall-clock sections of one module step (host side), cfg2."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg
from doubly_contrastive_semseg_b200 import loss as L, _lib
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
d = make_inputs(wl, seed=1, device="cuda")
crit = pkg.PixelContrastLoss(device="cuda")
crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
feats = d["feats"][: wl.B].contiguous().requires_grad_(True)
labels, predict = d["labels"], d["predict"]
lib = _lib.load()
pc = time.perf_counter
acc = {}
def add(k, t): acc[k] = acc.get(k, 0.0) + t
import ctypes
for it in range(25):
    feats.grad = None
    torch.cuda.synchronize()
    t0 = pc()
    B, C, h, w = feats.shape
    hb = crit._host_buffers(B, feats.device)
    code, chunk, counts = L.classify(labels, predict, h, w); t1 = pc()
    hb["counts"].copy_(counts.view(-1), non_blocking=True); hb["event"].record()
    dz = torch.zeros_like(feats); t2 = pc()
    hb["event"].synchronize(); t3 = pc()
    st = torch.get_rng_state(); sbuf = st.numpy()
    cap, stage, an, rows, info = hb["cap"], hb["stage_np"], hb["anchors"], hb["rows"], hb["info"]
    rc = lib.dcl_host_plan_rows(hb["counts_np"].ctypes.data, B, 255, wl.max_samples, wl.max_views, sbuf.ctypes.data, sbuf.nbytes,
                                info.ctypes.data, an[0].ctypes.data, an[1].ctypes.data, an[2].ctypes.data, an[3].ctypes.data,
                                an[4].ctypes.data, hb["ranks"].ctypes.data, stage.ctypes.data, stage[cap * 4:].ctypes.data,
                                rows[0].ctypes.data, rows[1].ctypes.data); t4 = pc()
    torch.set_rng_state(st)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
    e0.record()
    A, n_view, n, n_pad = (int(v) for v in info)
    packed = hb["stage"].to(feats.device, non_blocking=True)
    pix = L.select_pixels(code, chunk, B, h * w, packed[: n_pad * 4], n_pad); t5 = pc()
    loss = L._PixelContrastFn.apply(feats, pix, packed[cap * 4: cap * 4 + n_pad], n, 0.07, 0.07, dz); t6 = pc()
    e1.record()
    loss.backward(); t7 = pc()
    e2.record()
    torch.cuda.synchronize(); t8 = pc()
    if it >= 5:
        add("gpu: plan-end -> fwd done", e0.elapsed_time(e1) * 1e-3)
        add("gpu: fwd done -> bwd done", e1.elapsed_time(e2) * 1e-3)
    if it >= 5:
        for k, v in (("classify launch", t1 - t0), ("copy+zero launch", t2 - t1), ("sync wait", t3 - t2), ("C plan", t4 - t3),
                     ("h2d+select launch", t5 - t4), ("fn.forward launch", t6 - t5), ("backward launch", t7 - t6), ("drain", t8 - t7)):
            add(k, v)
for k, v in acc.items(): print(f"{k:22s} {v / 20 * 1e6:8.1f} us")
print("total", sum(acc.values()) / 20 * 1e6)
stats = (ctypes.c_longlong * 4)(); lib.dcl_host_lookahead_stats(stats); print("lookahead stats (stream, inline, starts, drops):", list(stats))
