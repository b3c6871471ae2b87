"""Per-stage CUDA-event timing of the hot path (run on the B200 box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import doubly_contrastive_semseg_b200 as pkg                     # noqa: E402
from doubly_contrastive_semseg_b200 import loss as L             # noqa: E402
from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs   # noqa: E402


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


def contrast_only(n, K=16):
    g = torch.Generator(device="cuda").manual_seed(n)
    y = torch.randint(0, K, (n,), generator=g, device="cuda").sort().values.int()
    Z = torch.randn(n, 128, generator=g, device="cuda")
    n_pad = (n + 127) // 128 * 128
    tiles, sq = L.pack_rows(Z, n_pad)
    nJ = n_pad // 128
    out = {}
    out["fwd"] = timed(lambda: L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07))
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
    out["bwd"] = timed(lambda: L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0))
    fl = 6.0 * n * n * 128
    tot = out["fwd"][0] + out["bwd"][0]
    print(f"contrast n={n}: fwd {out['fwd'][0]:.1f}us bwd {out['bwd'][0]:.1f}us total {tot:.1f}us -> "
          f"{fl / tot / 1e6:.1f} TFLOP/s algorithmic ({fl / tot / 1e6 / 1654.3 * 100:.1f}% of 1654)", flush=True)


def module(wl_name, iters=10):
    wl = WORKLOADS[wl_name]
    d = make_inputs(wl, seed=1, device="cuda")
    crit = pkg.PixelContrastLoss(device="cuda")
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    feats = d["feats"][: wl.B].contiguous().requires_grad_(True)

    def step():
        feats.grad = None
        loss = crit(feats, labels=d["labels"], predict=d["predict"])
        loss.backward()
        return loss
    med, mn = timed(step, iters=iters, warm=3)
    n = crit.last_layout.n
    print(f"module {wl_name}: N={n} fwd+bwd median {med:.0f}us min {mn:.0f}us -> {n / med * 1e6:.3e} anchors/s", flush=True)
    # stage split
    t = {}
    t["classify"] = timed(lambda: L.classify(d["labels"], d["predict"], wl.h, wl.w))
    code, chunk, counts = L.classify(d["labels"], d["predict"], wl.h, wl.w)
    t0 = time.perf_counter()
    for _ in range(5):
        ch = counts.cpu().numpy().reshape(wl.B, 256, 2)
        plan = L.plan_anchors(ch, 255, wl.max_samples, wl.max_views)
        lay = L.layout_rows(plan, np.arange(plan.A), 0)
    t["host_plan_ms"] = (time.perf_counter() - t0) / 5 * 1e3
    host = torch.from_numpy(np.concatenate([lay.req.reshape(-1), lay.y])).cuda()
    req, y = host[: lay.n_pad * 4], host[lay.n_pad * 4:]
    t["select"] = timed(lambda: L.select_pixels(code, chunk, wl.B, wl.h * wl.w, req, lay.n_pad))
    pix = L.select_pixels(code, chunk, wl.B, wl.h * wl.w, req, lay.n_pad)
    fd = feats.detach()
    t["gather"] = timed(lambda: L.gather_tiles(fd, pix, lay.n_pad))
    tiles, sq = L.gather_tiles(fd, pix, lay.n_pad)
    nJ = lay.n_pad // 128
    t["contrast_fwd"] = timed(lambda: L.contrast_forward(tiles, y, sq, nJ, 0, nJ, lay.n, 0, 0.07, 0.07))
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, lay.n, 0, 0.07, 0.07)
    t["contrast_bwd"] = timed(lambda: L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0))
    dF = L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
    gout = torch.ones((), device="cuda")
    dfe = torch.empty_like(fd)
    from doubly_contrastive_semseg_b200 import _lib
    t["scatter+zero"] = timed(lambda: _lib.call("dcl_scatter_grad", L._p(dF), L._p(pix), lay.n_pad, L._p(gout),
                                               L._p(dfe), wl.B, wl.h * wl.w, 1, L._stream()))
    print("   stages (median us):", {k: (round(v[0], 1) if isinstance(v, tuple) else round(v, 2)) for k, v in t.items()}, flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for n in (1024, 8192, 16384, 65536):
        contrast_only(n)
    module("cfg1")
    module("cfg2")
    module("cfg4", iters=5)
