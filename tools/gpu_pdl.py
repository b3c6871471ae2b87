"""GPU-side time of the contrast forward / backward chains with and without programmatic dependent launch.
The host is kept ahead of the device (a spin kernel is queued first) so the events bracket device time only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib

def timed(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400_000)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts)), float(np.min(ts))

def main():
    K = 16
    lib = _lib.load()
    for n in [int(a) for a in sys.argv[1:]] or [1024, 8192, 16384, 65536]:
        g = torch.Generator(device="cuda").manual_seed(n)
        y = torch.randint(0, K, (n,), generator=g, device="cuda").sort().values.int()
        Z = torch.randn(n, 128, generator=g, device="cuda")
        n_pad = (n + 127) // 128 * 128
        tiles, sq = L.pack_rows(Z, n_pad); nJ = n_pad // 128
        ref = None
        for flags in (16, 0, 16, 0):
            lib.dcl_debug_flags(flags)
            f = timed(lambda: L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07))
            colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
            b = timed(lambda: L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0))
            dF = L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
            torch.cuda.synchronize()
            if ref is None: ref = (ls.clone(), dF.clone())
            same = torch.equal(ref[0], ls) and torch.equal(ref[1], dF)
            tot = f[0] + b[0]
            fl = 6.0 * n * n * 128
            print(f"n={n:6d} pdl={'off' if flags else 'on '}: fwd {f[0]:7.1f} (min {f[1]:7.1f}) bwd {b[0]:7.1f} (min {b[1]:7.1f}) "
                  f"total {tot:7.1f} us -> {fl / tot / 1e6 / 1654.3 * 100:5.1f}% ; identical to plain: {same}", flush=True)
        lib.dcl_debug_flags(0)

if __name__ == "__main__":
    main()
