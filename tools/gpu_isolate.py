"""Component isolation timing of the contrast kernels via dcl_debug_flags (diagnostics)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
from tools.gpu_time import timed

lib = _lib.load()
for n in (16384, 65536):
    g = torch.Generator(device="cuda").manual_seed(n)
    y = torch.randint(0, 16, (n,), generator=g, device="cuda").sort().values.int()
    Z = torch.randn(n, 128, generator=g, device="cuda")
    tiles, sq = L.pack_rows(Z, n)
    nJ = n // 128
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
    for flags, name in ((0, "full"), (1, "no epilogue math"), (2, "no S MMA"), (3, "no epi, no S MMA (load pipeline only)"),
                        (4, "no dF MMA"), (6, "no MMA at all"), (7, "nothing but loads+barriers")):
        lib.dcl_debug_flags(flags)
        f = timed(lambda: L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07), iters=5, warm=2)
        b = timed(lambda: L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0), iters=5, warm=2)
        print(f"n={n} flags={flags} ({name}): fwd {f[0]:.0f} us  bwd {b[0]:.0f} us", flush=True)
    lib.dcl_debug_flags(0)
