"""Backward / forward time when no tile (distinct labels), every tile (one class) or half of the tiles (two classes)
can hold same-class pairs: isolates the cost of the mixed-tile paths."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L
from tools.gpu_pdl import timed
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(n)
Z = torch.randn(n, 128, generator=g, device="cuda")
tiles, sq = L.pack_rows(Z, n); nJ = n // 128
for name, y in (("distinct", torch.arange(n, device="cuda").int() % 200 + 0 * 0),   # labels 0..199 cyclic: unsorted, every tile overlaps
                ("sorted-many", (torch.arange(n, device="cuda") // 128).int() % 256),
                ("one class", torch.zeros(n, device="cuda").int()),
                ("two classes", (torch.arange(n, device="cuda") >= n // 2).int()),
                ("16 classes", (torch.arange(n, device="cuda") * 16 // n).int())):
    f = timed(lambda: L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07))
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
    b = timed(lambda: L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0))
    print(f"n={n} {name:12s}: fwd {f[0]:8.1f} us  bwd {b[0]:8.1f} us", flush=True)
