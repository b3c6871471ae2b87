"""Minimal driver for ncu: pack -> contrast fwd -> contrast bwd on n random rows."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(n)
y = torch.randint(0, 16, (n,), generator=g, device="cuda").sort().values.int()
Z = torch.randn(n, 128, generator=g, device="cuda")
n_pad = (n + 127) // 128 * 128
nJ = n_pad // 128
for _ in range(reps):
    tiles, sq = L.pack_rows(Z, n_pad)
    colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
    dF = L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
torch.cuda.synchronize()
print("loss", float(ls[0].item()) / n, "dF", float(dF.abs().max()))
