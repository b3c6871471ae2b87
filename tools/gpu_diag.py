"""Row-level diagnostic of one contrast case: which rows deviate from the oracle, v3 and legacy forward."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
from oracle import dcl_oracle as O
lib = _lib.load()
n, K = int(sys.argv[1]), int(sys.argv[2])
T = 0.07
g = torch.Generator().manual_seed(0)
y = torch.randint(0, K, (n,), generator=g).sort().values
cent = torch.randn(K, 128, generator=g)
Z = 0.5 * torch.randn(n, 128, generator=g) + 0.5 * cent[y]
Zb = Z.to(torch.bfloat16).float()
loss_o, dF_o, st = O.contrast_closed_form(Zb, y, T, T, 0)
n_pad = (n + 127) // 128 * 128
tiles, sq = L.pack_rows(Z.cuda().contiguous(), n_pad)
ypad = torch.full((n_pad,), -1, dtype=torch.int32, device="cuda"); ypad[:n] = y.cuda().int()
nJ = n_pad // 128
kap = 1.0 / (T * st["r"])
for flags, name in ((0, "v3"), (8, "legacy")):
    lib.dcl_debug_flags(flags)
    for rep in range(2):
        colA, colB, rl, ls = L.contrast_forward(tiles, ypad, sq, nJ, 0, nJ, n, 0, T, T)
        torch.cuda.synchronize()
        cA, cB = colA[:n].cpu().double(), colB[:n].cpu().double()
        for nm, dev, ref in (("a", cA[:, 0], kap * 1.4426950408889634), ("den", cB[:, 1], st["neg"]), ("rowloss", rl[:n].cpu().double(), st["rowloss"]),
                             ("q", cA[:, 3], -kap * st["Q"]), ("p", cA[:, 2], -kap * st["R"] * np.log(2.0))):
            err = ((dev - ref).abs() / (ref.abs() + 1e-30)).numpy()
            bad = np.nonzero(err > 1e-3)[0]
            blocks = sorted(set((bad // 128).tolist()))
            print(f"{name} rep{rep} {nm:8s} max rel {err.max():.2e}  bad rows {len(bad)}  blocks {blocks[:24]}  first rows {bad[:8].tolist()}")
lib.dcl_debug_flags(0)
