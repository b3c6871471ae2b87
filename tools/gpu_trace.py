"""Per-role timeline of CTA 0 of k_backward (diagnostics)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(n)
y = torch.randint(0, 16, (n,), generator=g, device="cuda").sort().values.int()
Z = torch.randn(n, 128, generator=g, device="cuda")
tiles, sq = L.pack_rows(Z, n)
nJ = n // 128
colA, colB, rl, ls = L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
for _ in range(2):
    L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
lib.dcl_debug_flags(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
buf = torch.zeros(5 * 32 * 8, dtype=torch.int64, device="cuda")
lib.dcl_debug_trace(buf.data_ptr())
L.contrast_backward(tiles, y, colA, colB, nJ, 0, nJ, 0)
torch.cuda.synchronize()
lib.dcl_debug_trace(None)
t = buf.cpu().view(5, 32, 8)
t0 = int(t[t > 0].min())
print("kernel entry", int(t[0, 0, 7]) - t0, "set-up done", int(t[0, 2, 7]) - t0, "all roles done", int(t[0, 1, 7]) - t0)
print("burst| prod: wait_e got_e | issuer(k%2): wait_full got_full got_pfull got_turn issued | g0: top start got_tfull done | g1: top start got_tfull done")
for it in range(28):
    r = lambda role, ev: (int(t[role, it, ev]) - t0) if int(t[role, it, ev]) else -1
    i = 1 + (it & 1)
    print(f"{it:3d} | {r(0,0):6d} {r(0,1):6d} | {r(i,0):6d} {r(i,1):6d} {r(i,2):6d} {r(i,4):6d} {r(i,3):6d} | "
          f"{r(3,3):6d} {r(3,0):6d} {r(3,1):6d} {r(3,2):6d} | {r(4,3):6d} {r(4,0):6d} {r(4,1):6d} {r(4,2):6d}")
