"""Per-role timeline of CTA 0 of the forward power-sum sweep (sweep C only overwrites the first few tiles: labels
are all distinct, so only the diagonal tile overlaps)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(n)
K = int(sys.argv[3]) if len(sys.argv) > 3 else 0
y = torch.arange(n, device="cuda").int() if K == 0 else torch.randint(0, K, (n,), generator=g, device="cuda").sort().values.int()
Z = torch.randn(n, 128, generator=g, device="cuda")
tiles, sq = L.pack_rows(Z, n)
nJ = n // 128
for _ in range(2):
    L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
lib.dcl_debug_flags(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
buf = torch.zeros(5 * 32 * 8, dtype=torch.int64, device="cuda")
lib.dcl_debug_trace(buf.data_ptr())
L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
torch.cuda.synchronize()
lib.dcl_debug_trace(None)
t = buf.cpu().view(5, 32, 8)
base = int(t[0, 0, 7])
print("kernel entry 0, set-up done", int(t[0, 2, 7]) - base, ", all roles done", int(t[0, 1, 7]) - base)
print("g0 first unit: iterator built", int(t[3,31,3])-base, "loads issued", int(t[3,31,0])-base, "loads arrived", int(t[3,31,1])-base, "A stored", int(t[3,31,2])-base)
print(f"forward sweep, n={n}")
print("tile | prod: wait_e got_e | issuer(it%2): wait_full got_full got_te0 got_turn got_te1 issued | g0: start got_tfull done | g1: start got_tfull done")
for it in range(0, 30):
    r = lambda role, ev: (int(t[role, it, ev]) - base) if int(t[role, it, ev]) else -1
    i = 1 + (it & 1)
    print(f"{it:3d} | {r(0,0):6d} {r(0,1):6d} | {r(i,0):6d} {r(i,1):6d} {r(i,2):6d} {r(i,4):6d} {r(i,5):6d} {r(i,3):6d} | "
          f"{r(3,0):6d} {r(3,1):6d} {r(3,2):6d} | {r(4,0):6d} {r(4,1):6d} {r(4,2):6d}")
rows = [int(t[0, k, 6]) for k in range(7)]
if rows[0]:
    print("k_rows block 0 (clk from its start): after cmax", rows[1] - rows[0], "partials", rows[2] - rows[0], "row_scale", rows[3] - rows[0],
          "poly sums", rows[4] - rows[0], "finalize+stores", rows[5] - rows[0], "loss sum", rows[6] - rows[0])
