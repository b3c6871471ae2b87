"""Per-role timeline of CTA 0 of the LAST forward kernel that stamps (sweep C overwrites A/B; run with
sweep selection through separate launches is not possible via the ABI, so this traces the whole
forward and reports what the final stamping kernel left: use n with few relevant tiles)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L, _lib
lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = torch.Generator(device="cuda").manual_seed(n)
# all-distinct labels: no tile pair overlaps except the diagonal -> sweep C stamps at most 1 tile of CTA 0
y = torch.arange(n, device="cuda").int()
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
base = int(t[:, 2:, :][t[:, 2:, :] > 0].min())
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
print(f"sweep B, n={n} (tiles >= 2 keep B's stamps)")
print("tile | prod: wait_e got_e | mma: wait_full got_full got_te0 commit0 got_te1 issued | g0: start got_y got_tfull done | g1: start got_y got_tfull done")
for it in range(2, 26):
    r = lambda role, ev: (int(t[role, it, ev]) - base) if int(t[role, it, ev]) else -1
    print(f"{it:3d} | {r(0,0):6d} {r(0,1):6d} | {r(1,0):6d} {r(1,1):6d} {r(1,2):6d} {r(1,4):6d} {r(1,5):6d} {r(1,3):6d} | "
          f"{r(2,0):6d} {r(2,3):6d} {r(2,1):6d} {r(2,2):6d} | {r(3,0):6d} {r(3,3):6d} {r(3,1):6d} {r(3,2):6d}")
