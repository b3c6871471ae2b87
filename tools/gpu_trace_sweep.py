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
buf = torch.zeros(3 * 32 * 4, dtype=torch.int64, device="cuda")
lib.dcl_debug_trace(buf.data_ptr())
L.contrast_forward(tiles, y, sq, nJ, 0, nJ, n, 0, 0.07, 0.07)
torch.cuda.synchronize()
lib.dcl_debug_trace(None)
t = buf.cpu().view(3, 32, 4)
base = int(t[:, 2:, :][t[:, 2:, :] > 0].min())
print("sweep B (tiles >= 2 keep B's stamps; C only overwrites tile 0/1)")
print("tile | producer: wait_empty got_empty | mma: wait_full got_full got_tempty issued | epilogue(g0,w0): start got_tfull done")
for it in range(2, 26):
    r = lambda role, ev: (int(t[role, it, ev]) - base) if int(t[role, it, ev]) else -1
    print(f"{it:4d} | {r(0,0):7d} {r(0,1):7d} | {r(1,0):7d} {r(1,1):7d} {r(1,2):7d} {r(1,3):7d} | {r(2,0):7d} {r(2,1):7d} {r(2,2):7d}")
