"""Offline regression of per-CTA kernel durations (tools/gpu_cta_times.py dumps) on tiles / mixed tiles per CTA."""
import sys, math
import numpy as np
def part(nU, nJ, G0, sw):
    total = nU * nJ; G = min(total, G0); excl = False; base, extra = 1, 0
    if nU <= G:
        b, e = G // nU, G % nU
        if (nJ + b - 1) // b <= (total + G - 1) // G + (0 if G % nU == 0 else sw):
            excl, base, extra = True, b, e
    def begin(c):
        if not excl: return c * total // G
        wide = extra * (base + 1)
        if c < wide: u = c // (base + 1); k = c - u * (base + 1); parts = base + 1
        else:
            d = c - wide; u = extra + d // base; k = d - (u - extra) * base; parts = base
        return u * nJ + k * nJ // parts
    return G, begin, excl
for f in sys.argv[1:]:
    d = np.load(f); t = d["t"]; y = d["y"]; n = int(d["n"]); nJ = n // 128
    yb = y.reshape(nJ, 128); lo = yb.min(1); hi = yb.max(1)
    ov = (lo[:, None] <= hi[None, :]) & (lo[None, :] <= hi[:, None])
    s = int(nJ * 0.381966 + 0.5)
    while math.gcd(s, nJ) != 1: s += 1
    js = s % nJ
    for slot, name, unit_rows, sw in ((0, "sweep P", 2, 4), (1, "backward", 1, 12)):
        nU = nJ // unit_rows
        G, begin, excl = part(nU, nJ, 148, sw)
        a = t[slot]; clk = (a[:, 3] - a[:, 1])[:G]
        rows = []
        for b in range(G):
            t0, t1 = begin(b), begin(b + 1); m = 0; segs = set()
            for x in range(t0, t1):
                U = x // nJ; J = ((x % nJ) * js) % nJ
                m += sum(ov[unit_rows * U + g, J] for g in range(unit_rows)); segs.add(U)
            rows.append((t1 - t0, m, len(segs)))
        r = np.array(rows, float)
        A = np.c_[r, np.ones(G)]
        coef = np.linalg.lstsq(A, clk, rcond=None)[0]
        print(f"{f.split('/')[-1]:28s} {name:9s} excl={int(excl)} clk min/med/max {clk.min()}/{int(np.median(clk))}/{clk.max()}  fit: {coef[0]:.0f}/tile + {coef[1]:.0f}/mixed + {coef[2]:.0f}/unit + {coef[3]:.0f}  (resid {np.std(A @ coef - clk):.0f})")
