"""CPU prototype (numpy) of k_focal's work decomposition, checked against the fp64 closed form on the golden fixtures
before the kernel was written: a "thread" owns the run of label columns whose FIRST bilinear tap is its low-resolution
column c, adds the (1 - lx) parts to column c and hands the lx parts to the owner of column c + 1; along Y a thread
walks the label rows whose first tap lies in its strip of low-resolution rows and spills what it collected for the
first row of the next strip.  Every pixel is evaluated exactly once.

    python tools/proto_focal.py            # all tests/golden/focal_*.npz, several strip heights
"""
import glob
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import dcl_oracle as O  # noqa: E402


def first_with_tap_ge(t, i0, out_size):
    """first output index whose first tap is >= t (out_size if none); i0 is monotone"""
    return int(np.searchsorted(i0, t, side="left")) if t > 0 else 0


def emulate(logits, target, alpha, weight, gamma, mode, strip, ignore_id=255):
    x = logits.astype(np.float64)
    B, C, h, w = x.shape
    H, W = target.shape[1:]
    t = target.astype(np.int64).copy()
    t[t == ignore_id] = 0
    y0, y1, ly = O.bilinear_source(H, h)
    x0, x1, lx = O.bilinear_source(W, w)
    G = np.zeros_like(x)
    nst = (h + strip - 1) // strip
    spill = np.zeros((B, nst, C, w))
    loss, cnt = 0.0, 0
    for b in range(B):
        for k in range(nst):
            ys0, ys1 = k * strip, min((k + 1) * strip, h)
            Ylo, Yhi = first_with_tap_ge(ys0, y0, H), first_with_tap_ge(ys1, y0, H)
            accA = np.zeros((C, w))
            accB = np.zeros((C, w))
            ra = ys0
            for Y in range(Ylo, Yhi):
                while y0[Y] > ra:
                    G[b, :, ra, :] = accA
                    accA, accB = accB, np.zeros((C, w))
                    ra += 1
                same = y1[Y] == y0[Y]
                w0, w1 = (1.0, 0.0) if same else (1.0 - float(ly[Y]), float(ly[Y]))
                SR = np.zeros((C, w))
                for c in range(w):                       # one "thread" per low-resolution column
                    c1 = min(c + 1, w - 1)
                    Xlo, Xhi = first_with_tap_ge(c, x0, W), first_with_tap_ge(c + 1, x0, W)
                    r = x[b, :, y0[Y], c] + float(ly[Y]) * (x[b, :, y1[Y], c] - x[b, :, y0[Y], c])
                    r1 = x[b, :, y0[Y], c1] + float(ly[Y]) * (x[b, :, y1[Y], c1] - x[b, :, y0[Y], c1])
                    d = r1 - r
                    for X in range(Xlo, Xhi):
                        assert x0[X] == c
                        l = float(lx[X])
                        z = r + l * d
                        m = z.max()
                        e = np.exp(z - m)
                        s = e.sum()
                        tc = t[b, Y, X]
                        logpt = (z[tc] - m) - np.log(s)
                        pt = np.exp(logpt)
                        focal = np.exp(gamma * (1 - pt))
                        a = float(alpha[b, Y, X])
                        wt = float(weight[tc])
                        kk = {"plain_focal": focal, "no_class_weights": a * focal, "no_EDT": wt * focal}.get(mode, wt * a * focal)
                        loss -= kk * logpt
                        cnt += a > 0
                        e[tc] -= s
                        q = kk / s
                        accA[:, c] += (1 - l) * w0 * q * e
                        accB[:, c] += (1 - l) * w1 * q * e
                        SR[:, c] += l * q * e
                # the lx parts go to the owner of the next column (the last column keeps its own: x1 == x0 there)
                recv = np.zeros((C, w))
                recv[:, 1:] = SR[:, :-1]
                recv[:, w - 1] += SR[:, w - 1]
                accA += w0 * recv
                accB += w1 * recv
            while ra < ys1:
                G[b, :, ra, :] = accA
                accA, accB = accB, np.zeros((C, w))
                ra += 1
            if ys1 < h:
                spill[b, k] = accA
            else:
                assert np.abs(accA).max() == 0.0
        for k in range(1, nst):
            G[b, :, k * strip, :] += spill[b, k - 1]
    return loss / max(cnt, 1), G / max(cnt, 1), t


def main():
    root = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    worst = 0.0
    for path in sorted(glob.glob(os.path.join(root, "focal_*.npz"))):
        g = np.load(path)
        mode, gamma = str(g["mode"]), float(g["gamma"])
        h = g["logits"].shape[2]
        ref_loss, ref_grad, _ = O.focal_closed_form(g["logits"], g["target"], g["alpha"], g["weight"], gamma, mode)
        for strip in (1, 2, 3, 4, h):
            loss, grad, t = emulate(g["logits"], g["target"], g["alpha"], g["weight"], gamma, mode, strip)
            el = abs(loss - ref_loss) / abs(ref_loss)
            eg = np.abs(grad - ref_grad).max() / np.abs(ref_grad).max()
            worst = max(worst, el, eg)
            print(f"{os.path.basename(path):22s} strip {strip:3d}: loss rel {el:.2e} grad rel {eg:.2e}")
            assert el < 1e-12 and eg < 1e-12
            assert np.array_equal(t, g["target_after"].astype(np.int64))
    print("ok, worst", worst)


if __name__ == "__main__":
    main()
