"""Stage-by-stage GPU diagnostic (run on the B200 box): compares every intermediate the CUDA
path exposes with the CPU oracle.  Not a test: prints numbers, never asserts."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from doubly_contrastive_semseg_b200 import loss as L          # noqa: E402
from oracle import dcl_oracle as O                           # noqa: E402

LOG2E = 1.4426950408889634


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def check_rows(n, K, mode, seed=0, T=0.07, offset=0.0, sort=True):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, K, (n,), generator=g)
    if mode == L.MODE_SUPCON:
        half = n // 2
        y = torch.cat([y[:half], y[:half]])[:n] if n % 2 == 0 else y
    elif sort:
        y = y.sort().values
    cent = torch.randn(K, 128, generator=g)
    Z = 0.5 * torch.randn(n, 128, generator=g) + 0.5 * cent[y] + offset
    Zb = bf16_round(Z)
    t0 = time.time()
    loss_o, dF_o, st = O.contrast_closed_form(Zb, y, T, T, mode)
    t_or = time.time() - t0
    Zd = Z.cuda().requires_grad_(True)
    yd = y.cuda()
    n_pad = (n + 127) // 128 * 128
    # stage 1: pack
    tiles, sqnorm = L.pack_rows(Zd.detach().contiguous(), n_pad)
    sq_o = (Zb.double() ** 2).sum(1)
    e_sq = rel(sqnorm[:n].cpu(), sq_o)
    # stage 2: forward
    ypad = torch.full((n_pad,), -1, dtype=torch.int32, device="cuda")
    ypad[:n] = yd.int()
    nJ = n_pad // 128
    colA, colB, rowloss, loss_sum = L.contrast_forward(tiles, ypad, sqnorm, nJ, 0, nJ, n, mode, T, T)
    torch.cuda.synchronize()
    cA, cB = colA[:n].cpu().double(), colB[:n].cpu().double()
    kappa_o = 1.0 / (T * st["r"])
    a_o = kappa_o * LOG2E
    b_o = -(st["m"] * T) * a_o
    e_a, e_b = rel(cA[:, 0], a_o), rel(cA[:, 1], b_o)
    e_den = rel(cB[:, 1], st["neg"])
    e_rl = rel(rowloss[:n].cpu(), st["rowloss"])
    p_o = -kappa_o * st["R"] * np.log(2.0)
    q_o = -kappa_o * st["Q"]
    e_p, e_q = rel(cA[:, 2], p_o), rel(cA[:, 3], q_o)
    loss_d = float(loss_sum[0].item()) / n
    # stage 3: backward
    dF = L.contrast_backward(tiles, ypad, colA, colB, nJ, 0, nJ, mode)
    torch.cuda.synchronize()
    e_g = rel(dF[:n].cpu(), dF_o)
    # vs un-rounded fp32 inputs (the tolerance the north star states)
    loss_f, dF_f, _ = O.contrast_closed_form(Z, y, T, T, mode)
    print(f"rows n={n:6d} K={K:3d} mode={mode} off={offset}: sq {e_sq:.1e} a {e_a:.1e} b {e_b:.1e} den {e_den:.1e} "
          f"rowloss {e_rl:.1e} p {e_p:.1e} q {e_q:.1e} | loss dev {loss_d:.7f} oracle {loss_o:.7f} "
          f"rel {abs(loss_d - loss_o) / abs(loss_o):.1e} | grad(bf16-in) {e_g:.1e} | vs fp32-in: loss "
          f"{abs(loss_d - loss_f) / abs(loss_f):.1e} grad {rel(dF[:n].cpu(), dF_f):.1e}  (oracle {t_or:.1f}s)",
          flush=True)


def main():
    print(torch.cuda.get_device_name(0), flush=True)
    cases = [(40, 2, 0), (130, 4, 0), (192, 5, 0), (128, 3, 0), (256, 4, 0), (1000, 7, 0), (2048, 16, 0),
             (32, 4, 1), (6, 3, 1), (300, 8, 1), (4096, 16, 0)]
    for n, K, mode in cases:
        try:
            check_rows(n, K, mode)
        except Exception:
            traceback.print_exc()
            print("FAILED case", n, K, mode, flush=True)
            return 1
    try:
        check_rows(1000, 7, 0, offset=2.0)
        check_rows(777, 5, 0, sort=False)
    except Exception:
        traceback.print_exc()
    return 0


if __name__ == "__main__":
    sys.exit(main())
