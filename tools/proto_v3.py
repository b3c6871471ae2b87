"""CPU prototype (numpy fp64/fp32) of the v3 contrast algorithm: closed-form row norms from a
centred Gram matrix, per-row economised polynomials for exp on the row's logit range, polynomial
backward.  Validates the formulas the CUDA kernels implement against the oracle's closed form."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dcl_oracle as O
from scipy.special import iv


def cheb_exp_coeffs(r, deg):
    """monomial coefficients c_n (z^n) of the degree-`deg` truncated Chebyshev series of e^z on [-r, r]"""
    if r < 1e-12:
        c = np.zeros(deg + 1); c[0] = 1.0
        if deg >= 1: c[1] = 1.0
        if deg >= 2: c[2] = 0.5
        return c
    # T_k(x) monomials
    T = [np.array([1.0]), np.array([0.0, 1.0])]
    for k in range(2, deg + 1):
        t = np.zeros(k + 1)
        t[1:] += 2 * T[k - 1]
        t[:k - 1] -= T[k - 2]
        T.append(t)
    c = np.zeros(deg + 1)
    for k in range(deg + 1):
        a = iv(k, r) * (1.0 if k == 0 else 2.0)
        for n, tn in enumerate(T[k]):
            c[n] += a * tn / r ** n
    return c


def run(N=1000, K=7, D=128, deg=2, seed=0, normalize=True, scale=1.0, T=0.07):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, K, (N,), generator=g).sort().values
    cent = torch.randn(K, D, generator=g)
    Z = 0.5 * torch.randn(N, D, generator=g) + 0.5 * cent[y]
    if normalize:
        Z = torch.nn.functional.normalize(Z, dim=1)
    Z = (Z * scale).to(torch.bfloat16).to(torch.float64)
    loss_o, dF_o, st = O.contrast_closed_form(Z, y, T, T, O.PIXEL)
    F = Z.numpy(); yn = y.numpy()
    n = N
    c = (F * F).sum(1)
    # --- closed-form norms
    ref = F[:128].mean(0)
    Fr = F - ref
    Mr = Fr.T @ Fr
    fs = Fr.sum(0)
    delta = fs / n
    Mc = Mr - n * np.outer(delta, delta)
    mu = ref + delta
    qf = np.einsum('id,de,ie->i', F, Mc, F)
    fm = F @ mu
    m = c.copy()                       # speculative max = diagonal
    S = F @ F.T
    assert np.all(S.max(1) <= m * (1 + 2 ** -8)), "speculation fails on this data"
    nrm2 = qf + n * (fm - m) ** 2
    kap = 1.0 / np.maximum(np.sqrt(nrm2), T * 1e-12)
    print("kappa rel err vs oracle (r*T):", np.abs(kap * (st['r'].numpy() * T) - 1).max())
    cmax = c.max()
    L = np.minimum(1.0, kap * (m + np.sqrt(c * cmax))) * (1 + 1e-6)
    # --- per-row polynomial in s:  E(s) = exp(kap (s - m))
    d = np.zeros((n, deg + 1))
    for i in range(n):
        r = L[i] / 2
        cz = cheb_exp_coeffs(r, deg) * np.exp(-r)
        al, be = kap[i], -kap[i] * m[i] + r
        # expand sum_n cz[n] (al s + be)^n
        for nn in range(deg + 1):
            for j in range(nn + 1):
                from math import comb
                d[i, j] += cz[nn] * comb(nn, j) * al ** j * be ** (nn - j)
    Sf = S.astype(np.float32)
    E = np.zeros_like(S)
    for j in range(deg, -1, -1):
        E = E * S + d[:, j:j + 1]
    Eex = np.exp(kap[:, None] * (S - m[:, None]))
    print(f"deg {deg}: Lmax {L.max():.4f}  poly rel err max {np.abs(E / Eex - 1).max():.2e}")
    neg = yn[:, None] != yn[None, :]
    Den = (E * neg).sum(1)
    SEs = (E * S * neg).sum(1)
    print("Den rel err vs oracle:", np.abs(Den / st['neg'].numpy() - 1).max())
    # --- rest of forward exactly (positives)
    l = kap[:, None] * (S - m[:, None])
    same = ~neg
    pos = same & ~np.eye(n, dtype=bool)
    P = pos.sum(1)
    lp = l - np.log(np.exp(l) + Den[:, None])
    rowloss = -(lp * pos).sum(1) / P
    loss = rowloss.mean()
    print("loss rel err:", abs(loss / loss_o - 1))
    # --- backward constants
    cc = 1.0 / n
    w = -cc / P
    inv = 1.0 / (np.exp(l) + Den[:, None])
    Q = w * (inv * pos).sum(1)
    El = kap * (SEs - m * Den)                      # sum_neg E l
    R = w * Den * (inv * l * pos).sum(1) - Q * El
    print("Q rel err", np.abs(Q / st['Q'].numpy() - 1).max(), " R rel err", np.abs(R / st['R'].numpy() - 1).max())
    # dS_ik = kap_i (g_ik - l_ik R_i);  negatives: g = -E Q ;  row polynomial in s:
    #   rowpoly_i(s) = kap_i ( -Q_i E_i(s) - R_i kap_i (s - m_i) )
    dr = -kap[:, None] * Q[:, None] * d
    dr[:, 1] += -kap * kap * R
    dr[:, 0] += kap * kap * R * m
    Rm = np.zeros_like(S)
    for j in range(deg, -1, -1):
        Rm = Rm * S + dr[:, j:j + 1]
    G = Rm + Rm.T                                   # negatives
    # positives / self: exact
    gpos = w[:, None] * Den[:, None] * inv
    lin = kap[:, None] * (-l * R[:, None])
    dSpos = kap[:, None] * gpos + lin
    Gs = np.where(pos, dSpos + dSpos.T, 0.0) + np.where(np.eye(n, dtype=bool), 2 * lin, 0.0)
    G = np.where(neg, G, Gs)
    Gb = torch.tensor(G).to(torch.bfloat16).to(torch.float64).numpy()
    dF = Gb @ F
    dFo = dF_o.numpy()
    print("grad err / max-abs:", np.abs(dF - dFo).max() / np.abs(dFo).max(), " (fp64 G:", np.abs(G @ F - dFo).max() / np.abs(dFo).max(), ")")


if __name__ == "__main__":
    for kw in (dict(N=1000, K=7, deg=2), dict(N=1000, K=7, deg=3), dict(N=200, K=4, deg=2), dict(N=200, K=4, deg=3), dict(N=200, K=4, deg=4),
               dict(N=64, K=3, deg=4), dict(N=2000, K=19, deg=2, normalize=False), dict(N=2000, K=19, deg=2, scale=3.0)):
        print("==", kw)
        run(**kw)
