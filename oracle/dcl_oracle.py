"""CPU oracle for the doubly contrastive loss hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the algorithm of the reference's `utils/loss.py`
(`PixelContrastLoss` :250-415 and `SupConLoss` :84-205).  It exists so that `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs can check and
time against it.  Nothing under `doubly_contrastive_semseg_b200/` may import it: the product path
is the CUDA library and fails loudly without it.

Parity status: PINNED.  The reference has no tests or golden vectors of its own,
so the pin is `tests/golden/*.npz`, produced by importing the real reference in the build
container (`tests/golden/make_golden.py`) and checked by `tests/test_oracle_golden.py`.

Three layers, each citing the reference lines it follows:
  * `MT19937` / `randperm_prefix`  - the CPU generator behind `torch.randperm` (loss.py:327,329)
  * `sample_anchors`               - label down-sampling, argmax, hard-anchor sampling
                                     (loss.py:264-337, :396-410), integer-exact, numpy
  * `pixel_contrast_closed_form` / `supcon_closed_form` - fp64 forward + analytic backward of
                                     the N x N contrast (loss.py:339-389, :161-204), chunked so it
                                     also works where the reference cannot allocate N x N
  * `PixelContrastPort` / `SupConPort` - fp32 torch-autograd restatement with the reference's
                                     op sequence, used as the timed CPU baseline ("port")
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# RNG: torch's CPU generator is a 32-bit Mersenne Twister (ATen mt19937, not vendored in the
# reference; semantics re-verified against torch 2.11 in tests/test_oracle_golden.py).
# --------------------------------------------------------------------------------------


class MT19937:
    """MT19937 seeded like `torch.manual_seed(seed)` (init_genrand on the low 32 bits)."""

    N, M = 624, 397

    def __init__(self, seed: int):
        s = np.zeros(self.N, dtype=np.uint64)
        s[0] = seed & 0xFFFFFFFF
        for i in range(1, self.N):
            s[i] = (1812433253 * (int(s[i - 1]) ^ (int(s[i - 1]) >> 30)) + i) & 0xFFFFFFFF
        self.state = s.astype(np.uint32)
        self.pos = self.N  # force a twist before the first draw

    def _twist(self):
        s = self.state
        N, M = self.N, self.M
        up, lo = np.uint32(0x80000000), np.uint32(0x7FFFFFFF)
        mag = np.uint32(0x9908B0DF)
        # three dependency-free segments of the standard in-place recurrence
        for a, b in ((0, N - M), (N - M, 2 * (N - M)), (2 * (N - M), N - 1)):
            y = (s[a:b] & up) | (s[a + 1 : b + 1] & lo)
            src = s[a + M : b + M] if b + M <= N else s[a + M - N : b + M - N]
            s[a:b] = src ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), mag, np.uint32(0))
        y = (s[N - 1] & up) | (s[0] & lo)
        s[N - 1] = s[M - 1] ^ (y >> np.uint32(1)) ^ (mag if (y & np.uint32(1)) else np.uint32(0))
        self.pos = 0

    def draw(self, n: int) -> np.ndarray:
        """Next n tempered 32-bit outputs."""
        out = np.empty(n, dtype=np.uint32)
        filled = 0
        while filled < n:
            if self.pos >= self.N:
                self._twist()
            take = min(n - filled, self.N - self.pos)
            y = self.state[self.pos : self.pos + take].copy()
            y ^= y >> np.uint32(11)
            y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
            y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
            y ^= y >> np.uint32(18)
            out[filled : filled + take] = y
            self.pos += take
            filled += take
        return out


def randperm_prefix(rng: MT19937, n: int, k: int) -> np.ndarray:
    """First k entries of `torch.randperm(n)` (CPU path, n < 2^32/20): Fisher-Yates with
    z = u32 % (n - i) for i < n-1.  Entry i is final after step i, but the generator always
    advances n-1 draws (0 draws when n <= 1).  Reference use: loss.py:327-330."""
    if n <= 1:
        return np.arange(min(n, k), dtype=np.int64)
    z = rng.draw(n - 1).astype(np.int64)
    k = min(k, n)
    moved = {}
    out = np.empty(k, dtype=np.int64)
    for i in range(k):
        if i < n - 1:
            j = i + int(z[i] % (n - i))
        else:
            j = i
        vi = moved.get(i, i)
        vj = moved.get(j, j)
        moved[i], moved[j] = vj, vi
        out[i] = vj
    return out


def torch_randperm_prefix(n: int, k: int) -> np.ndarray:
    """Same quantity drawn from torch's *global* CPU generator (what the reference consumes)."""
    return torch.randperm(n)[:k].numpy().astype(np.int64)


# --------------------------------------------------------------------------------------
# Sampler (integer-exact)
# --------------------------------------------------------------------------------------


def nearest_src_index(out_size: int, in_size: int) -> np.ndarray:
    """Legacy 'nearest' source index used by F.interpolate(mode='nearest') (loss.py:401-402):
    src = min(floor(dst * float32(in/out)), in - 1), all in float32."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def downsample_labels(labels: np.ndarray, h: int, w: int) -> np.ndarray:
    """labels [B,H,W] int -> [B,h*w] (loss.py:400-408)."""
    B, H, W = labels.shape
    iy, ix = nearest_src_index(h, H), nearest_src_index(w, W)
    return labels[:, iy][:, :, ix].reshape(B, h * w)


def argmax_first(predict: np.ndarray) -> np.ndarray:
    """predict [B,C,h,w] -> first-index argmax over C, flattened [B,h*w] (loss.py:396,409)."""
    B, C, h, w = predict.shape
    return np.argmax(predict, axis=1).reshape(B, h * w)


@dataclass
class AnchorPlan:
    """Everything the sampler decides, in the reference's anchor order (image asc, class asc)."""

    A: int = 0                      # number of (image, class) anchors == total_classes
    n_view: int = 0
    image: List[int] = field(default_factory=list)       # [A]
    cls: List[int] = field(default_factory=list)         # [A]  == y_
    num_hard: List[int] = field(default_factory=list)
    num_easy: List[int] = field(default_factory=list)
    keep_hard: List[int] = field(default_factory=list)
    keep_easy: List[int] = field(default_factory=list)
    pixels: Optional[np.ndarray] = None                  # [A, n_view] flat pixel index p = y*w+x


def split_rule(num_hard: int, num_easy: int, n_view: int) -> Tuple[int, int]:
    """(keep_hard, keep_easy), loss.py:314-325."""
    if num_hard >= n_view / 2 and num_easy >= n_view / 2:
        kh = n_view // 2
        return kh, n_view - kh
    if num_hard >= n_view / 2:
        return n_view - num_easy, num_easy
    if num_easy >= n_view / 2:
        return num_hard, n_view - num_hard
    raise Exception("this shoud be never touched! {} {} {}".format(num_hard, num_easy, n_view))


def sample_anchors(lab: np.ndarray, pred: np.ndarray, ignore_label: int, max_samples: int,
                   max_views: int,
                   randperm: Callable[[int, int], np.ndarray]) -> Optional[AnchorPlan]:
    """Hard-anchor sampling (loss.py:264-337) on down-sampled labels `lab` [B,hw] and argmax
    predictions `pred` [B,hw].  `randperm(n, k)` must return torch.randperm(n)[:k] and is called
    hard-then-easy for every anchor in reference order.  Returns None when no class qualifies
    (the reference returns (None, None), loss.py:287-288)."""
    B = lab.shape[0]
    plan = AnchorPlan()
    per_image = []
    for b in range(B):
        vals, counts = np.unique(lab[b], return_counts=True)          # ascending, loss.py:280
        keep = [int(v) for v, c in zip(vals, counts) if v != ignore_label and c > max_views]
        per_image.append(keep)
        plan.A += len(keep)
    if plan.A == 0:
        return None
    plan.n_view = min(max_samples // plan.A, max_views)               # loss.py:290-291
    rows = []
    for b in range(B):
        for c in per_image[b]:
            is_c = lab[b] == c
            hard = np.nonzero(is_c & (pred[b] != c))[0]               # raster order, loss.py:308
            easy = np.nonzero(is_c & (pred[b] == c))[0]
            kh, ke = split_rule(len(hard), len(easy), plan.n_view)
            ph = randperm(len(hard), kh)
            pe = randperm(len(easy), ke)
            rows.append(np.concatenate([hard[ph], easy[pe]]))
            plan.image.append(b)
            plan.cls.append(c)
            plan.num_hard.append(len(hard))
            plan.num_easy.append(len(easy))
            plan.keep_hard.append(kh)
            plan.keep_easy.append(ke)
    plan.pixels = np.stack(rows).astype(np.int64) if plan.n_view > 0 else np.zeros((plan.A, 0), np.int64)
    return plan


def gather_anchor_rows(feats: np.ndarray, plan: AnchorPlan) -> Tuple[np.ndarray, np.ndarray]:
    """feats [B,C,h,w] -> contrast rows F [N,C] in the reference's view-major order
    row = v*A + a (loss.py:347), and their labels y [N]."""
    B, C, h, w = feats.shape
    flat = feats.reshape(B, C, h * w)
    A, V = plan.A, plan.n_view
    F = np.empty((V * A, C), dtype=feats.dtype)
    y = np.empty(V * A, dtype=np.int64)
    for a in range(A):
        for v in range(V):
            F[v * A + a] = flat[plan.image[a], :, plan.pixels[a, v]]
            y[v * A + a] = plan.cls[a]
    return F, y


# --------------------------------------------------------------------------------------
# N x N contrast, fp64 closed form (validated against autograd of the
# reference in tests/test_oracle_golden.py)
# --------------------------------------------------------------------------------------

PIXEL, SUPCON = 0, 1


def _row_chunks(n: int, chunk: int):
    for s in range(0, n, chunk):
        yield s, min(n, s + chunk)


def contrast_closed_form(F, y, temperature=0.07, base_temperature=0.07, mode=PIXEL,
                         want_grad=True, chunk=2048, dtype=torch.float64, device=None):
    """Loss and dLoss/dF for the row-normalised contrast.

    mode=PIXEL  : loss.py:339-389   lp_ij = l_ij - log(exp(l_ij) + sum_{neg k} exp(l_ik))
    mode=SUPCON : loss.py:161-204   lp_ij = l_ij - log(sum_{k != i} exp(l_ik))
    Common part : a = F F^T / T, m_i = row max (detached), l = normalize(a - m) (L2 over the
    row, eps 1e-12), positives = same label minus the diagonal, loss = mean_i
    [-(T/T_b) * mean_{j in pos(i)} lp_ij].
    Returns (loss float, dF [N,D] or None, stats dict of per-row tensors).  `device`: where the (plain torch)
    arithmetic runs; tests pass a CUDA device for N = 65536, where the fp64 chunks take minutes on host cores.
    """
    dev = torch.device(device) if device is not None else torch.device("cpu")
    F = torch.as_tensor(F).to(device=dev, dtype=dtype)
    y = torch.as_tensor(y).to(device=dev, dtype=torch.int64)
    N = F.shape[0]
    T, Tb = float(temperature), float(base_temperature)
    c = (T / Tb) / N
    m = torch.empty(N, dtype=dtype, device=dev)
    r = torch.empty(N, dtype=dtype, device=dev)
    neg = torch.empty(N, dtype=dtype, device=dev)       # PIXEL: sum over negatives; SUPCON: sum over k != i
    P = torch.empty(N, dtype=dtype, device=dev)
    rowloss = torch.empty(N, dtype=dtype, device=dev)
    Q = torch.empty(N, dtype=dtype, device=dev)
    R = torch.empty(N, dtype=dtype, device=dev)
    ar = torch.arange(N, device=dev)
    for s, e in _row_chunks(N, chunk):
        a = (F[s:e] @ F.T) / T
        mi = a.max(dim=1).values
        u = a - mi[:, None]
        ri = torch.clamp(u.norm(dim=1), min=1e-12)
        l = u / ri[:, None]
        E = torch.exp(l)
        same = y[s:e, None] == y[None, :]
        diag = ar[s:e, None] == ar[None, :]
        pos = same & ~diag
        Pi = pos.sum(dim=1).to(dtype)
        if mode == PIXEL:
            den_mask = ~same
            negi = (E * den_mask).sum(dim=1)
            lp = l - torch.log(E + negi[:, None])
        else:
            den_mask = ~diag
            negi = (E * den_mask).sum(dim=1)
            lp = l - torch.log(negi)[:, None]
        rowloss[s:e] = -(T / Tb) * (lp * pos).sum(dim=1) / Pi
        m[s:e], r[s:e], neg[s:e], P[s:e] = mi, ri, negi, Pi
        if want_grad:
            w = (-c / Pi)[:, None] * pos                              # dL/dlp_ij
            if mode == PIXEL:
                inv = 1.0 / (E + negi[:, None])
                Qi = (w * inv).sum(dim=1)
                g = w * negi[:, None] * inv - den_mask * E * Qi[:, None]
            else:
                Qi = -c / negi                                        # so that g = w - den*E*Q
                g = w - den_mask * E * Qi[:, None]
            Q[s:e] = Qi
            R[s:e] = (g * l).sum(dim=1)
    loss = float(rowloss.mean())
    stats = dict(m=m, r=r, neg=neg, P=P, rowloss=rowloss, Q=Q, R=R)
    if not want_grad:
        return loss, None, stats
    dF = torch.zeros_like(F)
    for s, e in _row_chunks(N, chunk):
        a = (F[s:e] @ F.T) / T
        l = (a - m[s:e, None]) / r[s:e, None]
        E = torch.exp(l)
        same = y[s:e, None] == y[None, :]
        diag = ar[s:e, None] == ar[None, :]
        pos = same & ~diag
        w = (-c / P[s:e])[:, None] * pos
        if mode == PIXEL:
            g = w * neg[s:e, None] / (E + neg[s:e, None]) - (~same) * E * Q[s:e, None]
        else:
            g = w - (~diag) * E * Q[s:e, None]
        dS = (g - l * R[s:e, None]) / (r[s:e, None] * T)
        dF[s:e] += dS @ F
        dF += dS.T @ F[s:e]
    return loss, dF, stats


def pixel_contrast_closed_form(F, y, temperature=0.07, base_temperature=0.07, **kw):
    return contrast_closed_form(F, y, temperature, base_temperature, PIXEL, **kw)


def supcon_closed_form(Z, y, temperature=0.07, base_temperature=0.07, **kw):
    return contrast_closed_form(Z, y, temperature, base_temperature, SUPCON, **kw)


# --------------------------------------------------------------------------------------
# fp32 torch-autograd restatement with the reference's op sequence ("port"): the timed CPU
# baseline.  Same module surface as the reference.
# --------------------------------------------------------------------------------------


class PixelContrastPort(torch.nn.Module):
    """CPU restatement of PixelContrastLoss (loss.py:250-415): same attributes, same forward
    signature, same consumption of the global CPU RNG."""

    def __init__(self, device=None):
        super().__init__()
        self.device = device
        self.temperature = 0.07
        self.base_temperature = 0.07
        self.ignore_label = 255
        self.max_samples = 1024
        self.max_views = 2
        self.loss_weight = 1
        self.contrast_mode = "all"
        self.last_plan: Optional[AnchorPlan] = None

    def forward(self, feats, labels=None, predict=None):
        B, C, h, w = feats.shape
        pred = predict.argmax(dim=1).reshape(B, -1)                                  # :396
        lab = torch.nn.functional.interpolate(labels.unsqueeze(1).float(), (h, w),
                                              mode="nearest").squeeze(1).long()      # :400-403
        assert lab.shape[-1] == feats.shape[-1], "{} {}".format(lab.shape, feats.shape)
        lab = lab.reshape(B, -1)
        X = feats.permute(0, 2, 3, 1).contiguous().view(B, -1, C).float()            # :409-410
        plan = sample_anchors(lab.numpy(), pred.numpy(), self.ignore_label, self.max_samples,
                              self.max_views, torch_randperm_prefix)
        self.last_plan = plan
        if plan is None:
            raise RuntimeError("no class qualifies for anchor sampling (reference returns None)")
        rows = []
        for a in range(plan.A):                                                      # :297-335
            idx = torch.from_numpy(plan.pixels[a])
            rows.append(X[plan.image[a], idx, :])
        X_ = torch.stack(rows)                                                       # [A,V,C]
        y_ = torch.tensor(plan.cls, dtype=torch.float32)
        return self._contrastive(X_, y_)

    def _contrastive(self, feats_, labels_):                                         # :339-389
        A, V = feats_.shape[0], feats_.shape[1]
        T, Tb = self.temperature, self.base_temperature
        y = labels_.reshape(-1, 1)
        same = (y == y.T).float().repeat(V, V)
        Fm = torch.cat(torch.unbind(feats_, dim=1), dim=0)
        a = (Fm @ Fm.T) / T
        l = torch.nn.functional.normalize(a - a.max(dim=1, keepdim=True).values.detach())
        not_self = 1.0 - torch.eye(A * V)
        pos = same * not_self
        E = torch.exp(l)
        negsum = (E * (1.0 - same)).sum(dim=1, keepdim=True)
        lp = l - torch.log(E + negsum)
        per_row = -(T / Tb) * (pos * lp).sum(dim=1) / pos.sum(dim=1)
        return per_row.mean()


class SupConPort(torch.nn.Module):
    """CPU restatement of SupConLoss (loss.py:84-205), `opts.deeplab` False => dim_in 128."""

    def __init__(self, temperature=0.07, contrast_mode="all", base_temperature=0.07, weight=None,
                 device=None, opts=None):
        super().__init__()
        self.temperature = temperature
        self.base_temperature = base_temperature
        self.device = device
        self.weight = weight
        self.opts = opts
        dim_in = 2048 if (opts is not None and getattr(opts, "deeplab", False)) else 128
        self.avgpool = torch.nn.AdaptiveAvgPool2d((1, 1))
        self.projection = torch.nn.Sequential(torch.nn.Linear(dim_in, dim_in),
                                              torch.nn.ReLU(inplace=True),
                                              torch.nn.Linear(dim_in, 128))
        self.contrast_mode = "all"

    def forward(self, features, class_labels=None, mask=None):
        x = torch.flatten(self.avgpool(features), 1)                                  # :115-116
        bsz = x.shape[0] // 2
        x = torch.stack([x[:bsz], x[bsz:]], dim=1)                                    # :117-119
        z = self.projection(x)                                                        # :120
        if z.dim() < 3:
            raise ValueError("`features` needs to be [bsz, n_views, ...],"
                             "at least 3 dimensions are required")
        if class_labels is not None and mask is not None:
            raise ValueError("Cannot define both `labels` and `mask`")
        if class_labels is None and mask is None:
            mask = torch.eye(bsz)
        elif class_labels is not None:
            lab = class_labels.contiguous().view(-1, 1)
            if lab.shape[0] != bsz:
                raise ValueError("Num of labels does not match num of features")
            mask = (lab == lab.T).float()
        else:
            mask = mask.float()
        T, Tb = self.temperature, self.base_temperature
        V = z.shape[1]
        Z = torch.cat(torch.unbind(z, dim=1), dim=0)                                  # :161
        a = (Z @ Z.T) / T
        l = a - a.max(dim=1, keepdim=True).values.detach()
        not_self = 1.0 - torch.eye(bsz * V)
        pos = mask.repeat(V, V) * not_self
        l = torch.nn.functional.normalize(l)                                          # :194
        lp = l - torch.log((torch.exp(l) * not_self).sum(dim=1, keepdim=True))
        per_row = -(T / Tb) * (pos * lp).sum(dim=1) / pos.sum(dim=1)
        return per_row.view(V, bsz).mean()


# --------------------------------------------------------------------------------------
# Segmentation-loss neighbour of the hot path: BoundaryAwareFocalLoss, loss.py:27-80
# --------------------------------------------------------------------------------------


def bilinear_source(out_size: int, in_size: int):
    """F.interpolate(mode='bilinear', align_corners=False) source taps of every output index (loss.py:5):
    src = max((dst + 0.5) * (in/out) - 0.5, 0) in float32; i0 = floor(src), i1 = min(i0 + 1, in - 1), lambda = src - i0."""
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.maximum((np.arange(out_size, dtype=np.float32) + np.float32(0.5)) * scale - np.float32(0.5), np.float32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    lam = (src - i0.astype(np.float32)).astype(np.float32)
    return i0, i1, lam


class BoundaryFocalPort(torch.nn.Module):
    """CPU restatement of BoundaryAwareFocalLoss (loss.py:27-80): bilinear up-sampling of the logits to the label
    size when the sizes differ, log-softmax, gather at the target, class weight x EDT weight x detached focal factor
    exp(gamma (1 - p_t)), sum / #(EDT weight > 0).  Like the reference it rewrites `target` in place (ignore -> 0)."""

    def __init__(self, gamma=0, num_classes=19, ignore_id=19, print_each=20, weight=None, device=None, opts=None):
        super().__init__()
        self.num_classes, self.ignore_id, self.print_each = num_classes, ignore_id, print_each
        self.step_counter = 0
        self.gamma, self.weight, self.device, self.opts = gamma, weight, device, opts

    def forward(self, input, target, batch, **kwargs):
        if input.shape[-2:] != target.shape[-2:]:
            input = torch.nn.functional.interpolate(input, target.shape[-2:], mode="bilinear", align_corners=False)
        target[target == self.ignore_id] = 0                                         # :43
        alpha = batch["label_distance_weight"]
        N = (alpha.data > 0.).sum()                                                   # :45
        if N.le(0):
            return torch.zeros(size=(0,), requires_grad=True).sum()
        x = input.view(input.size(0), input.size(1), -1).transpose(1, 2).contiguous().view(-1, input.size(1))
        t = target.view(-1, 1)
        alphas = alpha.view(-1)
        w = self.weight[t].view(-1)                                                   # :54 (TypeError when weight is None)
        logpt = torch.nn.functional.log_softmax(x.to(torch.float32), dim=-1).gather(1, t).view(-1)
        pt = logpt.detach().exp()
        focal = torch.exp(self.gamma * (1 - pt))
        crit = getattr(self.opts, "criterion", None)
        if crit == "plain_focal":
            loss = -1 * focal * logpt
        elif getattr(self.opts, "no_class_weights", False):
            loss = -1 * alphas * focal * logpt
        elif getattr(self.opts, "no_EDT", False):
            loss = -1 * w * focal * logpt
        else:
            loss = -1 * w * alphas * focal * logpt
        self.step_counter += 1
        return loss.sum() / N


def focal_closed_form(logits, target, alpha, weight, gamma, mode="full", ignore_id=255):
    """fp64 forward + analytic backward of BoundaryAwareFocalLoss (loss.py:27-80) on pre-upsample logits [B,C,h,w]
    (h == H, w == W allowed): what the fused CUDA kernel computes.  With the up-sampled logits z = U x (U = the
    separable bilinear operator of `bilinear_source`), p = softmax(z), k = coefficient of the mode
    (class weight x EDT weight x exp(gamma (1 - p_t)), the focal factor detached):
        loss = - sum_i k_i log p_{i,t_i} / N,     N = #(alpha > 0)
        d loss / d z_{i,c} = - k_i (delta_{c,t_i} - p_{i,c}) / N,     d loss / d x = U^T (d loss / d z).
    Returns (loss, dlogits [B,C,h,w], target with ignore -> 0)."""
    x = np.asarray(logits, dtype=np.float64)
    t = np.asarray(target).astype(np.int64).copy()
    a = np.asarray(alpha, dtype=np.float64)
    B, C, h, w = x.shape
    H, W = t.shape[1:]
    t[t == ignore_id] = 0
    N = int((a > 0).sum())
    if N <= 0:
        return 0.0, np.zeros_like(x), t
    y0, y1, ly = bilinear_source(H, h)
    x0, x1, lx = bilinear_source(W, w)
    ly = ly.astype(np.float64)[None, None, :, None]
    lx = lx.astype(np.float64)[None, None, None, :]
    rows = (1 - ly) * x[:, :, y0, :] + ly * x[:, :, y1, :]                     # [B,C,H,w]
    z = (1 - lx) * rows[:, :, :, x0] + lx * rows[:, :, :, x1]                   # [B,C,H,W]
    m = z.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(z - m).sum(axis=1, keepdims=True))
    logp = z - lse
    p = np.exp(logp)
    bi, yi, xi = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")
    logpt = logp[bi, t, yi, xi]
    pt = np.exp(logpt)
    wt = np.asarray(weight, dtype=np.float64)[t]
    focal = np.exp(gamma * (1 - pt))
    k = {"plain_focal": focal, "no_class_weights": a * focal, "no_EDT": wt * focal}.get(mode, wt * a * focal)
    loss = float(-(k * logpt).sum() / N)
    onehot = np.zeros_like(z)
    onehot[bi, t, yi, xi] = 1.0
    dz = -(k[:, None] * (onehot - p)) / N
    # U^T: scatter columns, then rows
    drows = np.zeros((B, C, H, w))
    np.add.at(drows, (slice(None), slice(None), slice(None), x0), dz * (1 - lx))
    np.add.at(drows, (slice(None), slice(None), slice(None), x1), dz * lx)
    dx = np.zeros_like(x)
    np.add.at(dx, (slice(None), slice(None), y0, slice(None)), drows * (1 - ly))
    np.add.at(dx, (slice(None), slice(None), y1, slice(None)), drows * ly)
    return loss, dx, t
