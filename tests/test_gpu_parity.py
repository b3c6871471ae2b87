"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through the C ABI
(ctypes) behind the drop-in modules, against the CPU oracle and the golden vectors generated from
the real reference.  Tolerances are BASELINE.json's: loss <= 1e-3 relative, gradients <= 1e-2
relative to the max-abs gradient; sampled indices bit-exact."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from oracle import dcl_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_RTOL = 1e-3
GRAD_RTOL = 1e-2


def _gold(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def _cases(prefix):
    return sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


def _relmax(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


@pytest.fixture(scope="module")
def dcl():
    import doubly_contrastive_semseg_b200 as pkg
    from doubly_contrastive_semseg_b200 import _lib
    _lib.load()                                   # a missing .so is a hard failure on the GPU box
    assert _lib.load().dcl_check_device() == 0, _lib.load().dcl_last_error()
    return pkg


# ------------------------------------------------------------------ N x N contrast via the ABI
@pytest.mark.parametrize("n,K,mode,offset,sort", [
    (40, 2, 0, 0.0, True), (130, 4, 0, 0.0, True), (128, 3, 0, 0.0, True), (257, 4, 0, 0.0, True),
    (1000, 7, 0, 0.0, True), (1000, 7, 0, 2.0, True), (777, 5, 0, 0.0, False), (2048, 16, 0, 0.0, True),
    (5000, 19, 0, 0.0, True), (6, 3, 1, 0.0, False), (32, 4, 1, 0.0, False), (300, 8, 1, 0.0, False),
    (64, 1, 1, 0.0, False)])
def test_contrast_rows_vs_oracle(dcl, n, K, mode, offset, sort):
    g = torch.Generator().manual_seed(n * 7 + K)
    y = torch.randint(0, K, (n,), generator=g)
    if mode == 1:
        y = torch.cat([y[: n // 2], y[: n // 2]])
    elif sort:
        y = y.sort().values
    cent = torch.randn(K, 128, generator=g)
    Z = 0.5 * torch.randn(n, 128, generator=g) + 0.5 * cent[y] + offset
    Zd = Z.cuda().requires_grad_(True)
    loss = dcl.contrast_rows(Zd, y.cuda(), mode)
    loss.backward()
    loss_o, dZ_o, _ = O.contrast_closed_form(Z, y, 0.07, 0.07, mode)
    assert abs(loss.item() - loss_o) <= LOSS_RTOL * abs(loss_o)
    assert _relmax(Zd.grad.cpu(), dZ_o) <= GRAD_RTOL
    # tighter: same bf16-rounded inputs the tensor cores see -> only kernel arithmetic differs
    Zb = Z.to(torch.bfloat16).float()
    loss_b, dZ_b, _ = O.contrast_closed_form(Zb, y, 0.07, 0.07, mode)
    assert abs(loss.item() - loss_b) <= 2e-5 * abs(loss_b)
    assert _relmax(Zd.grad.cpu(), dZ_b) <= 5e-3


@pytest.mark.parametrize("case", _cases("contrast"))
def test_contrast_rows_vs_reference_golden(dcl, case):
    g = _gold(case)
    X, y = g["X"], g["y"]
    A, V, D = X.shape
    F = np.concatenate([X[:, v] for v in range(V)], axis=0)
    yy = np.tile(y, V)
    Zd = torch.from_numpy(F).cuda().requires_grad_(True)
    loss = dcl.contrast_rows(Zd, torch.from_numpy(yy).cuda(), 0)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    dX = Zd.grad.cpu().numpy().reshape(V, A, D).transpose(1, 0, 2)
    assert _relmax(dX, g["dX"]) <= GRAD_RTOL


@pytest.mark.parametrize("case", ["contrastL_n2048.npz", "contrastL_n8192.npz"])
def test_contrast_rows_vs_reference_golden_large(dcl, case):
    """The reference's own `_contrastive` (utils/loss.py:339-389, fp32 CPU autograd) at N = 2048 / 8192: the sizes
    where the CUDA path runs its degree-1 polynomial exp and the positive-pair series (the small fixtures only reach
    the exact fallback), so the fast path is pinned to the reference and not only to the oracle."""
    from test_oracle_golden import _large_contrast_case
    F, yy, loss_ref, dX_ref, (A, V) = _large_contrast_case(case)
    Zd = F.cuda().requires_grad_(True)
    loss = dcl.contrast_rows(Zd, yy.cuda(), 0)
    loss.backward()
    assert abs(loss.item() - loss_ref) <= LOSS_RTOL * abs(loss_ref)
    dX = Zd.grad.cpu().numpy().reshape(V, A, 128).transpose(1, 0, 2)
    assert _relmax(dX, dX_ref) <= GRAD_RTOL
    # the rows in class-sorted order (what the sampler hands the kernels): same numbers, permuted
    order = torch.argsort(yy, stable=True)
    Zs = F[order].cuda().requires_grad_(True)
    loss_s = dcl.contrast_rows(Zs, yy[order].cuda(), 0)
    loss_s.backward()
    assert abs(loss_s.item() - loss_ref) <= LOSS_RTOL * abs(loss_ref)
    back = torch.empty_like(Zs.grad)
    back[order.cuda()] = Zs.grad
    assert _relmax(back.cpu().numpy().reshape(V, A, 128).transpose(1, 0, 2), dX_ref) <= GRAD_RTOL


def test_contrast_n65536_vs_fp64_closed_form(dcl):
    """cfg4's size (65536 anchors, the size every multi-GPU number is quoted on).  The reference cannot allocate
    N x N there; the chunked fp64 closed form - pinned to the reference at N <= 8192 by
    tests/test_oracle_golden.py - runs on the GPU in plain torch as the checker."""
    n, K = 65536, 16
    g = torch.Generator().manual_seed(65536)
    y = torch.randint(0, K, (n,), generator=g).sort().values
    cent = 0.5 * torch.randn(19, 128, generator=g)
    Z = 0.5 * torch.randn(n, 128, generator=g) + cent[y]
    Zd = Z.cuda().requires_grad_(True)
    loss = dcl.contrast_rows(Zd, y.cuda(), 0)
    loss.backward()
    torch.cuda.synchronize()
    loss_o, dZ_o, _ = O.contrast_closed_form(Z, y, 0.07, 0.07, 0, chunk=2048, device="cuda")
    assert abs(loss.item() - loss_o) <= LOSS_RTOL * abs(loss_o)
    err = float((Zd.grad.double() - dZ_o).abs().max() / dZ_o.abs().max())
    assert err <= GRAD_RTOL, err
    # and against the same bf16-rounded rows the tensor cores see: only kernel arithmetic differs
    loss_b, dZ_b, _ = O.contrast_closed_form(Z.to(torch.bfloat16).float(), y, 0.07, 0.07, 0, chunk=2048, device="cuda")
    assert abs(loss.item() - loss_b) <= 2e-5 * abs(loss_b)
    assert float((Zd.grad.double() - dZ_b).abs().max() / dZ_b.abs().max()) <= 5e-3


def test_single_view_rows_give_nan_like_reference(dcl):
    """n_view == 1: no row has a positive pair, the reference divides by zero (loss.py:383) and returns NaN."""
    g = torch.Generator().manual_seed(3)
    Z = torch.randn(12, 128, generator=g)
    y = torch.arange(12)
    loss = dcl.contrast_rows(Z.cuda(), y.cuda(), 0)
    assert torch.isnan(loss).item()
    port = O.PixelContrastPort()
    ref = port._contrastive(Z[:, None, :], y.float())
    assert torch.isnan(ref).item()


def test_temperature_and_upstream_gradient(dcl):
    g = torch.Generator().manual_seed(5)
    y = torch.randint(0, 4, (200,), generator=g).sort().values
    Z = torch.randn(200, 128, generator=g)
    Zd = Z.cuda().requires_grad_(True)
    loss = dcl.contrast_rows(Zd, y.cuda(), 0, temperature=0.2, base_temperature=0.1)
    (3.0 * loss).backward()
    loss_o, dZ_o, _ = O.contrast_closed_form(Z, y, 0.2, 0.1, 0)
    assert abs(loss.item() - loss_o) <= LOSS_RTOL * abs(loss_o)
    assert _relmax(Zd.grad.cpu(), 3.0 * dZ_o) <= GRAD_RTOL


# ------------------------------------------------------------------ sampler: bit-exact indices
def _device_pixels_in_reference_order(crit, plan_A, n_view, hw):
    lay, pix = crit.last_layout, crit.last_pix.cpu().numpy()
    out = np.full((plan_A, n_view), -1, dtype=np.int64)
    for n in range(lay.n):
        v, a = divmod(int(lay.ref_row[n]), plan_A)
        out[a, v] = pix[n] % hw
    return out


@pytest.mark.parametrize("case", _cases("pixel"))
def test_sampler_bit_exact_vs_reference_golden(dcl, case):
    g = _gold(case)
    B, C, h, w = g["feats"].shape
    crit = dcl.PixelContrastLoss(device="cuda")
    crit.max_samples, crit.max_views = int(g["max_samples"]), int(g["max_views"])
    torch.manual_seed(int(g["call_seed"]))
    out = crit.sample(torch.from_numpy(g["feats"]).cuda(),
                      torch.from_numpy(g["labels"].astype(np.int64)).cuda(),
                      torch.from_numpy(g["predict"]).cuda())
    assert out is not None
    A, V = g["pixels"].shape
    assert crit.last_plan.A == A and crit.last_plan.n_view == V
    assert np.array_equal(crit.last_plan.cls, g["y"])
    got = _device_pixels_in_reference_order(crit, A, V, h * w)
    assert np.array_equal(got, g["pixels"])


@pytest.mark.parametrize("B,H,W,h,w,K,mv,ms", [
    (3, 100, 180, 25, 45, 4, 7, 1024),       # hw = 1125: not a multiple of 4, ragged scale
    (2, 512, 1024, 128, 256, 19, 2, 1024),   # cfg1 with the reference defaults
    (2, 512, 1024, 128, 256, 19, 32, 1024),  # cfg1, N = 988
    (2, 96, 96, 96, 96, 3, 5, 1024),         # scale 1
    (1, 64, 64, 8, 8, 2, 3, 1024),           # single chunk, tiny
    (5, 37, 53, 19, 31, 6, 4, 50),           # odd everything, sample-capped
])
def test_sampler_bit_exact_vs_oracle(dcl, B, H, W, h, w, K, mv, ms):
    g = torch.Generator().manual_seed(H * 31 + W)
    bh = max(1, H // 6)
    coarse = torch.randint(0, K, (B, (H + bh - 1) // bh, (W + bh - 1) // bh), generator=g)
    labels = coarse.repeat_interleave(bh, 1).repeat_interleave(bh, 2)[:, :H, :W].contiguous()
    labels[torch.rand(B, H, W, generator=g) < 0.05] = 255
    predict = torch.randn(B, 19, h, w, generator=g)
    lab = O.downsample_labels(labels.numpy(), h, w)
    boost = (torch.rand(B, h * w, generator=g) < 0.6) & torch.from_numpy(lab != 255)
    predict.view(B, 19, -1).scatter_add_(1, torch.from_numpy(lab).clamp(max=18)[:, None], 4.0 * boost[:, None].float())
    feats = torch.randn(B, 128, h, w, generator=g)
    pred = O.argmax_first(predict.numpy())
    torch.manual_seed(99)
    plan = O.sample_anchors(lab, pred, 255, ms, mv, O.torch_randperm_prefix)
    crit = dcl.PixelContrastLoss(device="cuda")
    crit.max_samples, crit.max_views = ms, mv
    torch.manual_seed(99)
    out = crit.sample(feats.cuda(), labels.long().cuda(), predict.cuda())
    assert (plan is None) == (out is None)
    if plan is None:
        return
    assert crit.last_plan.A == plan.A and crit.last_plan.n_view == plan.n_view
    got = _device_pixels_in_reference_order(crit, plan.A, plan.n_view, h * w)
    assert np.array_equal(got, plan.pixels)
    assert np.array_equal(crit.last_plan.num_hard, np.array(plan.num_hard))
    assert np.array_equal(crit.last_plan.num_easy, np.array(plan.num_easy))


@pytest.mark.parametrize("B,H,W,h,w,K,mv,ms", [
    (3, 100, 180, 25, 45, 4, 7, 1024),
    (2, 512, 1024, 128, 256, 19, 32, 1024),
    (5, 37, 53, 19, 31, 6, 4, 50),
    (4, 256, 512, 64, 128, 16, 32, 2048),
])
def test_fused_step_equals_stage_by_stage(dcl, B, H, W, h, w, K, mv, ms):
    """PixelContrastLoss through dcl_pixel_begin/fwd/bwd (one C call per direction) == the same module walking the
    stages from Python: identical sample, bit-identical loss and gradient, identical generator state afterwards."""
    from doubly_contrastive_semseg_b200 import loss as L
    g = torch.Generator().manual_seed(H * 7 + W)
    bh = max(1, H // 6)
    coarse = torch.randint(0, K, (B, (H + bh - 1) // bh, (W + bh - 1) // bh), generator=g)
    labels = coarse.repeat_interleave(bh, 1).repeat_interleave(bh, 2)[:, :H, :W].contiguous().long().cuda()
    predict = torch.randn(B, 19, h, w, generator=g).cuda()
    feats = torch.randn(B, 128, h, w, generator=g).cuda()
    out = {}
    saved = L._FUSED_STEP
    try:
        for fused in (True, False):
            L._FUSED_STEP = fused
            crit = dcl.PixelContrastLoss(device="cuda")
            crit.max_samples, crit.max_views = ms, mv
            x = feats.clone().requires_grad_(True)
            torch.manual_seed(5)
            res = []
            for _ in range(5):                       # consecutive steps: from the third on the look-ahead stream runs and
                                                     # the fused step replays the permutations on the GPU (k_plan)
                x.grad = None
                loss = crit(x, labels=labels, predict=predict)
                (loss * 1.5).backward()
                res.append((loss.detach().clone(), x.grad.clone(), crit.last_pix.clone(), crit.last_plan.ranks.copy(),
                            crit.last_layout.n))
            out[fused] = (res, torch.get_rng_state().clone())
    finally:
        L._FUSED_STEP = saved
    assert torch.equal(out[True][1], out[False][1])
    for a, b in zip(out[True][0], out[False][0]):
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
        assert np.array_equal(a[3], b[3]) and a[4] == b[4]


def test_shard_pack_unpack_roundtrip(dcl):
    """The sharded step's one pre-backward message (dcl_shard_pack -> all-gather -> dcl_shard_unpack), exercised on
    one GPU by playing every rank in turn: constants land in place, the loss is the rank-ordered sum / n_global."""
    from doubly_contrastive_semseg_b200 import _lib
    from doubly_contrastive_semseg_b200.loss import _p, _stream
    world, n_pad, n_global = 3, 256, 700
    g = torch.Generator(device="cuda").manual_seed(4)
    colA = torch.randn(world * n_pad, 4, generator=g, device="cuda")
    colB = torch.randn(world * n_pad, 4, generator=g, device="cuda")
    parts = torch.randn(world, generator=g, device="cuda")
    recv = torch.empty(world, 2 * n_pad * 4 + 4, device="cuda")
    for r in range(world):
        loss2 = torch.stack([parts[r], parts[r] * 0]).contiguous()
        _lib.call("dcl_shard_pack", _p(colA), _p(colB), _p(loss2), r, n_pad, _p(recv[r]), _stream())
    outA, outB = torch.zeros_like(colA), torch.zeros_like(colB)
    loss = torch.empty(1, device="cuda")
    _lib.call("dcl_shard_unpack", _p(recv), world, n_pad, _p(outA), _p(outB), n_global, _p(loss), _stream())
    torch.cuda.synchronize()
    assert torch.equal(outA, colA) and torch.equal(outB, colB)
    want = (parts[0] + parts[1] + parts[2]) / n_global
    assert abs(float(loss) - float(want)) <= 1e-6 * abs(float(want)) + 1e-9


# ------------------------------------------------------------------ whole modules vs golden
@pytest.mark.parametrize("case", _cases("pixel"))
def test_pixel_module_vs_reference_golden(dcl, case):
    g = _gold(case)
    crit = dcl.PixelContrastLoss(device="cuda")
    crit.max_samples, crit.max_views = int(g["max_samples"]), int(g["max_views"])
    x = torch.from_numpy(g["feats"]).cuda().requires_grad_(True)
    torch.manual_seed(int(g["call_seed"]))
    loss = crit(x, labels=torch.from_numpy(g["labels"].astype(np.int64)).cuda(),
                predict=torch.from_numpy(g["predict"]).cuda())
    loss.backward()
    assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.is_cuda
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert x.grad.shape == x.shape
    assert _relmax(x.grad.cpu(), g["dfeats"]) <= GRAD_RTOL
    # the gradient is non-zero exactly at the sampled pixels
    nz = (x.grad.abs().sum(1) != 0).sum().item()
    assert nz == g["pixels"].size


@pytest.mark.parametrize("case", _cases("supcon"))
def test_supcon_module_vs_reference_golden(dcl, case):
    g = _gold(case)
    crit = dcl.SupConLoss(temperature=0.07, contrast_mode="all", base_temperature=0.07, weight=None,
                          device="cuda", opts=types.SimpleNamespace(deeplab=False))
    sd = {k: torch.from_numpy(g[k.replace(".", "_")]) for k in crit.projection.state_dict()}
    crit.projection.load_state_dict(sd)
    x = torch.from_numpy(g["feats"]).cuda().requires_grad_(True)
    labels = torch.from_numpy(g["weather"]).cuda() if bool(g["use_labels"]) else None
    loss = crit(x, class_labels=labels, mask=None)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= LOSS_RTOL * abs(float(g["loss"]))
    assert _relmax(x.grad.cpu(), g["dfeats"]) <= GRAD_RTOL
    # projection-parameter gradients are sums of dZ over the batch computed by PyTorch autograd
    # downstream of our dZ; the cancellation in those sums amplifies the bf16 rounding of the
    # tensor-core operands, so they get their own (looser, stated) bound.
    for k, p in crit.projection.named_parameters():
        assert _relmax(p.grad.cpu(), g["g_" + k.replace(".", "_")]) <= 3e-2


def _doubly_inputs(B, H, W, h, w, K, seed):
    g = torch.Generator().manual_seed(seed)
    bh = max(1, H // 6)
    coarse = torch.randint(0, K, (B, (H + bh - 1) // bh, (W + bh - 1) // bh), generator=g)
    labels = coarse.repeat_interleave(bh, 1).repeat_interleave(bh, 2)[:, :H, :W].contiguous().long()
    labels[torch.rand(B, H, W, generator=g) < 0.05] = 255
    predict = torch.randn(B, 19, h, w, generator=g)
    feats = torch.randn(2 * B, 128, h, w, generator=g) + 0.3 * torch.randn(2 * B, 128, 1, 1, generator=g)
    weather = torch.randint(0, 4, (B, 1), generator=g)
    return feats, labels, predict, weather


@pytest.mark.parametrize("B,H,W,h,w,K,mv,ms", [(2, 64, 128, 16, 32, 5, 8, 1024), (3, 48, 80, 12, 20, 4, 5, 64)])
def test_doubly_combination_vs_ports(dcl, B, H, W, h, w, K, mv, ms):
    """The `supcon_pixelcontrast_focal` combination of trainer.py:143-158: SupConLoss on fine_feat [2B,...],
    PixelContrastLoss on fine_feat[:B], total = (supcon + pixel) / batch_size, ONE backward into the shared tensor.
    (a) the two drop-in modules called the trainer's way, (b) DoublyContrastiveLoss (fused touch points) and (c) the
    CPU ports must agree on both losses and on the summed gradient."""
    feats, labels, predict, weather = _doubly_inputs(B, H, W, h, w, K, seed=B * 100 + h)
    opts = types.SimpleNamespace(deeplab=False)
    torch.manual_seed(17)
    sup = dcl.SupConLoss(device="cuda", opts=opts)
    pix = dcl.PixelContrastLoss(device="cuda")
    pix.max_samples, pix.max_views = ms, mv
    lab_d, pred_d, w_d = labels.cuda(), predict.cuda(), weather.cuda()
    # (a) trainer's way
    xa = feats.cuda().requires_grad_(True)
    torch.manual_seed(5)
    la_s = sup(xa, class_labels=w_d)
    la_p = pix(xa[:B], labels=lab_d, predict=pred_d)
    ((la_s + la_p) / B).backward()
    ga = xa.grad.clone()
    wa = [p.grad.clone() for p in sup.projection.parameters()]
    for p in sup.projection.parameters():
        p.grad = None
    # (b) fused
    both = dcl.DoublyContrastiveLoss(pix, sup)
    xb = feats.cuda().requires_grad_(True)
    torch.manual_seed(5)
    lb_s, lb_p = both(xb, labels=lab_d, predict=pred_d, class_labels=w_d)
    ((lb_s + lb_p) / B).backward()
    assert lb_s.item() == la_s.item() and lb_p.item() == la_p.item()
    assert torch.allclose(xb.grad, ga, rtol=1e-5, atol=1e-7 * float(ga.abs().max()))
    for p, q in zip(sup.projection.parameters(), wa):
        assert torch.allclose(p.grad, q, rtol=1e-5, atol=1e-9)
    # (c) ports
    sup_o = O.SupConPort(opts=opts)
    sup_o.projection.load_state_dict({k: v.detach().cpu() for k, v in sup.projection.state_dict().items()})
    pix_o = O.PixelContrastPort()
    pix_o.max_samples, pix_o.max_views = ms, mv
    xc = feats.clone().requires_grad_(True)
    torch.manual_seed(5)
    lc_s = sup_o(xc, class_labels=weather)
    lc_p = pix_o(xc[:B], labels=labels, predict=predict)
    ((lc_s + lc_p) / B).backward()
    assert abs(lb_s.item() - lc_s.item()) <= LOSS_RTOL * abs(lc_s.item())
    assert abs(lb_p.item() - lc_p.item()) <= LOSS_RTOL * abs(lc_p.item())
    assert _relmax(xb.grad.cpu(), xc.grad) <= GRAD_RTOL


def test_supcon_mask_odd_batch_and_float_labels(dcl):
    """SupConLoss's remaining call forms (loss.py:149-159): an explicit same-group `mask`, float labels compared by
    value, and the failure on an odd number of crops."""
    opts = types.SimpleNamespace(deeplab=False)
    torch.manual_seed(3)
    sup = dcl.SupConLoss(device="cuda", opts=opts)
    g = torch.Generator().manual_seed(8)
    x = torch.randn(8, 128, 4, 6, generator=g).cuda()
    y = torch.tensor([[2], [0], [2], [1]])
    l_lab = sup(x, class_labels=y.cuda())
    mask = torch.eq(y, y.T).float().cuda()
    l_mask = sup(x, mask=mask)
    assert l_mask.item() == l_lab.item()
    l_float = sup(x, class_labels=torch.tensor([[0.5], [-1.25], [0.5], [7.0]]).cuda())
    assert l_float.item() == l_lab.item()
    assert sup(x, mask=torch.eye(4).cuda()).item() == sup(x).item()
    bad = mask.clone()
    bad[0, 1] = 1.0                                     # not an equivalence relation
    with pytest.raises(NotImplementedError):
        sup(x, mask=bad)
    with pytest.raises(ValueError):
        sup(torch.randn(7, 128, 4, 6).cuda())
    sup_o = O.SupConPort(opts=opts)
    sup_o.projection.load_state_dict({k: v.detach().cpu() for k, v in sup.projection.state_dict().items()})
    ref = sup_o(x.cpu(), mask=mask.cpu())
    assert abs(l_mask.item() - ref.item()) <= LOSS_RTOL * abs(ref.item())


def test_labels_outside_byte_range_raise(dcl):
    crit = dcl.PixelContrastLoss(device="cuda")
    labels = torch.zeros(1, 16, 16, dtype=torch.long)
    labels[0, 4, 8] = 300                      # a pixel the legacy-nearest down-sampling picks (rows/cols 0, 4, 8, 12)
    with pytest.raises(ValueError):
        crit(torch.randn(1, 128, 4, 4).cuda(), labels=labels.cuda(), predict=torch.randn(1, 19, 4, 4).cuda())
    with pytest.raises(dcl.loss._lib.DclError):
        crit(torch.randn(1, 128, 4, 4).cuda(), labels=labels, predict=torch.randn(1, 19, 4, 4).cuda())


def test_max_samples_raised_on_live_module(dcl):
    """`max_samples` is a public mutable attribute (loss.py:258): raising it between calls must re-size the buffers."""
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
    wl = WORKLOADS["small"]
    d = make_inputs(wl, seed=4, device="cuda")
    crit = dcl.PixelContrastLoss(device="cuda")
    crit.max_views = 64
    for ms in (1000, 1024, 1100, 2048):
        crit.max_samples = ms
        x = d["feats"].clone().requires_grad_(True)
        torch.manual_seed(1)
        loss = crit(x, labels=d["labels"], predict=d["predict"])
        loss.backward()
        plan = crit.last_plan
        assert plan.n_view == min(ms // plan.A, 64) and crit.last_layout.n == plan.A * plan.n_view
        assert int((x.grad.abs().sum(1) != 0).sum()) == crit.last_layout.n


def test_chunkwise_gather_scatter_equal_rowwise(dcl):
    """The pixel-ordered kernels (k_gather_px / k_scatter_px / k_dense_grad, driven by the pixel->row map that
    dcl_sample_select writes) move exactly the same values as the one-warp-per-row kernels."""
    from doubly_contrastive_semseg_b200 import _lib, loss as L
    from doubly_contrastive_semseg_b200.loss import _p, _stream
    for (B, h, w, n, n_pad) in [(3, 25, 45, 700, 768), (4, 64, 128, 4000, 4096), (2, 96, 96, 1, 128)]:
        hw = h * w
        g = torch.Generator(device="cuda").manual_seed(n)
        feats = torch.randn(B, 128, h, w, generator=g, device="cuda")
        pix = torch.full((n_pad,), -1, dtype=torch.int32, device="cuda")
        pix[:n] = torch.randperm(B * hw, generator=g, device="cuda")[:n].int()
        rowof = torch.full((B * hw,), -1, dtype=torch.int32, device="cuda")
        rowof[pix[:n].long()] = torch.arange(n, dtype=torch.int32, device="cuda")
        t1, s1 = L.gather_tiles(feats, pix, n_pad)
        t2, s2 = L.gather_tiles(feats, pix, n_pad, rowof)
        assert torch.equal(t1, t2) and torch.equal(s1, s2)
        dF = torch.randn(n_pad, 128, generator=g, device="cuda")
        up = torch.tensor(1.7, device="cuda")
        outs = []
        for rw in (None, rowof):
            for mode in (1, 2):
                d = torch.full((B, 128, h, w), 0.25, device="cuda")
                _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n_pad, _p(up), _p(d), B, hw, mode, _p(rw), _stream())
                outs.append(d)
        assert torch.equal(outs[0], outs[2]) and torch.equal(outs[1], outs[3])
        assert int((outs[0] != 0).any(1).sum()) == n
        # one-pass dense gradient == pooled broadcast followed by the scatter-add
        Ball = B + 2
        gap = torch.randn(Ball * 128, generator=g, device="cuda")
        want = torch.empty(Ball, 128, h, w, device="cuda")
        _lib.call("dcl_gap_bwd", _p(gap), Ball * 128, hw, _p(want), 0, _stream())
        _lib.call("dcl_scatter_grad", _p(dF), _p(pix), n_pad, _p(up), _p(want), B, hw, 2, None, _stream())
        got = torch.full_like(want, 9.0)
        _lib.call("dcl_dense_grad", _p(dF), _p(rowof), B, _p(up), _p(gap), _p(got), Ball, hw, _stream())
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-9)
        z = torch.full((1000 * 16,), 3.0, device="cuda")
        _lib.call("dcl_zero_fill", _p(z), z.numel() * 4, 1, _stream())
        assert float(z.abs().sum()) == 0.0


def test_gap_matches_torch(dcl):
    from doubly_contrastive_semseg_b200.loss import _GapFn
    for shape in [(4, 128, 6, 10), (2, 128, 5, 7), (6, 128, 64, 128)]:
        x = torch.randn(*shape, device="cuda", requires_grad=True)
        y = _GapFn.apply(x)
        ref = x.detach().mean(dim=(2, 3))
        assert torch.allclose(y, ref, rtol=1e-5, atol=1e-6)
        gout = torch.randn_like(y)
        y.backward(gout)
        ref_g = (gout / (shape[2] * shape[3]))[:, :, None, None].expand(shape)
        assert torch.allclose(x.grad, ref_g, rtol=1e-6, atol=0)


# ------------------------------------------------------------------ error behaviour
def test_error_contract(dcl):
    from doubly_contrastive_semseg_b200 import _lib
    crit = dcl.PixelContrastLoss(device="cuda")
    with pytest.raises(_lib.DclError):
        crit(torch.randn(1, 128, 4, 4), labels=torch.zeros(1, 16, 16, dtype=torch.long),
             predict=torch.randn(1, 19, 4, 4))
    with pytest.raises(ValueError):
        crit(torch.randn(1, 64, 4, 4).cuda(), labels=torch.zeros(1, 16, 16, dtype=torch.long).cuda(),
             predict=torch.randn(1, 19, 4, 4).cuda())
    sup = dcl.SupConLoss(device="cuda", opts=types.SimpleNamespace(deeplab=False))
    x = torch.randn(4, 128, 2, 2).cuda()
    with pytest.raises(ValueError):
        sup(x, class_labels=torch.zeros(2, 1, dtype=torch.long).cuda(), mask=torch.eye(2).cuda())
    with pytest.raises(ValueError):
        sup(x, class_labels=torch.zeros(3, 1, dtype=torch.long).cuda())
    # no class qualifies -> zero loss that still back-propagates (documented deviation)
    f = torch.randn(1, 128, 4, 4, device="cuda", requires_grad=True)
    loss = crit(f, labels=torch.full((1, 16, 16), 255, dtype=torch.long).cuda(),
                predict=torch.randn(1, 19, 4, 4).cuda())
    loss.backward()
    assert loss.item() == 0.0 and float(f.grad.abs().sum()) == 0.0


# ------------------------------------------------------------------ full-size properties (cfg2)
def test_full_size_cfg2_properties_and_oracle(dcl):
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
    wl = WORKLOADS["cfg2"]
    d = make_inputs(wl, seed=2, device="cuda")
    crit = dcl.PixelContrastLoss(device="cuda")
    crit.max_samples, crit.max_views = wl.max_samples, wl.max_views
    x = d["feats"].requires_grad_(True)
    torch.manual_seed(2)
    loss = crit(x, labels=d["labels"], predict=d["predict"])
    loss.backward()
    lay = crit.last_layout
    assert lay.n == 8192 and crit.last_plan.A == 128 and crit.last_plan.n_view == 64
    pix = crit.last_pix[: lay.n].long()
    hw = wl.h * wl.w
    b, p = pix // hw, pix % hw
    # sampled pixels carry the class they were sampled for and are all distinct
    iy = torch.floor(torch.arange(wl.h, dtype=torch.float32) * (torch.tensor(float(wl.H)) / wl.h)).long().clamp(max=wl.H - 1).cuda()
    ix = torch.floor(torch.arange(wl.w, dtype=torch.float32) * (torch.tensor(float(wl.W)) / wl.w)).long().clamp(max=wl.W - 1).cuda()
    lab_ds = d["labels"][:, iy][:, :, ix].reshape(wl.B, hw)
    assert torch.equal(lab_ds[b, p].cpu(), torch.from_numpy(lay.y[: lay.n]).long())
    assert torch.unique(pix).numel() == lay.n
    # gradient support == sampled pixels
    gr = x.grad
    rows_f = x.detach().reshape(wl.B, 128, hw)[b, :, p]
    rows_g = gr.reshape(wl.B, 128, hw)[b, :, p]
    assert int((gr.abs().sum(1) != 0).sum()) == lay.n
    # determinism (segment partials, no atomics) and linearity in the upstream gradient: the same
    # RNG state gives bit-identical loss and gradients; 2*loss gives exactly 2*grad
    x2 = d["feats"].detach().clone().requires_grad_(True)
    torch.manual_seed(2)
    loss2 = crit(x2, labels=d["labels"], predict=d["predict"])
    (2.0 * loss2).backward()
    assert loss2.item() == loss.item()
    assert torch.equal(x2.grad, 2.0 * gr)
    # full comparison against the fp64 closed form on the gathered rows
    loss_o, dF_o, _ = O.pixel_contrast_closed_form(rows_f.cpu(), torch.from_numpy(lay.y[: lay.n]).long())
    assert abs(loss.item() - loss_o) <= LOSS_RTOL * abs(loss_o)
    assert _relmax(rows_g.cpu(), dF_o) <= GRAD_RTOL


def test_full_size_cfg3_doubly_step(dcl):
    """BASELINE configs[2] at full size: the doubly contrastive step on one [32,128,256,512] tensor (2.15 GB) -
    image-level term on all 32 crops, pixel term on the first 16, (supcon + pixel) / batch_size, one backward
    (trainer.py:143-158).  Checked against the CPU port of the image-level term (dense part of the gradient) and the
    fp64 closed form of the pixel term on the gathered rows (the part at the sampled pixels)."""
    from doubly_contrastive_semseg_b200.synthetic import WORKLOADS, make_inputs
    wl = WORKLOADS["cfg3"]
    d = make_inputs(wl, seed=3, device="cuda")
    opts = types.SimpleNamespace(deeplab=False)
    torch.manual_seed(23)
    crit = dcl.DoublyContrastiveLoss(device="cuda", opts=opts)
    crit.pixel.max_samples, crit.pixel.max_views = wl.max_samples, wl.max_views
    x = d["feats"].requires_grad_(True)
    torch.manual_seed(3)
    l_sup, l_pix = crit(x, labels=d["labels"], predict=d["predict"], class_labels=d["weather"])
    ((l_sup + l_pix) / wl.B).backward()
    lay = crit.pixel.last_layout
    assert lay.n == 8192 and crit.pixel.last_plan.A == 256 and crit.pixel.last_plan.n_view == 32
    hw = wl.h * wl.w
    pixi = crit.pixel.last_pix[: lay.n].long()
    b, p = pixi // hw, pixi % hw
    assert int(b.max()) < wl.B and torch.unique(pixi).numel() == lay.n
    # image-level term alone on the GPU (same module, same kernels): its dense gradient is what the fused writer
    # broadcasts, so subtracting it isolates the pixel term exactly
    xs = x.detach().clone().requires_grad_(True)
    ls_gpu = crit.supcon(xs, class_labels=d["weather"])
    (ls_gpu / wl.B).backward()
    assert ls_gpu.item() == l_sup.item()
    dense = xs.grad
    # ... and against the CPU port with the same projection
    sup_o = O.SupConPort(opts=opts)
    sup_o.projection.load_state_dict({k: v.detach().cpu() for k, v in crit.supcon.projection.state_dict().items()})
    xc = x.detach().cpu().requires_grad_(True)
    ls_o = sup_o(xc, class_labels=d["weather"].cpu())
    (ls_o / wl.B).backward()
    assert abs(l_sup.item() - ls_o.item()) <= LOSS_RTOL * abs(ls_o.item())
    assert _relmax(dense.cpu(), xc.grad) <= GRAD_RTOL
    # pixel term: fp64 closed form on the gathered rows
    rows_f = x.detach().reshape(2 * wl.B, 128, hw)[b, :, p]
    lp_o, dF_o, _ = O.pixel_contrast_closed_form(rows_f.cpu(), torch.from_numpy(lay.y[: lay.n]).long())
    assert abs(l_pix.item() - lp_o) <= LOSS_RTOL * abs(lp_o)
    got_rows = (x.grad.reshape(2 * wl.B, 128, hw)[b, :, p].double() - dense.reshape(2 * wl.B, 128, hw)[b, :, p].double())
    # the difference of two fp32 numbers of the dense term's size carries that size's rounding
    slack = float(dense.abs().max()) * 2.0 ** -23 / float((dF_o / wl.B).abs().max())
    assert _relmax(got_rows.cpu(), dF_o / wl.B) <= GRAD_RTOL + slack
    # everywhere else the gradient is the image-level broadcast alone, bit for bit
    mask = torch.ones(2 * wl.B, hw, dtype=torch.bool, device="cuda")
    mask[b, p] = False
    gx = x.grad.reshape(2 * wl.B, 128, hw).permute(0, 2, 1)[mask]
    gd = dense.reshape(2 * wl.B, 128, hw).permute(0, 2, 1)[mask]
    assert torch.equal(gx, gd)


# ------------------------------------------------------------------ multi-GPU (needs >= 2 GPUs)
@pytest.mark.parametrize("workload", ["small"])
def test_sharded_equals_single_gpu(dcl, workload):
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box (gpurun --gpus 2); host logic is covered by the gloo test")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(root, "tools", "sharded_check.py"), workload]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


# ------------------------------------------------------------------ segmentation-loss neighbour
def _focal_opts(mode):
    return types.SimpleNamespace(with_depth_level_loss=False, criterion="plain_focal" if mode == "plain_focal" else "x",
                                 no_class_weights=mode == "no_class_weights", no_EDT=mode == "no_EDT")


@pytest.mark.parametrize("case", _cases("focal"))
def test_boundary_focal_vs_reference_golden(dcl, case):
    """BoundaryAwareFocalLoss (utils/loss.py:27-80) through the fused kernel against outputs of the real reference:
    loss, gradient w.r.t. the pre-upsample logits, and the in-place ignore -> 0 rewrite of the target."""
    g = _gold(case)
    mode, gamma = str(g["mode"]), float(g["gamma"])
    C = g["logits"].shape[1]
    crit = dcl.BoundaryAwareFocalLoss(gamma=gamma, num_classes=C, ignore_id=255, weight=torch.from_numpy(g["weight"]),
                                      device="cuda", opts=_focal_opts(mode))
    x = torch.from_numpy(g["logits"]).cuda().requires_grad_(True)
    t = torch.from_numpy(g["target"].astype(np.int64)).cuda()
    loss = crit(x, t, {"label_distance_weight": torch.from_numpy(g["alpha"])})          # CPU weight map, like the loader's
    (2.0 * loss).backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert _relmax(x.grad.cpu() / 2.0, g["dlogits"]) <= 1e-3
    assert np.array_equal(t.cpu().numpy(), g["target_after"].astype(np.int64))
    assert crit.step_counter == 1


def test_boundary_focal_full_size_vs_torch(dcl):
    """cfg2's shapes (batch 8, labels 1024x2048, logits 256x512): the fused kernel against the plain-PyTorch
    restatement of the reference running on the GPU (which materialises the 1.27 GB up-sampled logits), plus
    size-independent properties: every pixel's class gradients sum to zero, repeat is bit-identical, zero EDT weight
    gives a zero loss."""
    g = torch.Generator(device="cuda").manual_seed(77)
    B, C, h, w, H, W = 8, 19, 256, 512, 1024, 2048
    logits = 2.0 * torch.randn(B, C, h, w, generator=g, device="cuda")
    target = torch.randint(0, C, (B, H, W), generator=g, device="cuda")
    target[torch.rand(B, H, W, generator=g, device="cuda") < 0.05] = 255
    alpha = torch.rand(B, H, W, generator=g, device="cuda") * 3.0
    alpha[torch.rand(B, H, W, generator=g, device="cuda") < 0.4] = 0.0
    alpha[target == 255] = 0.0
    weight = 0.5 + 4.0 * torch.rand(C, generator=g, device="cuda")
    opts = _focal_opts("full")
    crit = dcl.BoundaryAwareFocalLoss(gamma=0.5, num_classes=C, ignore_id=255, weight=weight, device="cuda", opts=opts)
    x = logits.clone().requires_grad_(True)
    t = target.clone()
    loss = crit(x, t, {"label_distance_weight": alpha})
    loss.backward()
    port = O.BoundaryFocalPort(gamma=0.5, num_classes=C, ignore_id=255, weight=weight, device="cuda", opts=opts)
    xr = logits.clone().requires_grad_(True)
    tr = target.clone()
    ref = port(xr, tr, {"label_distance_weight": alpha})
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item())
    assert float((x.grad - xr.grad).abs().max() / xr.grad.abs().max()) <= 1e-3
    assert torch.equal(t, tr)
    assert float(x.grad.sum(dim=1).abs().max()) <= 1e-4 * float(x.grad.abs().max())
    x2 = logits.clone().requires_grad_(True)
    loss2 = crit(x2, target.clone(), {"label_distance_weight": alpha})
    loss2.backward()
    assert loss2.item() == loss.item() and torch.equal(x2.grad, x.grad)
    x3 = logits.clone().requires_grad_(True)
    loss3 = crit(x3, target.clone(), {"label_distance_weight": torch.zeros_like(alpha)})
    loss3.backward()
    assert loss3.item() == 0.0 and float(x3.grad.abs().max()) == 0.0
    # full-resolution logits (what trainer.py passes today): same value as up-sampling inside the loss
    up = torch.nn.functional.interpolate(logits[:2], (H, W), mode="bilinear", align_corners=False)
    xa = up.clone().requires_grad_(True)
    la = crit(xa, target[:2].clone(), {"label_distance_weight": alpha[:2]})
    xb = logits[:2].clone().requires_grad_(True)
    lb = crit(xb, target[:2].clone(), {"label_distance_weight": alpha[:2]})
    assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())


def test_boundary_focal_error_contract(dcl):
    crit = dcl.BoundaryAwareFocalLoss(gamma=0.5, weight=None, device="cuda", opts=_focal_opts("full"), ignore_id=255)
    x = torch.randn(1, 19, 4, 4).cuda()
    t = torch.zeros(1, 8, 8, dtype=torch.long).cuda()
    with pytest.raises(TypeError):                       # self.weight[target] with weight=None, loss.py:54
        crit(x, t, {"label_distance_weight": torch.ones(1, 8, 8)})
    crit.weight = torch.ones(19)
    with pytest.raises(KeyError):
        crit(x, t, {})
    from doubly_contrastive_semseg_b200 import _lib
    with pytest.raises(_lib.DclError):
        crit(x.cpu(), t, {"label_distance_weight": torch.ones(1, 8, 8)})


# ------------------------------------------------------------------ full training step (BASELINE config 5)
def _train_sample(B, H, W, K, seed):
    g = torch.Generator().manual_seed(seed)
    left = torch.rand(2 * B, 3, H, W, generator=g) * 255.0
    coarse = torch.randint(0, K, (B, H // 32, W // 32), generator=g)
    label = coarse.repeat_interleave(32, 1).repeat_interleave(32, 2).clone()
    label[torch.rand(B, H, W, generator=g) < 0.04] = 255
    alpha = torch.rand(B, H, W, generator=g) * 2.0
    alpha[torch.rand(B, H, W, generator=g) < 0.3] = 0.0
    return {"left": left, "label": label.long(), "weather": torch.arange(B) % 2, "label_distance_weight": alpha}


@pytest.mark.gpu
def test_train_step_vs_ports(dcl):
    """One optimisation step of `TrainStep` (trainer.py:60-215, criterion supcon_pixelcontrast_focal) in fp32 against
    the same network driven by the CPU ports of the three losses and a plain torch Adam: the three losses, the
    parameter gradients and the updated weights."""
    import copy
    from doubly_contrastive_semseg_b200.swiftnet import fill_deterministic
    B, H, W, K = 2, 128, 256, 5
    opts = types.SimpleNamespace(amp=False, batch_size=B, lr=4e-4, weight_decay=1e-4, deeplab=False)
    torch.manual_seed(3)
    step = dcl.TrainStep(opts, device="cuda", class_weights=0.5 + torch.rand(19))
    fill_deterministic(step.net, 11)
    step.pixelcontrast_criterion.max_samples, step.pixelcontrast_criterion.max_views = 1024, 8
    ref_net = copy.deepcopy(step.net)
    ref_net.upsample_logits = True
    groups = [{"params": list(ref_net.random_init_params()), "lr": 4e-4, "weight_decay": 1e-4},
              {"params": list(ref_net.fine_tune_params()), "lr": 1e-4, "weight_decay": 2.5e-5}]
    ref_opt = torch.optim.Adam(groups, betas=(0.9, 0.99))
    sample = _train_sample(B, H, W, K, seed=21)
    # ---- ours
    torch.manual_seed(9)
    out = step({k: v.clone() for k, v in sample.items()})
    # ---- the same step with the ports (losses on the CPU, network on the GPU)
    sup_o = O.SupConPort(opts=opts)
    sup_o.projection.load_state_dict({k: v.detach().cpu() for k, v in step.supcon_criterion.projection.state_dict().items()})
    pix_o = O.PixelContrastPort()
    pix_o.max_samples, pix_o.max_views = 1024, 8
    foc_o = O.BoundaryFocalPort(gamma=0.5, num_classes=19, ignore_id=255, weight=step.criterion.weight.detach().cpu(),
                                device="cpu", opts=step.opts)
    ref_net.train()
    seg, before, fine, fine0 = ref_net(sample["left"].cuda(), return_supcon_feature=True)
    labels = sample["label"].clone()
    torch.manual_seed(9)
    l_sup = sup_o(fine.cpu(), class_labels=sample["weather"])
    l_pix = pix_o(fine0.cpu(), labels=labels, predict=before.cpu())
    l_seg = foc_o(seg.cpu(), labels, {"label_distance_weight": sample["label_distance_weight"]})
    total = (l_sup + l_pix) * (1.0 / B) + l_seg * 1.2
    ref_opt.zero_grad()
    total.backward()
    ref_opt.step()
    assert abs(out["supcon_loss"].item() - l_sup.item()) <= LOSS_RTOL * abs(l_sup.item())
    assert abs(out["pixelcontrast_loss"].item() - l_pix.item()) <= LOSS_RTOL * abs(l_pix.item())
    assert abs(out["seg_loss"].item() - l_seg.item()) <= LOSS_RTOL * abs(l_seg.item())
    assert abs(out["total_loss"].item() - total.item()) <= LOSS_RTOL * abs(total.item())
    worst = 0.0
    for (name, p), q in zip(step.net.named_parameters(), ref_net.parameters()):
        assert p.grad is not None and q.grad is not None, name
        worst = max(worst, float((p.grad - q.grad).abs().max() / q.grad.abs().max().clamp_min(1e-20)))
    assert worst <= GRAD_RTOL, worst
    # Adam's first step moves a weight by lr * g / (|g| + eps): identical wherever the gradient is not tiny, at most
    # 2 lr apart anywhere (a sign flip of a near-zero gradient); the segmentation head is not optimised
    moved = 0
    with torch.no_grad():
        for (name, p), q in zip(step.net.named_parameters(), ref_net.parameters()):
            if name.startswith("segmentation."):
                assert torch.equal(p, q), name
                continue
            moved += 1
            lr = 4e-4 if "upsample_" in name else 1e-4
            d = (p - q).abs()
            assert float(d.max()) <= 2.05 * lr, name
            big = q.grad.abs() > 1e-2 * q.grad.abs().max()
            assert float(d[big].max()) <= 0.05 * lr, name
    assert moved == 83
    assert step.num_iter == 1


@pytest.mark.gpu
@pytest.mark.parametrize("labels_kind", ["int64", "int32", "none", "mask"])
def test_supcon_fused_head_equals_torch_head(dcl, labels_kind):
    """The image-level head as four launches (projection MLP forward / backward, group ids, fp32 contrast;
    csrc/dcl_contrast_small.cu) against the same module with the torch MLP and glue: loss, d features, d projection."""
    from doubly_contrastive_semseg_b200 import loss as L
    B = 6
    g = torch.Generator().manual_seed(40)
    feats = torch.randn(2 * B, 128, 6, 10, generator=g)
    weather = torch.tensor([0, 3, 1, 3, 0, 2])
    opts = types.SimpleNamespace(deeplab=False)
    torch.manual_seed(12)
    crit = dcl.SupConLoss(device="cuda", opts=opts)
    kw = {"int64": dict(class_labels=weather.cuda()), "int32": dict(class_labels=weather.int().cuda().view(-1, 1)),
          "none": {}, "mask": dict(mask=(weather[:, None] == weather[None, :]).float().cuda())}[labels_kind]
    res = []
    for fused in (True, False):
        L._FUSED_HEAD = fused
        try:
            for p in crit.projection.parameters():
                p.grad = None
            x = feats.cuda().requires_grad_(True)
            n0 = L.launch_count()
            loss = crit(x, **kw)
            (3.0 * loss).backward()
            res.append((loss.item(), x.grad.clone(), [p.grad.clone() for p in crit.projection.parameters()], L.launch_count() - n0))
        finally:
            L._FUSED_HEAD = True
    (la, ga, pa, na), (lb, gb, pb, nb) = res
    assert abs(la - lb) <= 1e-5 * abs(lb)
    assert _relmax(ga.cpu(), gb.cpu()) <= 1e-4
    for a, b in zip(pa, pb):
        assert _relmax(a.cpu(), b.cpu()) <= 1e-4
    assert na >= 5                                       # pool, (group ids), MLP, contrast, MLP backward x2, pool backward
