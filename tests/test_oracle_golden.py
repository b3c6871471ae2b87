"""Pins the CPU oracle (oracle/dcl_oracle.py) to vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from oracle import dcl_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


# ------------------------------------------------------------------ RNG restatement
@pytest.mark.parametrize("seed", [0, 1, 101, 2**31 + 7])
def test_mt19937_matches_torch_randperm(seed):
    rng = O.MT19937(seed)
    torch.manual_seed(seed)
    for n in [0, 1, 2, 3, 17, 623, 624, 625, 2000, 5, 1300]:
        want = torch.randperm(n).numpy()
        k = min(n, 7 if n % 2 else n)
        got = O.randperm_prefix(rng, n, k)
        assert np.array_equal(got, want[:k]), (seed, n)


# ------------------------------------------------------------------ sampler
PIXEL_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "pixel_*.npz")))


@pytest.mark.parametrize("case", PIXEL_CASES)
def test_sampler_bit_exact_vs_reference(case):
    g = _load(case)
    B, C, h, w = g["feats"].shape
    lab = O.downsample_labels(g["labels"].astype(np.int64), h, w)
    pred = O.argmax_first(g["predict"])
    assert np.array_equal(lab.reshape(B, h, w), g["lab_ds"])
    assert np.array_equal(pred.reshape(B, h, w), g["pred"])
    rng = O.MT19937(int(g["call_seed"]))
    plan = O.sample_anchors(lab, pred, 255, int(g["max_samples"]), int(g["max_views"]),
                            lambda n, k: O.randperm_prefix(rng, n, k))
    assert plan.A == g["pixels"].shape[0] and plan.n_view == g["pixels"].shape[1]
    assert np.array_equal(plan.pixels, g["pixels"])
    assert np.array_equal(np.array(plan.cls), g["y"])
    # and the same through torch's global generator
    torch.manual_seed(int(g["call_seed"]))
    plan2 = O.sample_anchors(lab, pred, 255, int(g["max_samples"]), int(g["max_views"]),
                             O.torch_randperm_prefix)
    assert np.array_equal(plan2.pixels, g["pixels"])


@pytest.mark.parametrize("case", PIXEL_CASES)
def test_pixel_closed_form_vs_reference(case):
    g = _load(case)
    B, C, h, w = g["feats"].shape
    lab = O.downsample_labels(g["labels"].astype(np.int64), h, w)
    pred = O.argmax_first(g["predict"])
    rng = O.MT19937(int(g["call_seed"]))
    plan = O.sample_anchors(lab, pred, 255, int(g["max_samples"]), int(g["max_views"]),
                            lambda n, k: O.randperm_prefix(rng, n, k))
    F, y = O.gather_anchor_rows(g["feats"], plan)
    loss, dF, _ = O.pixel_contrast_closed_form(F, y)
    assert abs(loss - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    dfe = np.zeros((B, C, h * w))
    for a in range(plan.A):
        for v in range(plan.n_view):
            dfe[plan.image[a], :, plan.pixels[a, v]] += dF[v * plan.A + a].numpy()
    ref = g["dfeats"].reshape(B, C, h * w)
    assert np.abs(dfe - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("case", PIXEL_CASES)
def test_pixel_port_vs_reference(case):
    g = _load(case)
    crit = O.PixelContrastPort(device="cpu")
    crit.max_samples, crit.max_views = int(g["max_samples"]), int(g["max_views"])
    x = torch.from_numpy(g["feats"]).requires_grad_(True)
    torch.manual_seed(int(g["call_seed"]))
    loss = crit(x, labels=torch.from_numpy(g["labels"].astype(np.int64)),
                predict=torch.from_numpy(g["predict"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.abs(x.grad.numpy() - g["dfeats"]).max() <= 1e-4 * np.abs(g["dfeats"]).max()


# ------------------------------------------------------------------ N x N contrast
@pytest.mark.parametrize("case", sorted(os.path.basename(p) for p in
                                        glob.glob(os.path.join(GOLDEN, "contrast_*.npz"))))
def test_contrast_closed_form_vs_reference(case):
    g = _load(case)
    X, y = g["X"], g["y"]
    A, V, D = X.shape
    F = np.concatenate([X[:, v] for v in range(V)], axis=0)             # row = v*A + a
    yy = np.tile(y, V)
    loss, dF, st = O.pixel_contrast_closed_form(F, yy, chunk=64)
    assert abs(loss - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    dX = dF.numpy().reshape(V, A, D).transpose(1, 0, 2)
    assert np.abs(dX - g["dX"]).max() <= 1e-4 * np.abs(g["dX"]).max()
    # row-permutation invariance (the CUDA path reorders rows by class)
    perm = np.random.default_rng(0).permutation(F.shape[0])
    loss_p, dF_p, _ = O.pixel_contrast_closed_form(F[perm], yy[perm])
    assert abs(loss_p - loss) <= 1e-12 * abs(loss)
    assert np.abs(dF_p.numpy() - dF.numpy()[perm]).max() <= 1e-12 * np.abs(dF.numpy()).max() + 1e-18


def _large_contrast_case(case):
    """Rows (reference order, row = v*A + a), labels and golden outputs of a contrastL_* fixture; the inputs are
    regenerated from the stored seed and checked against the stored digest."""
    import hashlib
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = _load(case)
    A, V, K = int(g["A"]), int(g["V"]), int(g["K"])
    X, y = mg.contrast_inputs(int(g["seed"]), A, V, K)
    digest = np.frombuffer(hashlib.sha256(X.numpy().tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(digest, g["x_sha256"]), "regenerated inputs differ from the ones the reference saw"
    assert np.array_equal(y.numpy().astype(np.int64), g["y"])
    F = torch.cat([X[:, v] for v in range(V)], dim=0)
    yy = torch.from_numpy(np.tile(g["y"], V))
    return F, yy, float(g["loss"]), g["dX"], (A, V)


@pytest.mark.parametrize("case", ["contrastL_n2048.npz", "contrastL_n8192.npz"])
def test_contrast_closed_form_vs_reference_large(case):
    """The reference's own `_contrastive` at N = 2048 / 8192 (fp32 autograd) pins the fp64 closed form at the sizes
    where the CUDA path switches to its degree-1 polynomial and positive-pair series."""
    F, yy, loss_ref, dX_ref, (A, V) = _large_contrast_case(case)
    loss, dF, _ = O.pixel_contrast_closed_form(F, yy, chunk=1024)
    assert abs(loss - loss_ref) <= 2e-6 * abs(loss_ref)
    dX = dF.numpy().reshape(V, A, 128).transpose(1, 0, 2)
    assert np.abs(dX - dX_ref).max() <= 2e-4 * np.abs(dX_ref).max()


# ------------------------------------------------------------------ image-level term
@pytest.mark.parametrize("case", sorted(os.path.basename(p) for p in
                                        glob.glob(os.path.join(GOLDEN, "supcon_*.npz"))))
def test_supcon_port_and_closed_form_vs_reference(case):
    g = _load(case)
    crit = O.SupConPort(opts=types.SimpleNamespace(deeplab=False))
    sd = {k: torch.from_numpy(g[k.replace(".", "_")]) for k in crit.projection.state_dict()}
    crit.projection.load_state_dict(sd)
    x = torch.from_numpy(g["feats"]).requires_grad_(True)
    labels = torch.from_numpy(g["weather"]) if bool(g["use_labels"]) else None
    loss = crit(x, class_labels=labels)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert np.abs(x.grad.numpy() - g["dfeats"]).max() <= 1e-4 * np.abs(g["dfeats"]).max()
    for k, p in crit.projection.named_parameters():
        ref = g["g_" + k.replace(".", "_")]
        assert np.abs(p.grad.numpy() - ref).max() <= 1e-4 * np.abs(ref).max() + 1e-12
    # closed form on the projected rows
    B2 = g["feats"].shape[0]
    B = B2 // 2
    with torch.no_grad():
        Z = crit.projection(torch.from_numpy(g["feats"]).mean(dim=(2, 3)))
    y = np.tile(g["weather"].reshape(-1), 2) if bool(g["use_labels"]) else np.tile(np.arange(B), 2)
    loss_cf, dZ, _ = O.supcon_closed_form(Z.numpy(), y)
    assert abs(loss_cf - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    # dZ against autograd of the port
    Zt = Z.clone().double().requires_grad_(True)
    a = (Zt @ Zt.T) / 0.07
    l = torch.nn.functional.normalize(a - a.max(dim=1, keepdim=True).values.detach())
    ns = 1.0 - torch.eye(B2, dtype=torch.float64)
    yt = torch.from_numpy(y)
    pos = (yt[:, None] == yt[None, :]).double() * ns
    lp = l - torch.log((torch.exp(l) * ns).sum(1, keepdim=True))
    (-(pos * lp).sum(1) / pos.sum(1)).mean().backward()
    assert (Zt.grad - dZ).abs().max() <= 1e-9 * Zt.grad.abs().max()


def test_supcon_port_errors_match_reference_contract():
    crit = O.SupConPort(opts=types.SimpleNamespace(deeplab=False))
    x = torch.randn(4, 128, 2, 2)
    with pytest.raises(ValueError):
        crit(x, class_labels=torch.zeros(2, 1, dtype=torch.long), mask=torch.eye(2))
    with pytest.raises(ValueError):
        crit(x, class_labels=torch.zeros(3, 1, dtype=torch.long))


# ------------------------------------------------------------------ segmentation-loss neighbour
FOCAL_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "focal_*.npz")))


def _focal_opts(mode):
    return types.SimpleNamespace(with_depth_level_loss=False, criterion="plain_focal" if mode == "plain_focal" else "x",
                                 no_class_weights=mode == "no_class_weights", no_EDT=mode == "no_EDT")


@pytest.mark.parametrize("case", FOCAL_CASES)
def test_boundary_focal_port_and_closed_form_vs_reference(case):
    """BoundaryAwareFocalLoss (loss.py:27-80): the torch port and the fp64 closed form (the math of the fused CUDA
    kernel, including the bilinear taps and their transpose) against outputs of the real reference."""
    g = _load(case)
    mode, gamma = str(g["mode"]), float(g["gamma"])
    C = g["logits"].shape[1]
    crit = O.BoundaryFocalPort(gamma=gamma, num_classes=C, ignore_id=255, weight=torch.from_numpy(g["weight"]),
                               device="cpu", opts=_focal_opts(mode))
    x = torch.from_numpy(g["logits"]).requires_grad_(True)
    t = torch.from_numpy(g["target"].astype(np.int64))
    loss = crit(x, t, {"label_distance_weight": torch.from_numpy(g["alpha"])})
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert np.abs(x.grad.numpy() - g["dlogits"]).max() <= 1e-5 * np.abs(g["dlogits"]).max()
    assert np.array_equal(t.numpy(), g["target_after"].astype(np.int64))          # the in-place ignore -> 0 rewrite
    loss_cf, dx, t_cf = O.focal_closed_form(g["logits"], g["target"], g["alpha"], g["weight"], gamma, mode)
    assert abs(loss_cf - float(g["loss"])) <= 2e-6 * abs(float(g["loss"]))
    assert np.abs(dx - g["dlogits"]).max() <= 2e-5 * np.abs(g["dlogits"]).max()
    assert np.array_equal(t_cf, g["target_after"].astype(np.int64))


def test_boundary_focal_port_zero_weight_returns_zero():
    crit = O.BoundaryFocalPort(gamma=0.5, weight=torch.ones(19), opts=_focal_opts("full"), ignore_id=255)
    x = torch.randn(1, 19, 4, 4, requires_grad=True)
    loss = crit(x, torch.zeros(1, 8, 8, dtype=torch.long), {"label_distance_weight": torch.zeros(1, 8, 8)})
    assert loss.item() == 0.0
    loss.backward()


def test_focal_kernel_decomposition_prototype():
    """The work decomposition of k_focal (csrc/dcl_focal.cu) restated in numpy (tools/proto_focal.py): runs of label
    columns by first tap, hand-over of the lx parts to the next column's owner, strips of low-resolution rows with a
    spill row - against the fp64 closed form on two reference-generated fixtures, for strip heights that exercise
    the spill path and the single-strip path."""
    from tools import proto_focal as P
    for name in ("focal_ragged", "focal_no_edt"):
        g = np.load(os.path.join(GOLDEN, name + ".npz"))
        mode, gamma = str(g["mode"]), float(g["gamma"])
        ref_loss, ref_grad, _ = O.focal_closed_form(g["logits"], g["target"], g["alpha"], g["weight"], gamma, mode)
        for strip in (2, g["logits"].shape[2]):
            loss, grad, t = P.emulate(g["logits"], g["target"], g["alpha"], g["weight"], gamma, mode, strip)
            assert abs(loss - ref_loss) <= 1e-12 * abs(ref_loss)
            assert np.abs(grad - ref_grad).max() <= 1e-12 * np.abs(ref_grad).max()
            assert np.array_equal(t, g["target_after"].astype(np.int64))
