"""Generate golden vectors from the REAL reference (`/root/reference/utils/loss.py`).

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/pixel_*.npz, contrast_*.npz, supcon_*.npz.  The reference module is
loaded by file path (its package __init__ needs matplotlib) and executed
unmodified; stdout is silenced because loss.py:270 prints on every call.
"""
import contextlib
import importlib.util
import io
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/utils/loss.py"


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_loss", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def blocky_labels(g, B, H, W, K, block, ignore_frac=0.05):
    bh, bw = (H + block - 1) // block, (W + block - 1) // block
    coarse = torch.randint(0, K, (B, bh, bw), generator=g)
    lab = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].clone()
    lab[torch.rand(B, H, W, generator=g) < ignore_frac] = 255
    return lab.long()


def make_pixel_inputs(seed, B, H, W, h, w, K, block, correct_frac):
    g = torch.Generator().manual_seed(seed)
    labels = blocky_labels(g, B, H, W, K, block)
    feats = torch.randn(B, 128, h, w, generator=g)
    predict = torch.randn(B, 19, h, w, generator=g)
    # push the true class on a fraction of pixels so hard and easy sets are both populated
    lab_ds = torch.nn.functional.interpolate(labels[:, None].float(), (h, w), mode="nearest")[:, 0].long()
    boost = (torch.rand(B, h, w, generator=g) < correct_frac) & (lab_ds != 255)
    idx = lab_ds.clamp(max=18)
    predict.scatter_add_(1, idx[:, None], 4.0 * boost[:, None].float())
    return feats, labels, predict


def run_pixel(ref, name, seed, B, H, W, h, w, K, block, correct_frac, max_samples, max_views,
              call_seed):
    feats, labels, predict = make_pixel_inputs(seed, B, H, W, h, w, K, block, correct_frac)
    crit = ref.PixelContrastLoss(device="cpu")
    crit.max_samples, crit.max_views = max_samples, max_views
    x = feats.clone().requires_grad_(True)
    torch.manual_seed(call_seed)
    with contextlib.redirect_stdout(io.StringIO()):
        loss = crit(x, labels=labels, predict=predict)
    loss.backward()
    # recover the sampled pixel indices: replay the sampler with index-encoded features
    pix = torch.arange(h * w, dtype=torch.float32).view(1, h * w, 1).expand(B, h * w, 128).contiguous()
    lab_ds = torch.nn.functional.interpolate(labels[:, None].float().clone(), (h, w), mode="nearest")[:, 0].long()
    pred = predict.max(1)[1]
    torch.manual_seed(call_seed)
    with contextlib.redirect_stdout(io.StringIO()):
        X_, y_ = crit._hard_anchor_sampling(pix, lab_ds.view(B, -1).float(), pred.view(B, -1).float())
    np.savez_compressed(
        os.path.join(HERE, f"pixel_{name}.npz"),
        feats=feats.numpy(), labels=labels.numpy().astype(np.int16), predict=predict.numpy(),
        max_samples=max_samples, max_views=max_views, call_seed=call_seed,
        loss=loss.item(), dfeats=x.grad.numpy(),
        pixels=X_[:, :, 0].numpy().astype(np.int64), y=y_.numpy().astype(np.int64),
        lab_ds=lab_ds.numpy().astype(np.int16), pred=pred.numpy().astype(np.int16))
    print(f"pixel_{name}: A={X_.shape[0]} n_view={X_.shape[1]} loss={loss.item():.6f}")


def run_contrast(ref, name, seed, A, V, K, offset):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, K, (A,), generator=g).float()
    cent = torch.randn(K, 128, generator=g)
    X = (0.5 * torch.randn(A, V, 128, generator=g) + 0.5 * cent[y.long()][:, None, :] + offset)
    X = X.clone().requires_grad_(True)
    crit = ref.PixelContrastLoss(device="cpu")
    loss = crit._contrastive(X, y)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, f"contrast_{name}.npz"), X=X.detach().numpy(),
                        y=y.numpy().astype(np.int64), loss=loss.item(), dX=X.grad.numpy())
    print(f"contrast_{name}: N={A * V} loss={loss.item():.6f}")


def contrast_inputs(seed, A, V, K):
    """Inputs of the large contrast fixtures, regenerated from the seed by the tests (torch's CPU generator is
    deterministic for a given torch build; the fixture stores a digest of the bytes to prove it)."""
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, K, (A,), generator=g).float()
    cent = torch.randn(K, 128, generator=g)
    X = 0.5 * torch.randn(A, V, 128, generator=g) + 0.5 * cent[y.long()][:, None, :]
    return X, y


def run_contrast_large(ref, name, seed, A, V, K):
    """Reference `_contrastive` (loss.py:339-389) at N = A*V in the thousands: the size class where the CUDA path
    takes its degree-1 polynomial / positive-pair series route, pinned to the reference's own output."""
    import hashlib
    X, y = contrast_inputs(seed, A, V, K)
    digest = hashlib.sha256(X.numpy().tobytes()).hexdigest()
    X = X.clone().requires_grad_(True)
    crit = ref.PixelContrastLoss(device="cpu")
    loss = crit._contrastive(X, y)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, f"contrastL_{name}.npz"), seed=seed, A=A, V=V, K=K,
                        x_sha256=np.frombuffer(bytes.fromhex(digest), dtype=np.uint8),
                        y=y.numpy().astype(np.int64), loss=loss.item(), dX=X.grad.numpy())
    print(f"contrastL_{name}: N={A * V} loss={loss.item():.6f}")


def run_supcon(ref, name, seed, B, h, w, use_labels):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)                       # projection init (nn.Linear default init)
    crit = ref.SupConLoss(temperature=0.07, contrast_mode="all", base_temperature=0.07, weight=None,
                          device="cpu", opts=types.SimpleNamespace(deeplab=False))
    feats = torch.randn(2 * B, 128, h, w, generator=g) + 0.3 * torch.randn(2 * B, 128, 1, 1, generator=g)
    weather = torch.randint(0, 4, (B, 1), generator=g)
    x = feats.clone().requires_grad_(True)
    loss = crit(x, class_labels=weather if use_labels else None, mask=None)
    loss.backward()
    sd = {k.replace(".", "_"): v.detach().numpy() for k, v in crit.projection.state_dict().items()}
    gd = {"g_" + k.replace(".", "_"): p.grad.numpy() for k, p in crit.projection.named_parameters()}
    np.savez_compressed(os.path.join(HERE, f"supcon_{name}.npz"), feats=feats.numpy(),
                        weather=weather.numpy(), use_labels=use_labels, loss=loss.item(),
                        dfeats=x.grad.numpy(), **sd, **gd)
    print(f"supcon_{name}: rows={2 * B} loss={loss.item():.6f}")


def focal_inputs(seed, B, C, h, w, H, W, ignore_frac=0.08):
    g = torch.Generator().manual_seed(seed)
    logits = 2.0 * torch.randn(B, C, h, w, generator=g)
    target = torch.randint(0, C, (B, H, W), generator=g)
    target[torch.rand(B, H, W, generator=g) < ignore_frac] = 255
    # EDT-style weight: positive near "boundaries", exactly zero on a part of the image and on ignore pixels
    alpha = torch.rand(B, H, W, generator=g) * 3.0
    alpha[torch.rand(B, H, W, generator=g) < 0.3] = 0.0
    alpha[target == 255] = 0.0
    weight = 0.5 + torch.rand(C, generator=g) * 4.0
    return logits, target.long(), alpha, weight


def run_focal(ref, name, seed, B, C, h, w, H, W, gamma, mode):
    """BoundaryAwareFocalLoss (loss.py:27-80) on pre-upsample logits [B,C,h,w] (the loss up-samples them itself when the
    label size differs, loss.py:41-42) or on full-resolution logits (h == H)."""
    logits, target, alpha, weight = focal_inputs(seed, B, C, h, w, H, W)
    opts = types.SimpleNamespace(with_depth_level_loss=False, criterion="plain_focal" if mode == "plain_focal" else "x",
                                 no_class_weights=mode == "no_class_weights", no_EDT=mode == "no_EDT")
    crit = ref.BoundaryAwareFocalLoss(gamma=gamma, num_classes=C, ignore_id=255, weight=weight, device="cpu", opts=opts)
    x = logits.clone().requires_grad_(True)
    t = target.clone()
    loss = crit(x, t, {"label_distance_weight": alpha})
    loss.backward()
    np.savez_compressed(os.path.join(HERE, f"focal_{name}.npz"), logits=logits.numpy(), target=target.numpy().astype(np.int16),
                        alpha=alpha.numpy(), weight=weight.numpy(), gamma=gamma, mode=mode, loss=loss.item(),
                        dlogits=x.grad.numpy(), target_after=t.numpy().astype(np.int16))
    print(f"focal_{name}: loss={loss.item():.6f}")


def main():
    ref = load_reference()
    # name, seed, B, H, W, h, w, K, block, correct_frac, max_samples, max_views, call_seed
    run_pixel(ref, "defaults", 11, 2, 32, 64, 8, 16, 19, 8, 0.6, 1024, 2, 101)
    run_pixel(ref, "views6", 12, 2, 32, 64, 8, 16, 5, 8, 0.6, 1024, 6, 102)
    run_pixel(ref, "oddview_ragged", 13, 3, 30, 50, 8, 16, 4, 7, 0.5, 1024, 5, 103)
    run_pixel(ref, "hard_scarce", 14, 2, 32, 64, 8, 16, 3, 16, 0.97, 1024, 8, 104)
    run_pixel(ref, "easy_scarce", 15, 2, 32, 64, 8, 16, 3, 16, 0.03, 1024, 8, 105)
    run_pixel(ref, "sample_capped", 16, 2, 32, 64, 8, 16, 6, 8, 0.6, 40, 8, 106)
    run_contrast(ref, "n192", 21, 24, 8, 5, 0.0)
    run_contrast(ref, "n130_offset", 22, 26, 5, 4, 1.5)
    run_contrast(ref, "n40_twoclass", 23, 20, 2, 2, 0.0)
    run_contrast_large(ref, "n2048", 41, 64, 32, 16)
    run_contrast_large(ref, "n8192", 42, 128, 64, 16)
    run_focal(ref, "x4", 51, 2, 19, 6, 10, 24, 40, 0.5, "full")
    run_focal(ref, "ragged", 52, 2, 19, 5, 7, 18, 30, 0.5, "full")
    run_focal(ref, "fullres", 53, 1, 19, 12, 20, 12, 20, 0.5, "full")
    run_focal(ref, "no_edt", 54, 2, 7, 6, 10, 24, 40, 2.0, "no_EDT")
    run_focal(ref, "no_cw", 55, 2, 7, 6, 10, 24, 40, 0.0, "no_class_weights")
    run_focal(ref, "plain", 56, 1, 19, 8, 8, 32, 32, 0.5, "plain_focal")
    run_supcon(ref, "labels_b4", 31, 4, 4, 6, True)
    run_supcon(ref, "simclr_b3", 32, 3, 4, 6, False)
    run_supcon(ref, "labels_b16", 33, 16, 2, 3, True)


if __name__ == "__main__":
    main()
