"""Drop-in `BoundaryAwareFocalLoss` (reference utils/loss.py:27-80), the segmentation-loss neighbour of the
contrastive path: same constructor, same `forward(input, target, batch)` and the same side effects
(`target` is rewritten in place, `step_counter` advances), with forward and gradient computed by one fused kernel
(csrc/dcl_focal.cu) that never forms the up-sampled logits.  Pass the PRE-upsample logits (`left_seg_beforeup`,
network/weathernet.py:87) to get the fusion; full-resolution logits (`left_seg`) work too - the reference up-samples
inside the loss with the very function the model uses (loss.py:5,41-42; weathernet.py:88), so both give the same value.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _lib
from .loss import _count, _on_device, _p, _require_cuda, _stream

MODE_FULL, MODE_PLAIN, MODE_NO_CLASS_WEIGHTS, MODE_NO_EDT = 0, 1, 2, 3


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, alpha, weight, gamma, mode, ignore_id):
        B, C, h, w = logits.shape
        H, W = target.shape[1], target.shape[2]
        dev = logits.device
        unscaled = torch.empty_like(logits)
        loss_n = torch.empty(2, dtype=torch.float32, device=dev)
        nbytes = int(_lib.load().dcl_focal_workspace_bytes(B, h, w))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
        _lib.call("dcl_focal_fwd", _p(logits), _p(target), _p(alpha), _p(weight), B, C, h, w, H, W, int(ignore_id),
                  float(gamma), int(mode), _p(unscaled), _p(loss_n), _p(ws), nbytes, _stream())
        _count(2)
        ctx.save_for_backward(unscaled, loss_n)
        return loss_n[0]

    @staticmethod
    def backward(ctx, grad_out):
        unscaled, loss_n = ctx.saved_tensors
        g = grad_out.to(torch.float32).contiguous()
        out = torch.empty_like(unscaled)
        _lib.call("dcl_focal_bwd", _p(unscaled), _p(loss_n), _p(g), _p(out), ctypes.c_size_t(unscaled.numel()), _stream())
        _count(1)
        return out, None, None, None, None, None, None


class BoundaryAwareFocalLoss(nn.Module):
    """Reference: utils/loss.py:27-80; constructed at utils/init_trainer.py:216-218, called as
    `self.criterion(left_seg, labels, sample)` (trainer.py:139,155).

    Deviations (documented in DESIGN.md): CUDA sm_100 only; when no pixel has a positive EDT weight the reference
    returns a constant zero that is detached from `input` (loss.py:46-47) - here the zero loss back-propagates zeros;
    labels must be class ids below `input.shape[1]` or `ignore_id`; up to 32 classes.
    """

    def __init__(self, gamma=0, num_classes=19, ignore_id=19, print_each=20, weight=None, device=None, opts=None):
        super().__init__()
        self.num_classes = num_classes
        self.ignore_id = ignore_id
        self.print_each = print_each
        self.step_counter = 0
        self.gamma = gamma
        self.weight = weight
        self.device = device
        self.opts = opts

    def forward(self, input, target, batch, **kwargs):
        _require_cuda(input, "input")
        if input.dim() != 4 or target.dim() != 3 or input.shape[0] != target.shape[0]:
            raise ValueError("input must be [B,C,h,w] and target [B,H,W]")
        dev = input.device
        _on_device(target, dev, "target")
        if target.dtype != torch.int64 or not target.is_contiguous():
            raise ValueError("target must be a contiguous int64 tensor (it is rewritten in place: ignore_id -> 0)")
        if target.shape[1] < input.shape[2] or target.shape[2] < input.shape[3]:
            raise NotImplementedError("logits larger than the label map (down-sampling inside the loss) are not supported")
        alpha = batch["label_distance_weight"].to(dev)                          # loss.py:44
        if tuple(alpha.shape[-3:]) != tuple(target.shape) and alpha.numel() != target.numel():
            raise ValueError("label_distance_weight must match target")
        alpha = alpha.reshape(target.shape).contiguous().to(torch.float32)
        if self.weight is None:
            # `self.weight[target]` (loss.py:54) runs in every mode
            raise TypeError("'NoneType' object is not subscriptable")
        if getattr(self.opts, "with_depth_level_loss", False):
            batch["disp_distance_weight"].to(dev)                               # loss.py:56-58 (read, never used)
        weight = self.weight.to(device=dev, dtype=torch.float32).contiguous()
        crit = getattr(self.opts, "criterion", None)
        if crit == "plain_focal":
            mode = MODE_PLAIN
        elif getattr(self.opts, "no_class_weights", False):
            mode = MODE_NO_CLASS_WEIGHTS
        elif getattr(self.opts, "no_EDT", False):
            mode = MODE_NO_EDT
        else:
            mode = MODE_FULL
        x = input.contiguous().to(torch.float32)
        with torch.cuda.device(dev):
            loss = _FocalFn.apply(x, target, alpha, weight, float(self.gamma), mode, int(self.ignore_id))
        self.step_counter += 1
        return loss
