"""One optimisation step of the reference trainer with this framework's losses (BASELINE config 5):
trainer.py:60-215 for the criteria that use the contrastive terms, init_trainer.py:160-177 (Adam, two learning-rate
groups), init_trainer.py:299-306 (cosine schedule), trainer.py:407-421 (checkpoint dictionary).

    step = TrainStep(opts)                       # opts: the reference's option names (see _DEFAULTS)
    out = step(sample)                           # sample = the loader's dict, or the (sample0, sample1) pair of the
                                                 # two-crop loader ('supcon' criteria), trainer.py:64-71
    step.save_checkpoint(path, score)            # {'epoch','num_iter','model_state','optimizer_state','score',...}

What runs where: the network (swiftnet.WeatherNet) is cuDNN under bf16 autocast in channels_last; everything after it
is this repository's CUDA - DoublyContrastiveLoss (one pass over fine_feat for both contrastive terms) or its parts,
and the fused BoundaryAwareFocalLoss on the PRE-upsample logits (the [B,19,H,W] tensor the reference feeds its loss is
never formed).  The weather classifier of trainer.py:109-114 only feeds a logged accuracy (its loss is not part of
total_loss, trainer.py:204) and is not built.

Several GPUs: one process per GPU under DistributedDataParallel (the reference uses nn.DataParallel, i.e. the losses
see the global batch).  The pixel term keeps that meaning through ShardedPixelContrastLoss - anchors of the global
batch sharded over the ranks, contrast set all-gathered - whose gradient is already the global one, so it is scaled by
the world size to survive DDP's averaging; the image-level and focal terms are per-rank means (a 2B x 2B contrast and
a per-pixel mean: DDP averages them).  Every rank seeds torch's CPU generator with `seed + num_iter` before the
sampler draws, which the sharded sampler requires (all ranks replay one stream).
"""
import os
import types
from typing import Optional

import torch
import torch.nn as nn

from .focal import BoundaryAwareFocalLoss
from .loss import DoublyContrastiveLoss, PixelContrastLoss, ShardedPixelContrastLoss, SupConLoss
from .swiftnet import WeatherNet

_DEFAULTS = dict(lr=4e-4, weight_decay=1e-4, last_lr=1e-6, epochs=400, batch_size=8, num_classes=19,
                 criterion="supcon_pixelcontrast_focal", backbone="resnet18", deeplab=False, amp=True,
                 with_depth_level_loss=False, no_class_weights=False, no_EDT=False, seed=0,
                 channels_last=True)
_CRITERIA = ("supcon_focal", "supcon_simclr_focal", "pixelcontrast_focal", "supcon_pixelcontrast_focal",
             "supcon_simclr_pixelcontrast_focal", "focal")


def _opts(opts):
    o = types.SimpleNamespace(**_DEFAULTS)
    if opts is not None:
        for k, v in (vars(opts) if not isinstance(opts, dict) else opts).items():
            setattr(o, k, v)
    return o


class TrainStep:
    def __init__(self, opts=None, device="cuda", class_weights: Optional[torch.Tensor] = None, process_group=None,
                 model: Optional[nn.Module] = None):
        self.opts = o = _opts(opts)
        if o.criterion not in _CRITERIA:
            raise NotImplementedError("criterion %r (contrastive / focal criteria of trainer.py:115-181 only)" % (o.criterion,))
        self.device = torch.device(device)
        self.world = torch.distributed.get_world_size(process_group) if torch.distributed.is_initialized() else 1
        self.process_group = process_group
        net = model if model is not None else WeatherNet(o, num_classes=o.num_classes, backbone=o.backbone,
                                                         upsample_logits=False, amp=o.amp)
        net = net.to(self.device)
        if o.channels_last:
            net = net.to(memory_format=torch.channels_last)
        self.net = net
        self.model = net
        if self.world > 1:
            self.model = nn.parallel.DistributedDataParallel(net, device_ids=[self.device.index], process_group=process_group,
                                                             gradient_as_bucket_view=True)
        weight = class_weights if class_weights is not None else torch.ones(o.num_classes)
        self.criterion = BoundaryAwareFocalLoss(gamma=0.5, num_classes=o.num_classes, ignore_id=255, weight=weight,
                                                device=self.device, opts=o)                    # init_trainer.py:216-218
        self.supcon_criterion = SupConLoss(temperature=0.07, contrast_mode="all", base_temperature=0.07, weight=weight,
                                           device=self.device, opts=o)                          # :221
        if self.world > 1:
            self.pixelcontrast_criterion = ShardedPixelContrastLoss(device=self.device, process_group=process_group)
            self.doubly = None
        else:
            self.pixelcontrast_criterion = PixelContrastLoss(device=self.device)                # :222
            self.doubly = DoublyContrastiveLoss(pixel=self.pixelcontrast_criterion, supcon=self.supcon_criterion,
                                                device=self.device, opts=o)
        # init_trainer.py:168-177: decoder at lr, trunk at lr / 4 with weight decay / 4; nothing else is optimised
        fine_tune_factor = 4
        groups = [{"params": list(net.random_init_params()), "lr": o.lr, "weight_decay": o.weight_decay},
                  {"params": list(net.fine_tune_params()), "lr": o.lr / fine_tune_factor,
                   "weight_decay": o.weight_decay / fine_tune_factor}]
        self.optimizer = torch.optim.Adam(groups, betas=(0.9, 0.99), fused=self.device.type == "cuda")
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, o.epochs, o.last_lr)
        self.cur_epochs, self.num_iter, self.best_score, self.best_score_epoch = 0, 0, 0.0, 0

    # ------------------------------------------------------------------ one step (trainer.py:60-215)
    def __call__(self, sample):
        return self.step(sample)

    def step(self, sample):
        o = self.opts
        supcon = "supcon" in o.criterion
        if isinstance(sample, (tuple, list)):                                   # two-crop loader, trainer.py:64-71
            s0, s1 = sample
            sample = dict(s0)
            sample["left"] = torch.cat([s0["left"], s1["left"]], dim=0)
        elif supcon and sample["left"].shape[0] != 2 * sample["label"].shape[0]:
            raise ValueError("'supcon' criteria need the two crops of every image: 'left' must hold 2B images")
        self.num_iter += 1
        left = sample["left"].to(self.device, dtype=torch.float32, non_blocking=True)
        if o.channels_last:
            left = left.contiguous(memory_format=torch.channels_last)
        labels = sample["label"].to(self.device, dtype=torch.long, non_blocking=True)
        self.model.train()
        _, before_up, fine_feat, fine_feat0 = self.model(left, return_supcon_feature=supcon)
        zero = torch.zeros((), device=self.device)
        supcon_loss, pixel_loss = zero, zero
        class_labels = None
        if supcon and "simclr" not in o.criterion:
            class_labels = sample["weather"].to(self.device)
        pixel = "pixelcontrast" in o.criterion
        if self.world > 1:
            torch.manual_seed(o.seed + self.num_iter)                           # one generator stream on every rank
        if supcon and pixel and self.doubly is not None:
            supcon_loss, pixel_loss = self.doubly(fine_feat, labels=labels, predict=before_up, class_labels=class_labels)
        else:
            if supcon:
                supcon_loss = self.supcon_criterion(fine_feat, class_labels=class_labels, mask=None)
            if pixel:
                pixel_loss = self.pixelcontrast_criterion(fine_feat0, labels=labels, predict=before_up)
                if self.world > 1:
                    pixel_loss = pixel_loss * self.world                        # survives DDP's mean over ranks
        seg_loss = self.criterion(before_up, labels, sample)                    # rewrites labels in place (ignore -> 0)
        if supcon or pixel:
            total = (supcon_loss + pixel_loss) * (1.0 / o.batch_size) + seg_loss * 1.2          # trainer.py:115-181
        else:
            total = seg_loss
        self.optimizer.zero_grad(set_to_none=True)
        total.backward()
        self.optimizer.step()
        if self.world > 1 and pixel:
            pixel_loss = pixel_loss / self.world
        return {"total_loss": total.detach(), "supcon_loss": supcon_loss.detach(), "pixelcontrast_loss": pixel_loss.detach(),
                "seg_loss": seg_loss.detach()}

    def end_epoch(self):
        self.scheduler.step()
        self.cur_epochs += 1

    # ------------------------------------------------------------------ checkpoint (trainer.py:407-421, saver)
    def checkpoint(self, score=None):
        return {"epoch": self.cur_epochs, "num_iter": self.num_iter, "model_state": self.net.state_dict(),
                "optimizer_state": self.optimizer.state_dict(), "score": score, "best_score": self.best_score,
                "best_score_epoch": self.best_score_epoch}

    def save_checkpoint(self, path, score=None):
        if self.world > 1 and torch.distributed.get_rank(self.process_group) != 0:
            return
        tmp = path + ".tmp"
        torch.save(self.checkpoint(score), tmp)
        os.replace(tmp, path)

    def load_checkpoint(self, path_or_dict, load_optimizer=True):
        ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location=self.device)
        self.net.load_state_dict(ck["model_state"])
        if load_optimizer and ck.get("optimizer_state") is not None:
            self.optimizer.load_state_dict(ck["optimizer_state"])                # init_trainer.py:255
        self.cur_epochs = int(ck.get("epoch", 0))
        self.num_iter = int(ck.get("num_iter", 0))
        self.best_score = ck.get("best_score", 0.0)
        self.best_score_epoch = ck.get("best_score_epoch", 0)
        return ck
