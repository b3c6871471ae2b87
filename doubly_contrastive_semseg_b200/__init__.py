"""B200-native doubly contrastive loss (pixel-level + image-level) behind the reference's
loss-module interface.  See DESIGN.md / INTEGRATION.md."""
from .loss import (MODE_PIXEL, MODE_SUPCON, DoublyContrastiveLoss, PixelContrastLoss, ShardedPixelContrastLoss, SupConLoss,
                   contrast_rows, layout_rows, plan_anchors, shard_plan)

from .focal import BoundaryAwareFocalLoss
from .swiftnet import WeatherNet
from .train_step import TrainStep

__all__ = ["BoundaryAwareFocalLoss", "WeatherNet", "TrainStep", "PixelContrastLoss", "DoublyContrastiveLoss", "ShardedPixelContrastLoss", "SupConLoss", "shard_plan", "contrast_rows", "plan_anchors", "layout_rows",
           "MODE_PIXEL", "MODE_SUPCON"]
