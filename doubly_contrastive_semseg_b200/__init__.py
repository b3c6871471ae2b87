"""B200-native doubly contrastive loss (pixel-level + image-level) behind the reference's
loss-module interface.  See DESIGN.md / INTEGRATION.md."""
from .loss import (MODE_PIXEL, MODE_SUPCON, PixelContrastLoss, SupConLoss, contrast_rows,
                   plan_anchors, layout_rows)

__all__ = ["PixelContrastLoss", "SupConLoss", "contrast_rows", "plan_anchors", "layout_rows",
           "MODE_PIXEL", "MODE_SUPCON"]
