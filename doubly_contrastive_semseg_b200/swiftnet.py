"""The producer of the embeddings the contrastive losses consume: pyramid SwiftNet on a ResNet-18/34
trunk with the 1x1 segmentation head, as the reference's `WeatherNet` builds it (network/weathernet.py:14-86,
network/backbone/resnet_pyramid.py:109-310, network/utils.py:35-106).  The convolutions are library calls (cuDNN) -
this module exists so that the full training step of BASELINE config 5 can run on this framework's loss kernels; it is
written for the B200 box, not transcribed:

  * parameter / buffer names and shapes are the reference's (`state_dict()` round-trips with a reference checkpoint,
    trainer.py:413-421), the forward graph is built from a table of stages instead of the reference's unrolled code;
  * memory format is channels_last and the trunk runs under bf16 autocast when asked to (`amp=True`): 180 GB of HBM
    means no activation checkpointing (the reference wraps every conv-bn in torch.utils.checkpoint);
  * the three pyramid levels share the trunk's weights but keep their own stem batch-norm (`bn1_0..2`), exactly like
    the reference; the up-sampled full-resolution logits are only produced on request (the fused focal loss consumes
    the pre-upsample logits, focal.py).

forward(left_img, return_supcon_feature) -> (pred_segmap | None, pred_segmap_beforeup, fine_feat, fine_feat0), the
reference's 4-tuple (weathernet.py:68-92): with `return_supcon_feature` the batch holds the two views back to back and
the segmentation head only sees the first half.
"""
from itertools import chain

import torch
import torch.nn as nn
import torch.nn.functional as F

ACDC_MEAN = (73.15, 82.90, 72.3)          # weathernet.py:37-38
ACDC_STD = (47.67, 48.49, 47.73)
_LAYERS = {"resnet18": (2, 2, 2, 2), "resnet34": (3, 4, 6, 3)}


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=k // 2, bias=bias)


class _Residual(nn.Module):
    """conv3x3-bn-relu-conv3x3-bn + shortcut, relu (resnet_pyramid.py:55-90)"""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1, self.bn1 = _conv(cin, cout, 3, stride), nn.BatchNorm2d(cout)
        self.conv2, self.bn2 = _conv(cout, cout, 3), nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x)), inplace=True)))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class _NormReluConv(nn.Sequential):
    """bn -> relu -> conv, the decoder's unit (network/utils.py:35-51); sub-module names norm / relu / conv"""

    def __init__(self, cin, cout, k, bias=False):
        super().__init__()
        self.add_module("norm", nn.BatchNorm2d(cin))
        self.add_module("relu", nn.ReLU(inplace=True))
        self.add_module("conv", _conv(cin, cout, k, bias=bias))


class _Blend(nn.Module):
    """bilinear up to the skip's size, add, bn-relu-conv3x3 (network/utils.py:83-106)"""

    def __init__(self, ch, k):
        super().__init__()
        self.blend_conv = _NormReluConv(ch, ch, k)

    def forward(self, x, skip):
        x = F.interpolate(x, skip.shape[2:], mode="bilinear", align_corners=False)
        return self.blend_conv(x + skip)


class PyramidResNet(nn.Module):
    """Shared ResNet trunk over an image pyramid, skips of equal stride summed, blended coarse to fine."""

    def __init__(self, layers, num_features=128, pyramid_levels=3, k_upsample=3, mean=ACDC_MEAN, std=ACDC_STD):
        super().__init__()
        self.pyramid_levels, self.num_features = pyramid_levels, num_features
        self.register_buffer("img_mean", torch.tensor(mean).view(1, -1, 1, 1))
        self.register_buffer("img_std", torch.tensor(std).view(1, -1, 1, 1))
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        for lvl in range(pyramid_levels):                             # one stem batch-norm per pyramid level
            setattr(self, "bn1_%d" % lvl, nn.BatchNorm2d(64))
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin = 64
        for i, (width, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            blocks = [_Residual(cin if j == 0 else width, width, (1 if i == 1 else 2) if j == 0 else 1) for j in range(n)]
            setattr(self, "layer%d" % i, nn.Sequential(*blocks))
            setattr(self, "upsample_bottlenecks%d" % i, _conv(width, num_features, 1))
            cin = width
        self.n_blends = pyramid_levels + 2                             # skip levels - 1 (output stride 4)
        for i in range(1, self.n_blends + 1):
            setattr(self, "upsample_blends%d" % i, _Blend(num_features, k_upsample))
        for m in self.modules():                                       # resnet_pyramid.py:218-223
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    # the reference's two optimiser groups (resnet_pyramid.py:189-191, 211-214, 225-229)
    def random_init_params(self):
        mods = [getattr(self, "upsample_bottlenecks%d" % i) for i in range(1, 5)]
        mods += [getattr(self, "upsample_blends%d" % i) for i in range(1, self.n_blends + 1)]
        return chain(*[m.parameters() for m in mods])

    def fine_tune_params(self):
        mods = [self.conv1] + [getattr(self, "layer%d" % i) for i in range(1, 5)]
        mods += [getattr(self, "bn1_%d" % l) for l in range(self.pyramid_levels)]
        return chain(*[m.parameters() for m in mods])

    def forward(self, image):
        image = (image - self.img_mean) / self.img_std
        skips = [[] for _ in range(self.pyramid_levels + 3)]           # index = log2(stride) - 2
        for lvl in range(self.pyramid_levels):
            x = image if lvl == 0 else F.interpolate(image, scale_factor=1 / 2 ** lvl, mode="bicubic", align_corners=None)
            x = self.maxpool(F.relu(getattr(self, "bn1_%d" % lvl)(self.conv1(x)), inplace=True))
            for i in range(1, 5):
                x = getattr(self, "layer%d" % i)(x)
                skips[lvl + i - 1].append(getattr(self, "upsample_bottlenecks%d" % i)(x))
        x = skips[-1][0]
        coarse = x
        for i in range(1, self.n_blends + 1):
            x = getattr(self, "upsample_blends%d" % i)(x, sum(skips[-1 - i]))
        return x, {"skips_0": coarse}


class WeatherNet(nn.Module):
    """Drop-in for network/weathernet.py:14 with a ResNet trunk (`backbone` = 'resnet18' | 'resnet34'; the reference's
    efficientnet option is not built).  `opts` is kept for signature compatibility."""

    def __init__(self, opts=None, num_classes=19, backbone="resnet18", upsample_logits=True, amp=False, **_unused):
        super().__init__()
        if backbone not in _LAYERS:
            raise NotImplementedError(backbone)
        self.opts, self.num_classes, self.upsample_logits, self.amp = opts, num_classes, upsample_logits, amp
        self.feature_extractor = PyramidResNet(_LAYERS[backbone])
        self.segmentation = _NormReluConv(self.feature_extractor.num_features, num_classes, 1, bias=True)

    def forward(self, left_img, return_supcon_feature=False):
        with torch.autocast(device_type=left_img.device.type, dtype=torch.bfloat16, enabled=self.amp):
            fine_feat, _ = self.feature_extractor(left_img)
            fine_feat0 = fine_feat[: fine_feat.shape[0] // 2] if return_supcon_feature else fine_feat
            before_up = self.segmentation(fine_feat0)
        # the losses work in fp32 on contiguous NCHW tensors
        fine_feat = fine_feat.to(dtype=torch.float32, memory_format=torch.contiguous_format)      # one pass (bf16 NHWC -> f32 NCHW)
        fine_feat0 = fine_feat[: fine_feat.shape[0] // 2] if return_supcon_feature else fine_feat
        before_up = before_up.to(dtype=torch.float32, memory_format=torch.contiguous_format)
        seg = None
        if self.upsample_logits:
            seg = F.interpolate(before_up, left_img.shape[2:], mode="bilinear", align_corners=False)
        return seg, before_up, fine_feat, fine_feat0

    # the reference optimises the extractor's two groups only: the segmentation head is in neither
    # (weathernet.py:98-104, its chain() with self.segmentation is commented out)
    def random_init_params(self):
        return self.feature_extractor.random_init_params()

    def fine_tune_params(self):
        return self.feature_extractor.fine_tune_params()


def fill_deterministic(module: nn.Module, seed: int = 0) -> None:
    """Repeatable stand-in for a checkpoint (there is no network for the ImageNet weights): every state_dict entry,
    in key order, from its own numpy stream - convolutions at He scale, batch-norm affine / running statistics
    perturbed around their defaults.  The same call on the reference's WeatherNet gives identical weights (the key
    sets are equal), which is how tests/golden/swiftnet_*.npz were made."""
    import numpy as np
    sd = module.state_dict()
    with torch.no_grad():
        for i, key in enumerate(sorted(sd)):
            t = sd[key]
            rs = np.random.RandomState((seed * 100003 + i) % (2 ** 31))
            if key.endswith("num_batches_tracked") or key.endswith("img_mean") or key.endswith("img_std"):
                continue
            if t.dim() == 4:
                fan_out = t.shape[0] * t.shape[2] * t.shape[3]
                v = rs.standard_normal(tuple(t.shape)) * np.sqrt(2.0 / fan_out)
            elif key.endswith("running_var"):
                v = 1.0 + 0.2 * rs.random_sample(tuple(t.shape))
            elif key.endswith("norm.weight") or ".bn" in key and key.endswith("weight") or key.endswith("downsample.1.weight"):
                v = 1.0 + 0.1 * rs.standard_normal(tuple(t.shape))
            else:
                v = 0.1 * rs.standard_normal(tuple(t.shape))
            t.copy_(torch.from_numpy(np.asarray(v, dtype=np.float32)))
