"""Golden vectors for the embedding producer from the REAL reference network
(`/root/reference/network/weathernet.py` on the resnet18 pyramid trunk), build container only:
    python tests/golden/make_golden_net.py
Writes tests/golden/swiftnet_rn18.npz (outputs of an eval forward, of a train-mode forward / backward, the updated
batch-norm statistics), swiftnet_rn18_keys.json and swiftnet_rn34_keys.json (state_dict key -> shape, optimiser
group shapes).  The reference
package needs matplotlib (stubbed) and downloads ImageNet weights (stubbed to an empty dict: strict=False makes that a
no-op); weights come from swiftnet.fill_deterministic, applied identically on both sides."""
import contextlib
import io
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm"):
    sys.modules.setdefault(m, types.ModuleType(m))
warnings.filterwarnings("ignore")
import torch.utils.model_zoo as model_zoo  # noqa: E402

model_zoo.load_url = lambda *a, **k: {}


def main():
    from doubly_contrastive_semseg_b200.swiftnet import fill_deterministic
    with contextlib.redirect_stdout(io.StringIO()):
        from network.weathernet import WeatherNet
        net = WeatherNet(types.SimpleNamespace(deeplab=False), backbone="resnet18")
    fill_deterministic(net, 5)
    keys = {k: list(v.shape) for k, v in net.state_dict().items()}
    groups = {"random_init": [list(p.shape) for p in net.random_init_params()],
              "fine_tune": [list(p.shape) for p in net.fine_tune_params()]}
    json.dump({"state_dict": keys, "groups": groups}, open(os.path.join(HERE, "swiftnet_rn18_keys.json"), "w"), indent=0)
    # the ResNet-34 trunk (WeatherNet's default backbone): names, shapes and optimiser groups only
    with contextlib.redirect_stdout(io.StringIO()):
        net34 = WeatherNet(types.SimpleNamespace(deeplab=False), backbone="resnet34")
    json.dump({"state_dict": {k: list(v.shape) for k, v in net34.state_dict().items()},
               "groups": {"random_init": [list(p.shape) for p in net34.random_init_params()],
                          "fine_tune": [list(p.shape) for p in net34.fine_tune_params()]}},
              open(os.path.join(HERE, "swiftnet_rn34_keys.json"), "w"), indent=0)
    del net34
    g = torch.Generator().manual_seed(9)
    img = torch.rand(2, 3, 64, 128, generator=g) * 255.0                   # two views of one image
    out = {"image": img.numpy()}
    net.eval()
    with torch.no_grad():
        seg, before, fine, fine0 = net(img, return_supcon_feature=True)
    out.update(eval_seg=seg.numpy(), eval_before=before.numpy(), eval_fine=fine.numpy())
    with torch.no_grad():
        seg1, before1, fine1, fine01 = net(img[:1], return_supcon_feature=False)
    out.update(eval_single_before=before1.numpy())
    net.train()
    seg, before, fine, fine0 = net(img, return_supcon_feature=True)
    wf = torch.linspace(-1, 1, fine.numel()).view_as(fine)
    wb = torch.linspace(1, -1, before.numel()).view_as(before)
    loss = (fine * wf).sum() + (before * wb).sum() + 1e-3 * (seg ** 2).sum()
    loss.backward()
    sd = net.state_dict()
    named = dict(net.named_parameters())
    out.update(train_fine=fine.detach().numpy(), train_before=before.detach().numpy(), train_loss=np.float64(loss.item()),
               g_conv1=named["feature_extractor.conv1.weight"].grad.numpy(),
               g_seg=named["segmentation.conv.weight"].grad.numpy(),
               g_l4=named["feature_extractor.layer4.1.conv2.weight"].grad.numpy()[:8],
               g_blend5=named["feature_extractor.upsample_blends5.blend_conv.conv.weight"].grad.numpy()[:8],
               g_bn1_2=named["feature_extractor.bn1_2.weight"].grad.numpy(),
               rm_bn1_1=sd["feature_extractor.bn1_1.running_mean"].numpy(),
               rv_seg=sd["segmentation.norm.running_var"].numpy())
    np.savez_compressed(os.path.join(HERE, "swiftnet_rn18.npz"), **out)
    print("wrote swiftnet_rn18.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
