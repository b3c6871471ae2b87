"""CPU tests of the embedding producer and the training-step host logic: the network against outputs of
the REAL reference network (tests/golden/swiftnet_rn18.npz, made by tests/golden/make_golden_net.py), the optimiser
groups of init_trainer.py:168-177 and the checkpoint dictionary of trainer.py:407-421."""
import json
import os
import types
import warnings

import numpy as np
import pytest
import torch

warnings.filterwarnings("ignore")
from doubly_contrastive_semseg_b200.swiftnet import WeatherNet, fill_deterministic  # noqa: E402
from doubly_contrastive_semseg_b200.train_step import TrainStep  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "swiftnet_rn18.npz")), json.load(open(os.path.join(GOLD, "swiftnet_rn18_keys.json")))


def test_state_dict_and_groups_are_the_reference_s(gold):
    _, keys = gold
    net = WeatherNet(None, backbone="resnet18")
    sd = net.state_dict()
    assert list(sd.keys()) == list(keys["state_dict"].keys())                     # same names, same order
    assert all(list(sd[k].shape) == v for k, v in keys["state_dict"].items())
    assert [list(p.shape) for p in net.random_init_params()] == keys["groups"]["random_init"]
    assert [list(p.shape) for p in net.fine_tune_params()] == keys["groups"]["fine_tune"]
    # the segmentation head is in neither group (weathernet.py:98-104)
    n_opt = len(keys["groups"]["random_init"]) + len(keys["groups"]["fine_tune"])
    assert n_opt == len(list(net.parameters())) - 4
    with pytest.raises(NotImplementedError):
        WeatherNet(None, backbone="efficientnetb0")


def test_resnet34_trunk_has_the_reference_s_names_and_groups():
    keys = json.load(open(os.path.join(GOLD, "swiftnet_rn34_keys.json")))
    net = WeatherNet(None, backbone="resnet34")
    sd = net.state_dict()
    assert list(sd.keys()) == list(keys["state_dict"].keys())
    assert all(list(sd[k].shape) == v for k, v in keys["state_dict"].items())
    assert [list(p.shape) for p in net.random_init_params()] == keys["groups"]["random_init"]
    assert [list(p.shape) for p in net.fine_tune_params()] == keys["groups"]["fine_tune"]


def test_forward_matches_the_reference_network(gold):
    g, _ = gold
    net = WeatherNet(None, backbone="resnet18")
    fill_deterministic(net, 5)
    img = torch.from_numpy(g["image"])
    net.eval()
    with torch.no_grad():
        seg, before, fine, fine0 = net(img, return_supcon_feature=True)
        _, before1, fine1, fine01 = net(img[:1], return_supcon_feature=False)
    assert fine.shape == (2, 128, 16, 32) and fine0.shape == (1, 128, 16, 32) and before.shape == (1, 19, 16, 32)
    assert fine01.data_ptr() == fine1.data_ptr()
    assert _rel(seg, g["eval_seg"]) <= 1e-5 and _rel(before, g["eval_before"]) <= 1e-5 and _rel(fine, g["eval_fine"]) <= 1e-5
    assert _rel(before1, g["eval_single_before"]) <= 1e-5
    # no up-sampled logits unless asked for (the fused focal loss takes the pre-upsample ones)
    lean = WeatherNet(None, backbone="resnet18", upsample_logits=False)
    lean.load_state_dict(net.state_dict())
    lean.eval()
    with torch.no_grad():
        seg2, before2, _, _ = lean(img, return_supcon_feature=True)
    assert seg2 is None and torch.equal(before2, before)


def test_train_mode_forward_backward_matches_the_reference_network(gold):
    g, _ = gold
    net = WeatherNet(None, backbone="resnet18")
    fill_deterministic(net, 5)
    net.train()
    seg, before, fine, _ = net(torch.from_numpy(g["image"]), return_supcon_feature=True)
    wf = torch.linspace(-1, 1, fine.numel()).view_as(fine)
    wb = torch.linspace(1, -1, before.numel()).view_as(before)
    loss = (fine * wf).sum() + (before * wb).sum() + 1e-3 * (seg ** 2).sum()
    loss.backward()
    named, sd = dict(net.named_parameters()), net.state_dict()
    assert _rel(fine.detach(), g["train_fine"]) <= 1e-5 and _rel(before.detach(), g["train_before"]) <= 1e-5
    assert abs(loss.item() - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    assert _rel(named["feature_extractor.conv1.weight"].grad, g["g_conv1"]) <= 1e-4
    assert _rel(named["segmentation.conv.weight"].grad, g["g_seg"]) <= 1e-4
    assert _rel(named["feature_extractor.layer4.1.conv2.weight"].grad[:8], g["g_l4"]) <= 1e-4
    assert _rel(named["feature_extractor.upsample_blends5.blend_conv.conv.weight"].grad[:8], g["g_blend5"]) <= 1e-4
    assert _rel(named["feature_extractor.bn1_2.weight"].grad, g["g_bn1_2"]) <= 1e-4
    assert _rel(sd["feature_extractor.bn1_1.running_mean"], g["rm_bn1_1"]) <= 1e-5
    assert _rel(sd["segmentation.norm.running_var"], g["rv_seg"]) <= 1e-5


def test_optimizer_groups_schedule_and_checkpoint(tmp_path):
    opts = types.SimpleNamespace(amp=False, lr=4e-4, weight_decay=1e-4, epochs=10, last_lr=1e-6)
    st = TrainStep(opts, device="cpu")
    gs = st.optimizer.param_groups
    assert [len(g["params"]) for g in gs] == [19, 64]
    assert gs[0]["lr"] == 4e-4 and gs[1]["lr"] == 1e-4 and gs[0]["weight_decay"] == 1e-4 and gs[1]["weight_decay"] == 2.5e-5
    assert gs[0]["betas"] == (0.9, 0.99)
    st.end_epoch()
    assert st.cur_epochs == 1 and st.optimizer.param_groups[0]["lr"] < 4e-4
    fill_deterministic(st.net, 3)
    st.num_iter, st.best_score, st.best_score_epoch = 7, 0.61, 1
    path = str(tmp_path / "ck.pth")
    st.save_checkpoint(path, score={"Mean IoU": 0.6})
    ck = torch.load(path, weights_only=False)
    assert list(ck.keys()) == ["epoch", "num_iter", "model_state", "optimizer_state", "score", "best_score", "best_score_epoch"]
    assert len(ck["model_state"]) == 173 and ck["num_iter"] == 7 and ck["epoch"] == 1
    other = TrainStep(opts, device="cpu")
    other.load_checkpoint(path)
    assert other.num_iter == 7 and other.cur_epochs == 1 and other.best_score == 0.61
    for (k, a), (_, b) in zip(st.net.state_dict().items(), other.net.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(NotImplementedError):
        TrainStep(types.SimpleNamespace(criterion="crossentropy"), device="cpu")


def test_train_step_has_no_cpu_fallback():
    """The step's losses are CUDA kernels: on a machine without a GPU the step fails loudly instead of computing them
    some other way (the network itself is plain torch and runs anywhere)."""
    from doubly_contrastive_semseg_b200 import _lib
    st = TrainStep(types.SimpleNamespace(amp=False, batch_size=1), device="cpu")
    g = torch.Generator().manual_seed(2)
    sample = {"left": torch.rand(2, 3, 64, 128, generator=g) * 255.0, "label": torch.randint(0, 19, (1, 64, 128), generator=g),
              "weather": torch.zeros(1, dtype=torch.long), "label_distance_weight": torch.rand(1, 64, 128, generator=g)}
    with pytest.raises(_lib.DclError):
        st(sample)
    with pytest.raises(ValueError):                                   # 'supcon' criteria need both crops
        st({**sample, "left": sample["left"][:1]})
