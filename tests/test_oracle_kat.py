"""Known-answer tests that pin the CPU oracle: MONAI's published parameter
count, the reference configs' counts, the state-dict key list, closed-form gradients vs autograd,
Adam vs torch.optim.Adam, Dice vs outputs of the reference's own _binary_dice_iou."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.dice_oracle import binary_dice_iou, evaluate_logits
from oracle.sliding_window_oracle import gaussian_importance, scan_interval, sliding_window_oracle, window_starts
from oracle.tent_oracle import (TentOracle, adam_reference, bernoulli_entropy, collect_params, configure_model,
                                entropy_loss, softmax_entropy)
from oracle.unet_oracle import (BARE_DEFAULT_MODEL_CFG, BRATS_MODEL_CFG, HECKTOR_MODEL_CFG, OracleUNet, norm_modules)

HERE = os.path.dirname(os.path.abspath(__file__))

APPENDIX_A_FIRST = ["model.0.conv.unit0.conv.weight", "model.0.conv.unit0.conv.bias", "model.0.conv.unit1.conv.weight",
                    "model.0.conv.unit1.conv.bias", "model.0.residual.weight", "model.0.residual.bias",
                    "model.1.submodule.0.conv.unit0.conv.weight"]
APPENDIX_A_LAST = ["model.2.0.conv.weight", "model.2.0.conv.bias", "model.2.1.conv.unit0.conv.weight",
                   "model.2.1.conv.unit0.conv.bias"]


def test_monai_published_parameter_count():
    # MONAI spleen tutorial: UNet(3, 1, 2, (16,32,64,128,256), (2,2,2,2), num_res_units=2, norm=BATCH)
    # has 4 808 917 parameters with MONAI's default PReLU activation = 17 PReLU scalars on top of
    # the conv/norm parameters restated here with ReLU (the reference passes act="relu").
    m = OracleUNet(1, 2, (16, 32, 64, 128, 256), (2, 2, 2, 2), num_res_units=2, norm="BATCH")
    n_adn = len(norm_modules(m))
    assert n_adn == 17
    assert sum(p.numel() for p in m.parameters()) + n_adn == 4_808_917


def test_reference_config_parameter_counts_and_keys():
    m = OracleUNet.from_cfg(BRATS_MODEL_CFG)
    sd = m.state_dict()
    assert sum(p.numel() for p in m.parameters()) == 19_223_961 and len(sd) == 46
    keys = list(sd.keys())
    assert keys[:7] == APPENDIX_A_FIRST and keys[-4:] == APPENDIX_A_LAST
    assert sd["model.1.submodule.1.submodule.1.submodule.2.0.conv.weight"].shape == (768, 128, 3, 3, 3)
    assert sd["model.1.submodule.1.submodule.1.submodule.1.submodule.residual.weight"].shape == (512, 256, 1, 1, 1)
    assert [mm.num_features for _, mm in norm_modules(m)] == [32, 32, 64, 64, 128, 128, 256, 256, 512, 512, 128, 128,
                                                              64, 64, 32, 32, 3]
    bare = OracleUNet.from_cfg(BARE_DEFAULT_MODEL_CFG)
    assert sum(p.numel() for p in bare.parameters()) == 7_915_297
    assert "model.0.adn.N.running_mean" in bare.state_dict()


def test_forward_shape_and_divisibility():
    m = OracleUNet.from_cfg(HECKTOR_MODEL_CFG).eval()
    with torch.no_grad():
        assert m(torch.zeros(1, 2, 16, 32, 16)).shape == (1, 1, 16, 32, 16)
        with pytest.raises(RuntimeError):      # 24 is not divisible by 16: the skip concat mismatches
            m(torch.zeros(1, 2, 24, 16, 16))
    with pytest.raises(ValueError):
        OracleUNet.from_cfg(dict(BRATS_MODEL_CFG, in_channels="auto"))


def test_tent_configure_materialises_affine_and_freezes():
    m = configure_model(OracleUNet.from_cfg(BRATS_MODEL_CFG))
    params, names = collect_params(m)
    assert len(params) == 34 and sum(p.numel() for p in params) == 4870
    assert all(p.requires_grad for p in params)
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 4870
    bn = configure_model(OracleUNet.from_cfg(dict(BARE_DEFAULT_MODEL_CFG, in_channels=4)))
    assert all(mm.running_mean is None for _, mm in norm_modules(bn))
    assert sum(p.numel() for p in bn.parameters() if p.requires_grad) == 2432


def test_entropy_closed_form_gradients():
    torch.manual_seed(0)
    z = (torch.randn(2, 3, 4, 5, 6) * 3).requires_grad_(True)
    softmax_entropy(z).sum().backward()
    p = torch.softmax(z.detach(), 1)
    pz = (p * z.detach()).sum(1, keepdim=True)
    assert torch.allclose(z.grad, -p * (z.detach() - pz), atol=1e-6)
    z2 = (torch.randn(2, 1, 4, 5, 6) * 3).requires_grad_(True)
    bernoulli_entropy(z2).sum().backward()
    s = torch.sigmoid(z2.detach())
    assert torch.allclose(z2.grad, -z2.detach() * s * (1 - s), atol=1e-6)
    with pytest.raises(ValueError):
        entropy_loss(torch.zeros(1, 1, 2, 2, 2), "softmax")


def test_instance_norm_backward_formula():
    torch.manual_seed(1)
    y = torch.randn(2, 5, 4, 4, 4, requires_grad=True)
    g, b = torch.rand(5) + 0.5, torch.randn(5)
    gr, br = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    dz = torch.randn(2, 5, 4, 4, 4)
    F.instance_norm(y, weight=gr, bias=br, eps=1e-5).backward(dz)
    mu = y.detach().mean((2, 3, 4), keepdim=True)
    r = 1 / torch.sqrt(y.detach().var((2, 3, 4), unbiased=False, keepdim=True) + 1e-5)
    xh = (y.detach() - mu) * r
    M = 64
    s1, s2 = dz.sum((2, 3, 4), keepdim=True), (dz * xh).sum((2, 3, 4), keepdim=True)
    dy = g.view(1, 5, 1, 1, 1) * r / M * (M * dz - s1 - xh * s2)
    assert torch.allclose(dy, y.grad, atol=1e-5)
    assert torch.allclose(s2.sum(0).flatten(), gr.grad, atol=1e-4) and torch.allclose(s1.sum(0).flatten(), br.grad, atol=1e-4)


def test_adam_reference_matches_torch():
    torch.manual_seed(2)
    p0 = torch.randn(100)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-3)
    q, m, v = p0.clone(), torch.zeros(100), torch.zeros(100)
    for t in range(1, 6):
        g = torch.randn(100)
        p.grad = g.clone(); opt.step()
        adam_reference(q, g, m, v, t, 1e-3, 0.9, 0.999, 1e-8)
        assert torch.allclose(q, p.detach(), atol=1e-7)


def test_dice_oracle_equals_reference_outputs():
    gold = np.load(os.path.join(HERE, "golden", "dice_golden.npz"))
    for i in range(4):
        d, j, v = binary_dice_iou(torch.from_numpy(gold[f"pred{i}"]), torch.from_numpy(gold[f"gt{i}"]))
        assert torch.equal(d, torch.from_numpy(gold[f"dice{i}"]))
        assert torch.equal(j, torch.from_numpy(gold[f"iou{i}"]))
        assert torch.equal(v, torch.from_numpy(gold[f"valid{i}"]))
    # aggregation: GT-empty regions are skipped, avg over valid regions only
    logits = torch.full((1, 3, 4, 4, 4), 5.0)
    y = torch.zeros(1, 3, 4, 4, 4); y[:, 0] = 1
    out = evaluate_logits([logits], [y])
    assert out["et_dc"] == pytest.approx(1.0) and out["tc_dc"] == 0.0 and out["avg_dc"] == pytest.approx(1.0)


def test_sliding_window_tiling_known_answers():
    # BraTS full volume 240x240x155 (D,H,W = 155,240,240), roi 128^3: 2x3x3 = 18 windows at 0.25 and 0.5
    for ov in (0.25, 0.5):
        iv = scan_interval((155, 240, 240), (128, 128, 128), ov)
        assert len(window_starts((155, 240, 240), (128, 128, 128), iv)) == 18
    iv = scan_interval((144, 144, 144), (96, 96, 96), 0.5)
    st = window_starts((144, 144, 144), (96, 96, 96), iv)
    assert len(st) == 8 and st[0] == (0, 0, 0) and st[-1] == (48, 48, 48)
    imp = gaussian_importance((8, 8, 8))
    assert float(imp.min()) >= 1e-3 and float(imp.max()) <= 1.0 and imp[3, 3, 3] == imp[4, 4, 4]


def test_sliding_window_blend_is_exact_for_pointwise_predictor():
    torch.manual_seed(3)
    x = torch.randn(2, 2, 20, 30, 17)
    out = sliding_window_oracle(x, (16, 16, 16), 3, lambda w: w[:, :1] * 2 + 1, overlap=0.5)
    assert torch.allclose(out, x[:, :1] * 2 + 1, atol=1e-5)
    small = torch.randn(1, 1, 10, 12, 9)      # smaller than the roi: symmetric zero pad, cropped back
    out = sliding_window_oracle(small, (16, 16, 16), 1, lambda w: w + 0.5, overlap=0.25)
    assert out.shape == small.shape and torch.allclose(out, small + 0.5, atol=1e-6)


def test_tent_oracle_step_returns_pre_update_logits_and_moves_only_norm_params():
    torch.manual_seed(4)
    m = OracleUNet.from_cfg(HECKTOR_MODEL_CFG)
    conv_w = m.model[0].conv.unit0.conv.weight.detach().clone()
    t = TentOracle(m, mode="sigmoid")
    x = torch.randn(1, 2, 32, 32, 32)
    with torch.no_grad():
        before = t.model(x).clone()
    logits, loss = t.step(x)
    assert torch.allclose(logits, before, atol=1e-6) and loss > 0
    assert torch.equal(conv_w, m.model[0].conv.unit0.conv.weight)
    g = t.model.model[0].conv.unit0.adn.N.weight
    assert float((g - 1).abs().max()) == pytest.approx(1e-3, rel=1e-3)     # Adam's first step = lr*sign(g)
    t2 = TentOracle(OracleUNet.from_cfg(HECKTOR_MODEL_CFG), episodic=True)
    a, _ = t2.step(x); b, _ = t2.step(x)
    assert torch.equal(a, b)
