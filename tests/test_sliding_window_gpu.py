"""Sliding-window TTA: on-device gather / TENT step / Gaussian blend against the CPU oracle."""
import pytest
import torch

from multimodal_tta_b200 import SlidingWindowTTA, TentB200
from multimodal_tta_b200.synthetic import brats_volume, hecktor_volume
from oracle.sliding_window_oracle import sliding_window_oracle
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import BRATS_MODEL_CFG, HECKTOR_MODEL_CFG
from tests.util import make_pair, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dims,roi,swb", [((40, 48, 36), (32, 32, 32), 2), ((24, 40, 32), (32, 32, 32), 1)])
def test_sliding_window_tta_matches_oracle(cuda, dims, roi, swb):
    oracle, prod = make_pair(BRATS_MODEL_CFG, seed=31)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    vol = brats_volume(1, dims, seed=77)
    ref = sliding_window_oracle(vol, roi, swb, lambda w: to.step(w)[0], overlap=0.5)
    sw = SlidingWindowTTA(tp, roi, sw_batch=swb, overlap=0.5)
    got = sw(vol.cuda()).cpu()
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < 1e-3
    assert ((got >= 0) == (ref >= 0)).float().mean().item() >= 0.9999
    # parameters adapted through the same number of steps
    p_o, p_p = flat_gamma_beta(to.model), prod.engine.flat_params().cpu()
    assert (p_p - p_o).abs().median() < 1e-4          # north-star tolerance on the typical scalar


def test_missing_modality_dropout(cuda):
    oracle, prod = make_pair(HECKTOR_MODEL_CFG, seed=32)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": False})
    vol, keep = hecktor_volume(2, (32, 48, 32), seed=5, p_drop=1.0)
    ref = sliding_window_oracle(vol * keep.view(2, 2, 1, 1, 1), (32, 32, 32), 2, lambda w: to.step(w)[0], overlap=0.5)
    sw = SlidingWindowTTA(tp, (32, 32, 32), sw_batch=2, overlap=0.5)
    got = sw(vol.cuda(), chan_scale_per_volume=keep).cpu()
    assert rel_l2(got, ref) < 1e-3


def test_hecktor_pipeline_raw_intensities_to_blended_logits(cuda):
    """BASELINE config 3 end to end on the device: RAW CT (HU) / PET (SUV) values -> intensity policy
    (clip + masked z-score, configs/_global_patches/hecktor21.yaml:27-46) -> missing-modality dropout ->
    sliding-window TENT -> Gaussian blend, against the oracle pipeline (oracle.intensity_oracle is pinned
    to the reference's own transform).  Same tolerances as the other sliding-window cases."""
    from multimodal_tta_b200 import IntensityPolicy
    from oracle.intensity_oracle import normalize_img
    from tests.golden.gen_intensity_golden import HECKTOR
    oracle, prod = make_pair(HECKTOR_MODEL_CFG, seed=33)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    g = torch.Generator().manual_seed(9)
    dims = (32, 48, 40)
    raw = torch.stack([torch.stack([(torch.randn(dims, generator=g) * 300 - 200).clamp(-1500, 1500),
                                    torch.empty(dims).exponential_(1.0 / 1.5, generator=g).clamp(0, 25)])
                       for _ in range(2)])
    keep = torch.tensor([[1.0, 0.0], [1.0, 1.0]])                      # volume 0 lost its PET
    norm_ref = torch.stack([normalize_img(v, intensity_policy=HECKTOR) for v in raw])
    ref = sliding_window_oracle(norm_ref * keep.view(2, 2, 1, 1, 1), (32, 32, 32), 2, lambda w: to.step(w)[0],
                                overlap=0.5)
    vol = IntensityPolicy(HECKTOR)(raw.cuda())
    sw = SlidingWindowTTA(tp, (32, 32, 32), sw_batch=2, overlap=0.5)
    got = sw(vol, chan_scale_per_volume=keep).cpu()
    assert rel_l2(got, ref) < 1e-3
    assert ((got >= 0) == (ref >= 0)).float().mean().item() >= 0.9999
    # fused variant: raw intensities in, clip + z-score applied inside the window gather.  The operand planes
    # are bit-identical to the two-pass path (tests/test_intensity.py); the logits agree to the run-to-run
    # noise of the split-K float atomics in the deep layers.
    _, prod2 = make_pair(HECKTOR_MODEL_CFG, seed=33)
    tp2 = TentB200(prod2, {"cuda_graph": True})
    sw2 = SlidingWindowTTA(tp2, (32, 32, 32), sw_batch=2, overlap=0.5)
    got2 = sw2(raw.cuda(), chan_scale_per_volume=keep, intensity_policy=IntensityPolicy(HECKTOR)).cpu()
    assert rel_l2(got2, got) < 5e-4, rel_l2(got2, got)      # measured 3e-5 (atomics order x several Adam steps)
