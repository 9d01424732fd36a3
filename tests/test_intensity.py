"""Input intensity policy: the oracle restatement against golden vectors produced by
the reference's own `_normalize_img` (tests/golden/gen_intensity_golden.py), the host rule table, and --
on the GPU -- the CUDA kernels against the same golden vectors."""
import os

import numpy as np
import pytest
import torch

from tests.golden.gen_intensity_golden import HECKTOR, MIXED

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "intensity_golden.npz"))
CASES = {"hecktor": dict(intensity_policy=HECKTOR), "mixed": dict(intensity_policy=MIXED),
         "legacy": dict(mean=[0.5, -1.0, 2.0], std=[2.0, 0.5, 4.0]), "identity": dict()}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_golden(name):
    from oracle.intensity_oracle import normalize_img
    got = normalize_img(torch.from_numpy(GOLD[f"{name}_in"]), **CASES[name])
    assert torch.equal(got, torch.from_numpy(GOLD[f"{name}_out"]))      # same torch ops in the same order: bit exact


def test_rule_table_and_errors():
    from multimodal_tta_b200.intensity import IntensityPolicy
    r, mc = IntensityPolicy(HECKTOR).rules(2)
    assert mc == 16
    assert r[0].tolist() == [1.0, -1000.0, 1000.0, 1.0, -900.0, pytest.approx(1e-6), 0.0, 1.0]
    assert r[1].tolist()[:5] == [1.0, 0.0, 15.0, 1.0, 0.0]
    r, _ = IntensityPolicy(MIXED).rules(4)
    assert r[1, 3] == 2.0 and r[2, 3] == 0.0 and r[2, 0] == 1.0 and r[0, 0] == 0.0
    r, _ = IntensityPolicy(None, mean=[0.5, -1.0, 2.0], std=[2.0, 0.5, 4.0]).rules(3)
    assert r[:, 3].tolist() == [3.0, 3.0, 3.0] and r[:, 6].tolist() == [0.5, -1.0, 2.0]
    with pytest.raises(RuntimeError):       # reference: len(channel_names) != C (transforms.py:154-158)
        IntensityPolicy(HECKTOR).rules(3)
    with pytest.raises(RuntimeError):       # CUDA only, no CPU fallback
        IntensityPolicy(HECKTOR)(torch.zeros(2, 4, 4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_intensity_policy_matches_reference_golden(cuda, name):
    """Tolerance: the reference reduces in fp32 (torch mean/std), the kernel in fp64 and multiplies by
    1/sd instead of dividing: <= 2e-6 relative to the largest normalised value, stated here."""
    from multimodal_tta_b200.intensity import IntensityPolicy
    x = torch.from_numpy(GOLD[f"{name}_in"])
    ref = torch.from_numpy(GOLD[f"{name}_out"])
    pol = IntensityPolicy(CASES[name].get("intensity_policy"), mean=CASES[name].get("mean"), std=CASES[name].get("std"))
    got = pol(x.to(cuda)).cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max()))
    # a batch of two different volumes in one launch, second call (self-resetting counters), in place
    xb = torch.stack([x, x.flip(-1) * 0.5 + 0.1]).to(cuda).contiguous()
    from oracle.intensity_oracle import normalize_img
    refb = torch.stack([normalize_img(v.cpu(), **CASES[name]) for v in xb])
    for _ in range(2):
        work = xb.clone()
        pol(work, out=work)
        assert float((work.cpu() - refb).abs().max()) <= 2e-6 * max(1.0, float(refb.abs().max()))


@pytest.mark.gpu
def test_cuda_intensity_policy_large_volume(cuda):
    """HECKTOR-shaped 2 x 144^3 (BASELINE config 3 input): many blocks per slab, odd split sizes."""
    from multimodal_tta_b200.intensity import IntensityPolicy
    from oracle.intensity_oracle import normalize_img
    g = torch.Generator().manual_seed(3)
    dims = (72, 72, 70)
    ct = torch.clamp(torch.randn(dims, generator=g) * 300 - 200, -1500, 1500)
    pt = torch.clamp(torch.empty(dims).exponential_(1.5, generator=g), 0, 25)
    pt[:, :20] = 0
    x = torch.stack([ct, pt])
    ref = normalize_img(x, intensity_policy=HECKTOR)
    got = IntensityPolicy(HECKTOR)(x.to(cuda)).cpu()
    assert float((got - ref).abs().max()) <= 5e-6 * float(ref.abs().max())


@pytest.mark.gpu
def test_gather_with_fused_intensity_policy_is_bit_identical(cuda):
    """tta_gather_pack_norm(raw volume, affine) == tta_gather_pack(tta_intensity_apply(raw volume)): window origins
    inside / partly outside the volume, modality dropout scale, both the 4-wide and the scalar kernel."""
    from multimodal_tta_b200._lib import check, lib as load
    from multimodal_tta_b200.intensity import IntensityPolicy
    lib = load()
    st = torch.cuda.current_stream().cuda_stream
    x = torch.from_numpy(GOLD["hecktor_in"])
    raw = torch.stack([x, x.flip(-1) * 0.7 + 3.0]).to(cuda).contiguous()          # [2, 2, 10, 12, 14]
    pol = IntensityPolicy(HECKTOR)
    affine = pol.stats(raw)
    norm = pol(raw)
    wins = torch.tensor([[0, 0, 0, 0], [1, 2, 4, 6], [1, -3, -2, -5]], dtype=torch.int32, device=cuda)
    scale = torch.tensor([[1., 1.], [1., 0.], [0.5, 1.]], device=cuda)
    for roi in ((8, 8, 8), (8, 8, 6)):
        V = roi[0] * roi[1] * roi[2]
        outs = []
        for vol, aff in ((raw, affine), (norm, None)):
            hi = torch.zeros((3, 1, *roi, 8), dtype=torch.int16, device=cuda); lo = torch.zeros_like(hi)
            check(lib.tta_gather_pack_norm(vol.data_ptr(), 2, 2, 10, 12, 14, wins.data_ptr(), scale.data_ptr(),
                                           aff.data_ptr() if aff is not None else 0, 3, *roi, hi.data_ptr(),
                                           lo.data_ptr(), V * 8, 1, 0, st), "gather_pack_norm")
            outs.append((hi, lo))
        torch.cuda.synchronize()
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert int(outs[0][0].abs().max()) > 0
