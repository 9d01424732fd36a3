"""Generate tests/golden/intensity_golden.npz by running the REFERENCE's own intensity policy
(`_build_3d_seg_transforms(...)._apply`, src/datasets/transforms.py:44-300) in this container.

    python tests/golden/gen_intensity_golden.py          # needs /root/reference (not on the GPU box)

`src/datasets/transforms.py` imports `monai.transforms` at module level (not installed here); stub
classes are pre-seeded -- with split="val" no augmentation object is ever built or called, so the
stubs do not influence the numbers.  omegaconf is absent, which the reference handles itself
(transforms.py:15-21).  Cases: the HECKTOR policy (configs/_global_patches/hecktor21.yaml:27-46) on
CT/PET-shaped values, a channel whose mask keeps fewer than min_count voxels (falls back to all
voxels), an unmasked z-score without clip, a channel without any rule, and the legacy mean/std branch.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

HECKTOR = {"enabled": True, "channel_names": ["ct", "pt"],
           "channels": {"ct": {"clip": [-1000, 1000], "zscore": {"masked": True, "mask_gt": -900, "eps": 1.0e-6}},
                        "pt": {"clip": [0.0, 15.0], "zscore": {"masked": True, "mask_gt": 0.0, "eps": 1.0e-6}}}}
MIXED = {"enabled": True, "channel_names": ["a", "b", "c", "d"],
         "channels": {"a": {"zscore": {"masked": True, "mask_gt": 5.0, "eps": 1.0e-6}},       # < 16 voxels pass
                      "b": {"zscore": {"masked": False, "eps": 1.0e-6}},
                      "c": {"clip": [-0.5, 0.5]},
                      "d": {"clip": [-2.0, 2.0], "zscore": {"masked": True, "mask_gt": 0.0, "eps": 10.0}}}}


def cases():
    g = torch.Generator().manual_seed(42)
    dims = (10, 12, 14)
    ct = torch.clamp(torch.randn(dims, generator=g) * 300 - 200, -1500, 1500)
    ct[:3] = -1024.0                                                    # air
    pt = torch.clamp(torch.empty(dims).exponential_(1.5, generator=g), 0, 25)
    pt[:, :4] = 0.0                                                     # background
    out = {"hecktor": (torch.stack([ct, pt]), dict(intensity_policy=HECKTOR))}
    x = torch.randn((4, 6, 7, 9), generator=g)
    x[0, 0, 0, :5] = 7.0                                                # 5 voxels above mask_gt = 5
    out["mixed"] = (x, dict(intensity_policy=MIXED))
    out["legacy"] = (torch.randn((3, 5, 6, 8), generator=g) * 2 + 1, dict(mean=[0.5, -1.0, 2.0], std=[2.0, 0.5, 4.0]))
    out["identity"] = (torch.randn((2, 4, 4, 4), generator=g), dict())
    return out


def main():
    sys.path.insert(0, REF)
    mt = types.ModuleType("monai.transforms")
    for n in ("Compose", "RandAxisFlipd", "RandRotate90d", "RandScaleIntensity", "RandShiftIntensity"):
        setattr(mt, n, type(n, (), {"__init__": lambda self, *a, **k: None}))
    sys.modules["monai"] = types.ModuleType("monai")
    sys.modules["monai.transforms"] = mt
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_transforms", os.path.join(REF, "src/datasets/transforms.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    blobs = {}
    for name, (img, kw) in cases().items():
        fn = mod._build_3d_seg_transforms("val", normalize=True, geom_aug=False, intensity_aug=False, **kw)
        label = torch.zeros(img.shape[1:], dtype=torch.long)
        got, _ = fn(img.clone(), label)
        blobs[f"{name}_in"] = img.numpy()
        blobs[f"{name}_out"] = got.numpy()
        print(name, tuple(img.shape), float(got.mean()), float(got.std()))
    np.savez_compressed(os.path.join(HERE, "intensity_golden.npz"), **blobs)


if __name__ == "__main__":
    main()
