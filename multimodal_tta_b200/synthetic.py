"""Seeded synthetic inputs of the shapes BASELINE.json names: there are no
dataset files in the container, so every benchmark/test volume is generated here."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def ellipsoid_mask(D: int, H: int, W: int, scale: float = 0.85) -> torch.Tensor:
    z = torch.linspace(-1, 1, D).view(D, 1, 1)
    y = torch.linspace(-1, 1, H).view(1, H, 1)
    x = torch.linspace(-1, 1, W).view(1, 1, W)
    return ((z / scale) ** 2 + (y / scale) ** 2 + (x / scale) ** 2) <= 1.0


def brats_volume(B: int, dims: Tuple[int, int, int], seed: int = 42, channels: int = 4) -> torch.Tensor:
    """[B,4,D,H,W]: per-channel N(0,1) inside an ellipsoidal "brain", 0 outside (BraTS background
    is exactly 0 -- /root/reference/src/datasets/brats.py:7)."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    x = torch.randn((B, channels, D, H, W), generator=g)
    return x * ellipsoid_mask(D, H, W).to(x.dtype)


def _masked_zscore(x: torch.Tensor, mask: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    v = x[mask]
    mu, sd = (v.mean(), v.std(unbiased=False)) if v.numel() > 0 else (x.mean(), x.std(unbiased=False))
    return (x - mu) / (sd + eps)


def hecktor_volume(B: int, dims: Tuple[int, int, int], seed: int = 42, p_drop: float = 0.5):
    """[B,2,D,H,W] CT/PET-like volume after the reference intensity policy
    (configs/_global_patches/hecktor21.yaml:33-46: clip then masked z-score) plus a per-volume
    modality keep mask [B,2] (with probability p_drop one uniformly chosen modality is zeroed)."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    out = torch.empty((B, 2, D, H, W))
    keep = torch.ones((B, 2))
    for b in range(B):
        ct = (torch.randn((D, H, W), generator=g) * 300 - 200).clamp(-1000, 1000)
        pt = (torch.empty((D, H, W)).exponential_(1.0 / 1.5, generator=g)).clamp(0, 15)
        out[b, 0] = _masked_zscore(ct, ct > -900)
        out[b, 1] = _masked_zscore(pt, pt > 0)
        if float(torch.rand((), generator=g)) < p_drop:
            keep[b, int(torch.randint(0, 2, (), generator=g))] = 0.0
    return out, keep


def region_labels(B: int, R: int, dims: Tuple[int, int, int], seed: int = 7) -> torch.Tensor:
    """Nested blob labels [B,R,D,H,W] in {0,1} (ET inside TC inside WT for R=3)."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = dims
    out = torch.zeros((B, R, D, H, W))
    for b in range(B):
        c = (torch.rand(3, generator=g) - 0.5) * 0.6
        z = torch.linspace(-1, 1, D).view(D, 1, 1) - c[0]
        y = torch.linspace(-1, 1, H).view(1, H, 1) - c[1]
        x = torch.linspace(-1, 1, W).view(1, 1, W) - c[2]
        r2 = z ** 2 + y ** 2 + x ** 2
        for r in range(R):
            out[b, r] = (r2 <= (0.15 + 0.12 * r) ** 2).float()
    return out


def domain_shift(x: torch.Tensor, domain: int, seed: int = 123) -> torch.Tensor:
    """Per-channel gain U(0.7,1.3), bias U(-0.3,0.3), noise sigma U(0,0.2) for a domain id."""
    g = torch.Generator().manual_seed(seed + 1000 * domain)
    C = x.shape[1]
    gain = 0.7 + 0.6 * torch.rand(C, generator=g)
    bias = -0.3 + 0.6 * torch.rand(C, generator=g)
    sig = 0.2 * torch.rand(C, generator=g)
    noise = torch.randn(x.shape, generator=g)
    v = lambda t: t.view(1, C, 1, 1, 1)
    return x * v(gain) + v(bias) + noise * v(sig)
