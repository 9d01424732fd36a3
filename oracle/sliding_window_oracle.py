"""ORACLE (test infrastructure only) -- sliding-window inference with Gaussian blending, CPU.

PARITY UNPINNED: the reference has no sliding-window code (it forwards the whole pre-cropped
volume, /root/reference/src/evaluation/seg_eval.py:300).  This restates MONAI's published
``sliding_window_inference(roi, sw_batch, overlap, mode="gaussian", sigma_scale=0.125)``
(SURVEY.md section 8c-5): scan interval int(roi*(1-overlap)), last window shifted back to fit,
separable Gaussian importance centred at (roi-1)/2 with sigma = 0.125*roi clamped from below
at max(min, 1e-3), out = sum(w*pred)/sum(w); symmetric constant pad when the image is smaller
than the roi.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import torch
import torch.nn.functional as F


def scan_interval(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> list[int]:
    out = []
    for i in range(len(roi)):
        if roi[i] == image_size[i]:
            out.append(int(roi[i]))
        else:
            iv = int(roi[i] * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return out


def window_starts(image_size: Sequence[int], roi: Sequence[int], interval: Sequence[int]) -> list[tuple[int, ...]]:
    """All window origins, last axis fastest (MONAI ``dense_patch_slices`` with ij meshgrid)."""
    per_dim = []
    for d in range(len(roi)):
        if interval[d] == 0:
            num = 1
        else:
            n = int(math.ceil(float(image_size[d]) / interval[d]))
            first = next((k for k in range(n) if k * interval[d] + roi[d] >= image_size[d]), None)
            num = first + 1 if first is not None else 1
        starts = []
        for k in range(num):
            s = k * interval[d]
            s -= max(s + roi[d] - image_size[d], 0)
            starts.append(s)
        per_dim.append(starts)
    out: list[tuple[int, ...]] = [()]
    for starts in per_dim:
        out = [o + (s,) for o in out for s in starts]
    return out


def gaussian_importance(roi: Sequence[int], sigma_scale: float = 0.125) -> torch.Tensor:
    imp = None
    for i, n in enumerate(roi):
        sigma = n * sigma_scale
        x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float)
        g = torch.exp(x ** 2 / (-2 * sigma ** 2))
        imp = g if imp is None else imp.unsqueeze(-1) * g[(None,) * i]
    mn = max(float(imp.min().item()), 1e-3)
    return imp.clamp_(min=mn).float()


def sliding_window_oracle(inputs: torch.Tensor, roi: Sequence[int], sw_batch: int,
                          predictor: Callable[[torch.Tensor], torch.Tensor], overlap: float = 0.5,
                          sigma_scale: float = 0.125) -> torch.Tensor:
    """inputs [B,C,D,H,W] -> blended predictor output [B,R,D,H,W].  ``predictor`` is called
    on window batches in scan order (so a TENT predictor adapts window-batch by window-batch)."""
    B = inputs.shape[0]
    orig = list(inputs.shape[2:])
    roi = [int(r) for r in roi]
    pads = []
    for k in range(4, 1, -1):
        diff = max(roi[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pads.extend([half, diff - half])
    x = F.pad(inputs, pads, mode="constant", value=0.0)
    size = list(x.shape[2:])
    iv = scan_interval(size, roi, overlap)
    starts = window_starts(size, roi, iv)
    nwin = len(starts)
    imp = gaussian_importance(roi, sigma_scale)
    out = cnt = None
    total = nwin * B
    for g in range(0, total, sw_batch):
        idxs = list(range(g, min(g + sw_batch, total)))
        wins = []
        for idx in idxs:
            b, s = idx // nwin, starts[idx % nwin]
            wins.append(x[b:b + 1, :, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]])
        pred = predictor(torch.cat(wins, 0))
        if out is None:
            out = torch.zeros((B, pred.shape[1], *size), dtype=pred.dtype)
            cnt = torch.zeros((B, 1, *size), dtype=pred.dtype)
        for j, idx in enumerate(idxs):
            b, s = idx // nwin, starts[idx % nwin]
            sl = (slice(b, b + 1), slice(None), slice(s[0], s[0] + roi[0]),
                  slice(s[1], s[1] + roi[1]), slice(s[2], s[2] + roi[2]))
            out[sl] += imp * pred[j:j + 1]
            cnt[sl] += imp
    out = out / cnt
    # crop the symmetric pad back off
    sl = [slice(None), slice(None)]
    for d in range(3):
        diff = max(roi[d] - orig[d], 0)
        half = diff // 2
        sl.append(slice(half, half + orig[d]))
    return out[tuple(sl)]
