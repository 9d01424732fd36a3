"""On-device sliding-window TTA with Gaussian blending (north-star feature; the reference itself
forwards whole pre-cropped volumes -- /root/reference/src/evaluation/seg_eval.py:300 -- and cannot
run sizes not divisible by 16).

Tiling follows MONAI's published ``sliding_window_inference``: scan interval
int(roi*(1-overlap)), last window shifted back to fit, windows enumerated volume-major with the
last spatial axis fastest, symmetric zero pad when the image is smaller than the roi, separable
Gaussian importance sigma = 0.125*roi clamped at max(min, 1e-3), out = sum(w*pred)/sum(w).

Patches are gathered straight from the resident volume into the conv operand layout
(tta_gather_pack), predictions are blended by a deterministic gather-form kernel (tta_sw_blend).
Multi-GPU: every global step adapts on ``sw_batch * world`` consecutive windows, rank r taking
the r-th block of ``sw_batch``; the tail is padded with zero-weight windows so all ranks join the
gradient all-reduce; blended accumulators are all-reduced once per volume batch.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check
from .tent import TentB200


def scan_interval(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> List[int]:
    out = []
    for n, r in zip(image_size, roi):
        if r == n:
            out.append(int(r))
        else:
            iv = int(r * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return out


def window_starts(image_size: Sequence[int], roi: Sequence[int], interval: Sequence[int]) -> List[Tuple[int, ...]]:
    per_dim = []
    for n, r, iv in zip(image_size, roi, interval):
        if iv == 0:
            num = 1
        else:
            cnt = int(math.ceil(float(n) / iv))
            first = next((k for k in range(cnt) if k * iv + r >= n), None)
            num = first + 1 if first is not None else 1
        per_dim.append([k * iv - max(k * iv + r - n, 0) for k in range(num)])
    out: List[Tuple[int, ...]] = [()]
    for starts in per_dim:
        out = [o + (s,) for o in out for s in starts]
    return out


def gaussian_factors(roi: Sequence[int], sigma_scale: float = 0.125) -> Tuple[List[torch.Tensor], float]:
    """Separable 1-D factors and the clamp floor max(min(product), 1e-3)."""
    fs, mn = [], 1.0
    for n in roi:
        sigma = n * sigma_scale
        x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float)
        g = torch.exp(x ** 2 / (-2 * sigma ** 2))
        fs.append(g)
        mn *= float(g.min().item())
    return fs, max(mn, 1e-3)


def plan_windows(vol_dims: Sequence[int], roi: Sequence[int], overlap: float):
    """-> (padded dims, pad_lo per axis, window origins in padded coordinates)."""
    pad_lo, padded = [], []
    for n, r in zip(vol_dims, roi):
        diff = max(r - n, 0)
        pad_lo.append(diff // 2)
        padded.append(max(n, r))
    iv = scan_interval(padded, roi, overlap)
    return padded, pad_lo, window_starts(padded, roi, iv)


def shard_schedule(total: int, sw_batch: int, world: int, rank: int):
    """Window indices this rank adapts on at every global step: step t covers the ``sw_batch*world``
    consecutive windows [t*G, (t+1)*G), rank r takes the r-th block of ``sw_batch``; indices past
    ``total`` are None (zero-weight padding so that every rank joins every all-reduce).
    Yields (list of index-or-None of length sw_batch, number of valid windows in the global step)."""
    G = sw_batch * world
    for g0 in range(0, total, G):
        idxs = []
        for j in range(sw_batch):
            i = g0 + rank * sw_batch + j
            idxs.append(i if i < total else None)
        yield idxs, min(G, total - g0)


class SlidingWindowTTA:
    def __init__(self, tent: TentB200, roi: Sequence[int], sw_batch: int = 1, overlap: float = 0.5,
                 sigma_scale: float = 0.125):
        self.tent, self.roi = tent, [int(r) for r in roi]
        self.sw_batch, self.overlap, self.sigma_scale = int(sw_batch), float(overlap), float(sigma_scale)
        self._bufs = {}
        self.lib = _lib.lib()
        self.last_num_windows = 0
        self.last_steps = 0

    def _state(self, device, key):
        if key not in self._bufs:
            NB = self.sw_batch
            fs, wmin = gaussian_factors(self.roi, self.sigma_scale)
            self._bufs[key] = dict(
                win=torch.zeros((NB, 4), dtype=torch.int32, device=device),
                win_host=torch.zeros((NB, 4), dtype=torch.int32).pin_memory(),
                sw=torch.ones(NB, dtype=torch.float32, device=device),
                sw_host=torch.ones(NB, dtype=torch.float32).pin_memory(),
                g=[f.to(device) for f in fs], wmin=wmin)
        return self._bufs[key]

    @torch.no_grad()
    def __call__(self, vol: torch.Tensor, chan_scale_per_volume: Optional[torch.Tensor] = None,
                 intensity_policy=None) -> torch.Tensor:
        """vol [B,C,Ds,Hs,Ws] fp32 cuda -> blended logits [B,R,Ds,Hs,Ws]; adapts window-batch by
        window-batch (non-episodic unless the TentB200 is episodic).  ``intensity_policy``
        (``IntensityPolicy``): ``vol`` holds RAW intensities; one statistics pass over the volume, then
        clip + z-score are applied while the windows are gathered (no normalised copy of the volume)."""
        import torch.distributed as dist

        tent = self.tent
        eng = tent.model.engine
        vol = eng._check_input(vol)
        B, C = int(vol.shape[0]), int(vol.shape[1])
        vd = [int(s) for s in vol.shape[2:]]
        padded, pad_lo, starts = plan_windows(vd, self.roi, self.overlap)
        nwin = len(starts)
        total = nwin * B
        ws = tent.world_size
        rank = dist.get_rank(tent.pg) if ws > 1 else 0
        NB, R = self.sw_batch, tent.model.out_channels
        st = self._state(vol.device, (vol.device, NB))
        acc = torch.zeros((B, R, *padded), dtype=torch.float32, device=vol.device)
        wsum = torch.zeros((B, *padded), dtype=torch.float32, device=vol.device)
        affine = None
        if intensity_policy is not None:
            akey = ("affine", vol.device, B, C)
            if akey not in self._bufs:
                self._bufs[akey] = torch.empty((B, C, 4), dtype=torch.float32, device=vol.device)
            affine = intensity_policy.stats(vol, out=self._bufs[akey])
        cs_dev = None
        if chan_scale_per_volume is not None:
            cs_dev = torch.ones((NB, C), dtype=torch.float32, device=vol.device)
        steps = 0
        for idxs, n_valid in shard_schedule(total, NB, ws, rank):
            for j, idx in enumerate(idxs):
                valid = idx is not None
                b, s = (idx // nwin, starts[idx % nwin]) if valid else (0, starts[0])
                # origins in UNPADDED volume coordinates (gather zero-fills outside the volume)
                st["win_host"][j] = torch.tensor([b, s[0] - pad_lo[0], s[1] - pad_lo[1], s[2] - pad_lo[2]],
                                                 dtype=torch.int32)
                st["sw_host"][j] = 1.0 if valid else 0.0
            st["win"].copy_(st["win_host"], non_blocking=True)
            st["sw"].copy_(st["sw_host"], non_blocking=True)
            if cs_dev is not None:
                cs_dev.copy_(chan_scale_per_volume.to(vol.device)[st["win_host"][:, 0].long()])
            logits = tent.step_windows(vol, st["win"], self.roi, sample_w=st["sw"], chan_scale=cs_dev,
                                       n_valid_global=n_valid if ws > 1 else None, affine=affine)
            # blend needs origins in PADDED coordinates
            winp = st["win"].clone()
            winp[:, 1:] += torch.tensor(pad_lo, dtype=torch.int32, device=vol.device)
            check(self.lib.tta_sw_blend(logits.data_ptr(), NB, R, *self.roi, winp.data_ptr(),
                                        st["sw"].data_ptr(), st["g"][0].data_ptr(), st["g"][1].data_ptr(),
                                        st["g"][2].data_ptr(), float(st["wmin"]), acc.data_ptr(),
                                        wsum.data_ptr(), B, *padded,
                                        torch.cuda.current_stream().cuda_stream), "sw_blend")
            steps += 1
        if ws > 1:
            dist.all_reduce(acc, group=tent.pg)
            dist.all_reduce(wsum, group=tent.pg)
        out = torch.empty_like(acc)
        Vs = padded[0] * padded[1] * padded[2]
        check(self.lib.tta_sw_normalise(acc.data_ptr(), wsum.data_ptr(), B, R, Vs, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream), "sw_normalise")
        self.last_num_windows, self.last_steps = total, steps
        if padded != vd:
            out = out[:, :, pad_lo[0]:pad_lo[0] + vd[0], pad_lo[1]:pad_lo[1] + vd[1],
                      pad_lo[2]:pad_lo[2] + vd[2]].contiguous()
        return out
