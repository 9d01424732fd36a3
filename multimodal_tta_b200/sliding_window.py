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
the r-th block of ``sw_batch``; the tail batch (single rank too: the plan has a fixed batch of
``sw_batch`` windows) is padded with zero-weight windows, the loss/gradient are rescaled to the
mean over the REAL windows (``n_valid_global``), and all ranks join the gradient all-reduce;
blended accumulators are all-reduced once per volume batch.  With BatchNorm the padding windows
(copies of window 0) still enter the batch statistics -- InstanceNorm, the reference's setting,
is unaffected.
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
        self.time_collectives = False        # bench: CUDA events around the accumulator all-reduce
        self.last_allreduce_ms: Optional[float] = None

    def _state(self, device, key):
        if key not in self._bufs:
            NB = self.sw_batch
            fs, wmin = gaussian_factors(self.roi, self.sigma_scale)
            self._bufs[key] = dict(
                win=torch.zeros((NB, 4), dtype=torch.int32, device=device),
                sw=torch.ones(NB, dtype=torch.float32, device=device),
                g=[f.to(device) for f in fs], wmin=wmin)
        return self._bufs[key]

    def _schedule(self, device, B, C, vd, world, rank, chan_scale_per_volume):
        """The whole window schedule of one call, uploaded ONCE: per step the gather origins (unpadded
        volume coordinates), the blend origins (padded coordinates), the window weights (0 = padding) and
        the per-(window, channel) modality scales.  The per-step loop then only issues stream-ordered
        device copies into the persistent buffers the CUDA graph reads -- no pinned buffer is reused
        while a copy may still be in flight, no host synchronisation per window batch."""
        padded, pad_lo, starts = plan_windows(vd, self.roi, self.overlap)
        nwin, NB = len(starts), self.sw_batch
        total = nwin * B
        rows_g, rows_b, rows_w, n_valid = [], [], [], []
        for idxs, nv in shard_schedule(total, NB, world, rank):
            for idx in idxs:
                valid = idx is not None
                b, s = (idx // nwin, starts[idx % nwin]) if valid else (0, starts[0])
                rows_g.append([b, s[0] - pad_lo[0], s[1] - pad_lo[1], s[2] - pad_lo[2]])
                rows_b.append([b, s[0], s[1], s[2]])
                rows_w.append(1.0 if valid else 0.0)
            n_valid.append(nv)
        steps = len(n_valid)
        win_g = torch.tensor(rows_g, dtype=torch.int32).view(steps, NB, 4)
        cs = None
        if chan_scale_per_volume is not None:
            cs = chan_scale_per_volume.detach().to("cpu", torch.float32)[win_g[..., 0].long()].contiguous()
            cs = cs.to(device)                                             # [steps][NB][C]
        return dict(padded=padded, pad_lo=pad_lo, total=total, steps=steps, n_valid=n_valid,
                    win_g=win_g.to(device), win_b=torch.tensor(rows_b, dtype=torch.int32).view(steps, NB, 4).to(device),
                    sw=torch.tensor(rows_w, dtype=torch.float32).view(steps, NB).to(device), cs=cs)

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
        eng._ensure_device(vol.device)
        with eng.on_device():
            return self._sweep(vol, chan_scale_per_volume, intensity_policy, dist)

    def _sweep(self, vol, chan_scale_per_volume, intensity_policy, dist):
        tent = self.tent
        B, C = int(vol.shape[0]), int(vol.shape[1])
        vd = [int(s) for s in vol.shape[2:]]
        ws = tent.world_size
        rank = dist.get_rank(tent.pg) if ws > 1 else 0
        NB, R = self.sw_batch, tent.model.out_channels
        dev = vol.device
        st = self._state(dev, (dev, NB))
        sched = self._schedule(dev, B, C, vd, ws, rank, chan_scale_per_volume)
        padded, pad_lo = sched["padded"], sched["pad_lo"]
        # the captured step bakes the volume pointer in: keep ONE resident volume buffer per shape and copy
        # the caller's tensor into it (143 MB for a BraTS volume, ~50 us) instead of re-capturing per volume
        vkey = ("vol", dev, tuple(vol.shape), vol.dtype)
        if vkey not in self._bufs:
            self._bufs[vkey] = torch.empty_like(vol)
        if vol.data_ptr() != self._bufs[vkey].data_ptr():
            self._bufs[vkey].copy_(vol, non_blocking=True)
        vol = self._bufs[vkey]
        acc = torch.zeros((B, R, *padded), dtype=torch.float32, device=dev)
        wsum = torch.zeros((B, *padded), dtype=torch.float32, device=dev)
        affine = None
        if intensity_policy is not None:
            akey = ("affine", dev, B, C)
            if akey not in self._bufs:
                self._bufs[akey] = torch.empty((B, C, 4), dtype=torch.float32, device=dev)
            affine = intensity_policy.stats(vol, out=self._bufs[akey])
        cs_dev = None
        if sched["cs"] is not None:
            ckey = ("cs", dev, NB, C)
            if ckey not in self._bufs:
                self._bufs[ckey] = torch.ones((NB, C), dtype=torch.float32, device=dev)
            cs_dev = self._bufs[ckey]
        stream = torch.cuda.current_stream(dev).cuda_stream
        for t in range(sched["steps"]):
            st["win"].copy_(sched["win_g"][t], non_blocking=True)
            st["sw"].copy_(sched["sw"][t], non_blocking=True)
            if cs_dev is not None:
                cs_dev.copy_(sched["cs"][t], non_blocking=True)
            logits = tent.step_windows(vol, st["win"], self.roi, sample_w=st["sw"], chan_scale=cs_dev,
                                       n_valid_global=sched["n_valid"][t], affine=affine)
            check(self.lib.tta_sw_blend(logits.data_ptr(), NB, R, *self.roi, sched["win_b"][t].data_ptr(),
                                        st["sw"].data_ptr(), st["g"][0].data_ptr(), st["g"][1].data_ptr(),
                                        st["g"][2].data_ptr(), float(st["wmin"]), acc.data_ptr(),
                                        wsum.data_ptr(), B, *padded, stream), "sw_blend")
        if ws > 1:
            ev = None
            if self.time_collectives:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            dist.all_reduce(acc, group=tent.pg)
            dist.all_reduce(wsum, group=tent.pg)
            if ev is not None:
                ev[1].record()
                ev[1].synchronize()
                self.last_allreduce_ms = ev[0].elapsed_time(ev[1])
        out = torch.empty_like(acc)
        Vs = padded[0] * padded[1] * padded[2]
        check(self.lib.tta_sw_normalise(acc.data_ptr(), wsum.data_ptr(), B, R, Vs, out.data_ptr(), stream),
              "sw_normalise")
        self.last_num_windows, self.last_steps = sched["total"], sched["steps"]
        if padded != vd:
            out = out[:, :, pad_lo[0]:pad_lo[0] + vd[0], pad_lo[1]:pad_lo[1] + vd[1],
                      pad_lo[2]:pad_lo[2] + vd[2]].contiguous()
        return out
