"""On-device input intensity policy (per-channel clip + masked z-score) -- the step immediately in
front of the TTA hot path, SURVEY.md section 8(f) row 3.

Mirrors the configuration the reference's dataset transform reads
(/root/reference/src/datasets/transforms.py:118-127,147-217 and
configs/_global_patches/hecktor21.yaml:27-46):

    intensity_policy:
      enabled: true
      channel_names: ["ct", "pt"]
      channels:
        ct: {clip: [-1000, 1000], zscore: {masked: true, mask_gt: -900, eps: 1.0e-6}}
        pt: {clip: [0.0, 15.0],   zscore: {masked: true, mask_gt: 0.0,  eps: 1.0e-6}}
    mean: [...]; std: [...]          # legacy branch when the policy is disabled

The reference normalises on the CPU, one channel at a time with a host sync per channel; here a
batch of volumes is reduced in ONE streaming pass (fp64 partial sums, deterministic finalize) and
normalised in a second (csrc/tta_intensity.cu).  CUDA only: no CPU fallback.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import torch

from ._lib import check, lib as _load_lib


def _plain(x) -> Dict[str, Any]:
    if x is None:
        return {}
    if isinstance(x, dict):
        return {k: (_plain(v) if hasattr(v, "items") else v) for k, v in x.items()}
    return {k: (_plain(v) if hasattr(v, "items") else v) for k, v in dict(x).items()}


class IntensityPolicy:
    def __init__(self, intensity_policy=None, channel_names: Optional[Sequence[str]] = None,
                 mean: Optional[Sequence[float]] = None, std: Optional[Sequence[float]] = None):
        self.ip = _plain(intensity_policy)
        self.enabled = bool(self.ip.get("enabled", False))
        if channel_names is None:
            cn = self.ip.get("channel_names", None)
            channel_names = [str(x) for x in cn] if cn else None
        self.channel_names = None if channel_names is None else [str(x) for x in channel_names]
        self.mean, self.std = mean, std
        self._rules: Dict[Any, torch.Tensor] = {}
        self._ws: Dict[Any, torch.Tensor] = {}

    # ------------------------------------------------------------------ rule table [C][8]
    def rules(self, C: int) -> tuple[torch.Tensor, int]:
        """Host rule table + min_count (transforms.py:147-193 / 196-214)."""
        r = torch.zeros((C, 8), dtype=torch.float32)
        r[:, 7] = 1.0
        min_count = 16
        if self.enabled:
            names = [str(i) for i in range(C)] if self.channel_names is None else self.channel_names
            if len(names) != C:
                raise RuntimeError(f"[3DTransforms] len(channel_names)={len(names)} != C={C}. Please set "
                                   "dataset.modality_order (or transforms.channel_names) to match channels.")
            chans = self.ip.get("channels", {})
            chans = chans if isinstance(chans, dict) else {}
            counts = set()
            for ci, name in enumerate(names):
                rule = chans.get(name, {})
                rule = rule if isinstance(rule, dict) else {}
                clip = rule.get("clip", None)
                if isinstance(clip, (list, tuple)) and len(clip) == 2:
                    r[ci, 0], r[ci, 1], r[ci, 2] = 1.0, float(clip[0]), float(clip[1])
                zc = rule.get("zscore", None)
                if isinstance(zc, dict):
                    r[ci, 3] = 1.0 if bool(zc.get("masked", True)) else 2.0
                    r[ci, 4] = float(zc.get("mask_gt", float("-inf")))
                    r[ci, 5] = float(zc.get("eps", 1.0e-6))
                    counts.add(int(zc.get("min_count", 16)))
            if len(counts) > 1:
                raise ValueError("intensity_policy: one min_count for all channels")
            if counts:
                min_count = counts.pop()
        else:
            def vec(v, default):
                t = torch.full((C,), default) if v is None else torch.as_tensor(v, dtype=torch.float32).reshape(-1)
                if t.numel() == 1:
                    t = t.repeat(C)
                if t.numel() != C:
                    raise RuntimeError(f"[3DTransforms] len(mean/std)={t.numel()} != C={C}")
                return t
            r[:, 3] = 3.0
            r[:, 6] = vec(self.mean, 0.0)
            r[:, 7] = vec(self.std, 1.0)
        return r, min_count

    # ------------------------------------------------------------------ device passes
    def stats(self, vol: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """vol [B,C,D,H,W] (or [C,D,H,W]) fp32 CUDA -> affine [B,C,4] = (lo, hi, mu, 1/sd) on the device
        (written into ``out`` when given: a persistent buffer for CUDA-graph consumers such as
        ``SlidingWindowTTA(..., intensity_policy=...)``)."""
        v5 = self._check(vol)
        B, C = int(v5.shape[0]), int(v5.shape[1])
        V = int(v5[0, 0].numel())
        key = (v5.device, C)
        if key not in self._rules:
            table, mc = self.rules(C)
            self._rules[key] = (table.to(v5.device), mc)
        rules, min_count = self._rules[key]
        lib = _load_lib()
        nbytes = int(lib.tta_intensity_workspace_bytes(B, C, V))
        wkey = (v5.device, nbytes)
        if wkey not in self._ws:
            self._ws[wkey] = torch.zeros(nbytes, dtype=torch.uint8, device=v5.device)
        affine = torch.empty((B, C, 4), dtype=torch.float32, device=v5.device) if out is None else out
        if tuple(affine.shape) != (B, C, 4) or affine.dtype != torch.float32 or not affine.is_cuda:
            raise ValueError("intensity stats: out must be a CUDA float32 tensor of shape [B, C, 4]")
        with torch.cuda.device(v5.device):
            check(lib.tta_intensity_stats(v5.data_ptr(), B, C, V, rules.data_ptr(), min_count, affine.data_ptr(),
                                          self._ws[wkey].data_ptr(), torch.cuda.current_stream(v5.device).cuda_stream),
                  "intensity_stats")
        return affine

    def __call__(self, vol: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Normalised copy (or ``out``; ``out=vol`` normalises in place)."""
        v5 = self._check(vol)
        affine = self.stats(v5)
        o5 = torch.empty_like(v5) if out is None else self._check(out)
        B, C = int(v5.shape[0]), int(v5.shape[1])
        with torch.cuda.device(v5.device):
            check(_load_lib().tta_intensity_apply(v5.data_ptr(), o5.data_ptr(), B, C, int(v5[0, 0].numel()),
                                                  affine.data_ptr(), torch.cuda.current_stream(v5.device).cuda_stream),
                  "intensity_apply")
        return o5.view(vol.shape) if out is None else out

    @staticmethod
    def _check(vol: torch.Tensor) -> torch.Tensor:
        if vol.dim() == 4:
            vol = vol.unsqueeze(0)
        if vol.dim() != 5:
            raise ValueError(f"[3DTransforms] expect image [C,D,H,W] or [B,C,D,H,W], got {tuple(vol.shape)}")
        if not vol.is_cuda:
            raise RuntimeError("multimodal_tta_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if vol.dtype != torch.float32 or not vol.is_contiguous():
            raise ValueError("intensity policy expects a contiguous float32 volume")
        return vol
