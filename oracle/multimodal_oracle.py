"""ORACLE (test infrastructure only) -- CPU restatement of the reference's multimodal model
``MultimodalUNetDeepFusion`` (/root/reference/src/models/unet_multimodal_midfusion.py:16-267,
registered as ``unet_multimodal_deepfusion`` / ``unet_multimodal_midfusion``;
/root/reference/configs/model/unet_multimodal_midfusion.yaml:6-26).

PARITY UNPINNED: like the plain UNet the arithmetic lives in MONAI blocks that are neither vendored nor
pinned -- ``ResidualUnit`` / ``Convolution`` (restated in oracle/unet_oracle.py) and
``UpSample(mode="nontrainable")``, which under MONAI's published defaults (``pre_conv="default"``,
``interp_mode="linear"``, ``align_corners=True``, ``bias=True``) is a 1x1x1 ``Conv3d`` named ``preconv`` when the
channel count changes, followed by ``nn.Upsample(scale_factor, mode="trilinear", align_corners=True)`` named
``upsample_non_trainable``.  The forward below follows the reference file line by line (:204-262): per-modality
encoders on ``torch.split(x, 1, dim=1)``, bottleneck fusion ``f_shared + Convolution(cat[f_shared, f_specific])``
with ONE fusion layer shared by all modalities, ``bottleneck_reduce`` (1x1x1, no bias), modality-mean skips,
four decoder stages (the last one concatenates the modality-mean of the INPUT), ``final_conv`` (1x1x1).
Module names equal the reference's so that ``state_dict()`` keys match its checkpoints.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

from typing import Any, List, Sequence

import torch
import torch.nn as nn

from .unet_oracle import Convolution, ResidualUnit, _cfg_get


class SpecificEncoder(nn.Module):
    """unet_multimodal_midfusion.py:16-78: ResidualUnits ``channels`` x ``strides + [1]``; skips = outputs of all
    levels but the last."""

    def __init__(self, in_channels, channels, strides, num_res_units, act, norm, dropout):
        super().__init__()
        self.layers = nn.ModuleList()
        cur = in_channels
        for out_ch, s in zip(channels, list(strides) + [1]):
            self.layers.append(ResidualUnit(cur, out_ch, s, 3, num_res_units, norm, act, dropout))
            cur = out_ch

    def forward(self, x):
        skips = []
        for i, layer in enumerate(self.layers):
            x = layer(x)
            if i < len(self.layers) - 1:
                skips.append(x)
        return x, torch.mean(x, dim=[2, 3, 4], keepdim=True), skips


class CompositionalLayer(nn.Module):
    """:81-97 -- ``Convolution(2C -> C, k3, act, norm)`` (no dropout argument -> no Dropout module)."""

    def __init__(self, channels, norm, act):
        super().__init__()
        self.fusion_conv = Convolution(channels * 2, channels, 1, 3, norm, act, None)

    def forward(self, f_shared, f_specific):
        return f_shared + self.fusion_conv(torch.cat([f_shared, f_specific], dim=1))


class UpSampleNonTrainable(nn.Sequential):
    """MONAI ``UpSample(mode="nontrainable")`` with its defaults (see the module docstring)."""

    def __init__(self, cin, cout, scale):
        super().__init__()
        if cin != cout:
            self.add_module("preconv", nn.Conv3d(cin, cout, kernel_size=1, bias=True))
        self.add_module("upsample_non_trainable", nn.Upsample(scale_factor=scale, mode="trilinear", align_corners=True))


class DecoderStage(nn.Module):
    """:100-136."""

    def __init__(self, cin, skip, cout, stride, num_res_units, act, norm, dropout):
        super().__init__()
        self.upsample = UpSampleNonTrainable(cin, cout, stride)
        self.conv = ResidualUnit(cout + skip, cout, 1, 3, num_res_units, norm, act, dropout)

    def forward(self, x, skip):
        return self.conv(torch.cat([self.upsample(x), skip], dim=1))


class OracleMultimodalUNet(nn.Module):
    def __init__(self, num_modalities=4, num_classes=3, channels: Sequence[int] = (32, 64, 128, 256, 512),
                 strides: Sequence[int] = (2, 2, 2, 2), num_res_units=2, act="RELU", norm="INSTANCE", dropout=0.0,
                 domain_enabled=True):
        super().__init__()
        channels, strides = [int(c) for c in channels], [int(s) for s in strides]
        self.num_modalities = int(num_modalities)
        self.in_channels, self.out_channels = self.num_modalities, int(num_classes)
        self.specific_encoders = nn.ModuleList([
            SpecificEncoder(1, channels, strides, num_res_units, act, norm, dropout) for _ in range(self.num_modalities)])
        self.fusion_layer = CompositionalLayer(channels[-1], norm, act)
        self.bottleneck_reduce = nn.Conv3d(channels[-1] * self.num_modalities, channels[-1], 1, bias=False)
        self.decoder_stages = nn.ModuleList()
        skip_channels = [channels[2], channels[1], channels[0], 1]          # :178 (hard-wired to five levels)
        for i in range(len(channels) - 1):
            idx = len(channels) - 1 - i
            self.decoder_stages.append(DecoderStage(channels[idx], skip_channels[i], channels[idx - 1],
                                                    strides[idx - 1], num_res_units, act, norm, dropout))
        self.final_conv = nn.Conv3d(channels[0], self.out_channels, kernel_size=1)
        self.domain_enabled = bool(domain_enabled)
        if self.domain_enabled:
            self.domain_classifier = nn.Linear(channels[-1], self.num_modalities)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        modalities = torch.split(x, 1, dim=1)
        feats, all_skips = [], []
        for enc, modal in zip(self.specific_encoders, modalities):
            f, _, skips = enc(modal)
            feats.append(f)
            all_skips.append(skips)
        shared = torch.stack(feats, dim=1).mean(dim=1)
        fused = [self.fusion_layer(shared, f) for f in feats]
        x_dec = self.bottleneck_reduce(torch.cat(fused, dim=1))
        fused_skips: List[torch.Tensor] = [torch.stack([m[i] for m in all_skips]).mean(dim=0)
                                           for i in range(len(all_skips[0]))]
        input_mean = torch.stack(modalities).mean(dim=0)
        for stage, skip in zip(self.decoder_stages, [fused_skips[2], fused_skips[1], fused_skips[0], input_mean]):
            x_dec = stage(x_dec, skip)
        return self.final_conv(x_dec)

    @classmethod
    def from_cfg(cls, cfg: Any) -> "OracleMultimodalUNet":
        """Fields / defaults of unet_multimodal_midfusion.py:148-161."""
        if int(_cfg_get(cfg, "spatial_dims", 3)) != 3:
            raise ValueError("oracle restates the 3-D path only")
        dom = _cfg_get(cfg, "domain_classifier", {})
        return cls(num_modalities=int(_cfg_get(cfg, "num_modalities", 4)), num_classes=int(_cfg_get(cfg, "num_classes", 3)),
                   channels=list(_cfg_get(cfg, "channels", [32, 64, 128, 256, 512])),
                   strides=list(_cfg_get(cfg, "strides", [2, 2, 2, 2])),
                   num_res_units=int(_cfg_get(cfg, "num_res_units", 2)), act=_cfg_get(cfg, "act", "RELU"),
                   norm=_cfg_get(cfg, "norm", "INSTANCE"), dropout=float(_cfg_get(cfg, "dropout", 0.0)),
                   domain_enabled=bool(_cfg_get(dom, "enabled", True)))


MULTIMODAL_MODEL_CFG = dict(name="unet_multimodal_deepfusion", num_modalities=4, num_classes=3, spatial_dims=3,
                            channels=[32, 64, 128, 256, 512], strides=[2, 2, 2, 2], num_res_units=2, norm="INSTANCE",
                            act="RELU", dropout=0.0)
