"""ORACLE (test infrastructure only) -- CPU restatement of the reference's 3D UNet.

PARITY UNPINNED: the reference model is a 69-line subclass of ``monai.networks.nets.UNet``
(/root/reference/src/models/unet.py:14-69) and MONAI is neither vendored nor version-pinned
(/root/reference/requirements.txt:1-9 does not list it; it cannot be installed here).  This
file restates MONAI's published ``UNet / Convolution / ADN / ResidualUnit / SkipConnection``
construction in plain ``torch.nn`` (fp32, CPU-runnable).  The only known-answer anchor is
MONAI's published parameter count 4 808 917 for
``UNet(3, 1, 2, (16,32,64,128,256), (2,2,2,2), num_res_units=2, norm=BATCH)`` (checked in
tests/test_oracle_kat.py) plus the counts derived for the reference's own configs
(19 223 961 for configs/_global_patches/brats.yaml:10-19; 7 915 297 for the bare
``model=unet`` defaults of src/models/unet.py:27-48).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this module.  The product path (``multimodal_tta_b200``) never does.

Module names follow MONAI so that ``state_dict()`` keys equal the reference's checkpoints
(SURVEY.md Appendix A): ``model.0.conv.unit0.conv.weight``, ``model.0.residual.weight``,
``model.1.submodule.…``, ``model.2.0.conv.weight``, ``…adn.N.weight`` …
"""
from __future__ import annotations

from typing import Any, Mapping, Sequence

import torch
import torch.nn as nn


def _cfg_get(cfg: Any, key: str, default: Any) -> Any:
    """Duck-typed read of a config node (dict, OmegaConf DictConfig or attribute bag)."""
    if cfg is None:
        return default
    if isinstance(cfg, Mapping):
        v = cfg.get(key, None)
    else:
        v = getattr(cfg, key, None)
        if v is None and hasattr(cfg, "get"):
            try:
                v = cfg.get(key, None)
            except Exception:  # pragma: no cover
                v = None
    return default if v is None else v


def _make_norm(norm: str, channels: int) -> nn.Module:
    """MONAI ``Norm[...]`` factory defaults: INSTANCE -> InstanceNorm3d(affine=False,
    track_running_stats=False, eps=1e-5); BATCH -> BatchNorm3d(affine=True, momentum=0.1)."""
    n = str(norm).upper()
    if n == "INSTANCE":
        return nn.InstanceNorm3d(channels)
    if n == "BATCH":
        return nn.BatchNorm3d(channels)
    raise ValueError(f"unsupported norm {norm!r} (oracle restates INSTANCE and BATCH)")


def _make_act(act: str) -> nn.Module:
    a = str(act).upper()
    if a == "RELU":
        return nn.ReLU()
    raise ValueError(f"unsupported act {act!r} (reference configs use RELU only)")


class ADN(nn.Sequential):
    """MONAI ADN with ordering "NDA": Norm -> Dropout -> Act.  The Dropout module exists
    because the reference passes ``dropout=0.0`` (not None) -- src/models/unet.py:45,65."""

    def __init__(self, channels: int, norm: str, act: str, dropout: float | None):
        super().__init__()
        self.add_module("N", _make_norm(norm, channels))
        if dropout is not None:
            self.add_module("D", nn.Dropout(float(dropout)))
        self.add_module("A", _make_act(act))


class Convolution(nn.Sequential):
    """MONAI ``Convolution``: conv (or transposed conv) k=3, padding=1, bias, then ADN unless
    ``conv_only``.  Transposed: ``output_padding = stride - 1``."""

    def __init__(self, cin: int, cout: int, stride: int, kernel_size: int, norm: str, act: str,
                 dropout: float | None, conv_only: bool = False, is_transposed: bool = False):
        super().__init__()
        pad = (kernel_size - 1) // 2
        if is_transposed:
            conv = nn.ConvTranspose3d(cin, cout, kernel_size, stride=stride, padding=pad,
                                      output_padding=stride - 1, bias=True)
        else:
            conv = nn.Conv3d(cin, cout, kernel_size, stride=stride, padding=pad, bias=True)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("adn", ADN(cout, norm, act, dropout))


class ResidualUnit(nn.Module):
    """MONAI ``ResidualUnit``: ``subunits`` Convolutions (first carries the stride) plus a
    shortcut: k3/stride conv when strided, 1x1 conv when only channels change, identity else."""

    def __init__(self, cin: int, cout: int, stride: int, kernel_size: int, subunits: int,
                 norm: str, act: str, dropout: float | None, last_conv_only: bool = False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual: nn.Module = nn.Identity()
        sch, sst = cin, stride
        subunits = max(1, subunits)
        for su in range(subunits):
            conv_only = last_conv_only and su == subunits - 1
            self.conv.add_module(f"unit{su:d}", Convolution(sch, cout, sst, kernel_size, norm, act,
                                                             dropout, conv_only=conv_only))
            sch, sst = cout, 1
        if stride != 1 or cin != cout:
            rk, rp = kernel_size, (kernel_size - 1) // 2
            if stride == 1:
                rk, rp = 1, 0
            self.residual = nn.Conv3d(cin, cout, rk, stride=stride, padding=rp, bias=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        res = self.residual(x)
        cx = self.conv(x)
        return cx + res


class SkipConnection(nn.Module):
    """MONAI ``SkipConnection(mode="cat")``: ``cat([x, submodule(x)], dim=1)``."""

    def __init__(self, submodule: nn.Module):
        super().__init__()
        self.submodule = submodule

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.cat([x, self.submodule(x)], dim=1)


class OracleUNet(nn.Module):
    """Restatement of ``monai.networks.nets.UNet`` as built by
    /root/reference/src/models/unet.py:56-66 (kernel_size=3, up_kernel_size=3, bias=True,
    adn_ordering="NDA")."""

    def __init__(self, in_channels: int, out_channels: int, channels: Sequence[int],
                 strides: Sequence[int], num_res_units: int = 0, act: str = "relu",
                 norm: str = "BATCH", dropout: float | None = 0.0, kernel_size: int = 3):
        super().__init__()
        if len(channels) < 2:
            raise ValueError("the length of `channels` should be no less than 2.")
        if len(strides) < len(channels) - 1:
            raise ValueError("the length of `strides` should equal to `len(channels) - 1`.")
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.channels, self.strides = [int(c) for c in channels], [int(s) for s in strides]
        self.num_res_units, self.act, self.norm, self.dropout = int(num_res_units), act, norm, dropout
        self.kernel_size = kernel_size

        def down_layer(cin, cout, stride):
            if self.num_res_units > 0:
                return ResidualUnit(cin, cout, stride, kernel_size, self.num_res_units, norm, act, dropout)
            return Convolution(cin, cout, stride, kernel_size, norm, act, dropout)

        def up_layer(cin, cout, stride, is_top):
            conv: nn.Module = Convolution(cin, cout, stride, kernel_size, norm, act, dropout,
                                          conv_only=is_top and self.num_res_units == 0,
                                          is_transposed=True)
            if self.num_res_units > 0:
                ru = ResidualUnit(cout, cout, 1, kernel_size, 1, norm, act, dropout, last_conv_only=is_top)
                conv = nn.Sequential(conv, ru)
            return conv

        def create_block(inc, outc, chs, sts, is_top):
            c, s = chs[0], sts[0]
            if len(chs) > 2:
                sub = create_block(c, c, chs[1:], sts[1:], False)
                upc = c * 2
            else:
                sub = down_layer(c, chs[1], 1)  # bottom layer
                upc = c + chs[1]
            down = down_layer(inc, c, s)
            up = up_layer(upc, outc, s, is_top)
            return nn.Sequential(down, SkipConnection(sub), up)

        self.model = create_block(self.in_channels, self.out_channels, self.channels, self.strides, True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)

    @classmethod
    def from_cfg(cls, cfg: Any, in_channels: int | None = None) -> "OracleUNet":
        """Same fields and defaults as /root/reference/src/models/unet.py:27-48."""
        c_in_cfg = _cfg_get(cfg, "in_channels", 3)
        c_in = in_channels if in_channels is not None else (None if c_in_cfg == "auto" else int(c_in_cfg))
        if c_in is None:
            raise ValueError("[UNet] in_channels is 'auto'; please pass in_channels at construction time.")
        spatial_dims = int(_cfg_get(cfg, "spatial_dims", 3))
        if spatial_dims != 3:
            raise ValueError("oracle restates the 3-D path only")
        return cls(
            in_channels=c_in,
            out_channels=int(_cfg_get(cfg, "num_classes", 1)),
            channels=list(_cfg_get(cfg, "channels", [32, 64, 128, 256, 512])),
            strides=list(_cfg_get(cfg, "strides", [2, 2, 2, 2])),
            num_res_units=int(_cfg_get(cfg, "num_res_units", 0)),
            act=_cfg_get(cfg, "act", "relu"),
            norm=_cfg_get(cfg, "norm", "BATCH"),
            dropout=float(_cfg_get(cfg, "dropout", 0.0)),
        )


BRATS_MODEL_CFG = dict(in_channels=4, num_classes=3, spatial_dims=3, channels=[32, 64, 128, 256, 512],
                       strides=[2, 2, 2, 2], num_res_units=2, norm="INSTANCE", act="RELU", dropout=0.0)
HECKTOR_MODEL_CFG = dict(in_channels=2, num_classes=1, spatial_dims=3, channels=[32, 64, 128, 256, 512],
                         strides=[2, 2, 2, 2], num_res_units=2, norm="INSTANCE", act="RELU", dropout=0.0)
BARE_DEFAULT_MODEL_CFG = dict(name="unet", num_classes=1)  # configs/model/unet.yaml + unet.py defaults


def norm_modules(model: nn.Module) -> list[tuple[str, nn.Module]]:
    """Norm layers in registration (= forward) order; defines the flat [gamma || beta] order."""
    return [(n, m) for n, m in model.named_modules()
            if isinstance(m, (nn.InstanceNorm3d, nn.BatchNorm3d))]
