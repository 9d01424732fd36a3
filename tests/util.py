"""Shared helpers for the parity tests (oracle side runs on the CPU in fp32)."""
import copy

import torch

from multimodal_tta_b200 import TentB200, UNetB200
from multimodal_tta_b200._lib import TTA_BF16, TTA_F16, check
from multimodal_tta_b200.layout import from_chunked, join_planes, split_planes, to_chunked
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import OracleUNet


def stream():
    return torch.cuda.current_stream().cuda_stream


def make_pair(cfg, seed=0, device="cuda", **tent_kw):
    """Oracle model + product model with identical weights."""
    torch.manual_seed(seed)
    oracle = OracleUNet.from_cfg(cfg)
    prod = UNetB200(dict(cfg))
    prod.load_state_dict(copy.deepcopy(oracle.state_dict()))
    prod.to(device)
    return oracle, prod


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def planes_from(x_ncdhw: torch.Tensor, dtype_tag: int):
    """NCDHW fp32 (cuda) -> (hi, lo) int16 planes in chunked layout, plus the value they encode."""
    ch = to_chunked(x_ncdhw)
    hi, lo = split_planes(ch, dtype_tag)
    val = from_chunked(join_planes(hi, lo, dtype_tag), x_ncdhw.shape[1])
    return hi.contiguous(), lo.contiguous(), val


def wsplit(planes: torch.Tensor) -> torch.Tensor:
    """[..., W, 8] chunked planes -> the w-parity-split layout [..., (even w | odd w), 8]."""
    return torch.cat([planes[..., 0::2, :], planes[..., 1::2, :]], dim=-2).contiguous()
