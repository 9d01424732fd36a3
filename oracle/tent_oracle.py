"""ORACLE (test infrastructure only) -- canonical TENT step on the oracle UNet, CPU fp32.

PARITY UNPINNED: the reference holds no TTA code (SURVEY.md "Read this first" #1); this
restates the canonical TENT recipe (Wang et al., ICLR 2021) on top of the reference's own
step shape and optimizer factory:
  * step order zero_grad -> forward -> loss -> backward -> optimizer.step -> float(loss)
    follows /root/reference/src/core/trainers/seg_trainer.py:105-145;
  * Adam construction (eps 1e-8, 1-D params get weight_decay 0) follows
    /root/reference/src/core/experiment_manager.py:199-237;
  * sigmoid heads (BraTS ET/TC/WT regions, HECKTOR GTVt) follow
    /root/reference/configs/_global_patches/brats.yaml:12 and seg_eval.py:304.

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

import copy
from typing import Iterable

import torch
import torch.nn as nn
import torch.nn.functional as F


def softmax_entropy(logits: torch.Tensor) -> torch.Tensor:
    """Per-voxel Shannon entropy of softmax over dim 1: H = lse(z) - sum_c p_c z_c. -> [B,D,H,W]"""
    logp = F.log_softmax(logits, dim=1)
    return -(logp.exp() * logp).sum(dim=1)


def bernoulli_entropy(logits: torch.Tensor) -> torch.Tensor:
    """Per-voxel sum over channels of the Bernoulli entropy of sigmoid(z):
    H = sum_c [softplus(z_c) - p_c z_c].  Mandatory when C == 1. -> [B,D,H,W]"""
    p = torch.sigmoid(logits)
    return (F.softplus(logits) - p * logits).sum(dim=1)


def entropy_loss(logits: torch.Tensor, mode: str) -> torch.Tensor:
    """Mean over batch and voxels of the per-voxel entropy (the 1/(B*D*H*W) scaling)."""
    if mode == "softmax":
        if logits.shape[1] < 2:
            raise ValueError("softmax entropy is degenerate for a single-channel head")
        return softmax_entropy(logits).mean()
    if mode == "sigmoid":
        return bernoulli_entropy(logits).mean()
    raise ValueError(f"unknown entropy mode {mode!r}")


def configure_model(model: nn.Module) -> nn.Module:
    """TENT configure: train mode, freeze everything, enable grads on norm weight/bias;
    materialise gamma=1, beta=0 where affine=False; BatchNorm always uses batch statistics."""
    model.train()
    model.requires_grad_(False)
    dev = next((p.device for p in model.parameters()), torch.device("cpu"))
    for m in model.modules():
        if isinstance(m, (nn.InstanceNorm3d, nn.BatchNorm3d)):
            c = m.num_features
            if m.weight is None:
                m.weight = nn.Parameter(torch.ones(c, device=dev))
                m.bias = nn.Parameter(torch.zeros(c, device=dev))
                m.affine = True
            m.weight.requires_grad_(True)
            m.bias.requires_grad_(True)
            if isinstance(m, nn.BatchNorm3d):
                m.track_running_stats = False
                m.running_mean = None
                m.running_var = None
    return model


def collect_params(model: nn.Module) -> tuple[list[nn.Parameter], list[str]]:
    """Norm affine params in forward order: [w0, b0, w1, b1, ...] with their names."""
    params, names = [], []
    for nm, m in model.named_modules():
        if isinstance(m, (nn.InstanceNorm3d, nn.BatchNorm3d)):
            for pn in ("weight", "bias"):
                params.append(getattr(m, pn))
                names.append(f"{nm}.{pn}")
    return params, names


def flat_gamma_beta(model: nn.Module) -> torch.Tensor:
    """Flat [gamma_0 .. gamma_L || beta_0 .. beta_L] buffer, the product's parameter order."""
    g, b = [], []
    for m in model.modules():
        if isinstance(m, (nn.InstanceNorm3d, nn.BatchNorm3d)):
            g.append(m.weight.detach().reshape(-1))
            b.append(m.bias.detach().reshape(-1))
    return torch.cat(g + b)


def flat_grads(model: nn.Module) -> torch.Tensor:
    g, b = [], []
    for m in model.modules():
        if isinstance(m, (nn.InstanceNorm3d, nn.BatchNorm3d)):
            g.append(m.weight.grad.detach().reshape(-1))
            b.append(m.bias.grad.detach().reshape(-1))
    return torch.cat(g + b)


class TentOracle:
    """Non-episodic TENT by default (state persists across batches); ``episodic=True`` resets
    model + optimizer state before every batch.  ``step`` returns the logits computed BEFORE
    the update, and the scalar loss."""

    def __init__(self, model: nn.Module, mode: str = "sigmoid", lr: float = 1e-3,
                 betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, steps: int = 1, episodic: bool = False):
        self.model = configure_model(model)
        self.mode = mode
        self.steps = int(steps)
        self.episodic = episodic
        params, self.param_names = collect_params(self.model)
        # experiment_manager.py:216-221 -- 1-D params go to the weight_decay=0 group.
        self.optimizer = torch.optim.Adam([{"params": params, "weight_decay": 0.0}], lr=lr,
                                          betas=betas, eps=eps)
        self._model_state = copy.deepcopy(self.model.state_dict())
        self._optim_state = copy.deepcopy(self.optimizer.state_dict())
        self.last_grads: torch.Tensor | None = None

    def reset(self) -> None:
        self.model.load_state_dict(self._model_state, strict=True)
        self.optimizer.load_state_dict(self._optim_state)

    def step(self, x: torch.Tensor) -> tuple[torch.Tensor, float]:
        if self.episodic:
            self.reset()
        out, loss_val = None, 0.0
        for _ in range(self.steps):
            self.optimizer.zero_grad()
            logits = self.model(x)
            loss = entropy_loss(logits, self.mode)
            loss.backward()
            self.last_grads = flat_grads(self.model)
            self.optimizer.step()
            if out is None:
                out, loss_val = logits.detach(), float(loss.item())
        return out, loss_val


def adam_reference(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
                   lr: float, b1: float, b2: float, eps: float) -> None:
    """In-place single-tensor Adam exactly as torch.optim.Adam (bias-corrected,
    denom = sqrt(v)/sqrt(bc2) + eps, step_size = lr/bc1, weight_decay 0)."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))
