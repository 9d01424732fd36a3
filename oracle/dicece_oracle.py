"""ORACLE (test infrastructure only) -- the supervised loss of the reference's trainer.

PARITY UNPINNED: ``SegTrainer._build_loss`` (/root/reference/src/core/trainers/seg_trainer.py:59-79) builds
``monai.losses.DiceCELoss(include_background, to_onehot_y, softmax, sigmoid, squared_pred, jaccard, lambda_dice,
lambda_ce, reduction="mean", weight)`` and MONAI is neither vendored nor pinned; the cross-entropy term for
``sigmoid=True`` with a multi-channel float target changed across MONAI releases.  This restates
MONAI >= 1.3: Dice on ``sigmoid(logits)`` per (sample, channel) with ``smooth_nr = smooth_dr = 1e-5``; CE =
``nn.CrossEntropyLoss`` with the float target as class probabilities when logits have more than one channel,
``BCEWithLogitsLoss`` for a single channel.  Effective BraTS values (configs/_global_patches/brats.yaml:46-55):
include_background true, sigmoid true, lambda_dice 1, lambda_ce 1 (``lambda_bce`` is never read).

The product does NOT implement this loss: the supervised path keeps the reference's own loss object and optimizer
(torch ops on the logits) and accelerates the model's forward / backward behind an autograd node.  Only tests/ import
this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def dice_ce_loss(logits: torch.Tensor, target: torch.Tensor, include_background: bool = True,
                 squared_pred: bool = False, jaccard: bool = False, lambda_dice: float = 1.0,
                 lambda_ce: float = 1.0) -> torch.Tensor:
    p, t = torch.sigmoid(logits), target.to(logits.dtype)
    if not include_background and logits.shape[1] > 1:
        p, t = p[:, 1:], t[:, 1:]
    dims = tuple(range(2, logits.ndim))
    inter = (p * t).sum(dims)
    go, po = ((t * t).sum(dims), (p * p).sum(dims)) if squared_pred else (t.sum(dims), p.sum(dims))
    denom = go + po
    if jaccard:
        denom = 2.0 * (denom - inter)
    dice = (1.0 - (2.0 * inter + 1e-5) / (denom + 1e-5)).mean()
    if logits.shape[1] == 1:
        ce = F.binary_cross_entropy_with_logits(logits, target.to(logits.dtype))
    else:
        ce = F.cross_entropy(logits, target.to(logits.dtype))      # float target = class probabilities
    return lambda_dice * dice + lambda_ce * ce


def make_optimizer(model: torch.nn.Module, name: str = "adam", lr: float = 1e-5, weight_decay: float = 5e-4,
                   betas=(0.9, 0.9999), eps: float = 1e-8) -> torch.optim.Optimizer:
    """ExperimentManager._build_optimizer_for (/root/reference/src/core/experiment_manager.py:199-237) with the
    defaults of configs/training/default.yaml:28-37,52-55: parameters whose name contains a no-decay key or that are
    1-D get weight_decay 0."""
    no_decay_keys = ("bias", "bn", "norm", "LayerNorm")
    decay, no_decay = [], []
    for n, p in model.named_parameters():
        if not p.requires_grad:
            continue
        (no_decay if (any(k in n for k in no_decay_keys) or p.ndim == 1) else decay).append(p)
    groups = [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]
    if name == "adam":
        return torch.optim.Adam(groups, lr=lr, betas=betas, eps=eps)
    if name == "adamw":
        return torch.optim.AdamW(groups, lr=lr, betas=betas, eps=eps)
    return torch.optim.SGD(groups, lr=lr, momentum=0.9)
