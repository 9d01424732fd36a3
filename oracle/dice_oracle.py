"""ORACLE (test infrastructure only) -- region Dice / IoU exactly as the reference evaluator.

PINNED: checked against outputs of the reference's own ``_binary_dice_iou`` imported from
/root/reference/src/evaluation/seg_eval.py:41-68 (tests/golden/gen_dice_golden.py ->
tests/golden/dice_golden.npz).  Aggregation follows seg_eval.py:304-308 (sigmoid, >= threshold,
label > 0.5), :363-392 (GT-empty regions skipped) and :402-423 (per-region means, avg_dc, miou).

Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may import this module.
"""
from __future__ import annotations

from typing import Sequence

import torch


def binary_dice_iou(pred: torch.Tensor, gt: torch.Tensor, eps: float = 1e-7):
    """pred, gt: [B,R,D,H,W] in {0,1} -> dice [B,R], iou [B,R], valid [B,R] (GT non-empty)."""
    assert pred.shape == gt.shape
    B, R = pred.shape[:2]
    p = pred.reshape(B, R, -1).float()
    g = gt.reshape(B, R, -1).float()
    inter = (p * g).sum(-1)
    ps, gs = p.sum(-1), g.sum(-1)
    valid = gs > 0
    dice = (2.0 * inter + eps) / (ps + gs + eps)
    iou = (inter + eps) / (ps + gs - inter + eps)
    return dice, iou, valid


def evaluate_logits(logits_list: Sequence[torch.Tensor], labels_list: Sequence[torch.Tensor],
                    region_order: Sequence[str] = ("ET", "TC", "WT"), threshold: float = 0.5) -> dict:
    """Reference metric dict for a sequence of (logits, label) batches: ``<region>_dc``,
    ``avg_dc``, ``miou``, ``jc``."""
    R = len(region_order)
    sd = torch.zeros(R, dtype=torch.float64)
    cd = torch.zeros(R, dtype=torch.float64)
    si = torch.zeros(R, dtype=torch.float64)
    for logits, y in zip(logits_list, labels_list):
        pred = (torch.sigmoid(logits) >= threshold).to(torch.uint8)
        gt = (y.float() > 0.5).to(torch.uint8)
        dice, iou, valid = binary_dice_iou(pred, gt)
        for i in range(pred.shape[0]):
            for c in range(R):
                if bool(valid[i, c]):
                    sd[c] += float(dice[i, c])
                    si[c] += float(iou[i, c])
                    cd[c] += 1.0
    md = [float(sd[c] / cd[c]) if cd[c] > 0 else 0.0 for c in range(R)]
    mi = [float(si[c] / cd[c]) if cd[c] > 0 else 0.0 for c in range(R)]
    vr = [c for c in range(R) if cd[c] > 0]
    out = {f"{n.lower()}_dc": v for n, v in zip(region_order, md)}
    out["avg_dc"] = float(sum(md[c] for c in vr) / max(1, len(vr)))
    out["miou"] = float(sum(mi[c] for c in vr) / max(1, len(vr)))
    out["jc"] = out["miou"]
    return out
