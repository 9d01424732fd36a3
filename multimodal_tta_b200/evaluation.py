"""``tta_seg_eval`` -- evaluation strategy that adapts while it predicts.

Drop-in beside the reference's ``@register_evaluation_strategy("seg_eval")``
(/root/reference/src/evaluation/seg_eval.py:151-479): same constructor (``cls(config)``), same
``evaluate_epoch(model, data_loader, device) -> Dict[str, float]`` contract and the same metric
keys (``<region>_dc``, ``avg_dc``, ``miou``, ``jc``, ``loss``, ``dom/<name>/...``), selected with
``task.eval_strategy=tta_seg_eval`` (/root/reference/src/core/experiment_manager.py:64,365-370).
Differences, all required by TTA: no ``torch.no_grad``/``model.eval()``;
each batch goes through one TENT step (or a sliding-window sweep of TENT steps) and is scored
with the logits computed before the update; sigmoid/threshold/Dice sums run in one on-device
kernel (tta_dice_counts) instead of a Python B x R loop of ``.item()`` syncs.  ``loss`` reports
the mean entropy (there is no supervised loss at test time).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import check
from .config import DictConfig, create, get_config
from .registry import register_evaluation_strategy
from .sliding_window import SlidingWindowTTA
from .tent import TentB200
from .unet_b200 import B200Model


def dice_iou_from_counts(counts: torch.Tensor, eps: float = 1e-7):
    """counts [B,R,3] = (inter, pred_sum, gt_sum) -> dice, iou, valid exactly as
    seg_eval.py:41-68 (float32 arithmetic on the integer sums)."""
    c = counts.to(torch.float32)
    inter, ps, gs = c[..., 0], c[..., 1], c[..., 2]
    valid = gs > 0
    dice = (2.0 * inter + eps) / (ps + gs + eps)
    iou = (inter + eps) / (ps + gs - inter + eps)
    return dice, iou, valid


def device_dice_counts(logits: torch.Tensor, labels: torch.Tensor, threshold: float,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One kernel: sigmoid >= threshold, label > 0.5, intersection / sums per (b, r).  ``out``: a ZEROED contiguous
    int64 [B, R, 3] device tensor to add into (a slot of a caller-owned ring: no allocation / memset per call)."""
    B, R = int(logits.shape[0]), int(logits.shape[1])
    V = logits[0, 0].numel()
    if out is not None:
        if out.dtype != torch.int64 or tuple(out.shape) != (B, R, 3) or not out.is_contiguous() or out.device != logits.device:
            raise ValueError("device_dice_counts: out must be a contiguous int64 [B, R, 3] tensor on the logits' device")
        counts = out
    else:
        counts = torch.zeros((B, R, 3), dtype=torch.int64, device=logits.device)
    lab = labels.to(device=logits.device, dtype=torch.float32).contiguous()
    if not logits.is_cuda:
        raise RuntimeError("multimodal_tta_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    with torch.cuda.device(logits.device):
        check(_lib.lib().tta_dice_counts(logits.contiguous().data_ptr(), lab.data_ptr(), B * R, V, float(threshold),
                                         counts.data_ptr(), torch.cuda.current_stream(logits.device).cuda_stream),
              "dice_counts")
    return counts


def _as_list_str(x, batch_size: int) -> List[str]:
    """'domain' batch entry -> one string per sample, with the reference's cases (seg_eval.py:20-38): missing -> ""
    (reported as ``unknown``), list/tuple -> str of every item, one str -> repeated, 0-d tensor -> its int for the
    whole batch, B-element tensor -> one int per sample, anything else -> str(x) repeated."""
    if x is None:
        return [""] * batch_size
    if isinstance(x, (list, tuple)):
        return [str(v) for v in x]
    if isinstance(x, str):
        return [x] * batch_size
    if torch.is_tensor(x):
        if x.ndim == 0:
            return [str(int(x.item()))] * batch_size
        if x.numel() == batch_size:
            return [str(int(v.item())) for v in x.reshape(-1)]
    return [str(x)] * batch_size


def load_source_checkpoint(model: B200Model, path: str, trusted: bool = False) -> None:
    """``method.checkpoint``: source-model weights written by the reference's CheckpointHook
    (/root/reference/src/core/hooks.py:53-70: ``{"epoch", "model_state_dict", "optimizer_state_dict",
    "best_metrics"}``; keys carry a ``module.`` prefix when the trainer wrapped the model in nn.DataParallel,
    experiment_manager.py:95-96).  A bare state dict is accepted too.  Unlike the reference's loader
    (hooks.py:72-76, warning + train from scratch) a missing file is an error: adapting random weights at
    test time is never what the user meant."""
    import os

    if not os.path.exists(path):
        raise FileNotFoundError(f"method.checkpoint: {path} does not exist")
    ck = torch.load(path, map_location="cpu", weights_only=not trusted)
    sd = ck["model_state_dict"] if isinstance(ck, dict) and "model_state_dict" in ck else ck
    model.load_state_dict(sd)


@register_evaluation_strategy("tta_seg_eval")
class TTASegmentationEvaluationStrategy:
    def __init__(self, config: Optional[DictConfig] = None):
        self.config = config if config is not None else create({})
        seg_cfg = get_config(self.config, "evaluation.seg", create({}))
        self.threshold = float(get_config(seg_cfg, "threshold", 0.5))
        self.region_order = list(get_config(seg_cfg, "region_order", ["ET", "TC", "WT"]))
        self.method_cfg = get_config(self.config, "method", create({}))
        sw = get_config(self.method_cfg, "sliding_window", create({}))
        self.sw_enabled = bool(get_config(sw, "enabled", False))
        self.sw_roi = list(get_config(sw, "roi", [128, 128, 128]))
        self.sw_batch = int(get_config(sw, "sw_batch", 1))
        self.sw_overlap = float(get_config(sw, "overlap", 0.5))
        self.checkpoint = get_config(self.method_cfg, "checkpoint", None)
        self.checkpoint_trusted = bool(get_config(self.method_cfg, "checkpoint_trusted", False))
        self._tent: Optional[TentB200] = None
        self._sw: Optional[SlidingWindowTTA] = None

    def _method(self, model) -> TentB200:
        core = model.module if hasattr(model, "module") else model  # nn.DataParallel wrapper
        if not isinstance(core, B200Model):
            raise TypeError("tta_seg_eval needs model.name=unet_b200 / unet_multimodal_deepfusion_b200 (got "
                            f"{type(core).__name__}); see configs/method/tent_b200.yaml")
        if self._tent is None or self._tent.model is not core:
            if self.checkpoint:      # source weights first: TENT snapshots the parameters it starts from
                load_source_checkpoint(core, str(self.checkpoint), self.checkpoint_trusted)
            self._tent = TentB200(core, self.method_cfg)
            self._sw = SlidingWindowTTA(self._tent, self.sw_roi, self.sw_batch, self.sw_overlap) \
                if self.sw_enabled else None
        return self._tent

    def evaluate_epoch(self, model, data_loader, device) -> Dict[str, float]:
        tent = self._method(model)
        tent.model.to(device)
        R_expected = len(self.region_order)
        f64 = lambda: torch.zeros(R_expected, dtype=torch.float64)
        sum_dice, cnt, sum_iou = f64(), f64(), f64()
        dom_sd, dom_c, dom_si = defaultdict(f64), defaultdict(f64), defaultdict(f64)
        total_loss, n_samples = 0.0, 0
        for batch in data_loader:
            x = batch["image"].to(device)
            B = x.size(0)
            if "label" not in batch:
                raise KeyError("[TTASegEval] batch must contain 'label' for region-based eval.")
            y = batch["label"]
            y = y.to(device) if torch.is_tensor(y) else torch.as_tensor(y, device=device)
            if y.ndim == 4:
                y = y.unsqueeze(0).expand(B, -1, -1, -1, -1)
            if y.ndim != 5:
                raise ValueError(f"[TTASegEval] label must be 5D, got {tuple(y.shape)}")
            if int(y.size(1)) != R_expected:
                raise ValueError(f"[TTASegEval] label channels={int(y.size(1))} but region_order={R_expected}")
            logits = self._sw(x) if self._sw is not None else tent.step(x)
            if logits.ndim != 5 or int(logits.size(1)) != R_expected:
                raise ValueError(f"[TTASegEval] model logits must be [B,{R_expected},D,H,W], got {tuple(logits.shape)}")
            counts = device_dice_counts(logits, y.float(), self.threshold)
            dice, iou, valid = dice_iou_from_counts(counts.cpu())
            domains = _as_list_str(batch.get("domain", None), B)
            for i in range(B):
                for c in range(R_expected):
                    if bool(valid[i, c]):
                        dv, iv = float(dice[i, c]), float(iou[i, c])
                        sum_dice[c] += dv; sum_iou[c] += iv; cnt[c] += 1.0
                        dom_sd[domains[i]][c] += dv; dom_si[domains[i]][c] += iv; dom_c[domains[i]][c] += 1.0
            total_loss += float(tent.last_loss.item()) * B
            n_samples += B

        def fin(s, c):
            return [float(s[i] / c[i]) if c[i] > 0 else 0.0 for i in range(R_expected)]

        md, mi = fin(sum_dice, cnt), fin(sum_iou, cnt)
        vr = [i for i in range(R_expected) if cnt[i] > 0]
        metrics: Dict[str, float] = {f"{n.lower()}_dc": v for n, v in zip(self.region_order, md)}
        metrics["avg_dc"] = float(sum(md[i] for i in vr) / max(1, len(vr)))
        metrics["miou"] = float(sum(mi[i] for i in vr) / max(1, len(vr)))
        metrics["jc"] = metrics["miou"]
        metrics["loss"] = float(total_loss / max(1, n_samples))
        for dom in sorted(dom_sd.keys()):
            safe = dom if dom != "" else "unknown"
            dm, di = fin(dom_sd[dom], dom_c[dom]), fin(dom_si[dom], dom_c[dom])
            dv = [i for i in range(R_expected) if dom_c[dom][i] > 0]
            for n, v in zip(self.region_order, dm):
                metrics[f"dom/{safe}/{n.lower()}_dc"] = v
            metrics[f"dom/{safe}/avg_dc"] = float(sum(dm[i] for i in dv) / max(1, len(dv)))
            metrics[f"dom/{safe}/miou"] = float(sum(di[i] for i in dv) / max(1, len(dv)))
        return metrics

    def is_best_model(self, eval_stats: Dict[str, float], best_metrics: Dict[str, float]) -> bool:
        return eval_stats.get("avg_dc", 0.0) > best_metrics.get("avg_dc", float("-inf"))
