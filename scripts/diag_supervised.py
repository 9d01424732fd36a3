"""Diagnostic (GPU): per-parameter gradient error of the supervised backward vs CPU autograd."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import UNetB200
from multimodal_tta_b200.presets import BRATS_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, region_labels
from oracle.dicece_oracle import dice_ce_loss
from oracle.unet_oracle import OracleUNet
from tests.util import rel_l2
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = brats_volume(2, (S, S, S), seed=5); y = region_labels(2, 3, (S, S, S), seed=6)
for opts in ({}, {"bwd_precision": "bf16x2"}, {"conv_backend": "simt"}):
    torch.manual_seed(81)
    oracle = OracleUNet.from_cfg(BRATS_MODEL_CFG).train()
    prod = UNetB200(dict(BRATS_MODEL_CFG, trainable=True, deterministic=True, **opts))
    prod.load_state_dict(copy.deepcopy(oracle.state_dict())); prod.to("cuda").train()
    dice_ce_loss(oracle(x), y).backward()
    dice_ce_loss(prod(x.cuda()), y.cuda()).backward()
    po, pp = dict(oracle.named_parameters()), dict(prod.named_parameters())
    rows, num, den = [], 0.0, 0.0
    for n, p in po.items():
        d = float((pp[n].grad.cpu().double() - p.grad.double()).pow(2).sum()); q = float(p.grad.double().pow(2).sum())
        num += d; den += q
        rows.append((d, n, (d / max(q, 1e-300)) ** 0.5, q ** 0.5))
    rows.sort(reverse=True)
    print(opts, f"total rel-L2 {(num / den) ** 0.5:.2e}")
    for d, n, r, q in rows[:6]:
        print(f"   {n:60s} rel {r:.2e} |g| {q:.2e} share of error {d / num:.2f}")
