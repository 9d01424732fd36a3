"""Diagnostic (GPU): how far is each conv backend from the fp32 CPU oracle on one TENT step?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import TentB200
from multimodal_tta_b200.synthetic import brats_volume
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import BRATS_MODEL_CFG
from tests.util import make_pair, rel_l2

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
x = brats_volume(2, (S, S, S), seed=42)
for backend in ("simt", "tc"):
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, conv_backend=backend), seed=11)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": False})
    for it in range(steps):
        lo, _ = to.step(x)
        lp = tp.step(x.cuda()).cpu()
        go, gp = to.last_grads, prod.engine.flat_grads().cpu()
        med = go.abs().median()
        flip = torch.sign(go) != torch.sign(gp)
        po, pp = flat_gamma_beta(to.model), prod.engine.flat_params().cpu()
        perr = (pp - po).abs()
        print(f"[{backend}] step {it}: logits relL2 {rel_l2(lp, lo):.2e} agree {((lp>=0)==(lo>=0)).float().mean():.6f} | "
              f"grad relL2 {rel_l2(gp, go):.2e} max|dg|/med {float((gp-go).abs().max()/med):.2e} flips {int(flip.sum())} "
              f"max|g|/med of flipped {float((go[flip].abs().max()/med) if flip.any() else 0):.2e} | "
              f"param err max {float(perr.max()):.2e} p99 {float(perr.quantile(0.99)):.2e} median {float(perr.median()):.2e} "
              f"n>1e-4 {int((perr>1e-4).sum())}")
