"""Parameter / logits agreement of the continual stream case (diagnostic for the test tolerances)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import TentB200
from multimodal_tta_b200.synthetic import brats_volume, domain_shift
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import BRATS_MODEL_CFG
from tests.util import make_pair, rel_l2
for dims in ((32, 32, 32), (48, 48, 48)):
    for B in (1, 2, 4):
        oracle, prod = make_pair(BRATS_MODEL_CFG, seed=41)
        to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"entropy": "sigmoid", "cuda_graph": True})
        stream = [domain_shift(brats_volume(B, dims, seed=300 + t), domain=t // 2, seed=7) for t in range(6)]
        ref = [to.step(x)[0] for x in stream]
        got = [o.clone().cpu() for o in tp.adapt_stream([x.pin_memory() for x in stream])]
        rl = [rel_l2(a, b) for a, b in zip(got, ref)]
        ag = [((a >= 0) == (b >= 0)).float().mean().item() for a, b in zip(got, ref)]
        perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
        print(dims, B, "rel", ["%.1e" % r for r in rl], "agree", ["%.5f" % a for a in ag],
              "median %.2e frac>1e-4 %.3f max %.2e" % (float(perr.median()), float((perr > 1e-4).float().mean()), float(perr.max())))
