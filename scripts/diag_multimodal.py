"""Diagnostic (GPU): gradient error of the multimodal TENT step vs the CPU oracle for different gradient formats."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import MultimodalUNetB200, TentB200
from multimodal_tta_b200.synthetic import brats_volume
from oracle.multimodal_oracle import MULTIMODAL_MODEL_CFG, OracleMultimodalUNet
from oracle.tent_oracle import TentOracle
from tests.util import rel_l2

S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = brats_volume(1, (S, S, S), seed=20)
torch.manual_seed(7)
oracle = OracleMultimodalUNet.from_cfg(MULTIMODAL_MODEL_CFG)
sd = copy.deepcopy(oracle.state_dict())
to = TentOracle(oracle, mode="sigmoid")
lo, _ = to.step(x)
go = to.last_grads
names = to.param_names
for opts in ({}, {"loss_scale_mult": 64.0}, {"loss_scale_mult": 4096.0}, {"bwd_precision": "bf16x2"}, {"conv_backend": "simt"}):
    prod = MultimodalUNetB200(dict(MULTIMODAL_MODEL_CFG, deterministic=True, **opts))
    prod.load_state_dict(copy.deepcopy(sd)); prod.to("cuda")
    tp = TentB200(prod, {"cuda_graph": False})
    lp = tp.step(x.cuda()).cpu()
    gp = prod.engine.flat_grads().cpu()
    # per-layer relative error (gamma grads of each norm layer, in flat order)
    off, worst = 0, []
    for nl in prod.engine.norm_layers:
        a, b = gp[off:off + nl.C], go[off:off + nl.C]
        worst.append((rel_l2(a, b), nl.name)); off += nl.C
    worst.sort(reverse=True)
    print(opts, f"logits {rel_l2(lp, lo):.2e} grad {rel_l2(gp, go):.2e} flips {int((torch.sign(gp) != torch.sign(go)).sum())}",
          "worst:", [(f"{e:.1e}", n[-40:]) for e, n in worst[:3]])
