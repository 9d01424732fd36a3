import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tests.test_step_parity_gpu as T
from tests.util import make_pair, rel_l2
from multimodal_tta_b200 import TentB200
from oracle.tent_oracle import TentOracle
from oracle.unet_oracle import BRATS_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume
cuda = torch.device("cuda")
pre = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] else []
for name in pre:
    fn = getattr(T, name)
    if name == "test_brats_64cube_tight":
        fn(cuda, "tc"); fn(cuda, "simt")
    elif name == "test_brats_resunit_instance_norm_sigmoid":
        fn(cuda, False); fn(cuda, True)
    else:
        fn(cuda)
    print("ran", name, flush=True)
for rep in range(3):
    x = brats_volume(1, (32, 32, 32), seed=45)
    cfg = dict(BRATS_MODEL_CFG, conv_backend="simt")
    oracle, prod = make_pair(cfg, seed=11)
    to = TentOracle(oracle, mode="sigmoid")
    tp = TentB200(prod, {"entropy": "sigmoid", "cuda_graph": False})
    lo, loss_o = to.step(x)
    lp = tp.step(x.cuda()).cpu()
    g_o, g_p = to.last_grads, prod.engine.flat_grads().cpu()
    print("rep", rep, "logits", rel_l2(lp, lo), "grads", rel_l2(g_p, g_o))
    eng = prod.engine
    off = 0
    nl = eng.norm_layers
    Cs = [n.C for n in nl]
    tot = sum(Cs)
    o = 0
    for i, n in enumerate(nl):
        dg_p, dg_o = g_p[o:o + n.C], g_o[o:o + n.C]
        db_p, db_o = g_p[tot + o:tot + o + n.C], g_o[tot + o:tot + o + n.C]
        print(f"  {i:2d} {n.name:45s} dgamma {rel_l2(dg_p, dg_o):.2e} dbeta {rel_l2(db_p, db_o):.2e}")
        o += n.C
