"""Round-1 verdict task 8: what does dropping the A_lo * B_hi product of the forward convs (tta_conv_tc flags bit 17:
activations rounded to fp16, weights still hi + lo) cost in parity and buy in time?  cfg 2 (2x4x128^3), two TENT
steps against the CPU oracle; then the device-timed step."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import TentB200, UNetB200
from multimodal_tta_b200.presets import BRATS_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import OracleUNet


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


xs = [brats_volume(2, (128, 128, 128), seed=42 + i) for i in range(2)]
for flags in (0, 131072):
    torch.manual_seed(11)
    oracle = OracleUNet.from_cfg(BRATS_MODEL_CFG)
    prod = UNetB200(dict(BRATS_MODEL_CFG, deterministic=True, tc_flags=flags))
    prod.load_state_dict(copy.deepcopy(oracle.state_dict()))
    prod.to("cuda")
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"entropy": "sigmoid", "cuda_graph": True})
    for it, x in enumerate(xs):
        lo, _ = to.step(x)
        lp = tp.step(x.cuda()).cpu()
        g_o, g_p = to.last_grads, prod.engine.flat_grads().cpu()
        perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
        print(f"flags {flags} step {it}: logits rel-L2 {rel(lp, lo):.2e} max {float((lp - lo).abs().max() / lo.abs().max()):.2e} "
              f"agreement {((lp >= 0) == (lo >= 0)).float().mean().item():.6f} grad rel-L2 {rel(g_p, g_o):.2e} "
              f"flips {int((torch.sign(g_o) != torch.sign(g_p)).sum())} params>1e-4 {100 * float((perr > 1e-4).float().mean()):.2f} %",
              flush=True)
    prod2 = UNetB200(dict(BRATS_MODEL_CFG, tc_flags=flags)).to("cuda")
    t2 = TentB200(prod2, {"entropy": "sigmoid", "cuda_graph": True})
    xd = [x.cuda() for x in xs]
    for i in range(5):
        t2.step(xd[i % 2])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(40):
        t2.step(xd[i % 2])
    b.record(); torch.cuda.synchronize()
    print(f"flags {flags}: {a.elapsed_time(b) / 40:.4f} ms per step", flush=True)
