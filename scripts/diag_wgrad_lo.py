"""Does the weight gradient need the lo plane of x?  All-parameter gradient error vs CPU autograd (64^3) and the
supervised backward time (2x4x128^3) with wgrad_x_lo = true / false, same process."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import UNetB200
from multimodal_tta_b200.presets import BRATS_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, region_labels
from oracle.dicece_oracle import dice_ce_loss
from oracle.unet_oracle import OracleUNet

dev = torch.device("cuda")
torch.manual_seed(81)
oracle = OracleUNet.from_cfg(BRATS_MODEL_CFG).train()
x = brats_volume(1, (64, 64, 64), seed=3); y = region_labels(1, 3, (64, 64, 64), seed=4)
dice_ce_loss(oracle(x), y).backward()
po = dict(oracle.named_parameters())
xb = brats_volume(2, (128, 128, 128), seed=1).to(dev); yb = region_labels(2, 3, (128, 128, 128), seed=2).to(dev)
for lo in (True, False, True, False):
    prod = UNetB200(dict(BRATS_MODEL_CFG, trainable=True, deterministic=True, wgrad_x_lo=lo))
    prod.load_state_dict(copy.deepcopy(oracle.state_dict()))
    prod.to(dev).train()
    dice_ce_loss(prod(x.to(dev)), y.to(dev)).backward()
    num = den = 0.0
    worst = (0.0, "")
    for k, p in prod.named_parameters():
        if k in po and po[k].grad is not None and p.grad is not None and k.endswith("weight") and p.grad.ndim == 5:
            d = (p.grad.cpu().double() - po[k].grad.double())
            num += float(d.pow(2).sum()); den += float(po[k].grad.double().pow(2).sum())
            r = float(d.norm() / po[k].grad.double().norm().clamp_min(1e-30))
            worst = max(worst, (r, k))
    for _ in range(2):
        prod.zero_grad(); dice_ce_loss(prod(xb), yb).backward()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        prod.zero_grad(); dice_ce_loss(prod(xb), yb).backward()
    b.record(); torch.cuda.synchronize()
    print(f"wgrad_x_lo={lo}: conv-weight gradient rel-L2 {(num / den) ** 0.5:.3e}, worst {worst[0]:.3e} ({worst[1]}); "
          f"forward + loss + backward at 2x4x128^3: {a.elapsed_time(b) / 5:.2f} ms", flush=True)
    del prod
