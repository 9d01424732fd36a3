"""one supervised step (reference step shape: zero_grad, forward, DiceCE, backward, Adam) of the BraTS UNet
on a 2x4x128^3 batch -- unet_b200(trainable) with the reference's torch loss / optimizer objects vs the same step
through torch eager (cuDNN) on the same GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import UNetB200
from multimodal_tta_b200.presets import BRATS_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, region_labels
from oracle.dicece_oracle import dice_ce_loss, make_optimizer
from oracle.unet_oracle import OracleUNet

dev = torch.device("cuda")
B, dims = 2, (128, 128, 128)
x = brats_volume(B, dims, seed=1).to(dev); y = region_labels(B, 3, dims, seed=2).to(dev)


def timed(model, opt, n=5):
    def step():
        opt.zero_grad()
        loss = dice_ce_loss(model(x), y)
        loss.backward()
        opt.step()
        return loss
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        loss = step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, float(loss)


torch.manual_seed(0)
prod = UNetB200(dict(BRATS_MODEL_CFG, trainable=True)).to(dev).train()
ms, loss = timed(prod, make_optimizer(prod))
print(f"unet_b200 (trainable) supervised step, 2x4x128^3: {ms:.2f} ms per step = {B * 1e3 / ms:.1f} volumes/s (loss {loss:.4f})")
# where the time goes: forward / backward (dgrad + norm + wgrad) / loss+optimizer in torch
eng = prod.engine
plan = eng.plans[(B, *dims)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    eng.forward_train(x)
torch.cuda.synchronize(); f_ms = (time.perf_counter() - t0) / 3 * 1e3
g = torch.randn(B, 3, *dims, device=dev) * 1e-7
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3):
    eng.backward_from_logits_grad(plan, g)
torch.cuda.synchronize(); b_ms = (time.perf_counter() - t0) / 3 * 1e3
print(f"   forward {f_ms:.2f} ms, backward incl. weight gradients {b_ms:.2f} ms ({plan.n_wgrad} wgrad/bias launches), "
      f"rest (torch loss + optimizer + device repack) {ms - f_ms - b_ms:.2f} ms")
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32; torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    ref = OracleUNet.from_cfg(BRATS_MODEL_CFG).to(dev).train()
    ms_r, _ = timed(ref, make_optimizer(ref), n=3)
    print(f"torch eager ({'TF32' if tf32 else 'fp32, TF32 off = reference setting'}): {ms_r:.2f} ms per step = {B * 1e3 / ms_r:.1f} volumes/s")
    del ref
# ---- in-stream duration of the backward ops (top 12)
ops = list(plan.bwd)
torch.cuda._sleep(40_000_000)
evs = []
for op in ops:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op(); e1.record(); evs.append((e0, e1))
torch.cuda.synchronize()
rows = sorted(((a.elapsed_time(b) * 1e3, getattr(op, "label", "?")) for (a, b), op in zip(evs, ops)), reverse=True)
print("backward ops, top 12 (us):")
for us, lab in rows[:12]:
    print(f"  {us:8.1f}  {lab}")
print(f"  wgrad total {sum(u for u, l in rows if l.startswith('wgrad')) / 1e3:.2f} ms")
