"""SURVEY 8f-2 at BraTS size: one TENT step of the multimodal model (4 modality encoders, bottleneck fusion,
trilinear-upsample decoder; 83 M parameters, 18 816 adaptable scalars) on a 4x128^3 volume, device-timed, with the
in-stream duration of every op; optional CPU-oracle time for the same step (--cpu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_tta_b200 import MultimodalUNetB200, TentB200
from multimodal_tta_b200.synthetic import brats_volume

CFG = dict(name="unet_multimodal_deepfusion", num_modalities=4, num_classes=3, spatial_dims=3,
           channels=[32, 64, 128, 256, 512], strides=[2, 2, 2, 2], num_res_units=2, norm="INSTANCE", act="RELU", dropout=0.0)
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1
dev = torch.device("cuda")
torch.manual_seed(0)
model = MultimodalUNetB200(dict(CFG)).to(dev)
tent = TentB200(model, {"entropy": "sigmoid"})
xs = [brats_volume(B, (128, 128, 128), seed=s).to(dev) for s in (1, 2)]
for i in range(3):
    tent.step(xs[i % 2])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
a.record()
for i in range(K):
    tent.step(xs[i % 2])
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / K
print(f"multimodal TENT step, batch {B} x 4x128^3: {ms:.2f} ms per step = {B * 1e3 / ms:.1f} adapted volumes/s, "
      f"{tent.gpu_launches_per_step} launches per step, loss {float(tent.last_loss):.5f}")
eng = model.engine
plan = eng.plans[(B, 128, 128, 128)]
ops = [lambda: eng._pack_input(plan, xs[0])] + list(plan.fwd) + [plan.head_train] + list(plan.bwd)
torch.cuda._sleep(40_000_000)
evs = []
for op in ops:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); op(); e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
rows = sorted(((e0.elapsed_time(e1) * 1e3, getattr(op, "label", getattr(op, "__qualname__", "?").split(".")[-1]))
               for (e0, e1), op in zip(evs, ops)), reverse=True)
tot = sum(r[0] for r in rows)
print(f"in-stream sum {tot / 1e3:.2f} ms over {len(rows)} ops; top 25:")
for us, lab in rows[:25]:
    print(f"  {us:8.1f} us  {lab}")
if "--cpu" in sys.argv:
    from oracle.multimodal_oracle import OracleMultimodalUNet
    from oracle.tent_oracle import TentOracle
    torch.set_num_threads(os.cpu_count() or 1)
    to = TentOracle(OracleMultimodalUNet.from_cfg(CFG), mode="sigmoid")
    x = brats_volume(1, (128, 128, 128), seed=1)
    to.step(x)
    t0 = time.perf_counter(); to.step(x); dt = time.perf_counter() - t0
    print(f"CPU oracle, same step, B=1, {os.cpu_count()} threads: {dt:.2f} s per step = {1 / dt:.3f} volumes/s")
