"""One backward of the supervised step (for ncu on the weight-gradient kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_tta_b200 import UNetB200  # noqa: E402
from multimodal_tta_b200.presets import BRATS_MODEL_CFG  # noqa: E402
from multimodal_tta_b200.synthetic import brats_volume  # noqa: E402

dev = torch.device("cuda", 0)
model = UNetB200(dict(BRATS_MODEL_CFG, trainable=True)).to(dev).train()
x = brats_volume(2, (128, 128, 128), seed=1).to(dev)
for _ in range(2):
    logits = model(x)
    logits.square().mean().backward()
torch.cuda.synchronize()
print("ok", sorted(set(model.engine.plans[(2, 128, 128, 128)].wgrad_backends.values())))
