"""Where does the end-to-end loop lose time against the device-resident loop?  Times, on one GPU and in one
process (same box): resident steps; adapt_stream alone; + loss D2H; + Dice counts + D2H (= bench.py's e2e)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_tta_b200 import TentB200, UNetB200  # noqa: E402
from multimodal_tta_b200.evaluation import device_dice_counts  # noqa: E402
from multimodal_tta_b200.presets import BRATS_MODEL_CFG  # noqa: E402
from multimodal_tta_b200.synthetic import brats_volume, region_labels  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B, DIMS, K, NROT = 2, (128, 128, 128), 40, 4
model = UNetB200(dict(BRATS_MODEL_CFG)).to(dev)
tent = TentB200(model, {"entropy": "sigmoid", "cuda_graph": True})
xs_host = [brats_volume(B, DIMS, seed=100 + i).pin_memory() for i in range(NROT)]
xs = [x.to(dev) for x in xs_host]
labels = [region_labels(B, 3, DIMS, seed=200 + i).to(dev) for i in range(NROT)]
print("labels dtype", labels[0].dtype)
loss_host = torch.zeros(K).pin_memory()
counts_host = torch.zeros((K, B, 3, 3), dtype=torch.int64).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(label, fn):
    fn(4)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0.record()
        fn(K)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / K)
    print(f"{label:50s} {best:.4f} ms/step = {B / best * 1e3:.1f} volumes/s", flush=True)


def resident(n):
    for i in range(n):
        tent.step(xs[i % NROT])


def stream_only(n):
    for _ in tent.adapt_stream([xs_host[i % NROT] for i in range(n)]):
        pass


def stream_loss(n):
    for j, _ in enumerate(tent.adapt_stream([xs_host[i % NROT] for i in range(n)])):
        loss_host[j:j + 1].copy_(tent.last_loss, non_blocking=True)


def stream_full(n):
    for j, logits in enumerate(tent.adapt_stream([xs_host[i % NROT] for i in range(n)])):
        loss_host[j:j + 1].copy_(tent.last_loss, non_blocking=True)
        counts_host[j].copy_(device_dice_counts(logits, labels[j % NROT], 0.5), non_blocking=True)


def resident_dice(n):
    for i in range(n):
        logits = tent.step(xs[i % NROT])
        counts_host[i].copy_(device_dice_counts(logits, labels[i % NROT], 0.5), non_blocking=True)


def h2d_only(n):
    st = torch.empty_like(xs[0])
    for i in range(n):
        st.copy_(xs_host[i % NROT], non_blocking=True)


timed("resident steps", resident)
timed("H2D copies alone (67 MB each)", h2d_only)
timed("adapt_stream (H2D prefetch) only", stream_only)
timed("adapt_stream + loss D2H", stream_loss)
timed("adapt_stream + loss D2H + dice counts (bench e2e)", stream_full)
timed("resident steps + dice counts", resident_dice)

side = torch.cuda.Stream()
counts_dev = torch.zeros((K, B, 3, 3), dtype=torch.int64, device=dev)
loss_dev = torch.zeros(K, device=dev)
events = [torch.cuda.Event() for _ in range(K)]


def stream_full_side(n):
    cur = torch.cuda.current_stream()
    counts_dev.zero_()
    for j, logits in enumerate(tent.adapt_stream([xs_host[i % NROT] for i in range(n)])):
        loss_dev[j:j + 1].copy_(tent.last_loss.reshape(1))
        device_dice_counts(logits, labels[j % NROT], 0.5, out=counts_dev[j])
        events[j].record(cur)
        with torch.cuda.stream(side):
            side.wait_event(events[j])
            loss_host[j:j + 1].copy_(loss_dev[j:j + 1], non_blocking=True)
            counts_host[j].copy_(counts_dev[j], non_blocking=True)
    cur.wait_stream(side)


timed("bench e2e with side-stream read-back", stream_full_side)
