#!/usr/bin/env python
"""Headline benchmark: adapted 3-D volumes / second for one full TTA step (forward + entropy
backward + norm-affine Adam update) of the BraTS-shaped 4x128^3 res-unit UNet, batch 2 per GPU
(BASELINE.json configs[1]); weak scaling over GPUs (each rank adapts its own 2 volumes, one
all-reduce of the 4 870 affine gradients per step keeps parameters identical).

    python bench.py --gpus N --steps K --warmup W            # product arm (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle port on the host cores

Prints ONE JSON line (contract in the task statement): value = device-timed whole-job throughput
with inputs resident in HBM; e2e = same metric through TentB200.step with pinned HOST input
(H2D inside the timed region) and a D2H read of the loss; roofline = conv kernels (tensor) plus
norm/entropy kernels (HBM) timed live with CUDA events; cpu_baseline = oracle on the host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "adapted 3D volumes/sec (TTA fwd+bwd+update)"
UNIT = "volumes/s"
DIMS = (128, 128, 128)
BATCH = 2
# algorithmic work per adapted 4x128^3 volume
CONV_GFLOP_PER_VOLUME = 190.8
NORM_ELEMS_PER_VOLUME = 51.1e6
LOGIT_ELEMS_PER_VOLUME = 3 * 128 ** 3


def measured_traffic():
    """Per-step DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, summed over the launches of
    a kernel family in ONE step) from the committed ncu launch list -- profiles/traffic_r1.json is
    written by scripts/summarize_step.py; None when the file is absent."""
    for name in ("traffic_r2.json", "traffic_r1.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            return d.get("conv_dram_bytes_per_step"), d.get("stream_dram_bytes_per_step"), name
    return None, None, None


def measured_tensor_pipe():
    """Duration-weighted tensor-pipe utilisation of the conv launches of one step (ncu, written by
    scripts/summarize_tensor_pipe.py to profiles/tensor_pipe_r1.json); None when the file is absent."""
    for name in ("tensor_pipe_r2.json", "tensor_pipe_r1.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            out = {k: d.get(k) for k in ("tensor_pipe_active_pct", "tc_pipe_active_pct", "sm_ghz_under_load")}
            out["source"] = "profiles/" + name
            return out
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sust=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Launch nvidia-smi (20 ms period) and wait for its first row, so that sampling is already
        running when the timed region starts (its start-up alone outlasts a 50 ms timed region)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t = time.time()
            while not self.rows and time.time() - t < 5.0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t))]
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1]); pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": statistics.median(pw) if pw else None,
                "power_w_max": max(pw) if pw else None}


def cpu_oracle_rate(steps: int, warmup: int, batch: int = 1, dims=DIMS):
    """Oracle (CPU port of the reference path) volumes/s on the host cores, all threads."""
    import torch
    from oracle.tent_oracle import TentOracle
    from oracle.unet_oracle import BRATS_MODEL_CFG, OracleUNet
    from multimodal_tta_b200.synthetic import brats_volume

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    tent = TentOracle(OracleUNet.from_cfg(BRATS_MODEL_CFG), mode="sigmoid")
    x = brats_volume(batch, dims, seed=42)
    for _ in range(warmup):
        tent.step(x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tent.step(x)
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return batch / sec, sec, cores, torch.get_num_threads()


def bind_to_gpu_numa(local: int):
    """Pin this rank's host threads to the CPUs next to its GPU (`nvidia-smi topo -m`, "CPU Affinity" column) BEFORE
    any pinned buffer is allocated, so the staging pages are first-touched on that NUMA node.  Eight ranks each
    pulling 67 MB per 2.3 ms step out of one node's memory is what held the 8-GPU end-to-end number at 0.86 of
    the device-timed one in round 1.  Best effort: returns what was done for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        import re
        ansi = re.compile(r"\x1b\[[0-9;]*m")
        lines = [ansi.sub("", ln) for ln in out.splitlines()]
        header = next(ln for ln in lines if "CPU Affinity" in ln)
        cols = [c.strip() for c in header.split("\t")]
        idx = cols.index("CPU Affinity")
        row = next(ln for ln in lines if ln.startswith(f"GPU{local}\t") or ln.startswith(f"GPU{local} "))
        cells = [c.strip() for c in row.split("\t")]
        spec = cells[idx]
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus |= set(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "affinity list empty after intersecting with the allowed CPUs"}
        os.sched_setaffinity(0, cpus)
        numa = cells[idx + 1] if idx + 1 < len(cells) else None
        return {"bound": True, "cpus": spec, "numa_node": numa}
    except Exception as e:  # no topology information: stay unbound
        return {"bound": False, "why": f"{type(e).__name__}: {e}"[:120]}


def gpu_eager_baselines(torch, dev, steps: int = 5, warmup: int = 2):
    """The bar a hand-written path has to beat on THIS GPU: the oracle's TENT step
    (plain torch.nn modules -> cuDNN / ATen on sm_100) on the same 2x4x128^3 batch, in the reference's precision
    setting (fp32, TF32 off: /root/reference/src/utils/metrics.py:60-67), with TF32 on, and as bf16 autocast in
    channels_last_3d.  Runs after every timed region of the product arm, is never imported by the package, and
    is a reported baseline like cpu_baseline -- not a parity check and not the product."""
    from oracle.tent_oracle import TentOracle
    from oracle.unet_oracle import BRATS_MODEL_CFG as CFG, OracleUNet
    from multimodal_tta_b200.synthetic import brats_volume

    x = brats_volume(BATCH, DIMS, seed=100).to(dev)
    out = {}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    for name, tf32, bf16 in (("fp32_no_tf32", False, False), ("tf32", True, False), ("bf16_channels_last_3d", True, True)):
        try:
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            torch.manual_seed(42)
            model = OracleUNet.from_cfg(CFG).to(dev)
            xi = x
            if bf16:
                model = model.to(memory_format=torch.channels_last_3d)
                xi = x.contiguous(memory_format=torch.channels_last_3d)
            tent = TentOracle(model, mode="sigmoid")

            def one():
                if bf16:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        tent.step(xi)
                else:
                    tent.step(xi)
            for _ in range(warmup):
                one()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": BATCH / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
            del model, tent
            torch.cuda.empty_cache()
        except Exception as e:
            out[name] = {"unavailable": f"{type(e).__name__}: {e}"[:160]}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["note"] = ("oracle TENT step through torch eager (cuDNN/ATen, autograd incl. the float(loss) sync of the "
                   f"reference's step shape), batch {BATCH} x 4x128^3, {warmup} warm-up + {steps} timed steps; "
                   "TentOracle.step's loss.item() is inside, as in seg_trainer.py:145")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    v, sec, cores, threads = cpu_oracle_rate(steps, warm, batch=1)
    sample = f"B=1 4x128^3 TENT step (Bernoulli entropy, Adam), {warm} warm-up + {steps} timed steps, median"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BraTS-shaped 4x128^3 res-unit 3D UNet (in 4, out 3, INSTANCE norm), TENT "
                               "norm-affine-only adaptation, CPU oracle port of the reference path",
                   "batch_per_step": 1},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "threads": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_product(args):
    import torch
    import torch.distributed as dist

    from multimodal_tta_b200 import TentB200, UNetB200
    from multimodal_tta_b200.synthetic import brats_volume
    from multimodal_tta_b200.presets import BRATS_MODEL_CFG

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if not args.no_numa_bind else {"bound": False, "why": "--no-numa-bind"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.manual_seed(42)
    overrides = {}
    for kv in args.set or []:     # experiment switches, e.g. --set fuse_stats=false
        k, v = kv.split("=", 1)
        overrides[k] = {"true": True, "false": False}.get(v.lower(), v)
    model = UNetB200(dict(BRATS_MODEL_CFG, conv_backend=args.conv_backend, **overrides)).to(dev)
    tent = TentB200(model, {"entropy": "sigmoid", "cuda_graph": not args.no_graph})
    eng = model.engine
    NROT = 4  # distinct resident input batches (4 x 67 MB > 126 MB L2)
    xs_host = [brats_volume(BATCH, DIMS, seed=100 + rank * 16 + i).pin_memory() for i in range(NROT)]
    xs = [x.to(dev) for x in xs_host]
    K, Wm = args.steps, max(3, args.warmup)

    for i in range(Wm):
        tent.step(xs[i % NROT])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for i in range(K):
        tent.step(xs[i % NROT])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * BATCH * K / (ms_max / 1e3)
    launches = tent.gpu_launches_per_step * K

    # ---- end to end through the public API: pinned HOST batches in, loss read back to the host.
    # TentB200.adapt_stream prefetches batch i+1 (H2D on a copy stream) while batch i adapts; both
    # the H2D copies and the D2H loss reads are inside the timed region.
    # The loss of every step is copied to pinned host memory by a stream-ordered asynchronous D2H copy
    # (one slot per step) and all of them are complete when the closing barrier returns: the host never
    # stalls the device between steps, which is how a streaming consumer uses adapt_stream.
    # What the consumer of an adapted volume needs -- the region Dice counts of its pre-update logits
    # (seg_eval.py:304-308, 41-68) -- is computed on the device by tta_dice_counts right behind every step and
    # leaves as 144 bytes ([B,R,3] int64); the labels of the rotating batches are resident (they are the
    # evaluator's input, not the adaptation step's).
    from multimodal_tta_b200.evaluation import device_dice_counts
    from multimodal_tta_b200.synthetic import region_labels
    labels = [region_labels(BATCH, 3, DIMS, seed=200 + i).to(dev) for i in range(NROT)]
    loss_host = torch.zeros(max(K, 4)).pin_memory()
    counts_host = torch.zeros((max(K, 4), BATCH, 3, 3), dtype=torch.int64).pin_memory()

    # The per-step read-back runs on a side stream: the step's loss (4-byte D2D snapshot) and Dice counts land in
    # per-step device slots, an event orders the two D2H copies behind them, and the adaptation stream goes on with
    # the next batch instead of idling through two DMA round trips (scripts/e2e_diag.py: ~25 us each).  The main stream
    # waits for the side stream before the closing event, so every copy is inside the timed region.
    side = torch.cuda.Stream(device=dev)
    n_slots = max(K, 4)
    counts_dev = torch.zeros((n_slots, BATCH, 3, 3), dtype=torch.int64, device=dev)
    loss_dev = torch.zeros(n_slots, device=dev)
    slot_events = [torch.cuda.Event() for _ in range(n_slots)]

    def e2e_loop(n, non_blocking, hosts=None):
        hosts = xs_host if hosts is None else hosts
        cur = torch.cuda.current_stream(dev)
        counts_dev.zero_()
        for j, logits in enumerate(tent.adapt_stream([hosts[i % NROT] for i in range(n)])):
            loss_dev[j:j + 1].copy_(tent.last_loss.reshape(1))
            device_dice_counts(logits, labels[j % NROT], 0.5, out=counts_dev[j])
            slot_events[j].record(cur)
            with torch.cuda.stream(side):
                side.wait_event(slot_events[j])
                loss_host[j:j + 1].copy_(loss_dev[j:j + 1], non_blocking=non_blocking)
                counts_host[j].copy_(counts_dev[j], non_blocking=non_blocking)
        cur.wait_stream(side)
    e2e_loop(4, False)
    barrier()
    e0.record()
    e2e_loop(K, True)
    e1.record()
    barrier()
    assert bool(torch.isfinite(loss_host[:K]).all()) and float(loss_host[:K].abs().min()) > 0.0
    assert int(counts_host[:K, :, :, 2].min()) > 0          # every region's ground truth is non-empty
    sampler.mark_end()     # clocks are sampled over both timed regions (device-resident and end-to-end)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * K / (float(t.item()) / 1e3)
    h2d = xs_host[0].numel() * 4
    d2h = 4 + counts_host[0].numel() * 8

    # ---- the same end-to-end loop with the host batches staged as fp16 (TentB200 accepts fp16 batches: the gather
    # kernel reads them directly): half the host -> device bytes.  Reported next to the fp32 number, not instead of
    # it -- fp32 is the reference's batch dtype; fp16 staging rounds the input to 11 bits.
    xs_host16 = [x.half().pin_memory() for x in xs_host]
    e2e_loop(4, False, xs_host16)
    barrier()
    e0.record()
    e2e_loop(K, True, xs_host16)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e16_value = world * BATCH * K / (float(t.item()) / 1e3)

    # ---- sustained: the same device-resident loop for >= args.sustain_s seconds (the 20-step headline region
    # lasts ~50 ms: a burst number), clocks and power sampled over exactly that region
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(K, int(args.sustain_s * 1e3 / (ms_max / K)) + 1)
        sam2 = ClockSampler(local)
        if rank == 0:
            sam2.start()
        barrier()
        sam2.mark_begin()
        e0.record()
        for i in range(n_sus):
            tent.step(xs[i % NROT])
        e1.record()
        barrier()
        sam2.mark_end()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sus_ms = float(t.item())
        sustained = {"value": world * BATCH * n_sus / (sus_ms / 1e3), "unit": UNIT, "steps": n_sus,
                     "seconds": sus_ms / 1e3, "ms_per_step": sus_ms / n_sus,
                     "clocks": sam2.stop() if rank == 0 else None}

    # ---- live per-kernel-family timing (eager, CUDA events on the launching stream)
    roof = kernel_family_times(torch, eng, tent, xs, min(K, 3), args.per_op if rank == 0 else None)
    pk = peaks()
    conv_tflops = BATCH * CONV_GFLOP_PER_VOLUME / 1e3 / (roof["conv_ms"] / 1e3)
    # algorithmic HBM bytes (SURVEY 8d at s = 4 B/element): norm fwd 2*N*s + bwd 3*N*s, entropy 2*C*V*4
    hbm_bytes = BATCH * (NORM_ELEMS_PER_VOLUME * (2 * 4 + 3 * 4) + 2 * LOGIT_ELEMS_PER_VOLUME * 4)
    hbm_gbs = hbm_bytes / 1e9 / (roof["stream_ms"] / 1e3)

    # the event pairs cost ~15 % (their family times sum to more than the graph-replayed step): the same two
    # rooflines with each family's time taken as its SHARE of the device-timed step are reported next to the
    # event-timed ones (`in_step`), so that a reader does not have to redo that division
    step_ms = ms_max / K
    conv_ms_in_step = step_ms * roof["conv_ms"] / roof["total_ms"]
    stream_ms_in_step = step_ms * roof["stream_ms"] / roof["total_ms"]
    conv_tflops_in_step = BATCH * CONV_GFLOP_PER_VOLUME / 1e3 / (conv_ms_in_step / 1e3)
    hbm_gbs_in_step = hbm_bytes / 1e9 / (stream_ms_in_step / 1e3)
    conv_traffic, stream_traffic, traffic_src = measured_traffic()
    if rank == 0:
        if args.skip_cpu:
            cpu = None
        else:
            v, sec, cores, threads = cpu_oracle_rate(8, 1, batch=1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "threads": threads, "kind": "port",
                   "sample": "oracle TENT step, B=1 4x128^3, 1 warm-up + 8 timed steps (median), all host threads"}
        gpu_base = None
        if world == 1 and not args.skip_gpu_baseline:
            gpu_base = gpu_eager_baselines(torch, dev)
        backends = sorted(set(eng.plans[(BATCH, *DIMS)].conv_backends.values()))
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 hi/lo split operands fwd (3 products), scaled fp16 bwd, f32 accumulate" if "tc" in backends else "f32",
            "data": "synthetic",
            "config": {"workload": "BraTS-shaped 4x128^3 res-unit 3D UNet (in 4, out 3, INSTANCE norm), TENT "
                                   "norm-affine-only adaptation (Bernoulli entropy, Adam lr 1e-3), batch 2 per GPU",
                       "batch_per_gpu": BATCH, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                       "conv_backends": backends, "cuda_graph": not args.no_graph,
                       "l2": "per-step working set ~1.6 GB >> 126 MB L2; inputs rotate over 4 resident "
                             "batches (268 MB)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "pinned host batch -> H2D (prefetched on a copy stream) -> TentB200.adapt_stream -> "
                            "tta_dice_counts on the pre-update logits -> async D2H (side stream, event-ordered) of every "
                            "step's loss and [B,R,3] Dice counts; all copies inside the timed region", "numa": numa},
            "e2e_fp16_staging": {"value": e2e16_value, "unit": UNIT, "h2d_bytes_per_step": h2d // 2,
                                 "d2h_bytes_per_step": d2h,
                                 "note": "same loop, host batches staged as fp16 (input rounded to 11 bits)"},
            "sustained": sustained,
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": f"conv3d fwd+dgrad ({roof['conv_launches']} launches/step)",
                         "achieved": conv_tflops, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                         "frac": conv_tflops / pk["tf_sust"], "traffic": conv_traffic, "peak_source": pk["src"],
                         "traffic_note": f"DRAM bytes of all conv launches of one step (ncu, profiles/{traffic_src})",
                         "ms_per_step": roof["conv_ms"], "share_of_step": roof["conv_ms"] / roof["total_ms"],
                         "in_step": {"ms_per_step": conv_ms_in_step, "achieved": conv_tflops_in_step,
                                     "frac": conv_tflops_in_step / pk["tf_sust"],
                                     "note": "family time = share_of_step x the device-timed graph step"},
                         "algorithmic_note": "achieved = algorithmic conv FLOPs (190.8 GFLOP/volume); the fp32-equivalent "
                                             "forward executes 3-4 fp16 products per algorithmic MAC",
                         "ncu_tensor_pipe": measured_tensor_pipe()},
            "roofline_hbm": {"bound": "hbm", "kernel": "norm stats/apply/bwd + fused entropy head",
                             "achieved": hbm_gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": hbm_gbs / pk["hbm"],
                             "traffic": stream_traffic, "algorithmic_bytes": hbm_bytes,
                             "ms_per_step": roof["stream_ms"],
                             "share_of_step": roof["stream_ms"] / roof["total_ms"],
                             "in_step": {"ms_per_step": stream_ms_in_step, "achieved": hbm_gbs_in_step,
                                         "frac": hbm_gbs_in_step / pk["hbm"]}},
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_base,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_sliding(args):
    """BASELINE configs[3] (and its 48-window variant): ONE BraTS full volume 4x155x240x240 through sliding-window
    TENT, the windows of every global step sharded over the ranks (rank r takes the r-th block of sw_batch),
    one all-reduce of the affine gradients per step and one of the blended accumulators per volume.  Strong
    scaling: the volume is fixed, so the ideal speed-up is windows / ceil(windows / ranks)."""
    import torch
    import torch.distributed as dist

    from multimodal_tta_b200 import SlidingWindowTTA, TentB200, UNetB200
    from multimodal_tta_b200.presets import BRATS_MODEL_CFG
    from multimodal_tta_b200.synthetic import brats_volume

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    roi = (96, 96, 96) if args.workload == "cfg4_96" else (128, 128, 128)
    torch.manual_seed(42)
    model = UNetB200(dict(BRATS_MODEL_CFG)).to(dev)
    tent = TentB200(model, {"entropy": "sigmoid", "cuda_graph": not args.no_graph})
    sw = SlidingWindowTTA(tent, roi, sw_batch=args.sw_batch, overlap=0.5)
    sw.time_collectives = True
    vols = [brats_volume(1, (160, 240, 240), seed=3 + i)[:, :, :155].contiguous().to(dev) for i in range(2)]
    K, Wm = args.steps, max(3, args.warmup)
    for i in range(Wm):
        sw(vols[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sw.time_collectives = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        out = sw(vols[i % 2])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    sw.time_collectives = True
    ar = []
    for i in range(3):
        sw(vols[i % 2])
        ar.append(sw.last_allreduce_ms)
    assert bool(torch.isfinite(out).all()) and tuple(out.shape) == (1, 3, 155, 240, 240)
    if rank == 0:
        nwin = sw.last_num_windows
        print(json.dumps({
            "metric": "adapted full volumes/sec (sliding-window TTA)", "value": 1e3 / ms, "unit": "volumes/s",
            "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "data": "synthetic",
            "dtype": "fp16 hi/lo split operands fwd (3 products), scaled fp16 bwd, f32 accumulate",
            "config": {"workload": f"BraTS full volume 4x155x240x240, roi {roi[0]}^3, overlap 0.5, {nwin} windows, "
                                   f"sw_batch {args.sw_batch} per rank, windows sharded over {world} ranks",
                       "windows": nwin, "tent_steps_per_volume": sw.last_steps,
                       "ideal_speedup": nwin / (-(-nwin // (world * args.sw_batch)) * args.sw_batch)},
            "windows_per_s": nwin * 1e3 / ms,
            "accumulator_allreduce_ms": (statistics.median(ar) if ar and ar[0] is not None else None),
            "accumulator_allreduce_bytes": 4 * 4 * 155 * 240 * 240,
            "gpu_launches": tent.gpu_launches_per_step * sw.last_steps * K}))
    if world > 1:
        dist.destroy_process_group()


def kernel_family_times(torch, eng, tent, xs, reps: int, per_op_path=None):
    """Eager replay with a CUDA event pair around every op; returns per-step ms by family."""
    plan = eng.plans[(BATCH, *DIMS)]
    ops = [("stream", lambda: eng._pack_input(plan, xs[0]))]
    ops += [("conv" if getattr(o, "__qualname__", "").find("_conv_call") >= 0 else "stream", o) for o in plan.fwd]
    ops += [("stream", plan.head_train)]
    ops += [("conv" if getattr(o, "__qualname__", "").find("_conv_call") >= 0 else "stream", o) for o in plan.bwd]
    acc = {"conv": 0.0, "stream": 0.0}
    per_op = [0.0] * len(ops)
    nconv = sum(1 for k, _ in ops if k == "conv")
    for _ in range(reps):
        evs = []
        # park the stream behind a ~20 ms spin kernel while the host enqueues every op with its event
        # pair: the GPU then runs them back to back and an event pair brackets the kernel only (in
        # plain eager order the host is slower than the small kernels and every pair would also time
        # the launch gap in front of its kernel)
        torch.cuda._sleep(40_000_000)
        for kind, op in ops:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); op(); b.record()
            evs.append((kind, a, b))
        torch.cuda.synchronize()
        for i, (kind, a, b) in enumerate(evs):
            acc[kind] += a.elapsed_time(b)
            per_op[i] += a.elapsed_time(b)
    if per_op_path:
        # in-stream duration of every op of the step (warm L2, back to back) -- what the graph replays
        with open(per_op_path, "w") as f:
            for i, (kind, op) in enumerate(ops):
                lab = getattr(op, "label", getattr(op, "__qualname__", "?").split(".")[-1])
                f.write(f"{i:3d} {kind:6s} {per_op[i] / reps * 1e3:8.1f} us  {lab}\n")
    return {"conv_ms": acc["conv"] / reps, "stream_ms": acc["stream"] / reps,
            "total_ms": (acc["conv"] + acc["stream"]) / reps, "conv_launches": nconv}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--conv-backend", default="auto", choices=["auto", "tc", "simt"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--per-op", default=None, help="write the in-stream duration of every op of one step to this file")
    ap.add_argument("--set", action="append", help="model config override key=value (experiments)")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained loop in seconds (0: skip)")
    ap.add_argument("--skip-gpu-baseline", action="store_true", help="skip the torch-eager GPU baselines")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin host threads next to the GPU")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg4_96"],
                    help="cfg2 = the headline (BASELINE configs[1]); cfg4 / cfg4_96 = full-volume sliding window, "
                         "18 / 48 windows sharded over the ranks (not a bench line: a scaling measurement)")
    ap.add_argument("--sw-batch", type=int, default=1, help="windows per rank and TENT step (cfg4 workloads)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "cfg2":
        run_sliding(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
