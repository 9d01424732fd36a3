"""``tent_b200`` -- the test-time-adaptation method plugin (TENT: entropy minimisation over the
norm-affine parameters only), executed by the sm_100a kernels behind ``TTAEngine``.

Step shape mirrors the reference's ``SegTrainer.run_step``
(/root/reference/src/core/trainers/seg_trainer.py:97-145: zero_grad -> forward -> loss ->
backward -> optimizer.step -> ``{"loss": float}``) with the canonical TENT recipe of SURVEY.md
section 8c-3: train-mode norm statistics, everything frozen except norm weight/bias (materialised as
gamma=1/beta=0 for ``InstanceNorm3d(affine=False)``), mean per-voxel entropy, Adam(lr=1e-3,
betas=(0.9, 0.999), eps=1e-8, no weight decay on 1-D params --
/root/reference/src/core/experiment_manager.py:199-237), logits returned from BEFORE the update.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import torch

from .config import DictConfig, create, get_config
from .registry import register_plugin
from .unet_b200 import B200Model, NormHolder

_MODES = {"softmax": 0, "sigmoid": 1}


@register_plugin("tent_b200")
class TentB200:
    def __init__(self, model: B200Model, cfg: Optional[DictConfig | Dict[str, Any]] = None, *,
                 process_group=None):
        if not isinstance(model, B200Model):
            raise TypeError("tent_b200 adapts a B200 model (set model.name=unet_b200 or "
                            f"unet_multimodal_deepfusion_b200); got {type(model).__name__}")
        cfg = cfg if isinstance(cfg, DictConfig) else create(dict(cfg or {}))
        self.model = model
        self.mode = str(get_config(cfg, "entropy", "auto"))
        if self.mode == "auto":  # sigmoid heads are what the reference trains (brats.yaml:12,59)
            self.mode = "sigmoid"
        if self.mode not in _MODES:
            raise ValueError(f"unknown entropy mode {self.mode!r} (softmax | sigmoid)")
        if self.mode == "softmax" and model.out_channels < 2:
            raise ValueError("softmax entropy is degenerate for a single-channel head; use sigmoid")
        self.lr = float(get_config(cfg, "lr", 1e-3))
        betas = list(get_config(cfg, "betas", [0.9, 0.999]))
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(get_config(cfg, "eps", 1e-8))
        self.steps = int(get_config(cfg, "steps", 1))
        self.episodic = bool(get_config(cfg, "episodic", False))
        self.use_graph = bool(get_config(cfg, "cuda_graph", model.use_cuda_graph))
        if self.steps < 1:
            raise ValueError("tent_b200: steps must be >= 1")
        self.pg = process_group
        self.configure_model()
        eng = model.engine
        eng.entropy_mode = _MODES[self.mode]
        eng.adam = dict(lr=self.lr, b1=self.betas[0], b2=self.betas[1], eps=self.eps)
        self._snapshot: Optional[torch.Tensor] = None
        self.last_loss: Optional[torch.Tensor] = None
        self.gpu_launches_per_step = 0

    # ------------------------------------------------------------------ TENT configure
    def configure_model(self) -> None:
        m = self.model
        m.train()
        m.requires_grad_(False)
        for h in m.norm_holders():
            h.materialize_affine()
            h.weight.requires_grad_(True)
            h.bias.requires_grad_(True)
            if h.kind == "batch":
                h.track_running_stats = False
        m._params_dirty = True

    def adaptable_parameters(self):
        return [p for h in self.model.norm_holders() for p in (h.weight, h.bias)]

    # ------------------------------------------------------------------ distributed
    @property
    def world_size(self) -> int:
        import torch.distributed as dist

        if self.pg is None and not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.pg)

    def _allreduce_grads(self, eng) -> float:
        """One SUM all-reduce of the flat [dgamma || dbeta] buffer; Adam divides by world size."""
        import torch.distributed as dist

        ws = self.world_size
        if ws > 1:
            dist.all_reduce(eng.dgb, op=dist.ReduceOp.SUM, group=self.pg)
        return 1.0 / ws

    # ------------------------------------------------------------------ state
    def reset(self) -> None:
        eng = self.model.engine
        if self._snapshot is not None:
            eng.gb.copy_(self._snapshot)
        eng.reset_optimizer()

    # ------------------------------------------------------------------ the step
    def _prepare(self, x: torch.Tensor):
        eng = self.model.engine
        x = eng._check_input(x)
        eng._ensure_device(x.device)
        if self._snapshot is None:
            self._snapshot = eng.gb.clone()
        if self.episodic:
            self.reset()
        return eng, x

    def _run(self, eng, plan, pack, gmul: float = 1.0, input_key=None) -> None:
        """``steps`` x (forward, entropy, backward, [all-reduce], Adam); logits of the first forward
        are kept in plan.logits_out.  ``input_key`` distinguishes graphs that bake different input
        pointers (the two staging buffers of adapt_stream); at most four are kept per plan."""
        with eng.on_device():
            self._run_on_device(eng, plan, pack, gmul, input_key)

    def _run_on_device(self, eng, plan, pack, gmul, input_key) -> None:
        ws = self.world_size
        single = ws == 1
        for it in range(self.steps):
            if self.use_graph:
                # the key names everything a captured step bakes in: who runs Adam, the entropy / Adam constants,
                # WHERE the input comes from (static copy, a caller-owned staging buffer, or a window gather with
                # its volume / window-table pointers) and the gradient multiplier of a padded tail batch
                key = (single, eng.entropy_mode, tuple(sorted(eng.adam.items())), input_key, round(float(gmul), 9))
                graphs = plan.graphs
                if plan.graph is None:          # invalidated by the caller (step_windows: pointers changed)
                    graphs.clear()
                if key not in graphs:
                    # warm-up run outside capture (also the first real step)
                    pack()
                    eng.run_step(plan, adam=False)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        pack()
                        eng.run_step(plan, adam=single, gscale=gmul)
                    if len(graphs) >= 4:
                        graphs.clear()
                    graphs[key] = g
                plan.graph, plan.graph_key = graphs[key], key
                plan.graph.replay()
                if not single:
                    eng.adam_step(self._allreduce_grads(eng) * gmul)
            else:
                pack()
                eng.run_step(plan, adam=False)
                eng.adam_step(self._allreduce_grads(eng) * gmul)
            if it == 0:
                if self.steps > 1:
                    plan.logits_out = plan.logits.clone()
                    plan.loss_out = plan.loss.clone()
                else:
                    plan.logits_out, plan.loss_out = plan.logits, plan.loss
        self.gpu_launches_per_step = (plan.launches_fwd + plan.launches_bwd) * self.steps
        self.last_loss = plan.loss_out

    def step(self, x: torch.Tensor, persistent_input: bool = False) -> torch.Tensor:
        """Adapt on one batch [B,C,D,H,W]; returns the pre-update logits [B,R,D,H,W] (a view of the
        engine's static output buffer -- clone it to keep it across steps).  ``persistent_input``: the
        caller keeps ``x`` alive at this address (a staging buffer it refills); the step then reads it
        in place -- its own CUDA graph per buffer -- instead of copying it into the plan's static input."""
        eng, x = self._prepare(x)
        plan = eng.get_plan(*[int(s) for s in (x.shape[0], *x.shape[2:])])
        if persistent_input:
            self._run(eng, plan, lambda: eng._pack_input(plan, x), input_key=("staging", x.data_ptr()))
            return plan.logits_out
        if plan.x_static is None or plan.x_static.shape != x.shape or plan.x_static.dtype != x.dtype:
            plan.x_static = torch.empty_like(x)
        if x.data_ptr() != plan.x_static.data_ptr():
            plan.x_static.copy_(x, non_blocking=True)
        self._run(eng, plan, lambda: eng._pack_input(plan, plan.x_static),
                  input_key=("static", plan.x_static.data_ptr()))
        return plan.logits_out

    __call__ = step

    def step_windows(self, vol: torch.Tensor, win: torch.Tensor, roi: Sequence[int],
                     sample_w: Optional[torch.Tensor] = None,
                     chan_scale: Optional[torch.Tensor] = None,
                     n_valid_global: Optional[int] = None,
                     affine: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Adapt on a batch of sliding-window patches gathered on device from ``vol``
        [n_vol,C,Ds,Hs,Ws].  ``win`` int32 [NB,4] = (volume index, d0, h0, w0) -- must live in a
        persistent buffer when CUDA graphs are on; ``sample_w`` [NB] zero-weights the padding windows of
        a tail batch; ``n_valid_global`` = real windows of this step over all ranks, so that the
        (all-reduced) gradient equals the mean over the real windows; ``affine`` [n_vol,C,4] from
        ``IntensityPolicy.stats`` normalises RAW intensities while the windows are gathered (persistent
        buffer, like ``win``)."""
        eng, vol = self._prepare(vol)
        NB = int(win.shape[0])
        plan = eng.get_plan(NB, int(roi[0]), int(roi[1]), int(roi[2]))
        key = (vol.data_ptr(), tuple(vol.shape), win.data_ptr(),
               None if chan_scale is None else chan_scale.data_ptr(),
               None if affine is None else affine.data_ptr())
        plan.win_key = key     # (part of the graph key below: a graph is replayed only for the pointers it baked in)
        if sample_w is not None:
            plan.sample_w.copy_(sample_w, non_blocking=True)
        gmul = 1.0
        if n_valid_global is not None:
            gmul = NB * self.world_size / float(n_valid_global)
        vd = tuple(int(s) for s in vol.shape[2:])
        self._run(eng, plan, lambda: eng._pack_input(plan, vol, win=win, chan_scale=chan_scale,
                                                     vol_dims=vd, n_vol=int(vol.shape[0]), affine=affine), gmul,
                  input_key=("win",) + key)
        if gmul != 1.0:            # padded tail batch: report the mean over the real windows, as the gradient is
            self.last_loss = plan.loss_out * gmul
        return plan.logits_out

    def adapt_stream(self, host_batches):
        """Adapt over an iterable of HOST batches [B,C,D,H,W] (pinned memory for true overlap).
        The H2D copy of batch i+1 runs on a side stream while batch i is being adapted, so the PCIe
        transfer (67 MB for 2x4x128^3) hides behind the ~3.5 ms step.  Yields the pre-update logits
        of each batch (static buffer: consume or clone before the next iteration)."""
        it = iter(host_batches)
        try:
            first = next(it)
        except StopIteration:
            return
        dev = self.model.engine.device or torch.device("cuda", torch.cuda.current_device())
        copy_stream = torch.cuda.Stream(device=dev)
        # fp16 host batches are staged as fp16 (half the PCIe bytes; the gather kernel reads them directly)
        sdt = torch.float16 if first.dtype == torch.float16 else torch.float32
        staging = [torch.empty(first.shape, dtype=sdt, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def enqueue(i, host):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i % 2])          # staging slot no longer read by the step
                staging[i % 2].copy_(host, non_blocking=True)
                ready[i % 2].record(copy_stream)

        cur = torch.cuda.current_stream(dev)
        for e in freed:
            e.record(cur)
        enqueue(0, first)
        i, nxt = 0, first
        while nxt is not None:
            try:
                following = next(it)
            except StopIteration:
                following = None
            if following is not None:
                if tuple(following.shape) != tuple(first.shape):
                    raise ValueError("adapt_stream: all batches must have the same shape")
                enqueue(i + 1, following)
            cur.wait_event(ready[i % 2])
            out = self.step(staging[i % 2], persistent_input=True)   # the graph of this staging buffer reads it in place
            freed[i % 2].record(cur)
            yield out
            i, nxt = i + 1, following

    def run_step(self, batch: Dict[str, torch.Tensor]) -> Dict[str, float]:
        """Reference trainer-step signature (seg_trainer.py:97): ``{"loss": float}`` (syncs)."""
        self.step(batch["image"].to(self.model.engine.device or "cuda"))
        return {"loss": float(self.last_loss.item())}
