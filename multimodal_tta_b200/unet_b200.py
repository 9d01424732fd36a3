"""``unet_b200`` -- drop-in for the reference's ``@register_model("unet")`` class
(/root/reference/src/models/unet.py:14-69) whose arithmetic runs in hand-written sm_100a CUDA.

Contract kept from the reference:
  * constructor ``cls(cfg)`` reading the same fields/defaults as src/models/unet.py:27-48;
  * ``forward(x[B,C,D,H,W] fp32 cuda) -> logits[B,R,D,H,W] fp32``;
  * ``state_dict()`` keys equal MONAI's (``model.0.conv.unit0.conv.weight`` ...), so checkpoints
    written by the reference's CheckpointHook (src/core/hooks.py:53-59) load unchanged
    (a ``module.`` DataParallel prefix is stripped).
The module tree below only HOLDS parameters under MONAI's names; no torch conv/norm op is ever
called -- ``forward`` goes through ``TTAEngine`` (engine.py) and fails loudly without CUDA.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Mapping, Optional, Sequence

import torch
import torch.nn as nn

from .config import DictConfig, create, get_config
from .registry import register_model


class ConvHolder(nn.Module):
    """Parameters of nn.Conv3d / nn.ConvTranspose3d (same shapes, same default init)."""

    def __init__(self, cin: int, cout: int, k: int, stride: int, transposed: bool):
        super().__init__()
        self.cin, self.cout, self.k, self.stride, self.transposed = cin, cout, k, stride, transposed
        shape = (cin, cout, k, k, k) if transposed else (cout, cin, k, k, k)
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.empty(cout))
        # identical call sequence to torch.nn.modules.conv._ConvNd.reset_parameters
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("ConvHolder only stores parameters; use UNetB200.forward")


class NormHolder(nn.Module):
    """Parameters/buffers of nn.InstanceNorm3d(affine=False) or nn.BatchNorm3d(affine=True)."""

    def __init__(self, channels: int, kind: str):
        super().__init__()
        self.num_features, self.kind, self.eps, self.momentum = channels, kind, 1e-5, 0.1
        if kind == "batch":
            self.weight = nn.Parameter(torch.ones(channels))
            self.bias = nn.Parameter(torch.zeros(channels))
            self.register_buffer("running_mean", torch.zeros(channels))
            self.register_buffer("running_var", torch.ones(channels))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
            self.track_running_stats = True
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
            self.track_running_stats = False

    def materialize_affine(self) -> None:
        """TENT needs gamma/beta on every norm; InstanceNorm3d(affine=False) gets gamma=1, beta=0
        (forward-identical, checkpoint-compatible)."""
        if self.weight is None:
            dev = next((b.device for b in self.buffers()), None) or torch.device("cpu")
            self.weight = nn.Parameter(torch.ones(self.num_features, device=dev))
            self.bias = nn.Parameter(torch.zeros(self.num_features, device=dev))


def _adn(channels: int, norm: str, act: str, dropout: Optional[float]) -> nn.Sequential:
    n = str(norm).upper()
    if n not in ("INSTANCE", "BATCH"):
        raise ValueError(f"unet_b200: norm {norm!r} unsupported (INSTANCE or BATCH)")
    if str(act).upper() != "RELU":
        raise ValueError(f"unet_b200: act {act!r} unsupported (RELU)")
    seq = nn.Sequential()
    seq.add_module("N", NormHolder(channels, "instance" if n == "INSTANCE" else "batch"))
    if dropout is not None:
        if float(dropout) != 0.0:
            raise ValueError("unet_b200: dropout > 0 unsupported (reference configs use 0.0)")
        seq.add_module("D", nn.Dropout(0.0))
    seq.add_module("A", nn.ReLU())
    return seq


class ConvolutionH(nn.Sequential):
    def __init__(self, cin, cout, stride, k, norm, act, dropout, conv_only=False, transposed=False):
        super().__init__()
        self.add_module("conv", ConvHolder(cin, cout, k, stride, transposed))
        self.conv_only = conv_only
        if not conv_only:
            self.add_module("adn", _adn(cout, norm, act, dropout))


class ResidualUnitH(nn.Module):
    def __init__(self, cin, cout, stride, k, subunits, norm, act, dropout, last_conv_only=False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual: nn.Module = nn.Identity()
        sch, sst = cin, stride
        subunits = max(1, subunits)
        for su in range(subunits):
            co = last_conv_only and su == subunits - 1
            self.conv.add_module(f"unit{su:d}", ConvolutionH(sch, cout, sst, k, norm, act, dropout, conv_only=co))
            sch, sst = cout, 1
        if stride != 1 or cin != cout:
            rk = k if stride != 1 else 1
            self.residual = ConvHolder(cin, cout, rk, stride, False)


class SkipConnectionH(nn.Module):
    def __init__(self, submodule: nn.Module):
        super().__init__()
        self.submodule = submodule


class _SupervisedFunction(torch.autograd.Function):
    """Autograd bridge of the supervised path: the reference's trainer computes its own loss on the logits (MONAI
    DiceCELoss) and calls ``loss.backward()``; this node turns dL/dlogits into the gradients of every parameter with
    the CUDA backward.  The input needs no gradient (it is data)."""

    @staticmethod
    def forward(ctx, x, model, *params):
        logits, plan = model.engine.forward_train(x)
        ctx.model, ctx.plan, ctx.params = model, plan, params
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        eng = ctx.model.engine
        eng.backward_from_logits_grad(ctx.plan, grad_logits)
        return (None, None, *eng.param_grads(ctx.params))


class B200Model(nn.Module):
    """Common part of the B200 model classes: backend options, holder plumbing, the lazily built TTAEngine.
    A subclass builds its holder tree, then implements ``build_graph(G)`` (walk the tree with the engine's
    emitters, return the final conv result) and optionally ``configure_layers(engine)``."""

    in_channels: int
    out_channels: int

    def _init_backend(self, cfg) -> None:
        get = lambda k, d: get_config(cfg, k, d)
        self._engine = None
        self._params_dirty = True
        self._backend_options(get)

    def build_graph(self, G):  # pragma: no cover - abstract
        raise NotImplementedError

    def configure_layers(self, engine) -> None:
        """Hook for per-layer launch options (effective-weight maps)."""

    # ------------------------------------------------------------------ plumbing
    def norm_holders(self) -> List[NormHolder]:
        return [m for m in self.modules() if isinstance(m, NormHolder)]

    def conv_holders(self) -> List[ConvHolder]:
        return [m for m in self.modules() if isinstance(m, ConvHolder)]

    def _apply(self, fn, *a, **k):
        self._params_dirty = True
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True, assign: bool = False):
        """Accepts reference checkpoints: plain MONAI keys or ``module.``-prefixed DataParallel
        keys (src/core/experiment_manager.py:95-96); norm affine keys materialise gamma/beta."""
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        for name, m in self.named_modules():
            if isinstance(m, NormHolder) and f"{name}.weight" in sd:
                m.materialize_affine()
        self._params_dirty = True
        out = super().load_state_dict(sd, strict=strict, assign=assign)
        if self._engine is not None:
            self._engine.invalidate()
        return out

    @property
    def engine(self):
        from .engine import TTAEngine

        if self._engine is None:
            self._engine = TTAEngine(self)
        return self._engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Raw logits [B,R,D,H,W].  train(): test-batch norm statistics; eval(): BatchNorm uses its
        running statistics (InstanceNorm always uses instance statistics, as in the reference).
        With ``trainable: true`` and gradients enabled the logits carry a grad_fn: ``loss.backward()`` of the
        reference's supervised step (src/core/trainers/seg_trainer.py:141-143) runs the CUDA backward (input
        gradients, norm backward, weight / bias gradients) and fills every parameter's ``.grad``."""
        if self.trainable and torch.is_grad_enabled() and self.training:
            params = [p for p in self.parameters() if p.requires_grad]
            if params:
                return _SupervisedFunction.apply(x, self, *params)
        return self.engine.forward(x)

    def _backend_options(self, get) -> None:
        """Backend options (not in the reference): conv kernel family, whole-step CUDA graph, A/B switches."""
        # supervised step: forward keeps an autograd edge, backward computes ALL parameter gradients
        self.trainable = bool(get("trainable", False))
        self.conv_backend = str(get("conv_backend", "auto"))
        # the launch list of a shape is recorded into a C object (tta_plan) and a step is one C call (tta_step);
        # false: the Python closures are called one by one (debugging)
        self.c_plan = bool(get("c_plan", True))
        self.use_cuda_graph = bool(get("cuda_graph", True))
        # deterministic=True disables split-K (float atomics) in the deep, SM-starved conv layers
        self.deterministic = bool(get("deterministic", False))
        # extra tta_conv_tc flag bits for A/B experiments (include/tta_b200.h), e.g. 32 = one-plane tiles
        self.tc_flags = int(get("tc_flags", 0))
        # head convT (Cout <= 4) as dense GEMM + shared-memory col2im instead of 28 small MMAs per block
        self.t2s_head = bool(get("t2s_head", True))
        # run a ResidualUnit's unit0 conv and its strided 3x3x3 shortcut conv as ONE launch
        self.fuse_shortcut = bool(get("fuse_shortcut", True))
        # run the full-resolution tail (norm apply + 3x3x3 conv + entropy, and its backward) as two
        # fused CUDA-core kernels instead of five streaming passes (csrc/tta_head.cu)
        self.fuse_head = bool(get("fuse_head", True))
        # the <= 4-channel tensors of the fused head (convT result, masked gradient) in a compact 4-channel layout
        # instead of 8-channel chunks: -0.4 GB of DRAM traffic per 2x4x128^3 step
        self.head_compact = bool(get("head_compact", True))
        # <= 4-channel operands of stride-2 convs (the packed network input, the head norm's gradient) in the compact
        # 4-channel layout: half the bytes of the 8-channel chunk planes and 9 instead of 15 MMAs per stem tile
        self.input_compact = bool(get("input_compact", True))
        # norm statistics (sum y, sum y^2) come out of the producing tcgen05 conv's epilogue instead of
        # a separate pass over y (only where the conv runs without split-K)
        self.fuse_stats = bool(get("fuse_stats", True))
        # OPT-IN: statistics from a conv epilogue finalized inside the apply pass's prologue instead of a separate
        # launch.  Measured slower (2.337 vs 2.326 ms per step: every apply block re-reduces 148 slots)
        self.stats_finalize_in_apply = bool(get("stats_finalize_in_apply", False))
        # ... also when the conv has a single TMEM accumulator buffer (the reduction is then not hidden behind the next
        # item's MMAs, but a statistics pass over the conv result disappears).  Same-box A/B: 2.2936 vs 2.2968 ms per
        # step in round 1 (within noise, default off); on the final round-2 build 2.1190 vs 2.1335 (two rounds each,
        # spread 0.002): default on
        self.fuse_stats_single_buffer = bool(get("fuse_stats_single_buffer", True))
        # ... and, OPT-IN, the norm-BACKWARD reductions (sum dz, sum dz*xhat) out of the epilogue of
        # the dgrad conv that completes the layer's incoming gradient.  Measured on B200 (2x4x128^3):
        # the four 64^3 / 32^3 layers it applies to lose 162 us in their dgrads (the epilogue has
        # 8 warps per SM for ~100 instructions per voxel-chunk) and save 144 us of streaming passes
        # (2048 threads per SM) -> default off; the streaming tta_norm_bwd_reduce stays the product path
        self.fuse_bwd_stats = bool(get("fuse_bwd_stats", False))
        # layers with <= 4096 voxels per instance (8^3, 16^3): statistics + apply, and backward
        # reduction + apply, as one launch each
        self.fuse_small_norm = bool(get("fuse_small_norm", True))
        # ... up to this many voxels per instance.  Slabs above 512 voxels are owned by a thread-block
        # cluster of up to 8 CTAs (totals through distributed shared memory).  Same-box A/B per step:
        # 512 -> 2.645 ms, 4096 (16^3 level too) -> 2.628 ms, 65536 (32^3 too) -> 2.641 ms
        self.small_norm_max_voxels = int(get("small_norm_max_voxels", 4096))
        # OPT-IN: InstanceNorm backward sample by sample when one sample's gradient + conv result fit
        # the 126 MB L2 but the batch does not, hoping the apply pass re-reads them from L2.  Measured
        # on the four 64^3 layers (67 MB per sample): 2.578 ms vs 2.556 ms per step -- the second pass
        # does not hit, the two extra launches per layer cost more; default off
        self.per_sample_norm_bwd = bool(get("per_sample_norm_bwd", False))
        # OPT-IN generalisation: (sample, channel-chunk range) groups whose reduction pass reads at most this many MB
        # (0 = off), small enough that the apply pass should find g and y in L2.  Measured (round 2, same box, ms per
        # step): off 2.233, 70 MB 2.302, 40 MB 2.371, 20 MB 2.422 -- more, smaller launches cost more than the
        # second pass gains; the re-read is not where the time goes
        self.norm_bwd_l2_mb = float(get("norm_bwd_l2_mb", 0))
        self.norm_bwd_l2_min_mb = float(get("norm_bwd_l2_min_mb", 100))   # only layers whose batch exceeds this
        # gradient operand format of the dgrad convs: "fp16" = one loss-scaled fp16 plane (1 MMA per
        # k-step), "bf16x2" = split bf16 planes (2 MMAs); DESIGN.md section 6 has the error budget
        self.bwd_precision = str(get("bwd_precision", "fp16"))
        # the fp16 gradient planes carry the loss times 2^ceil(log2(4 * voxels * loss_scale_mult)): large enough that
        # the deep layers' gradients stay out of the fp16 subnormals, small enough that the head's do not saturate
        self.loss_scale_mult = float(get("loss_scale_mult", 1.0))
        # activation gradients between a dgrad conv and the norm backward that consumes them as ONE loss-scaled fp16
        # plane instead of fp32 (tta_conv_tc flags bit 16 / the fp16-source bits of tta_norm_bwd_*), at the levels with
        # at least grad_f16_min_voxels voxels per instance (their dgrads run without split-K anyway).  Same-box A/B at
        # 2x4x128^3, ms per step: fp32 2.189, fp16 at the 64^3 and 32^3 levels 2.177, at the 64^3 level only 2.165
        self.grad_f16 = bool(get("grad_f16", True))
        self.grad_f16_min_voxels = int(get("grad_f16_min_voxels", 100000))
        # supervised step: weight gradients of the 3x3x3 layers on the tensor cores (tta_conv_wgrad_tc) where the
        # operand layouts allow it ("auto"), or always on the exact fp32 CUDA-core kernel ("simt")
        self.wgrad_backend = str(get("wgrad_backend", "auto"))
        # ... with x = fp16 hi + lo planes (two products, default) or the hi plane only (one product)
        self.wgrad_x_lo = bool(get("wgrad_x_lo", True))
        if self.wgrad_backend not in ("auto", "simt"):
            raise ValueError("unet_b200: wgrad_backend must be 'auto' or 'simt'")
        if self.bwd_precision not in ("fp16", "bf16x2"):
            raise ValueError("unet_b200: bwd_precision must be 'fp16' or 'bf16x2'")


@register_model("unet_b200")
class UNetB200(B200Model):
    def __init__(self, cfg: DictConfig | Dict[str, Any], in_channels: Optional[int] = None,
                 eps: Optional[float] = None):
        super().__init__()
        if not isinstance(cfg, DictConfig):
            cfg = create(dict(cfg))
        c_in_cfg = get_config(cfg, "in_channels", 3)
        c_in = in_channels if in_channels is not None else (None if c_in_cfg == "auto" else int(c_in_cfg))
        if c_in is None:
            raise ValueError("[UNet] in_channels is 'auto'; please pass in_channels at construction time.")
        self.in_channels = c_in
        self.out_channels = int(get_config(cfg, "num_classes", 1))
        self.channels = [int(c) for c in get_config(cfg, "channels", [32, 64, 128, 256, 512])]
        self.strides = [int(s) for s in get_config(cfg, "strides", [2, 2, 2, 2])]
        self.num_res_units = int(get_config(cfg, "num_res_units", 0))
        self.act = get_config(cfg, "act", "relu")
        self.norm = get_config(cfg, "norm", "BATCH")
        self.dropout = float(get_config(cfg, "dropout", 0.0))
        if int(get_config(cfg, "spatial_dims", 3)) != 3:
            raise ValueError("unet_b200 implements the 3-D path only")
        if len(self.channels) < 2:
            raise ValueError("the length of `channels` should be no less than 2.")
        if len(self.strides) < len(self.channels) - 1:
            raise ValueError("the length of `strides` should equal to `len(channels) - 1`.")
        for s in self.strides:
            if s not in (1, 2):
                raise ValueError(f"unet_b200: stride {s} unsupported (1 or 2)")
        self._init_backend(cfg)
        k, nru, norm, act, dr = 3, self.num_res_units, self.norm, self.act, self.dropout

        def down(cin, cout, stride):
            if nru > 0:
                return ResidualUnitH(cin, cout, stride, k, nru, norm, act, dr)
            return ConvolutionH(cin, cout, stride, k, norm, act, dr)

        def up(cin, cout, stride, is_top):
            conv: nn.Module = ConvolutionH(cin, cout, stride, k, norm, act, dr,
                                           conv_only=is_top and nru == 0, transposed=True)
            if nru > 0:
                conv = nn.Sequential(conv, ResidualUnitH(cout, cout, 1, k, 1, norm, act, dr, last_conv_only=is_top))
            return conv

        def block(inc, outc, chs, sts, is_top):
            c, s = chs[0], sts[0]
            if len(chs) > 2:
                sub, upc = block(c, c, chs[1:], sts[1:], False), c * 2
            else:
                sub, upc = down(c, chs[1], 1), c + chs[1]
            return nn.Sequential(down(inc, c, s), SkipConnectionH(sub), up(upc, outc, s, is_top))

        self.model = block(self.in_channels, self.out_channels, self.channels, self.strides, True)

    # ------------------------------------------------------------------ op graph
    def build_graph(self, G):
        """Walk MONAI's recursive ``_create_block`` structure (down, SkipConnection(sub), up) with the engine's
        emitters; ``torch.cat`` of the skip connection = channel slices of one buffer.  Returns the final conv
        result (src/models/unet.py:68-69 -> monai UNet.forward)."""
        def out_channels(mod) -> int:
            if isinstance(mod, ResidualUnitH):
                return list(mod.conv.children())[-1].conv.cout
            if isinstance(mod, ConvolutionH):
                return mod.conv.cout
            return out_channels(list(mod.children())[-1])

        def block(seq: nn.Sequential, inp, out):
            down, skip, up = seq[0], seq[1], seq[2]
            sub = skip.submodule
            c, cs = out_channels(down), out_channels(sub)
            if c % 8 or cs % 8:
                raise ValueError("unet_b200: skip-connection channel counts must be multiples of 8")
            d, h, w = inp.dims
            if isinstance(down, ResidualUnitH):
                s = list(down.conv.children())[0].conv.stride
            else:
                s = down.conv.stride
            if d % s or h % s or w % s:
                raise ValueError(f"unet_b200: spatial size {(d, h, w)} not divisible by stride {s} "
                                 "(the skip concat would mismatch, as in the reference)")
            cat = G.new_act(c + cs, (d // s, h // s, w // s), name="cat")
            dv = cat.view(0, c // 8)
            G.layer(down, inp, dv)
            sv = cat.view(c // 8, cs // 8)
            if isinstance(sub, nn.Sequential) and len(sub) == 3 and isinstance(sub[1], SkipConnectionH):
                block(sub, dv, sv)
            else:
                G.layer(sub, dv, sv)
            return G.layer(up, cat.view(), out)

        return block(self.model, G.x.view(), None)
