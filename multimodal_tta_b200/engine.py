"""Execution engine: turns a ``B200Model`` module tree (``UNetB200``, ``MultimodalUNetB200``) into a static list of C-ABI kernel
launches (forward, fused head, backward, Adam) over pre-allocated HBM buffers, optionally
replayed as one CUDA graph.

Path covered: MONAI-UNet forward (a2-a6), entropy loss + dlogits, backward
(dgrad only -- TENT freezes conv weights, so there is no wgrad), norm-affine gradients, Adam
(a10) on the flat [gamma || beta] buffer.

Layouts: activations that feed a conv are split 16-bit planes
(fp16 forward / bf16 backward, x ~= hi + lo) in the channel-blocked layout [N][C8][D][H][W][8];
conv results are fp32 in the same blocked layout.  ``torch.cat`` of the skip connection is free:
producers write into channel slices of one buffer.
"""
from __future__ import annotations

import contextlib
import ctypes
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import TTA_BF16, TTA_F16, TTA_F16_HI, check
from .layout import pack_bias, pack_weights_simt, pack_weights_small, pack_weights_tc, wg_dgrad, wg_forward
from .unet_b200 import (B200Model, ConvHolder, ConvolutionH, NormHolder, ResidualUnitH, SkipConnectionH)


_DEVICE: Optional[torch.device] = None   # device of the engine that is launching (set by TTAEngine._on_device)


def _stream() -> int:
    """Raw handle of torch's current stream ON THE ENGINE'S DEVICE (not the process-wide current device)."""
    return torch.cuda.current_stream(_DEVICE).cuda_stream


class NormBwdSeg(ctypes.Structure):
    """include/tta_b200.h: tta_norm_bwd_seg (HOST struct handed to tta_conv_tc_bwd_norm)."""
    _fields_ = [("c8_begin", ctypes.c_int), ("c8_count", ctypes.c_int), ("relu", ctypes.c_int), ("pad", ctypes.c_int),
                ("y", ctypes.c_void_p), ("y_ns", ctypes.c_longlong), ("mean", ctypes.c_void_p),
                ("rstd", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                ("partial", ctypes.c_void_p)]


class _NoGrad16(Exception):
    """A gradient tensor planned as one fp16 plane met a writer / reader that needs fp32: the plan is rebuilt
    with fp32 gradients."""


# ------------------------------------------------------------------------------ tensors
class Act:
    """Conv-operand activation: split fp16 planes [2][N][C8][D][H][W][8] (+ fp32 grad)."""

    def __init__(self, N, C, D, H, W, device, needs_grad: bool, name: str = "", g16: bool = False):
        self.N, self.C, self.C8, self.D, self.H, self.W = N, C, (C + 7) // 8, D, H, W
        self.V = D * H * W
        self.name, self.needs_grad = name, needs_grad
        self.planes = torch.zeros((2, N, self.C8, D, H, W, 8), dtype=torch.int16, device=device)
        # g16: the gradient is ONE loss-scaled fp16 plane (written by tcgen05 dgrads with flags bit 16, read by the
        # norm backward with its fp16-source bits): 2 instead of 4 bytes per element through three passes
        self.g16 = bool(g16 and needs_grad)
        self.grad = (torch.zeros((N, self.C8, D, H, W, 8), dtype=torch.int16 if self.g16 else torch.float32,
                                 device=device) if needs_grad else None)
        self.ns = self.C8 * self.V * 8  # elements between samples
        self.written: set = set()       # chunks of .grad written so far in this backward
        self.extra: List["ActView"] = []  # identity-residual gradient contributions
        # operands of a stride-2 tcgen05 conv are stored w-parity-split: either the
        # planes themselves (`wsplit`, single consumer: the network input) or a second copy written by
        # the producing norm (`ws_planes`, skip tensors that also feed the concat)
        self.writers: list = []         # dgrad launches that wrote .grad: (c8_begin, c8_end, record, accumulate)
        self.wsplit = False
        self.compact = False            # <= 4 channels stored [N][D][H][W][4] (8 B per voxel and plane): network input
        self.ws_planes: Optional[torch.Tensor] = None
        self.device = device
        self.peers: List["Act"] = []    # other shapes of the SAME storage (batched aliases): written-state is shared

    def batched_alias(self, factor: int, name: str = "") -> "Act":
        """The same storage seen as [N * factor][C / factor] instead of [N][C]: index ((n*f + m)*C8' + c)*V*8 equals
        (n*C8 + m*C8' + c)*V*8, so modality m of sample n is instance n*f + m.  Used where ONE layer (shared weights
        and norm parameters) is applied to every modality: the modalities become batch entries of a single launch."""
        if self.C8 % factor:
            raise ValueError("batched_alias: channel chunks not divisible")
        a = Act.__new__(Act)
        a.N, a.C8, a.D, a.H, a.W, a.V = self.N * factor, self.C8 // factor, self.D, self.H, self.W, self.V
        a.C = a.C8 * 8
        a.name, a.needs_grad = name or self.name + "/batched", self.needs_grad
        a.g16 = self.g16
        a.planes = self.planes.view(2, a.N, a.C8, self.D, self.H, self.W, 8)
        a.grad = self.grad.view(a.N, a.C8, self.D, self.H, self.W, 8) if self.grad is not None else None
        a.ns = a.C8 * a.V * 8
        a.written, a.extra, a.writers = set(), [], []
        a.wsplit, a.ws_planes, a.device = False, None, self.device
        a.peers = [self]
        self.peers.append(a)
        return a

    def mark_written(self, chunks: set):
        self.written |= chunks
        if len(self.written) == self.C8:          # fully written: so is every other shape of the storage
            for p in self.peers:
                p.written = set(range(p.C8))

    def need_ws_copy(self):
        if self.ws_planes is None:
            self.ws_planes = torch.zeros_like(self.planes)

    def view(self, c8_off: int = 0, c8_len: Optional[int] = None) -> "ActView":
        return ActView(self, c8_off, self.C8 - c8_off if c8_len is None else c8_len)


@dataclass
class ActView:
    parent: Act
    c8_off: int
    C8: int

    @property
    def dims(self):
        return self.parent.D, self.parent.H, self.parent.W

    @property
    def hi(self) -> int:
        return self.parent.planes[0].data_ptr() + self.c8_off * self.parent.V * 8 * 2

    @property
    def lo(self) -> int:
        return self.parent.planes[1].data_ptr() + self.c8_off * self.parent.V * 8 * 2

    @property
    def ws_hi(self) -> int:
        return self.parent.ws_planes[0].data_ptr() + self.c8_off * self.parent.V * 8 * 2

    @property
    def ws_lo(self) -> int:
        return self.parent.ws_planes[1].data_ptr() + self.c8_off * self.parent.V * 8 * 2

    @property
    def g(self) -> int:
        return self.parent.grad.data_ptr() + self.c8_off * self.parent.V * 8 * (2 if self.parent.g16 else 4)

    @property
    def ns(self) -> int:
        return self.parent.ns


class Res:
    """fp32 conv result [N][C8][D][H][W][8] (+ the 16-bit gradient planes dY that feed its dgrad)."""

    def __init__(self, N, C, D, H, W, device, name: str = ""):
        self.N, self.C, self.C8, self.D, self.H, self.W = N, C, (C + 7) // 8, D, H, W
        self.V = D * H * W
        self.name = name
        self.data = torch.zeros((N, self.C8, D, H, W, 8), dtype=torch.float32, device=device)
        self.ns = self.C8 * self.V * 8
        self.dy: Optional[torch.Tensor] = None
        self.dy_wsplit = False  # dY feeds a stride-2 tcgen05 dgrad: stored w-parity-split
        self.dy_c4 = False      # ... or compact ([N][D][H][W][4], <= 4 channels: the head norm's gradient)
        self.stats_c8 = 0       # > 0: the producing tcgen05 conv also emits norm statistics partials
        self.stats_grid = 0     #      ... with this many slots (= its CTA count) per (n, chunk)
        self.tc_query = None    # (ksplit, grid) of the producing tcgen05 launch, None for other backends
        self.c4 = False         # compact [N][D][H][W][4] fp32 layout (<= 4 channels, fused head only)
        self.t2s = False        # produced by the small-Cout transposed kernel (the only writer of the compact layout)
        self.device = device
        self.root, self.c8_off = self, 0

    def alloc_dy(self, planes: int = 2):
        if self.dy is None:
            shape = (self.N, self.C8, self.D, self.H, self.W, 8)
            self.dy = torch.zeros((2, *shape), dtype=torch.int16, device=self.device) if planes == 2 else \
                torch.zeros((1, *shape), dtype=torch.int16, device=self.device).expand(2, -1, -1, -1, -1, -1, -1)

    @property
    def ptr(self) -> int:
        return self.data.data_ptr()

    def dy_ptr(self, plane: int) -> int:
        return self.dy[plane].data_ptr()

    def view(self, c8_off: int, c8_len: int, C: int) -> "ResView":
        return ResView(self, c8_off, c8_len, C)


class ResView:
    """Channel slice of a Res (the two halves of a fused unit0 || shortcut convolution)."""

    def __init__(self, parent: Res, c8_off: int, C8: int, C: int):
        self.root, self.c8_off, self.C8, self.C = parent, c8_off, C8, C
        self.N, self.D, self.H, self.W, self.V, self.ns, self.name = (parent.N, parent.D, parent.H, parent.W,
                                                                      parent.V, parent.ns, parent.name)

    def alloc_dy(self, planes: int = 2):
        self.root.alloc_dy(planes)

    @property
    def ptr(self) -> int:
        return self.root.data.data_ptr() + self.c8_off * self.V * 8 * 4

    def dy_ptr(self, plane: int) -> int:
        return self.root.dy[plane].data_ptr() + self.c8_off * self.V * 8 * 2


# ------------------------------------------------------------------------------ layers
class ConvLayer:
    """One convolution launch.  ``extra`` fuses a second conv with the same input and geometry
    (ResidualUnit unit0 || strided shortcut) along the output channels: the forward reads the input
    once with N doubled, the dgrad is one conv over the concatenated gradients (K doubled)."""

    def __init__(self, holder: ConvHolder, name: str, fold_identity: bool = False, extra: Optional[ConvHolder] = None):
        self.h, self.name, self.fold_identity, self.extra = holder, name, fold_identity, extra
        self.K, self.stride = holder.k, holder.stride
        self.mode = 1 if holder.transposed else 0
        self.cin, self.cout = holder.cin, holder.cout + (extra.cout if extra is not None else 0)
        self.packed = {}
        # effective-weight hook (multimodal model): maps the holder's weight [co][ci][k,k,k] to the weight the
        # launch uses, e.g. a 1-channel encoder stem conv -> 4 input channels with one non-zero; sets self.cin
        self.w_map: Optional[Callable[[torch.Tensor], torch.Tensor]] = None
        self.dgrad_cin: Optional[int] = None    # input gradient only for the leading channels (rest needs none)

    def _canonical(self, w, b, w2=None, b2=None, add_identity=None):
        """Parameter tensors (or tensors of the same shapes holding source indices / additive constants) -> the
        canonical forward / dgrad weights Wg[T][ci][co] and the bias vector this launch uses."""
        if self.w_map is not None:
            w = self.w_map(w)
            if w2 is not None:
                w2 = self.w_map(w2)
        wf = wg_forward(w, self.h.transposed)
        wd = wg_dgrad(w, self.h.transposed)
        if w2 is not None:
            wf = torch.cat([wf, wg_forward(w2, False)], dim=2)      # [T][ci][co0 + co1]
            wd = torch.cat([wd, wg_dgrad(w2, False)], dim=1)        # [T][ci = dy0 || dy1][co = cin]
            b = torch.cat([b, b2])
        if add_identity is not None:
            c = self.K ** 3 // 2
            eye = torch.eye(self.cin, dtype=wf.dtype) * add_identity
            wf = wf.clone(); wd = wd.clone()
            wf[c] += eye
            wd[c] += eye
        if self.dgrad_cin is not None:
            wd = wd[:, :, : self.dgrad_cin].contiguous()
        return wf, wd, b

    def _pack_all(self, wf, wd, b, want_tc: bool, bwd_dtype: int, t2s: bool, small: bool = True) -> dict:
        """Every device format of the launch from the canonical tensors (values, or -- inside layout.index_mode() --
        source indices: the packers are pure rearrangements)."""
        out = {"simt_fwd": pack_weights_simt(wf), "simt_bwd": pack_weights_simt(wd), "bias": pack_bias(b)}
        lib = _lib.lib()
        if small and lib.tta_conv_small_supported(self.K, self.stride, self.cin, self.cout):
            # tiny-channel stride-1 convs (UNet head): direct CUDA-core kernel, HBM-bound (HOST weights = kernel params)
            out["small_fwd"] = pack_weights_small(wf, self.mode)
            out["small_bwd"] = pack_weights_small(wd, 1 - self.mode)
        if want_tc:
            self.tc_fwd_flags = 0
            if lib.tta_conv_tc_supported(self.mode, self.K, self.stride, self.cin, self.cout):
                # small-Cout transposed conv (the head convT): dense GEMM + col2im kernel, own weight
                # layout, real Cout in flags bits 8..10 (include/tta_b200.h)
                use_t2s = t2s and self.extra is None and bool(
                    lib.tta_conv_tc_t2s(self.mode, self.K, self.stride, self.cin, self.cout, 1))
                out["tc_fwd"] = pack_weights_tc(wf, self.mode, self.K, self.stride, TTA_F16, t2s=use_t2s)
                self.tc_fwd_flags = (self.cout << 8) if use_t2s else 0
                # stride-2 conv over <= 4 input channels: also the variant for a COMPACT input (network input)
                if lib.tta_conv_tc_s2c4(self.mode, self.K, self.stride, self.cin):
                    out["tc_fwd_c4"] = pack_weights_tc(wf, self.mode, self.K, self.stride, TTA_F16, s2c4=True)
            bmode = 1 - self.mode
            if lib.tta_conv_tc_supported(bmode, self.K, self.stride, self.cout, self.dgrad_cin or self.cin):
                out["tc_bwd"] = pack_weights_tc(wd, bmode, self.K, self.stride, bwd_dtype)
                if lib.tta_conv_tc_s2c4(bmode, self.K, self.stride, self.cout):   # dgrad of a <= 4-channel convT
                    out["tc_bwd_c4"] = pack_weights_tc(wd, bmode, self.K, self.stride, bwd_dtype, s2c4=True)
        return out

    def pack(self, device, want_tc: bool, bwd_dtype: int = TTA_BF16, t2s: bool = True, trainable: bool = False):
        """(Re)pack weights: fp32 for the CUDA-core kernels, split fp16 / single fp16 for tcgen05.
        Packing runs on the HOST (a few hundred small index ops per layer, once per weight load) and every
        blob reaches the device as one H2D copy -- the device only ever executes this library's kernels.
        ``trainable`` (supervised step): no host-parameter kernels (their weights could not follow the optimizer
        without a device -> host copy per step)."""
        host = torch.device("cpu")
        f32 = lambda t: t.detach().to(device=host, dtype=torch.float32)
        w = f32(self.h.weight)
        b = f32(self.h.bias) if self.h.bias is not None else torch.zeros(self.h.cout, dtype=torch.float32)  # bias=False
        w2 = f32(self.extra.weight) if self.extra is not None else None
        b2 = f32(self.extra.bias) if self.extra is not None else None
        wf, wd, bb = self._canonical(w, b, w2, b2, 1.0 if self.fold_identity else None)
        self.wg_fwd_host = wf          # canonical Wg[T][ci][co] (fp32, host): weight-gradient tests, diagnostics
        self.packed = {k: (v if k.startswith("small_") else v.to(device))
                       for k, v in self._pack_all(wf, wd, bb, want_tc, bwd_dtype, t2s, small=not trainable).items()}

    def index_maps(self, offsets: dict, device, want_tc: bool, bwd_dtype: int, t2s: bool) -> dict:
        """For every packed blob: (int32 map, optional int8 additive map, kind) with map[e] = +-(1 + flat index of the
        source parameter element) (sign = hi / lo plane, 0 = padding) -- what tta_repack_weights turns back into the
        blob from the flat parameter buffer after an optimizer step.  ``offsets[id(holder)] = (w_off, b_off)``."""
        from .layout import index_mode
        if self.w_map is not None:
            raise NotImplementedError("supervised step: layers with effective-weight maps (multimodal stems) are not supported")

        def idx(t, off):
            return torch.arange(t.numel(), dtype=torch.float64).view(t.shape) + (off + 1)
        wo, bo = offsets[id(self.h)]
        w = idx(self.h.weight, wo)
        b = idx(self.h.bias, bo) if self.h.bias is not None else torch.zeros(self.h.cout, dtype=torch.float64)
        w2 = b2 = None
        if self.extra is not None:
            wo2, bo2 = offsets[id(self.extra)]
            w2, b2 = idx(self.extra.weight, wo2), idx(self.extra.bias, bo2)
        maps = {}
        with index_mode():
            wf, wd, bb = self._canonical(w, b, w2, b2, None)
            pm = self._pack_all(wf, wd, bb, want_tc, bwd_dtype, t2s, small=False)
            pa = None
            if self.fold_identity:   # the additive identity of the folded shortcut, through the same rearrangement
                z = lambda t: torch.zeros(t.shape, dtype=torch.float64)
                af, ad, ab = self._canonical(z(self.h.weight), z(b), None, None, 1.0)
                pa = self._pack_all(af, ad, ab, want_tc, bwd_dtype, t2s, small=False)
        for k, m in pm.items():
            kind = 0 if k in ("simt_fwd", "simt_bwd", "bias") else (2 if (k.startswith("tc_bwd") and bwd_dtype == TTA_BF16) else 1)
            add = None
            if pa is not None and k != "bias" and bool((pa[k] != 0).any()):
                add = (pa[k] != 0).to(torch.int8).contiguous().to(device)
            maps[k] = (m.to(torch.int32).contiguous().to(device), add, kind)
        return maps


class NormLayer:
    def __init__(self, holder: NormHolder, name: str, off: int):
        self.h, self.name, self.off = holder, name, off
        self.C, self.C8 = holder.num_features, (holder.num_features + 7) // 8
        self.batch = 1 if holder.kind == "batch" else 0


# ------------------------------------------------------------------------------ plan
@dataclass
class Plan:
    N: int
    dims: Tuple[int, int, int]
    x: Act
    logits: torch.Tensor
    loss: torch.Tensor
    fwd: List[Callable[[], None]] = field(default_factory=list)
    head_infer: Optional[Callable[[], None]] = None
    head_train: Optional[Callable[[], None]] = None
    bwd: List[Callable[[], None]] = field(default_factory=list)
    keep: list = field(default_factory=list)
    launches_fwd: int = 0
    launches_bwd: int = 0
    conv_backends: dict = field(default_factory=dict)
    graph: Optional[torch.cuda.CUDAGraph] = None      # the graph replayed last (None: invalidated)
    graph_key: Optional[tuple] = None
    graphs: dict = field(default_factory=dict)        # key (mode, Adam, input pointer) -> captured step
    sample_w: Optional[torch.Tensor] = None
    win: Optional[torch.Tensor] = None
    chan_scale: Optional[torch.Tensor] = None
    x_static: Optional[torch.Tensor] = None
    stats_ops: list = field(default_factory=list)
    c_plans: dict = field(default_factory=dict)       # model.training -> CPlan (the launch list recorded in C)


class CPlan:
    """Handle of a ``tta_plan`` (include/tta_b200.h): the launch list of one shape, recorded once by running the
    Python closures while the C library is in recording mode; a step is then ONE C call (``tta_step``)."""
    SEC_FWD, SEC_HEAD_TRAIN, SEC_HEAD_INFER, SEC_BWD = 0, 1, 2, 3

    def __init__(self, lib, plan: "Plan"):
        self.lib = lib
        self.h = ctypes.c_void_p()
        check(lib.tta_plan_create(ctypes.byref(self.h)), "plan_create")
        for sec, ops in ((self.SEC_FWD, plan.fwd), (self.SEC_HEAD_TRAIN, [plan.head_train]),
                         (self.SEC_HEAD_INFER, [plan.head_infer]), (self.SEC_BWD, plan.bwd)):
            check(lib.tta_plan_begin(self.h, sec), "plan_begin")
            try:
                for op in ops:
                    op()
            finally:
                check(lib.tta_plan_end(), "plan_end")
        self.launches = [lib.tta_plan_num_launches(self.h, s_) for s_ in range(4)]

    def __del__(self):
        try:
            if self.h:
                self.lib.tta_plan_destroy(self.h)
                self.h = None
        except Exception:
            pass


class TTAEngine:
    """Owns packed weights, the flat norm-affine parameter/optimizer state and per-shape plans."""

    def __init__(self, model: B200Model):
        self.model = model
        self.lib = _lib.lib()  # raises if the CUDA library is missing
        self.plans = {}
        self.device: Optional[torch.device] = None
        self.conv_layers = {}
        self.norm_layers: List[NormLayer] = []
        self.P = 0
        self.gb = self.dgb = self.m = self.v = self.step_dev = None
        self.entropy_mode = 1
        self.adam = dict(lr=1e-3, b1=0.9, b2=0.999, eps=1e-8)
        self.bwd_dtype = TTA_F16_HI if model.bwd_precision == "fp16" else TTA_BF16
        self.last_inv_scale = 1.0
        # supervised step (model.trainable): every conv weight / bias lives in ONE flat fp32 buffer (the holders'
        # Parameters are views of it, their gradients views of gflat), so that the weight-gradient kernels write
        # into .grad directly and the packed operand blobs are rebuilt on the device after an optimizer step
        self.wflat = self.gflat = None
        self.w_off: dict = {}
        self._collect()

    # ---------------------------------------------------------------- structure
    def _collect(self):
        fold = set()
        for m in self.model.modules():
            if isinstance(m, ResidualUnitH):
                last = list(m.conv.children())[-1]
                if last.conv_only:
                    if not isinstance(m.residual, nn.Identity):
                        raise ValueError("unet_b200: conv-only residual unit with a conv shortcut is unsupported")
                    fold.add(id(last.conv))
        for name, m in self.model.named_modules():
            if isinstance(m, ConvHolder):
                self.conv_layers[id(m)] = ConvLayer(m, name, fold_identity=id(m) in fold)
        # unit0 || shortcut fusion candidates: 3x3x3 strided shortcut with unit0's geometry, and at
        # least two sub-units (the shortcut is added after the LAST unit's norm)
        self.fused_layers = {}
        if self.model.fuse_shortcut:
            for name, m in self.model.named_modules():
                if isinstance(m, ResidualUnitH) and isinstance(m.residual, ConvHolder):
                    units = list(m.conv.children())
                    u0, r = units[0].conv, m.residual
                    if (len(units) >= 2 and not units[0].conv_only and r.k == u0.k == 3 and r.stride == u0.stride
                            and r.cin == u0.cin and r.cout == u0.cout and u0.cout % 8 == 0 and not u0.transposed):
                        self.fused_layers[id(u0)] = ConvLayer(u0, name + ".conv.unit0.conv||residual", extra=r)
        off = 0
        for name, m in self.model.named_modules():
            if isinstance(m, NormHolder):
                nl = NormLayer(m, name, off)
                self.norm_layers.append(nl)
                off += nl.C8 * 8
        self.P = off
        self.model.configure_layers(self)

    def invalidate(self):
        self.model._params_dirty = True

    def active_layers(self) -> List["ConvLayer"]:
        fused_members = set()
        for fl in self.fused_layers.values():
            fused_members |= {id(fl.h), id(fl.extra)}
        return [cl for k, cl in self.conv_layers.items() if k not in fused_members] + list(self.fused_layers.values())

    def _bind_conv_params(self):
        """model.trainable: move every conv weight / bias into the flat buffer and rebind the Parameters as views
        (same Parameter objects: an optimizer built earlier keeps working)."""
        holders = self.model.conv_holders()
        n = sum(h.weight.numel() + (h.bias.numel() if h.bias is not None else 0) for h in holders)
        if self.wflat is not None and self.wflat.device == self.device and self.wflat.numel() == n:
            return
        wflat = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.gflat = torch.zeros(n, dtype=torch.float32, device=self.device)
        off = 0
        self.w_off = {}
        for h in holders:
            nw = h.weight.numel()
            wflat[off: off + nw].copy_(h.weight.detach().reshape(-1).to(self.device))
            h.weight.data = wflat[off: off + nw].view(h.weight.shape)
            bo = -1
            if h.bias is not None:
                bo = off + nw
                nb = h.bias.numel()
                wflat[bo: bo + nb].copy_(h.bias.detach().reshape(-1).to(self.device))
                h.bias.data = wflat[bo: bo + nb]
            self.w_off[id(h)] = (off, bo)
            off += nw + (h.bias.numel() if h.bias is not None else 0)
        self.wflat = wflat

    def repack_on_device(self):
        """Rebuild every packed operand blob from the flat parameter buffer (one gather per blob)."""
        with self.on_device():
            for cl in self.active_layers():
                for k, (mp, add, kind) in cl.maps.items():
                    out = cl.packed[k]
                    check(self.lib.tta_repack_weights(self.wflat.data_ptr(), mp.data_ptr(), add.data_ptr() if add is not None else 0,
                                                      mp.numel(), out.data_ptr(), kind, _stream()), "repack_weights")

    def param_grads(self, params) -> list:
        """Gradient tensors (fresh copies) for the given Parameters after a supervised backward: conv weights /
        biases from gflat, norm affines from the [dgamma || dbeta] buffer (loss scale divided out), None otherwise."""
        look = {}
        for h in self.model.conv_holders():
            wo, bo = self.w_off[id(h)]
            look[id(h.weight)] = self.gflat[wo: wo + h.weight.numel()].view(h.weight.shape)
            if h.bias is not None:
                look[id(h.bias)] = self.gflat[bo: bo + h.bias.numel()]
        P = self.P
        for nl in self.norm_layers:
            if nl.h.weight is not None:
                look[id(nl.h.weight)] = self.dgb[nl.off: nl.off + nl.C] * self.last_inv_scale
                look[id(nl.h.bias)] = self.dgb[P + nl.off: P + nl.off + nl.C] * self.last_inv_scale
        out = []
        for p_ in params:
            g = look.get(id(p_))
            out.append(None if g is None or not p_.requires_grad else g.clone())
        return out

    @contextlib.contextmanager
    def on_device(self):
        """Every launch of this engine runs with ITS device current and on that device's current stream, whatever
        the process-wide current device is (reference ``gpu_ids=[1]``, ``evaluate_epoch(device='cuda:1')``)."""
        global _DEVICE
        prev = _DEVICE
        _DEVICE = self.device
        try:
            with torch.cuda.device(self.device):
                yield
        finally:
            _DEVICE = prev

    @property
    def n_adaptable(self) -> int:
        return 2 * sum(n.C for n in self.norm_layers)

    def _ensure_device(self, device: torch.device, dry: bool = False):
        """``dry=True`` (tests only) builds buffers on the CPU to inspect the op graph; nothing can
        be launched from such a plan."""
        if device.type != "cuda" and not dry:
            raise RuntimeError("multimodal_tta_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self.device is not None and self.device != device:
            self.plans.clear()
            self.gb = None
            self._packed_fp = None
        self.device = device
        if self.gb is None:
            P = self.P
            self.gb = torch.zeros(2 * P, dtype=torch.float32, device=device)
            self.dgb = torch.zeros(2 * P, dtype=torch.float32, device=device)
            self.m = torch.zeros(2 * P, dtype=torch.float32, device=device)
            self.v = torch.zeros(2 * P, dtype=torch.float32, device=device)
            self.step_dev = torch.zeros(1, dtype=torch.int32, device=device)
            self.model._params_dirty = True
        trainable = bool(getattr(self.model, "trainable", False)) and not dry
        if trainable:
            self.model._params_dirty = True      # an optimizer updates the weights in place: check every time
        if self.model._params_dirty:
            self._bind_params()
            if trainable:
                self._bind_conv_params()
            # conv weights are baked into the plans (packed device blobs, host kernel parameters of the small /
            # head kernels, captured CUDA graphs): repack only when a weight tensor really changed
            # (storage, in-place version, device) and then drop every plan built on the old blobs
            fp = tuple((id(h), h.weight.data_ptr(), h.weight._version,
                        h.bias.data_ptr() if h.bias is not None else 0, h.bias._version if h.bias is not None else 0,
                        str(h.weight.device)) for h in self.model.conv_holders())
            fp = (fp, self.model.conv_backend, self.model.t2s_head, self.bwd_dtype)
            fp = (fp, trainable)
            prev = getattr(self, "_packed_fp", None)
            if fp == prev:
                self.model._params_dirty = False
                return
            want_tc = self.model.conv_backend in ("auto", "tc")
            # same storages, only in-place versions moved (an optimizer step, an in-place load_state_dict) and the
            # index maps exist: rebuild the blobs on the device, at their old addresses -- plans and graphs stay valid
            # fp = ((per-holder tuples, backend, t2s, bwd dtype), trainable)
            same_storage = (prev is not None and trainable and prev[1:] == fp[1:] and prev[0][1:] == fp[0][1:]
                            and len(prev[0][0]) == len(fp[0][0])
                            and all(a[:2] == b[:2] and a[3] == b[3] and a[5] == b[5] for a, b in zip(prev[0][0], fp[0][0]))
                            and all(hasattr(cl, "maps") for cl in self.active_layers()))
            self._packed_fp = fp
            if same_storage:
                self.repack_on_device()
                self.model._params_dirty = False
                return
            self.plans.clear()
            for cl in self.active_layers():
                cl.pack(device, want_tc, self.bwd_dtype, t2s=self.model.t2s_head, trainable=trainable)
                if trainable:
                    cl.maps = cl.index_maps(self.w_off, device, want_tc, self.bwd_dtype, self.model.t2s_head)
            self.model._params_dirty = False

    def _bind_params(self):
        """Copy norm affine values into the flat [gamma || beta] buffer and rebind the holders'
        Parameters as views of it, so ``state_dict()`` always shows the adapted values."""
        P = self.P
        for nl in self.norm_layers:
            g = self.gb[nl.off: nl.off + nl.C]
            b = self.gb[P + nl.off: P + nl.off + nl.C]
            if nl.h.weight is not None:
                if nl.h.weight.data_ptr() != g.data_ptr():
                    g.copy_(nl.h.weight.detach().to(self.device))
                    b.copy_(nl.h.bias.detach().to(self.device))
                    nl.h.weight.data = g
                    nl.h.bias.data = b
            else:
                g.fill_(1.0)
                b.zero_()

    def flat_params(self) -> torch.Tensor:
        """Compact [gamma_0..gamma_L || beta_0..beta_L] copy (real channels only)."""
        P = self.P
        gs = [self.gb[nl.off: nl.off + nl.C] for nl in self.norm_layers]
        bs = [self.gb[P + nl.off: P + nl.off + nl.C] for nl in self.norm_layers]
        return torch.cat(gs + bs)

    def flat_grads(self) -> torch.Tensor:
        P = self.P
        gs = [self.dgb[nl.off: nl.off + nl.C] for nl in self.norm_layers]
        bs = [self.dgb[P + nl.off: P + nl.off + nl.C] for nl in self.norm_layers]
        return torch.cat(gs + bs) * self.last_inv_scale

    def reset_optimizer(self):
        self.m.zero_(); self.v.zero_(); self.step_dev.zero_()

    # ---------------------------------------------------------------- op emitters
    def _uses_tc_s2(self, cl: ConvLayer, backward: bool) -> bool:
        """True when this launch is a stride-2 (non-transposed) conv on the tcgen05 kernel, whose
        input planes must be w-parity-split."""
        mode = (1 - cl.mode) if backward else cl.mode
        key = "bwd" if backward else "fwd"
        if ("small_" + key) in cl.packed and self.model.conv_backend != "simt":
            return False
        return (mode == 0 and cl.stride == 2 and cl.K == 3 and ("tc_" + key) in cl.packed
                and self.model.conv_backend in ("auto", "tc"))

    def _conv_call(self, plan: Plan, cl: ConvLayer, backward: bool, src, src_dtype, N, cin8, idims,
                   dst_ptr, dst_ns, cout8, odims, accumulate: bool, wsplit_in: bool = False,
                   stats_res: Optional["Res"] = None, bwd_rec: Optional[dict] = None, c4_in: bool = False,
                   out16: bool = False):
        """Returns a closure launching one conv (tcgen05 kernel when the geometry is supported,
        otherwise the fp32 CUDA-core kernel)."""
        lib = self.lib
        mode = (1 - cl.mode) if backward else cl.mode
        key = "bwd" if backward else "fwd"
        bias = 0 if backward else cl.packed["bias"].data_ptr()
        hi, lo, ns = src
        if ("small_" + key) in cl.packed and self.model.conv_backend != "simt":
            if out16:
                raise _NoGrad16(cl.name)
            wp = cl.packed["small_" + key]
            plan.keep.append(wp)
            plan.conv_backends[f"{cl.name}:{key}"] = "small"
            cin = cl.cout if backward else cl.cin
            cout = cl.cin if backward else cl.cout
            sargs = (hi, lo, ns, src_dtype, N, cin, *idims, wp.data_ptr(), bias, dst_ptr, dst_ns, cout,
                     int(accumulate))

            def run_small():
                check(lib.tta_conv_small(*sargs, _stream()), f"conv_small {cl.name}")
            run_small.label = f"conv_small {cl.name} {'bwd' if backward else 'fwd'}"
            return run_small
        use_tc = ("tc_" + key) in cl.packed and self.model.conv_backend in ("auto", "tc")
        if self.model.conv_backend == "tc" and not use_tc:
            raise RuntimeError(f"conv_backend=tc but {cl.name} ({key}) is not supported by the tcgen05 kernel")
        plan.conv_backends[f"{cl.name}:{key}"] = "tc" if use_tc else "simt"
        if out16 and not use_tc:
            raise _NoGrad16(cl.name)
        if use_tc:
            wp = cl.packed["tc_" + key + ("_c4" if c4_in else "")]
            plan.keep.append(wp)
            flags = (2 if self.model.deterministic else 0) | (8 if (wsplit_in and not c4_in) else 0) | self.model.tc_flags
            if c4_in:
                flags |= 32768
            if out16:
                flags |= 65536
            if not backward:
                flags |= getattr(cl, "tc_fwd_flags", 0)
            args = (hi, lo, ns, src_dtype, N, cin8, *idims, wp.data_ptr(), bias, dst_ptr, dst_ns, cout8,
                    *odims, mode, cl.K, cl.stride, int(accumulate), flags)
            if stats_res is not None:
                ks, grid, nbuf = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
                check(lib.tta_conv_tc_query(src_dtype, N, cin8, *idims, cout8, *odims, mode, cl.K, cl.stride,
                                            int(accumulate), flags, ctypes.byref(ks), ctypes.byref(grid),
                                            ctypes.byref(nbuf)), "conv_tc_query")
                stats_res.tc_query = (ks.value, grid.value, nbuf.value)
                stats_res.t2s = bool((flags >> 8) & 7)
                # the fused head may ask (later) for the compact layout of this result
                odv = odims[0] * odims[1] * odims[2]
                args_c4 = (hi, lo, ns, src_dtype, N, cin8, *idims, wp.data_ptr(), bias, dst_ptr, odv * 4, cout8,
                           *odims, mode, cl.K, cl.stride, int(accumulate), flags | 16384)
            if bwd_rec is not None:
                ks, grid, nbuf = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
                check(lib.tta_conv_tc_query(src_dtype, N, cin8, *idims, cout8, *odims, mode, cl.K, cl.stride,
                                            int(accumulate), flags, ctypes.byref(ks), ctypes.byref(grid),
                                            ctypes.byref(nbuf)), "conv_tc_query")
                bwd_rec["info"] = (ks.value, grid.value, nbuf.value)
                bn_args = (hi, lo, ns, src_dtype, N, cin8, *idims, wp.data_ptr(), dst_ptr, dst_ns, cout8,
                           *odims, mode, cl.K, cl.stride, int(accumulate), flags)

                def run_bwd():
                    # norm layers whose complete gradient this dgrad produces asked (later, at their own
                    # emission) for the backward reductions to come out of this epilogue
                    if bwd_rec["segs"]:
                        if "carr" not in bwd_rec:
                            arr = (NormBwdSeg * len(bwd_rec["segs"]))()
                            for i, sg in enumerate(bwd_rec["segs"]):
                                arr[i] = NormBwdSeg(*sg)
                            bwd_rec["carr"] = arr
                        check(lib.tta_conv_tc_bwd_norm(*bn_args, bwd_rec["carr"], len(bwd_rec["segs"]), _stream()),
                              f"conv_tc_bwd_norm {cl.name}")
                    else:
                        check(lib.tta_conv_tc(*args, 0, 0, _stream()), f"conv_tc {cl.name}")
                run_bwd.label = f"conv_tc {cl.name} bwd {cin8 * 8}->{cout8 * 8} in {tuple(idims)}"
                return run_bwd

            def run():
                # fused statistics are requested (later, by the norm that consumes this result) via the Res
                st = (plan.ws.data_ptr(), stats_res.stats_c8) if (stats_res is not None and stats_res.stats_c8) \
                    else (0, 0)
                a = args_c4 if (stats_res is not None and stats_res.c4) else args
                check(lib.tta_conv_tc(*a, *st, _stream()), f"conv_tc {cl.name}")
        else:
            wp = cl.packed["simt_" + key]
            plan.keep.append(wp)
            args = (hi, lo, ns, src_dtype, N, cin8, *idims, wp.data_ptr(), bias, dst_ptr, dst_ns, cout8,
                    *odims, mode, cl.K, cl.stride, int(accumulate))

            def run():
                check(lib.tta_conv_simt(*args, _stream()), f"conv_simt {cl.name}")
        run.label = (f"conv_{'tc' if use_tc else 'simt'} {cl.name} {'bwd' if backward else 'fwd'} "
                     f"{cin8 * 8}->{cout8 * 8} in {tuple(idims)}")
        return run

    def build_plan(self, N: int, D: int, H: int, W: int) -> Plan:
        model = self.model
        want = (getattr(model, "grad_f16", False) and getattr(model, "supports_grad_f16", True) and not model.trainable and self.bwd_dtype == TTA_F16_HI
                and model.conv_backend in ("auto", "tc") and not model.fuse_bwd_stats
                and not model.per_sample_norm_bwd and model.norm_bwd_l2_mb <= 0)
        if want:
            try:
                return self._build_plan(N, D, H, W, True)
            except _NoGrad16:
                pass
        return self._build_plan(N, D, H, W, False)

    def _build_plan(self, N: int, D: int, H: int, W: int, grad16: bool) -> Plan:
        dev = self.device
        lib = self.lib
        model = self.model
        P = self.P
        NB = N                       # the plan's batch (N is rebound per tensor further down: batched aliases)
        R = model.out_channels
        x = Act(N, model.in_channels, D, H, W, dev, needs_grad=False, name="x")
        plan = Plan(N=N, dims=(D, H, W), x=x,
                    logits=torch.zeros((N, R, D, H, W), dtype=torch.float32, device=dev),
                    loss=torch.zeros(1, dtype=torch.float32, device=dev))
        ops: list = []  # forward op records, replayed in reverse to emit the backward
        max_ws = [1]
        plan.grad16 = grad16

        def g16_for(V_: int) -> bool:
            # levels whose dgrads run without split-K anyway; the full-resolution tensors belong to the fused head
            return grad16 and model.grad_f16_min_voxels <= V_ < D * H * W

        def conv(cl: ConvLayer, inp: ActView) -> Res:
            N = inp.parent.N            # instances of THIS tensor (a batched alias has N * modalities)
            d, h, w = inp.dims
            s = cl.stride
            if cl.mode == 0:
                od, oh, ow = (d - 1) // s + 1, (h - 1) // s + 1, (w - 1) // s + 1
            else:
                od, oh, ow = d * s, h * s, w * s
            if inp.C8 != (cl.cin + 7) // 8:
                raise ValueError(f"{cl.name}: input has {inp.C8} channel chunks, layer expects {cl.cin} channels")
            y = Res(N, cl.cout, od, oh, ow, dev, name=cl.name)
            plan.keep.append(y)
            src = (inp.hi, inp.lo, inp.ns)
            ws_in = self._uses_tc_s2(cl, False) and w % 2 == 0
            if inp.parent is plan.x:
                # the packed network input has ONE copy: every consumer must agree on its layout
                prev = getattr(plan, "x_layout", None)
                if prev is not None and prev != ws_in:
                    raise ValueError("unet_b200: the network input feeds both a strided and an unstrided conv")
                plan.x_layout = ws_in
            c4_in = False
            if ws_in and inp.parent is plan.x and model.input_compact and not model.trainable and "tc_fwd_c4" in cl.packed \
                    and model.conv_backend in ("auto", "tc") and d % 2 == 0 and h % 2 == 0:
                # <= 4 input channels: the packed input in the compact layout (8 B per voxel and plane)
                c4_in = True
                plan.x.compact = True
                src = (plan.x.planes[0].data_ptr(), plan.x.planes[1].data_ptr(), plan.x.V * 4)
                prev_c = getattr(plan, "x_compact", None)
                if prev_c is False:
                    raise ValueError("unet_b200: the network input feeds convs that disagree on its layout")
            if inp.parent is plan.x:
                plan.x_compact = c4_in
            if ws_in and not c4_in:
                par = inp.parent
                if par is plan.x:
                    par.wsplit = True          # gather_pack writes the only copy parity-split
                else:
                    par.need_ws_copy()         # the producing norm writes a second, parity-split copy
                    src = (inp.ws_hi, inp.ws_lo, inp.ns)
            run = self._conv_call(plan, cl, False, src, TTA_F16, N, inp.C8, inp.dims,
                                  y.ptr, y.ns, y.C8, (od, oh, ow), False, wsplit_in=ws_in, stats_res=y, c4_in=c4_in)
            plan.fwd.append(run)
            ops.append(("conv", cl, inp, y))
            return y

        def normact(nl: NormLayer, y: Res, relu: bool, residual, out: Optional[ActView]) -> ActView:
            N = y.N
            if out is None:
                a = Act(N, y.C, y.D, y.H, y.W, dev, needs_grad=True, name=nl.name, g16=g16_for(y.V))
                plan.keep.append(a)
                out = a.view()
            if (out.C8, *out.dims) != (y.C8, y.D, y.H, y.W) or out.parent.N != N:
                raise ValueError(f"{nl.name}: output view shape mismatch")
            if y.V * (N if nl.batch else 1) == 1:
                plan.single_element_norm = nl.name
            mean = torch.zeros(N * y.C8 * 8, dtype=torch.float32, device=dev)
            rstd = torch.zeros_like(mean)
            sums = torch.zeros(N * y.C8 * 8 * 2, dtype=torch.float32, device=dev)
            plan.keep += [mean, rstd, sums]
            max_ws[0] = max(max_ws[0], lib.tta_norm_workspace_floats(N, y.C8, y.V),
                            lib.tta_norm_workspace_floats(1, y.C8, y.V))   # (per-sample backward launches)
            gptr = self.gb.data_ptr() + nl.off * 4
            bptr = self.gb.data_ptr() + (P + nl.off) * 4
            rec = dict(nl=nl, y=y, relu=relu, residual=residual, out=out, mean=mean, rstd=rstd, sums=sums,
                       gptr=gptr, bptr=bptr)
            if residual is None:
                rk, ra, rb, rns = 0, 0, 0, 0
            elif isinstance(residual, (Res, ResView)):
                rk, ra, rb, rns = 1, residual.ptr, 0, residual.ns
            else:
                rk, ra, rb, rns = 2, residual.hi, residual.lo, residual.ns
            # statistics from the producing conv's epilogue: tcgen05 launch without split-K, and this
            # norm reads the LEADING channel chunks of the conv result (whole result, or the unit0 half
            # of a fused unit0 || shortcut conv)
            q = y.root.tc_query
            # ... and double-buffered TMEM accumulators, so the reduction hides behind the next item's MMAs
            fuse_st = (model.fuse_stats and q is not None and q[0] == 1
                       and (q[2] == 2 or model.fuse_stats_single_buffer) and y.c8_off == 0
                       and y.root.stats_c8 == 0)
            if fuse_st:
                y.root.stats_c8, y.root.stats_grid = y.C8, q[1]
                max_ws[0] = max(max_ws[0], 1024 + N * y.C8 * q[1] * 16)
            rec["fused_stats"] = fuse_st
            st_args = (y.ptr, y.ns, N, y.C8, y.V, nl.batch, float(nl.h.eps), mean.data_ptr(), rstd.data_ptr())
            ap_args = (y.ptr, y.ns, N, y.C8, y.V, mean.data_ptr(), rstd.data_ptr(), gptr, bptr, int(relu),
                       rk, ra, rb, rns, out.hi, out.lo, out.ns, TTA_F16)
            plan.stats_ops.append((nl, mean, rstd, y))
            rec["st_args"] = st_args

            def ws_args():
                # resolved at launch time: a later strided conv may have asked for the parity-split copy
                if out.parent.ws_planes is None:
                    return (0, 0, 0, 0)
                return (out.ws_hi, out.ws_lo, out.ns, y.W)

            small = (model.fuse_small_norm and not fuse_st and y.V <= model.small_norm_max_voxels
                     and bool(lib.tta_norm_small_supported(N, y.V, nl.batch)))
            rec["small"] = small
            sm_args = (y.ptr, y.ns, N, y.C8, y.V, float(nl.h.eps), mean.data_ptr(), rstd.data_ptr(), gptr, bptr,
                       int(relu), rk, ra, rb, rns, out.hi, out.lo, out.ns, TTA_F16)

            def run():
                if small and not (nl.batch and not model.training and nl.h.track_running_stats):
                    check(lib.tta_norm_fwd_small(*sm_args, *ws_args(), _stream()), "norm_fwd_small")
                    return
                if nl.batch and not model.training and nl.h.track_running_stats:
                    # eval-mode BatchNorm: mean/rstd were filled from the running buffers
                    check(lib.tta_norm_apply(*ap_args, 0, 0, nl.batch, float(nl.h.eps), *ws_args(), _stream()),
                          "norm_apply")
                elif fuse_st and model.stats_finalize_in_apply:
                    # ... finalized by every block of the apply pass in its prologue (one launch less)
                    check(lib.tta_norm_apply(*ap_args, plan.ws.data_ptr(), y.root.stats_grid, nl.batch,
                                             float(nl.h.eps), *ws_args(), _stream()), "norm_apply")
                elif fuse_st:
                    # partial sums were left in the workspace by the conv epilogue: tiny parallel finalize
                    check(lib.tta_norm_stats_finalize(plan.ws.data_ptr(), N, y.C8, y.root.stats_grid, y.V, nl.batch,
                                                      float(nl.h.eps), mean.data_ptr(), rstd.data_ptr(), _stream()),
                          "norm_stats_finalize")
                    check(lib.tta_norm_apply(*ap_args, 0, 0, nl.batch, float(nl.h.eps), *ws_args(), _stream()),
                          "norm_apply")
                else:
                    # single-pass statistics: the last block of every chunk finalizes mean/rstd
                    check(lib.tta_norm_stats(*st_args, plan.ws.data_ptr(), 1, _stream()), "norm_stats")
                    check(lib.tta_norm_apply(*ap_args, 0, 0, nl.batch, float(nl.h.eps), *ws_args(), _stream()),
                          "norm_apply")
            run.label = f"norm fwd {nl.name} C={nl.C} V={y.V}"
            plan.fwd.append(run)
            ops.append(("norm", rec))
            return out

        def convolution(cv: ConvolutionH, inp: ActView, out: Optional[ActView]):
            y = conv(self.conv_layers[id(cv.conv)], inp)
            if cv.conv_only:
                return y
            return normact(self._nl(cv.adn.N), y, True, None, out)

        def residual_unit(ru: ResidualUnitH, inp: ActView, out: Optional[ActView]):
            units = list(ru.conv.children())
            u0 = units[0].conv
            fused = self.fused_layers.get(id(u0))
            y_first = None
            if fused is not None:
                # unit0 || strided shortcut share input and geometry: ONE conv with concatenated couts
                yf = conv(fused, inp)
                c = u0.cout
                y_first = yf.view(0, c // 8, c)
                res = yf.view(c // 8, c // 8, c)
            elif isinstance(ru.residual, ConvHolder):
                res = conv(self.conv_layers[id(ru.residual)], inp)
            else:
                res = inp
            cur = inp
            for i, u in enumerate(units):
                last = i == len(units) - 1
                y = y_first if (i == 0 and y_first is not None) else conv(self.conv_layers[id(u.conv)], cur)
                if u.conv_only:
                    return y  # identity shortcut folded into the centre tap (ConvLayer.fold_identity)
                cur = normact(self._nl(u.adn.N), y, True, res if last else None, out if last else None)
            return cur

        def layer(mod, inp: ActView, out: Optional[ActView]):
            if isinstance(mod, ResidualUnitH):
                return residual_unit(mod, inp, out)
            if isinstance(mod, ConvolutionH):
                return convolution(mod, inp, out)
            if isinstance(mod, nn.Sequential):  # up path: Sequential(convT-Convolution, ResidualUnit)
                cur = inp
                mods = list(mod.children())
                for i, sm in enumerate(mods):
                    cur = layer(sm, cur, out if i == len(mods) - 1 else None)
                return cur
            raise TypeError(type(mod))

        # ---- ops beyond the plain UNet (multimodal model): modality mean, trilinear upsample, fp32 -> operand cast
        def new_act(C, dims, needs_grad=True, name="", n=None) -> Act:
            a = Act(N if n is None else n, C, *dims, dev, needs_grad=needs_grad, name=name,
                    g16=g16_for(dims[0] * dims[1] * dims[2]))
            plan.keep.append(a)
            return a

        def mean(inputs: List[ActView], dst: ActView, rep: int = 1):
            """dst[(n*rep + r)] = mean_k inputs[k][n] for r < rep (operand planes)."""
            k = len(inputs)
            n_in = inputs[0].parent.N
            if any((v.C8, *v.dims, v.parent.N) != (dst.C8, *dst.dims, n_in) for v in inputs) or dst.parent.N != n_in * rep:
                raise ValueError("mean: shape mismatch")
            his = (ctypes.c_void_p * k)(*[v.hi for v in inputs])
            los = (ctypes.c_void_p * k)(*[v.lo for v in inputs])
            nss = (ctypes.c_longlong * k)(*[v.ns for v in inputs])
            V = dst.parent.V
            args = (his, los, nss, k, n_in, dst.C8, V, 1.0 / k, dst.hi, dst.lo, dst.ns, rep)

            def run():
                check(lib.tta_mean_planes(*args, _stream()), "mean_planes")
            run.label = f"mean of {k} x C={dst.C8 * 8} V={V} rep={rep}"
            run.launches = 1
            plan.fwd.append(run)
            plan.keep += [his, los, nss]
            ops.append(("mean", inputs, dst, rep))

        def upsample(y: Res, dst: ActView):
            """nn.Upsample(trilinear, align_corners=True) of a conv-only fp32 result into operand planes."""
            if y.C8 != dst.C8 or y.N != dst.parent.N:
                raise ValueError("upsample: shape mismatch")
            od = dst.dims
            args = (y.ptr, y.ns, y.N, y.C8, y.D, y.H, y.W, *od, dst.hi, dst.lo, dst.ns, TTA_F16)

            def run():
                check(lib.tta_upsample_fwd(*args, _stream()), "upsample_fwd")
            run.label = f"upsample C={y.C8 * 8} {(y.D, y.H, y.W)} -> {tuple(od)}"
            run.launches = 1
            plan.fwd.append(run)
            ops.append(("upsample", y, dst))

        def cast(y: Res, dst: Optional[ActView] = None) -> ActView:
            """fp32 conv-only result -> operand planes (two consecutive linear layers, no norm in between)."""
            if dst is None:
                dst = new_act(y.C, (y.D, y.H, y.W), name=y.name + "/cast", n=y.N).view()
            args = (y.ptr, y.ns, 0, 0, y.N, y.C8, y.V, dst.hi, dst.lo, dst.ns, TTA_F16)

            def run():
                check(lib.tta_split_f32(*args, _stream()), "split_f32")
            run.label = f"cast C={y.C8 * 8} V={y.V}"
            run.launches = 1
            plan.fwd.append(run)
            ops.append(("cast", y, dst))
            return dst

        import types
        G = types.SimpleNamespace(N=N, dims=(D, H, W), x=x, plan=plan, engine=self, conv=conv, normact=normact,
                                  convolution=convolution, residual_unit=residual_unit, layer=layer, new_act=new_act,
                                  mean=mean, upsample=upsample, cast=cast, conv_layer=lambda h: self.conv_layers[id(h)],
                                  norm_layer=self._nl)

        def final_dims(r):
            return r.D, r.H, r.W

        final = model.build_graph(G)        # the model walks its own module tree with the emitters above
        if not isinstance(final, Res):
            raise ValueError("unet_b200: the network must end in a conv (MONAI UNet does)")
        bdt = self.bwd_dtype
        nplanes = 1 if bdt == TTA_F16_HI else 2
        # ---- fused full-resolution tail: [norm apply -> small 3x3x3 conv -> entropy] as one kernel
        fused_head = None
        if (model.fuse_head and model.conv_backend != "simt" and len(ops) >= 2 and ops[-1][0] == "conv"
                and ops[-2][0] == "norm"):
            _, hcl, hinp, hy = ops[-1]
            hrec = ops[-2][1]
            if (hy is final and hrec["out"] is hinp and hrec["residual"] is None and hcl.mode == 0
                    and "small_fwd" in hcl.packed and hrec["y"].C8 == 1 and hcl.cout == R
                    and lib.tta_head_fused_supported(hcl.K, hcl.stride, hcl.cin, hcl.cout)):
                fused_head = (hcl, hrec, hinp)
                max_ws[0] = max(max_ws[0], lib.tta_head_fused_workspace_floats(N, *final_dims(final)))
        plan.fused_head = fused_head is not None
        plan.final = final
        # compact layout of the head's tensors (conv result fp32 x4, masked gradient fp16 x4 per voxel instead of
        # 8-channel chunks): when the small-Cout transposed kernel produces the result AND its statistics (no pass
        # of the generic 8-channel norm kernels ever touches it)
        head_c4 = False
        if fused_head is not None and model.head_compact:
            hrec_ = fused_head[1]
            hy_ = hrec_["y"]
            if isinstance(hy_, Res) and hy_.t2s and hrec_["fused_stats"]:
                head_c4 = hy_.c4 = True
        plan.head_c4 = head_c4
        plan.ws = torch.zeros(max_ws[0], dtype=torch.float32, device=dev)
        if fused_head is None:
            final.alloc_dy(nplanes)
        # power-of-two loss scale keeps the fp16 gradient planes in range; Adam divides it out
        plan.loss_scale = float(2 ** math.ceil(math.log2(4.0 * N * final.V * model.loss_scale_mult))) \
            if bdt == TTA_F16_HI else 1.0
        nblk = lib.tta_head_entropy_blocks(N, final.V)
        plan.partial = torch.zeros(nblk * N, dtype=torch.float32, device=dev)
        plan.sample_w = torch.ones(N, dtype=torch.float32, device=dev)
        plan.inv_count = 1.0 / (N * final.V)

        def head(train: bool):
            def run():
                check(lib.tta_head_entropy(
                    final.ptr, final.ns, N, R, final.V, self.entropy_mode, float(plan.inv_count),
                    float(plan.loss_scale), bdt, plan.sample_w.data_ptr(), plan.logits.data_ptr(),
                    final.dy_ptr(0) if train else 0, final.dy_ptr(1) if train else 0, final.ns,
                    plan.partial.data_ptr(), plan.loss.data_ptr(), _stream()), "head_entropy")
            return run

        if fused_head is not None:
            hcl, hrec, hinp = fused_head
            hy, hnl = hrec["y"], hrec["nl"]
            wp = hcl.packed["small_fwd"]          # HOST [27][8][8], identity shortcut folded in
            plan.keep.append(wp)
            plan.dlogits = torch.zeros((N, R, D, H, W), dtype=torch.float32, device=dev)
            plan.conv_backends[f"{hcl.name}:fwd"] = "head"
            plan.conv_backends[f"{hcl.name}:bwd"] = "head"
            # the norm's apply and the conv disappear from the forward list: statistics only
            del plan.fwd[-2:]

            def stats_only(st_args=hrec["st_args"], nl=hnl):
                if nl.batch and not model.training and nl.h.track_running_stats:
                    return
                if hrec["fused_stats"]:   # partials came out of the transposed conv's epilogue
                    check(lib.tta_norm_stats_finalize(plan.ws.data_ptr(), N, hy.C8, hy.root.stats_grid, hy.V,
                                                      nl.batch, float(nl.h.eps), hrec["mean"].data_ptr(),
                                                      hrec["rstd"].data_ptr(), _stream()), "norm_stats_finalize")
                else:
                    check(lib.tta_norm_stats(*st_args, plan.ws.data_ptr(), 1, _stream()), "norm_stats")
            plan.fwd.append(stats_only)

            def head(train: bool):  # noqa: F811  (fused variant replaces the streaming head)
                def run():
                    check(lib.tta_head_fused_fwd(
                        hy.ptr, hy.V * 4 if head_c4 else hy.ns, 4 if head_c4 else 8, N, R, hy.D, hy.H, hy.W,
                        hrec["mean"].data_ptr(), hrec["rstd"].data_ptr(),
                        hrec["gptr"], hrec["bptr"], int(hrec["relu"]), wp.data_ptr(),
                        hcl.packed["bias"].data_ptr(), self.entropy_mode, float(plan.inv_count),
                        float(plan.loss_scale), plan.sample_w.data_ptr(), plan.logits.data_ptr(),
                        plan.dlogits.data_ptr() if train else 0, plan.ws.data_ptr(), plan.loss.data_ptr(),
                        _stream()), "head_fused_fwd")
                return run
        plan.head_infer, plan.head_train = head(False), head(True)

        # ------------------------------------------------------------ backward emission
        bwd_apply_flags = []
        for op in reversed(ops):
            if op[0] == "mean":
                # d(mean)/d(input k) = g / K for every input; g = sum over the (replicated) destination's gradient
                # sources: its own grad slice (written by the dgrads of its consumers) and identity-residual extras
                _, inputs, dst, rep = op
                par = dst.parent
                if par.g16 or any(v.parent.g16 for v in inputs):
                    raise _NoGrad16("mean")
                chunks = set(range(dst.c8_off, dst.c8_off + dst.C8))
                srcs = []
                if chunks <= par.written:
                    srcs.append((dst.g, dst.ns))
                srcs += [(ev[0], ev[1]) for ev in par.extra if (ev[2], ev[3]) == (dst.c8_off, dst.C8)]
                if not srcs:
                    raise RuntimeError("mean: no gradient reaches the destination")
                n_in, V = inputs[0].parent.N, par.V
                tmp = torch.zeros((n_in, dst.C8, V, 8), dtype=torch.float32, device=dev)
                plan.keep.append(tmp)
                ps = (ctypes.c_void_p * len(srcs))(*[p_ for p_, _ in srcs])
                nss = (ctypes.c_longlong * len(srcs))(*[n_ for _, n_ in srcs])
                plan.keep += [ps, nss]
                sm_args = (ps, nss, len(srcs), rep, n_in, dst.C8, V, 1.0 / len(inputs), tmp.data_ptr(), dst.C8 * V * 8, 0)

                def run_mean_bwd(sm_args=sm_args):
                    check(lib.tta_sum_f32(*sm_args, _stream()), "sum_f32")
                run_mean_bwd.label = f"mean bwd C={dst.C8 * 8} V={V}"
                run_mean_bwd.launches = 1
                plan.bwd.append(run_mean_bwd)
                for v in inputs:
                    v.parent.extra.append((tmp.data_ptr(), dst.C8 * V * 8, v.c8_off, v.C8))
                continue
            if op[0] in ("upsample", "cast"):
                # gradient of the operand tensor (fp32, complete) -> 16-bit gradient plane(s) of the conv-only result
                _, y, dst = op
                par = dst.parent
                if par.g16:
                    raise _NoGrad16(op[0])
                chunks = set(range(dst.c8_off, dst.c8_off + dst.C8))
                if not chunks <= par.written or any((ev[2], ev[3]) == (dst.c8_off, dst.C8) for ev in par.extra):
                    raise RuntimeError(f"{op[0]}: the destination's gradient must come from dgrads only")
                y.alloc_dy(nplanes)
                if op[0] == "upsample":
                    u_args = (dst.g, dst.ns, y.N, y.C8, y.D, y.H, y.W, *dst.dims, y.dy_ptr(0), y.dy_ptr(1), y.ns, bdt)

                    def run_up_bwd(u_args=u_args):
                        check(lib.tta_upsample_bwd(*u_args, _stream()), "upsample_bwd")
                else:
                    u_args = (dst.g, dst.ns, 0, 0, y.N, y.C8, y.V, y.dy_ptr(0), y.dy_ptr(1), y.ns, bdt)

                    def run_up_bwd(u_args=u_args):
                        check(lib.tta_split_f32(*u_args, _stream()), "split_f32")
                run_up_bwd.label = f"{op[0]} bwd C={y.C8 * 8} V={y.V}"
                run_up_bwd.launches = 1
                plan.bwd.append(run_up_bwd)
                continue
            if op[0] == "conv":
                _, cl, inp, y = op
                N = y.N
                if model.trainable:
                    # supervised step: weight / bias gradients from the operands both passes hold (forward input
                    # planes, the output's gradient planes), straight into the parameters' .grad storage
                    if cl.w_map is not None or cl.dgrad_cin is not None:
                        raise NotImplementedError("supervised step: effective-weight layers are not supported")
                    y.alloc_dy(nplanes)
                    xpar = inp.parent
                    if xpar.compact:
                        raise RuntimeError("supervised step: compact input layout")
                    wo, bo = self.w_off[id(cl.h)]
                    wo2, bo2 = self.w_off[id(cl.extra)] if cl.extra is not None else (0, -1)
                    gbase = self.gflat.data_ptr()
                    wg_args = (inp.hi, inp.lo, inp.ns, *inp.dims, int(xpar.wsplit and xpar is plan.x),
                               y.dy_ptr(0), y.dy_ptr(1), y.ns, bdt, y.D, y.H, y.W, 0, N, cl.mode, cl.K, cl.stride,
                               cl.h.cin, cl.cout, 1.0 / plan.loss_scale, gbase + wo * 4, 1 if cl.h.transposed else 0,
                               cl.h.cout if cl.extra is not None else 0, gbase + wo2 * 4)
                    bg_args = None
                    if bo >= 0:
                        bg_args = (y.dy_ptr(0), y.dy_ptr(1), y.ns, bdt, N, cl.cout, y.V, 1.0 / plan.loss_scale,
                                   gbase + bo * 4, cl.h.cout if cl.extra is not None else 0,
                                   gbase + bo2 * 4 if bo2 >= 0 else 0)

                    # tensor-core variant (tta_conv_wgrad_tc): 3x3x3 layers with one scaled fp16 gradient plane; the
                    # finer operand of a stride-2 layer must be the w-parity-split copy its forward / dgrad conv reads
                    tc_args = None
                    if (model.wgrad_backend == "auto" and model.conv_backend in ("auto", "tc") and
                            lib.tta_conv_wgrad_tc_supported(cl.mode, cl.K, cl.stride, cl.h.cin, cl.cout, bdt)):
                        x_ws = self._uses_tc_s2(cl, False) and inp.dims[2] % 2 == 0
                        xp = None
                        if cl.mode == 0 and cl.stride == 2:
                            if x_ws and xpar is plan.x and xpar.wsplit:
                                xp = (inp.hi, inp.lo)
                            elif x_ws and xpar.ws_planes is not None:
                                xp = (inp.ws_hi, inp.ws_lo)
                        elif not (xpar.wsplit and xpar is plan.x):
                            xp = (inp.hi, inp.lo)
                        if xp is not None:
                            tc_args = (xp[0], xp[1], inp.ns, *inp.dims, int(cl.mode == 0 and cl.stride == 2),
                                       y.dy_ptr(0), y.ns, y.D, y.H, y.W, 0, N, cl.mode, cl.stride, cl.h.cin, cl.cout,
                                       1.0 / plan.loss_scale, gbase + wo * 4, 1 if cl.h.transposed else 0,
                                       cl.h.cout if cl.extra is not None else 0, gbase + wo2 * 4,
                                       0 if model.wgrad_x_lo else 1)
                    plan.wgrad_backends = getattr(plan, "wgrad_backends", {})

                    def run_wgrad(wg_args=wg_args, bg_args=bg_args, yroot=y.root, name=cl.name, tc_args=tc_args,
                                  tr=cl.mode == 1):
                        dyw = int(yroot.dy_wsplit)            # resolved at launch time, like the dgrad's flag
                        if tc_args is not None and dyw == int(tr):
                            a = list(tc_args)
                            a[12] = dyw
                            plan.wgrad_backends[name] = "tc"
                            check(lib.tta_conv_wgrad_tc(*a, _stream()), f"conv_wgrad_tc {name}")
                        else:
                            a = list(wg_args)
                            a[14] = dyw
                            plan.wgrad_backends[name] = "simt"
                            check(lib.tta_conv_wgrad(*a, _stream()), f"conv_wgrad {name}")
                        if bg_args is not None:
                            check(lib.tta_bias_grad(*bg_args, _stream()), f"bias_grad {name}")
                    run_wgrad.label = f"wgrad {cl.name} {cl.h.cin}->{cl.cout}"
                    run_wgrad.launches = 2 if bg_args is not None else 1
                    plan.bwd.append(run_wgrad)
                    plan.n_wgrad = getattr(plan, "n_wgrad", 0) + run_wgrad.launches
                if not inp.parent.needs_grad:
                    continue
                par = inp.parent
                c8o = (cl.dgrad_cin + 7) // 8 if cl.dgrad_cin is not None else inp.C8
                chunks = set(range(inp.c8_off, inp.c8_off + c8o))
                done = chunks & par.written
                if done and done != chunks:
                    raise RuntimeError("partial gradient accumulation state")
                acc = bool(done)
                par.mark_written(chunks)
                if fused_head is not None and cl is fused_head[0]:
                    # fused tail: dgrad of the small conv + ReLU mask + norm-backward reduction in ONE
                    # kernel; the masked gradient lands in inp.g, sums/dgamma/dbeta are finalized
                    hrec, hnl = fused_head[1], fused_head[1]["nl"]
                    hyv = hrec["y"].V
                    hb_args = (plan.dlogits.data_ptr(), N, R, y.D, y.H, y.W, cl.packed["small_fwd"].data_ptr(),
                               hrec["y"].ptr, hyv * 4 if head_c4 else hrec["y"].ns, 4 if head_c4 else 8,
                               hrec["mean"].data_ptr(), hrec["rstd"].data_ptr(),
                               hrec["gptr"], hrec["bptr"], int(hrec["relu"]), hnl.batch, inp.g,
                               hyv * 4 if head_c4 else inp.ns,
                               hrec["sums"].data_ptr(), self.dgb.data_ptr() + hnl.off * 4,
                               self.dgb.data_ptr() + (P + hnl.off) * 4)

                    def run_head_bwd(hb_args=hb_args):
                        check(lib.tta_head_fused_bwd(*hb_args, plan.ws.data_ptr(), _stream()), "head_fused_bwd")
                    plan.bwd.append(run_head_bwd)
                    continue
                y.alloc_dy(nplanes)
                crec = dict(segs=[], info=None)
                dyc4 = y.root.dy_c4
                plan.bwd.append(self._conv_call(
                    plan, cl, True, (y.dy_ptr(0), y.dy_ptr(1), y.V * 4 if dyc4 else y.ns), bdt, N, y.C8,
                    (y.D, y.H, y.W), inp.g, inp.ns, c8o, inp.dims, acc, wsplit_in=y.root.dy_wsplit,
                    bwd_rec=crec, c4_in=dyc4, out16=par.g16))
                par.writers.append((inp.c8_off, inp.c8_off + c8o, crec, acc))
            else:
                rec = op[1]
                nl, y, out = rec["nl"], rec["y"], rec["out"]
                N = y.N
                par = out.parent
                chunks = set(range(out.c8_off, out.c8_off + out.C8))
                srcs = []
                if chunks <= par.written:
                    srcs.append((out.g, out.ns, par.g16))
                for ev in par.extra:
                    if (ev[2], ev[3]) == (out.c8_off, out.C8):
                        srcs.append((ev[0], ev[1], len(ev) > 4 and ev[4]))
                if not srcs or len(srcs) > 2:
                    raise RuntimeError(f"{nl.name}: {len(srcs)} gradient sources (supported: 1 or 2)")
                g0, g0ns, g0h = srcs[0]
                g1, g1ns, g1h = srcs[1] if len(srcs) > 1 else (0, 0, False)
                # gradient source formats ride in the relu argument: bit 1 / bit 2 = g0 / g1 is one fp16 plane
                relu_flags = int(rec["relu"]) | (2 if g0h else 0) | (4 if g1h else 0)
                # backward reductions from the epilogue of the dgrad that COMPLETES this gradient: single
                # source (the parent's grad slice), every writer of the slice covers it, the last one is
                # a tcgen05 launch without split-K on one fp16 plane
                fuse_bwd = None
                cb0, cb1 = out.c8_off, out.c8_off + out.C8
                if (model.fuse_bwd_stats and not par.g16 and bdt == TTA_F16_HI and len(srcs) == 1 and chunks <= par.written
                        and not (fused_head is not None and rec is fused_head[1])):
                    touching = [w for w in par.writers if w[0] < cb1 and w[1] > cb0]
                    if touching and all(w[0] <= cb0 and w[1] >= cb1 for w in touching):
                        last = touching[-1]
                        info = last[2]["info"]
                        if (info is not None and info[0] == 1 and len(last[2]["segs"]) < 2
                                and last[3] == (len(touching) > 1)):
                            pbuf = torch.zeros(N * y.C8 * info[1] * 16, dtype=torch.float32, device=dev)
                            plan.keep.append(pbuf)
                            last[2]["segs"].append((cb0 - last[0], y.C8, int(rec["relu"]), 0, y.ptr, y.ns,
                                                    rec["mean"].data_ptr(), rec["rstd"].data_ptr(), rec["gptr"],
                                                    rec["bptr"], pbuf.data_ptr()))
                            fuse_bwd = (pbuf, info[1])
                # (supervised step: every conv needs its output gradient for the weight gradient)
                conv_in_needs = self._producer_input_needs_grad(ops, y) or model.trainable
                res = rec["residual"]
                aux = None
                if isinstance(res, (Res, ResView)) and (self._producer_input_needs_grad(ops, res) or model.trainable):
                    res.alloc_dy(nplanes)
                    aux = res
                elif isinstance(res, ActView):
                    # identity shortcut: this op's incoming gradient also flows into `res`
                    if len(srcs) != 1:
                        raise RuntimeError("identity residual with two incoming gradients is unsupported")
                    res.parent.extra.append((g0, g0ns, res.c8_off, res.C8, g0h))
                dg = self.dgb.data_ptr() + nl.off * 4
                db = self.dgb.data_ptr() + (P + nl.off) * 4
                rd_args = (g0, g0ns, g1, g1ns, y.ptr, y.ns, N, y.C8, nl.C, y.V, rec["mean"].data_ptr(),
                           rec["rstd"].data_ptr(), rec["gptr"], rec["bptr"], relu_flags, nl.batch,
                           rec["sums"].data_ptr(), dg, db)
                do_apply = conv_in_needs or aux is not None
                bwd_apply_flags.append(do_apply)
                dy_ws = 0
                if do_apply:
                    y.alloc_dy(nplanes)
                    pcl = self._producer_conv(ops, y)
                    if self._uses_tc_s2(pcl, True) and y.W % 2 == 0 and y.root is y:
                        y.dy_wsplit = True
                        dy_ws = y.W
                    ap_args = (g0, g0ns, g1, g1ns, y.ptr, y.ns, N, y.C8, y.V, rec["mean"].data_ptr(),
                               rec["rstd"].data_ptr(), rec["gptr"], rec["bptr"], relu_flags, nl.batch,
                               rec["sums"].data_ptr(), y.dy_ptr(0), y.dy_ptr(1), y.ns,
                               aux.dy_ptr(0) if aux else 0, aux.dy_ptr(1) if aux else 0,
                               aux.ns if aux else 0, bdt)

                skip_reduce = fused_head is not None and rec is fused_head[1]   # done by tta_head_fused_bwd
                c4_args = None
                if skip_reduce and head_c4 and do_apply:
                    pcl_ = self._producer_conv(ops, y)
                    if (model.input_compact and "tc_bwd_c4" in pcl_.packed and self._uses_tc_s2(pcl_, True)
                            and y.root is y and y.D % 2 == 0 and y.H % 2 == 0 and y.W % 2 == 0):
                        y.dy_c4, y.dy_wsplit = True, False       # its dgrad reads the compact gradient planes
                    c4_args = (g0, y.V * 4, y.ptr, y.V * 4, N, nl.C, y.V, rec["mean"].data_ptr(), rec["rstd"].data_ptr(),
                               rec["gptr"], rec["bptr"], nl.batch, rec["sums"].data_ptr(), y.dy_ptr(0), y.dy_ptr(1),
                               y.V * 4 if y.dy_c4 else y.ns, bdt, 0 if y.dy_c4 else dy_ws, int(y.dy_c4))
                fin_args = None
                if fuse_bwd is not None:
                    fin_args = (fuse_bwd[0].data_ptr(), N, y.C8, nl.C, fuse_bwd[1], nl.batch, rec["sums"].data_ptr(),
                                dg, db)
                rec["fused_bwd"] = fuse_bwd is not None

                sm_bwd = None
                if (model.fuse_small_norm and do_apply and fin_args is None and not skip_reduce
                        and y.V <= model.small_norm_max_voxels
                        and lib.tta_norm_small_supported(N, y.V, nl.batch)):
                    sm_bwd = (g0, g0ns, g1, g1ns, y.ptr, y.ns, N, y.C8, nl.C, y.V, rec["mean"].data_ptr(),
                              rec["rstd"].data_ptr(), rec["gptr"], rec["bptr"], relu_flags,
                              rec["sums"].data_ptr(), dg, db, y.dy_ptr(0), y.dy_ptr(1), y.ns,
                              aux.dy_ptr(0) if aux else 0, aux.dy_ptr(1) if aux else 0, aux.ns if aux else 0, bdt,
                              dy_ws)
                rec["small_bwd"] = sm_bwd is not None

                # L2-sized groups: reduce -> apply over one (sample, channel-chunk range) at a time, so that the apply
                # pass re-reads g and y from the 126 MB L2 instead of HBM.  per_sample_norm_bwd (round 1) used whole
                # samples (67 MB at the 64^3 level: the second pass did not hit); norm_bwd_l2_mb bounds the bytes a
                # group's reduction pass reads (g sources + y)
                per_n = None
                nsrc = 1 + (1 if g1 else 0)
                bytes_n = (nsrc + 1) * y.C8 * y.V * 32
                budget = model.norm_bwd_l2_mb * 1e6 if model.norm_bwd_l2_mb > 0 else (96e6 if model.per_sample_norm_bwd else 0)
                if (budget > 0 and do_apply and fin_args is None and not skip_reduce and sm_bwd is None
                        and not nl.batch and N * bytes_n > model.norm_bwd_l2_min_mb * 1e6):
                    parts = 1
                    while bytes_n / parts > budget and parts < y.C8:
                        parts += 1
                    while y.C8 % parts:
                        parts += 1
                    cc = y.C8 // parts
                    if parts * N > 1:
                        per_n = []
                        Cp = y.C8 * 8
                        for n_ in range(N):
                            for c0 in range(0, y.C8, cc):
                                o4, o2, oc = c0 * y.V * 8 * 4, c0 * y.V * 8 * 2, c0 * 8
                                mo = (n_ * Cp + oc) * 4
                                creal = max(0, min(nl.C - oc, cc * 8))
                                gp, bp = rec["gptr"] + oc * 4, rec["bptr"] + oc * 4
                                rd_n = (g0 + n_ * g0ns * 4 + o4, g0ns, (g1 + n_ * g1ns * 4 + o4) if g1 else 0, g1ns,
                                        y.ptr + n_ * y.ns * 4 + o4, y.ns, 1, cc, creal, y.V, rec["mean"].data_ptr() + mo,
                                        rec["rstd"].data_ptr() + mo, gp, bp, int(rec["relu"]), nl.batch,
                                        rec["sums"].data_ptr() + 2 * mo, dg + oc * 4, db + oc * 4)
                                ap_n = (g0 + n_ * g0ns * 4 + o4, g0ns, (g1 + n_ * g1ns * 4 + o4) if g1 else 0, g1ns,
                                        y.ptr + n_ * y.ns * 4 + o4, y.ns, 1, cc, y.V, rec["mean"].data_ptr() + mo,
                                        rec["rstd"].data_ptr() + mo, gp, bp, int(rec["relu"]), nl.batch,
                                        rec["sums"].data_ptr() + 2 * mo, y.dy_ptr(0) + n_ * y.ns * 2 + o2,
                                        (y.dy_ptr(1) + n_ * y.ns * 2 + o2) if nplanes == 2 else y.dy_ptr(1), y.ns,
                                        (aux.dy_ptr(0) + n_ * aux.ns * 2 + o2) if aux else 0,
                                        ((aux.dy_ptr(1) + n_ * aux.ns * 2 + o2) if nplanes == 2 else aux.dy_ptr(1)) if aux else 0,
                                        aux.ns if aux else 0, bdt)
                                per_n.append((rd_n, ap_n, int(n_ > 0), creal, dg + oc * 4, db + oc * 4))
                rec["per_sample_bwd"] = len(per_n) if per_n is not None else 0

                def run(rd_args=rd_args, ap_args=ap_args if do_apply else None, nl=nl, dg=dg, db=db, dy_ws=dy_ws,
                        skip_reduce=skip_reduce, fin_args=fin_args, sm_bwd=sm_bwd, per_n=per_n, c4_args=c4_args):
                    if c4_args is not None:
                        check(lib.tta_norm_bwd_apply_c4(*c4_args, _stream()), "norm_bwd_apply_c4")
                        return
                    if per_n is not None:
                        for rd_n, ap_n, accum, creal, dg_, db_ in per_n:
                            check(lib.tta_norm_bwd_reduce(*rd_n, plan.ws.data_ptr(), 1, accum, _stream()),
                                  "norm_bwd_reduce")
                            check(lib.tta_norm_bwd_apply(*ap_n, 0, creal, dg_, db_, dy_ws, _stream()), "norm_bwd_apply")
                        return
                    if sm_bwd is not None:
                        check(lib.tta_norm_bwd_small(*sm_bwd, plan.ws.data_ptr(), _stream()), "norm_bwd_small")
                        return
                    # single-pass reduction: the last block finalizes sums + dgamma/dbeta
                    if fin_args is not None:
                        check(lib.tta_norm_bwd_finalize(*fin_args, _stream()), "norm_bwd_finalize")
                    elif not skip_reduce:
                        check(lib.tta_norm_bwd_reduce(*rd_args, plan.ws.data_ptr(), 1, 0, _stream()),
                              "norm_bwd_reduce")
                    if ap_args is not None:
                        check(lib.tta_norm_bwd_apply(*ap_args, 0, nl.C, dg, db, dy_ws, _stream()), "norm_bwd_apply")
                run.label = f"norm bwd {nl.name} C={nl.C} V={y.V}"
                plan.bwd.append(run)
        n_conv = sum(1 for o in ops if o[0] == "conv")
        n_norm = sum(1 for o in ops if o[0] == "norm")
        plan.launches_fwd = 1 + n_conv + 2 * n_norm + 2          # gather + convs + (partial stats + apply) + head
        plan.launches_bwd = sum(1 for o in ops if o[0] == "conv" and o[2].parent.needs_grad) + \
            sum(2 if a else 1 for a in bwd_apply_flags) + 1       # dgrads + norm bwd (reduce [+ apply]) + adam
        if fused_head is not None:
            plan.launches_fwd -= 3   # norm apply + small conv + loss finalize folded into the fused head
            plan.launches_bwd -= 1   # small-conv dgrad + norm reduce are one kernel
        # statistics produced by the conv epilogue: no tta_norm_stats launch (the fused head still
        # needs mean/rstd up front: one tiny finalize launch takes its place)
        # (launch count unchanged: a tiny finalize launch takes the place of the statistics pass)
        plan.n_fused_stats = sum(1 for o in ops if o[0] == "norm" and o[1]["fused_stats"])
        plan.n_fused_bwd = sum(1 for o in ops if o[0] == "norm" and o[1].get("fused_bwd"))
        plan.n_small_fwd = sum(1 for o in ops if o[0] == "norm" and o[1].get("small")
                               and not (fused_head is not None and o[1] is fused_head[1]))
        plan.n_small_bwd = sum(1 for o in ops if o[0] == "norm" and o[1].get("small_bwd"))
        plan.launches_fwd -= plan.n_small_fwd      # statistics + apply in one launch
        plan.launches_bwd -= plan.n_small_bwd      # reduction + apply in one launch
        plan.launches_bwd += sum(2 * (o[1]["per_sample_bwd"] - 1) for o in ops if o[0] == "norm" and o[1].get("per_sample_bwd"))
        n_extra = sum(1 for o in ops if o[0] in ("mean", "upsample", "cast"))   # one launch each way
        plan.launches_fwd += n_extra + (1 if getattr(plan, "x2", None) is not None else 0)
        plan.launches_bwd += n_extra + getattr(plan, "n_wgrad", 0)
        return plan

    def _nl(self, holder: NormHolder) -> NormLayer:
        for nl in self.norm_layers:
            if nl.h is holder:
                return nl
        raise KeyError("norm holder not registered")

    @staticmethod
    def _producer_conv(ops, y) -> "ConvLayer":
        for op in ops:
            if op[0] == "conv" and op[3] is y.root:
                return op[1]
        raise KeyError("no producer conv for result tensor")

    @staticmethod
    def _producer_input_needs_grad(ops, y: Res) -> bool:
        for op in ops:
            if op[0] == "conv" and op[3] is y.root:
                return op[2].parent.needs_grad
        raise KeyError("no producer conv for result tensor")

    # ---------------------------------------------------------------- execution
    def get_plan(self, N, D, H, W) -> Plan:
        key = (N, D, H, W)
        if key not in self.plans:
            self.plans[key] = self.build_plan(N, D, H, W)
        plan = self.plans[key]
        bad = getattr(plan, "single_element_norm", None)
        if bad is not None and self.model.training:
            # same refusal as torch's instance/batch norm in training mode
            raise ValueError(f"Expected more than 1 spatial element when training, got a single element at {bad}")
        return plan

    def _load_running_stats(self, plan: Plan):
        """eval()-mode BatchNorm: statistics come from the running buffers, not the batch."""
        for nl, mean, rstd, y in plan.stats_ops:
            if nl.batch and not self.model.training and nl.h.track_running_stats:
                C8 = nl.C8 * 8
                mu = torch.zeros(C8, device=self.device); mu[:nl.C] = nl.h.running_mean.to(self.device)
                rs = torch.zeros(C8, device=self.device)
                rs[:nl.C] = torch.rsqrt(nl.h.running_var.to(self.device) + nl.h.eps)
                mean.copy_(mu.repeat(plan.N)); rstd.copy_(rs.repeat(plan.N))

    def _update_running_stats(self, plan: Plan):
        """train()-mode BatchNorm outside TENT keeps nn.BatchNorm3d's side effect: running stats
        move by ``momentum`` towards the batch statistics (unbiased variance)."""
        if not self.model.training:
            return
        for nl, mean, rstd, y in plan.stats_ops:
            h = nl.h
            if nl.batch and h.track_running_stats:
                M = plan.N * y.V
                mu = mean[:nl.C]
                var = (1.0 / (rstd[:nl.C] * rstd[:nl.C]) - h.eps) * (M / max(M - 1, 1))
                h.running_mean.mul_(1 - h.momentum).add_(mu.to(h.running_mean.device), alpha=h.momentum)
                h.running_var.mul_(1 - h.momentum).add_(var.to(h.running_var.device), alpha=h.momentum)
                h.num_batches_tracked += 1

    def _pack_input(self, plan: Plan, x: torch.Tensor, win: Optional[torch.Tensor] = None,
                    chan_scale: Optional[torch.Tensor] = None, vol_dims=None, n_vol=None,
                    affine: Optional[torch.Tensor] = None):
        N = plan.N
        D, H, W = plan.dims
        if win is None:
            if plan.win is None:
                w = torch.zeros((N, 4), dtype=torch.int32)
                w[:, 0] = torch.arange(N, dtype=torch.int32)
                plan.win = w.to(self.device)
            win = plan.win
            vol_dims, n_vol = (D, H, W), N
        a = plan.x
        gather = self.lib.tta_gather_pack_norm_f16 if x.dtype == torch.float16 else self.lib.tta_gather_pack_norm
        # affine [n_vol][C][4] (IntensityPolicy.stats): clip + z-score applied while gathering
        check(gather(x.data_ptr(), n_vol, self.model.in_channels, *vol_dims, win.data_ptr(),
                                            chan_scale.data_ptr() if chan_scale is not None else 0,
                                            affine.data_ptr() if affine is not None else 0, N, D, H, W,
                                            a.planes[0].data_ptr(), a.planes[1].data_ptr(), a.V * 4 if a.compact else a.ns,
                                            a.C8, 2 if a.compact else int(a.wsplit), _stream()),
              "gather_pack")
        x2 = getattr(plan, "x2", None)
        if x2 is not None:
            # a second consumer wants the other layout (multimodal model: the last decoder stage concatenates the
            # input in plain layout while the encoder stems read it w-parity-split)
            check(gather(x.data_ptr(), n_vol, self.model.in_channels, *vol_dims, win.data_ptr(),
                         chan_scale.data_ptr() if chan_scale is not None else 0,
                         affine.data_ptr() if affine is not None else 0, N, D, H, W,
                         x2.hi, x2.lo, x2.ns, x2.C8, 0, _stream()),
                  "gather_pack")

    def _check_input(self, x: torch.Tensor):
        if x.dim() != 5:
            raise ValueError(f"expected [B,C,D,H,W], got {tuple(x.shape)}")
        if x.shape[1] != self.model.in_channels:
            raise ValueError(f"expected {self.model.in_channels} input channels, got {x.shape[1]}")
        if not x.is_cuda:
            raise RuntimeError("multimodal_tta_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        # fp32 (the reference's batch dtype) or fp16 staging (half the host -> device bytes; the gather reads it
        # directly); anything else is converted to fp32 on the device
        if x.dtype not in (torch.float32, torch.float16):
            x = x.float()
        if not x.is_contiguous():
            x = x.contiguous()
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = self._check_input(x)
        self._ensure_device(x.device)
        with self.on_device():
            plan = self.get_plan(*[int(s) for s in (x.shape[0], *x.shape[2:])])
            self._load_running_stats(plan)
            self._pack_input(plan, x)
            if self.model.c_plan:
                cp = self.c_plan(plan)
                check(self.lib.tta_plan_run(cp.h, CPlan.SEC_FWD, _stream()), "plan_run(forward)")
                check(self.lib.tta_plan_run(cp.h, CPlan.SEC_HEAD_INFER, _stream()), "plan_run(head)")
            else:
                for op in plan.fwd:
                    op()
                plan.head_infer()
            self._update_running_stats(plan)
            return plan.logits.clone()

    def c_plan(self, plan: Plan) -> CPlan:
        """The plan's launch list as a C object; the closures branch on train / eval mode (BatchNorm running
        statistics), so one recording per mode."""
        key = bool(self.model.training)
        if key not in plan.c_plans:
            plan.c_plans[key] = CPlan(self.lib, plan)
        return plan.c_plans[key]

    def run_step(self, plan: Plan, adam: bool = True, gscale: float = 1.0):
        """forward + fused head + backward (+ Adam) on whatever is in plan.x."""
        self.last_inv_scale = 1.0 / plan.loss_scale
        if self.model.c_plan:
            check(self.lib.tta_step(self.c_plan(plan).h, _stream()), "tta_step")
        else:
            for op in plan.fwd:
                op()
            plan.head_train()
            for op in plan.bwd:
                op()
        if adam:
            self.adam_step(gscale)

    def forward_train(self, x: torch.Tensor):
        """Forward of the supervised path (model.trainable): train-mode norm statistics, logits only.  Returns
        (logits copy, plan); the plan's operand / result buffers are what backward_from_logits_grad reads."""
        x = self._check_input(x)
        self._ensure_device(x.device)
        with self.on_device():
            plan = self.get_plan(*[int(s_) for s_ in (x.shape[0], *x.shape[2:])])
            if plan.fused_head:
                raise RuntimeError("supervised step: the fused head holds host-side weights (internal error)")
            self._pack_input(plan, x)
            if self.model.c_plan:
                cp = self.c_plan(plan)
                check(self.lib.tta_plan_run(cp.h, CPlan.SEC_FWD, _stream()), "plan_run(forward)")
                check(self.lib.tta_plan_run(cp.h, CPlan.SEC_HEAD_INFER, _stream()), "plan_run(head)")
            else:
                for op in plan.fwd:
                    op()
                plan.head_infer()
            self._update_running_stats(plan)
            return plan.logits.clone(), plan

    def backward_from_logits_grad(self, plan: Plan, grad_logits: torch.Tensor):
        """dL/dlogits [N,R,D,H,W] fp32 (from ANY loss: the reference's DiceCELoss through autograd) -> gradients of every
        conv weight / bias (gflat) and norm affine (dgb, loss-scaled).  The gradient enters the 16-bit gradient planes
        of the last conv through tta_pack_grad (NCDHW -> chunk layout), times the plan's power-of-two loss scale."""
        if not self.model.trainable:
            raise RuntimeError("backward needs model.trainable = true")
        g = grad_logits.contiguous().float()
        R = self.model.out_channels
        N, (D, H, W) = plan.N, plan.dims
        final = plan.final
        with self.on_device():
            self.last_inv_scale = 1.0 / plan.loss_scale
            self.gflat.zero_()
            self.dgb.zero_()
            check(self.lib.tta_pack_grad(g.data_ptr(), N, R, final.V, float(plan.loss_scale), final.dy_ptr(0),
                                         final.dy_ptr(1), final.ns, self.bwd_dtype, _stream()), "pack_grad")
            if self.model.c_plan:
                check(self.lib.tta_plan_run(self.c_plan(plan).h, CPlan.SEC_BWD, _stream()), "plan_run(backward)")
            else:
                for op in plan.bwd:
                    op()

    def adam_step(self, gscale: float = 1.0):
        """``gscale`` multiplies the gradient (1/world for the all-reduced sum); the loss scale of
        the last backward is divided out here as well."""
        a = self.adam
        gscale = gscale * self.last_inv_scale
        check(self.lib.tta_adam_step(self.gb.data_ptr(), self.dgb.data_ptr(), self.m.data_ptr(),
                                     self.v.data_ptr(), 2 * self.P, a["lr"], a["b1"], a["b2"], a["eps"],
                                     float(gscale), self.step_dev.data_ptr(), _stream()), "adam")
