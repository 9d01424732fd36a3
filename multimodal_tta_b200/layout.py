"""HBM layouts of the B200 TTA path and the host-side weight packers.

* activations / results : channel-blocked ``[N][C8][D][H][W][8]`` (pad channels are zero)
* conv operands         : two 16-bit planes hi/lo (fp16 forward, bf16 backward), x ~= hi + lo
* canonical weights     : ``Wg[tap][ci][co]`` with tap = (kd*K + kh)*K + kw, see
                          csrc/tta_conv_simt.cu for the gather semantics of mode 0/1.
The torch ops here run once per weight load (plumbing), never inside the timed step.
"""
from __future__ import annotations

import torch

from ._lib import TTA_BF16, TTA_F16, TTA_F16_HI


# ---- index mode: the packers below are pure rearrangements (+ the hi/lo split) of the canonical weights.  Run on a
# tensor of SOURCE INDICES instead of values they yield, for every packed element, which parameter element it comes
# from -- the map the device-side repack kernel (tta_repack_weights) needs after every optimizer step of the
# supervised path, where re-running the host packers (1.3 s) is not an option.  In index mode values are float64
# (exact integers up to 2^53), packed outputs int64, and split_planes returns (+idx, -idx): the sign tells the
# repack kernel which plane (hi / lo) an element is.
_IDX = False


class index_mode:
    def __enter__(self):
        global _IDX
        self.prev, _IDX = _IDX, True

    def __exit__(self, *a):
        global _IDX
        _IDX = self.prev


def _i16():
    return torch.int64 if _IDX else torch.int16


def _f32():
    return torch.float64 if _IDX else torch.float32


def wg_forward(w: torch.Tensor, transposed: bool) -> torch.Tensor:
    """nn.Conv3d weight [co][ci][k,k,k] or nn.ConvTranspose3d weight [ci][co][k,k,k] -> Wg[T][ci][co]."""
    k = w.shape[-1]
    if transposed:
        return w.permute(2, 3, 4, 0, 1).reshape(k ** 3, w.shape[0], w.shape[1]).contiguous()
    return w.permute(2, 3, 4, 1, 0).reshape(k ** 3, w.shape[1], w.shape[0]).contiguous()


def wg_dgrad(w: torch.Tensor, transposed: bool) -> torch.Tensor:
    """Weights of the input-gradient conv.  dgrad(Conv3d) is a mode-1 (transposed) conv with
    Wg[k][ci=cout][co=cin] = w[cout][cin][k]; dgrad(ConvTranspose3d) is a mode-0 conv with
    Wg[k][ci=cout_f][co=cin_f] = w[cin_f][cout_f][k]."""
    k = w.shape[-1]
    if transposed:
        return w.permute(2, 3, 4, 1, 0).reshape(k ** 3, w.shape[1], w.shape[0]).contiguous()
    return w.permute(2, 3, 4, 0, 1).reshape(k ** 3, w.shape[0], w.shape[1]).contiguous()


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def pack_weights_simt(wg: torch.Tensor) -> torch.Tensor:
    """Wg[T][ci][co] -> Wp[T][C8in][C8out][8 ci][8 co] fp32 (zero padded)."""
    T, ci, co = wg.shape
    cip, cop = _pad8(ci), _pad8(co)
    p = torch.zeros((T, cip, cop), dtype=_f32(), device=wg.device)
    p[:, :ci, :co] = wg
    p = p.reshape(T, cip // 8, 8, cop // 8, 8).permute(0, 1, 3, 2, 4).contiguous()
    return p.round().long() if _IDX else p


def pack_weights_small(wg: torch.Tensor, mode: int) -> torch.Tensor:
    """Wg[27][ci][co] (ci, co <= 4) -> HOST fp32 [27][8][8] for tta_conv_small.  mode 1 (the transposed /
    input-gradient form, out[o] = sum_k in[o + 1 - k] Wg[k]) is turned into a plain correlation by
    flipping the tap order."""
    T, ci, co = wg.shape
    assert T == 27 and ci <= 8 and co <= 8
    p = torch.zeros((27, 8, 8), dtype=torch.float32)
    p[:, :ci, :co] = (wg.flip(0) if mode == 1 else wg).detach().cpu()
    return p.contiguous()   # HOST tensor: tta_conv_small passes the weights as kernel parameters


def pack_bias(b: torch.Tensor) -> torch.Tensor:
    p = torch.zeros(_pad8(b.numel()), dtype=_f32(), device=b.device)
    p[: b.numel()] = b
    if _IDX:
        return p.round().long()
    return p


def split_planes(x: torch.Tensor, dtype_tag: int) -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (hi, lo) 16-bit planes stored as int16, hi = rn16(x), lo = rn16(x - hi)."""
    if _IDX:
        h = x.round().long()
        return h, -h
    dt = torch.bfloat16 if dtype_tag == TTA_BF16 else torch.float16
    if dtype_tag != TTA_BF16:
        x = x.clamp(-65504.0, 65504.0)
    hi = x.to(dt)
    lo = (x - hi.float()).to(dt)
    return hi.view(torch.int16), lo.view(torch.int16)


def join_planes(hi: torch.Tensor, lo: torch.Tensor, dtype_tag: int) -> torch.Tensor:
    dt = torch.bfloat16 if dtype_tag == TTA_BF16 else torch.float16
    if dtype_tag == TTA_F16_HI:
        return hi.view(dt).float()
    return hi.view(dt).float() + lo.view(dt).float()


def to_chunked(x: torch.Tensor) -> torch.Tensor:
    """NCDHW fp32 -> [N][C8][D][H][W][8] fp32 (zero padded channels)."""
    N, C, D, H, W = x.shape
    Cp = _pad8(C)
    xp = torch.zeros((N, Cp, D, H, W), dtype=x.dtype, device=x.device)
    xp[:, :C] = x
    return xp.reshape(N, Cp // 8, 8, D, H, W).permute(0, 1, 3, 4, 5, 2).contiguous()


def from_chunked(x: torch.Tensor, C: int) -> torch.Tensor:
    """[N][C8][D][H][W][8] -> NCDHW (first C channels)."""
    N, C8, D, H, W, _ = x.shape
    return x.permute(0, 1, 5, 2, 3, 4).reshape(N, C8 * 8, D, H, W)[:, :C].contiguous()


# ---------------------------------------------------------------------------- tcgen05 packing
# Geometry tables shared with csrc/tta_conv_tc.cu (tta_conv_tc_plan): for every pipeline
# "group" g the kernel consumes, in order, the taps listed here.
def tc_groups(mode: int, K: int, stride: int) -> list[list[int]]:
    """Tap indices (kd*K*K + kh*K + kw) per pipeline group, in the kernel's MMA issue order."""
    if K == 1:
        return [[0]]
    taps = lambda kds: [kd * 9 + kh * 3 + kw for kd in kds for kh in range(3) for kw in range(3)]
    if mode == 0:                      # conv s1 / s2: one group per kd
        return [taps([0]), taps([1]), taps([2])]
    if stride == 1:                    # transposed s1: same structure, offsets mirrored in-kernel
        return [taps([0]), taps([1]), taps([2])]
    # transposed s2: group jd=0 holds kd in {1 (even out planes), 2 (odd)}, jd=1 holds kd=0
    return [taps([1, 2]), taps([0])]


def s2_pairs() -> list[tuple[int, int | None]]:
    """(kh*3 + kw) tap pairs of a stride-2 conv over a one-chunk input, in the kernel's entry order:
    both taps of a pair read the same (h, w)-parity sub-tile, one voxel or one row apart."""
    t = lambda kh, kw: kh * 3 + kw
    return [(t(0, 0), t(0, 2)), (t(2, 0), t(2, 2)), (t(0, 1), t(2, 1)), (t(1, 0), t(1, 2)), (t(1, 1), None)]


def t2_stacks() -> list[list[list[tuple[int, int]]]]:
    """Transposed stride-2 conv: per pipeline group, the weight STACKS the kernel multiplies with one
    shifted A tile (csrc/tta_conv_tc.cu, GEOM_T2 tables kG0 / kG1).  A stack is a list of
    (accumulator, tap) in accumulator order; accumulator a = qd*4 + qh*2 + qw is the output parity
    class, tap = kd*9 + kh*3 + kw.  Parity 0 along an axis uses k = 1 (shift 0), parity 1 uses k = 2
    (shift 0) and k = 0 (shift 1 = next input voxel)."""
    def tap(qd, qh, qw, jd, jh, jw):
        kd = 0 if jd else (2 if qd else 1)
        kh = 0 if jh else (2 if qh else 1)
        kw = 0 if jw else (2 if qw else 1)
        return kd * 9 + kh * 3 + kw

    def stacks(jd, tables):
        out = []
        for acc0, k, jh, jw in tables:
            out.append([(a, tap(a >> 2, (a >> 1) & 1, a & 1, jd, jh, jw)) for a in range(acc0, acc0 + k)])
        return out
    g0 = [(0, 8, 0, 0), (2, 2, 1, 0), (6, 2, 1, 0), (1, 1, 0, 1), (3, 1, 0, 1), (5, 1, 0, 1), (7, 1, 0, 1),
          (3, 1, 1, 1), (7, 1, 1, 1)]
    g1 = [(4, 4, 0, 0), (6, 2, 1, 0), (5, 1, 0, 1), (7, 1, 0, 1), (7, 1, 1, 1)]
    return [stacks(0, g0), stacks(1, g1)]


def pack_weights_t2s(wg: torch.Tensor) -> torch.Tensor:
    """Wg[27][ci][co] (transposed stride-2 conv, co <= 4, ci % 16 == 0) -> the resident B operand of
    csrc/tta_conv_tc.cu:conv_t2s_kernel: [k-step ci/16][kchunk 2][hi NP rows | lo NP rows][8 ci] fp16,
    row = tap * co + c (NP = 27*co rounded up to 16; pad rows zero)."""
    T, ci, co = wg.shape
    assert T == 27 and ci % 16 == 0 and 1 <= co <= 4
    npad = (27 * co + 15) // 16 * 16
    hi, lo = split_planes(wg.to(_f32()), TTA_F16)                    # [27][ci][co]
    out = torch.zeros((ci // 16, 2, 2, npad, 8), dtype=_i16(), device=wg.device)
    for pi, plane in enumerate((hi, lo)):
        rows = plane.permute(0, 2, 1).reshape(27 * co, ci // 16, 2, 8)   # [tap*co + c][ks][kc][8]
        out[:, :, pi, : 27 * co] = rows.permute(1, 2, 0, 3)
    return out.contiguous()


def pack_weights_tc(wg: torch.Tensor, mode: int, K: int, stride: int, dtype_tag: int, t2s: bool = False,
                    s2c4: bool = False) -> torch.Tensor:
    """Wg[T][ci][co] -> per-(n_tile, cblk, group) contiguous blobs
    [ntile][cblk][group][entry][kchunk 2][hi NT rows | lo NT rows][8 ci] of 16-bit values.
    A blob is what one pipeline stage of the tcgen05 kernel bulk-copies into shared memory as
    its B operand (K-major, no swizzle: 8 rows x 16 B core matrices, k-chunk pitch = 2*NT*16 B).
    Stacking the hi and lo rows lets ONE MMA with N = 2*NT compute A_hi*B_hi and A_hi*B_lo."""
    from . import _lib
    T, ci, co = wg.shape
    lib = _lib.lib()
    if t2s:   # dense-GEMM + col2im kernel for a small-Cout transposed conv (caller passes Cout in the flags)
        return pack_weights_t2s(wg)
    split = dtype_tag != TTA_F16_HI      # TTA_F16_HI: single fp16 plane, B rows are not stacked
    ntile = lib.tta_conv_tc_ntile(mode, K, stride, co, int(split))
    cip = (ci + 15) // 16 * 16
    cop = (co + ntile - 1) // ntile * ntile
    w = torch.zeros((T, cip, cop), dtype=_f32(), device=wg.device)
    w[:, :ci, :co] = wg
    hi, lo = split_planes(w, dtype_tag)
    planes = (hi, lo) if split else (hi,)
    groups = tc_groups(mode, K, stride)
    gmax = max(len(g) for g in groups)
    ncb, nnt = cip // 16, cop // ntile
    out = torch.zeros((nnt, ncb, len(groups), gmax, 2, len(planes), ntile, 8), dtype=_i16(), device=wg.device)
    if K == 3 and stride == 1 and lib.tta_conv_tc_stacked(mode, K, stride, ci, co, int(split)):
        # kd-stacked stride-1 conv (small n-tile, resident weights; csrc/tta_conv_tc.cu GEOM_S1K): one
        # blob per channel block, entry (kh, kw) = [kchunk][slot 0..2][hi NT | lo NT][8].  Slot s feeds
        # the accumulator of output plane j - 2 + s from input plane j: kd = 2 - s for a conv, kd = s
        # for its input-gradient dual (mirrored offsets).
        st = torch.zeros((1, ncb, 9, 2, 3, len(planes), ntile, 8), dtype=_i16(), device=wg.device)
        for sl in range(3):
            kd = 2 - sl if mode == 0 else sl
            for pi, plane in enumerate(planes):
                sel = plane[kd * 9: kd * 9 + 9].reshape(9, ncb, 2, 8, nnt, ntile)   # [E][cb][kc][8][nt][n]
                st[:, :, :, :, sl, pi] = sel.permute(4, 1, 0, 2, 5, 3)               # [nt][cb][E][kc][n][8]
        return st.contiguous()
    if K == 3 and mode == 1 and stride == 2:
        # transposed stride-2: per group a sequence of stacks, each [kchunk][accumulators of the stack]
        # [hi NT | lo NT][8]: ONE MMA covers all accumulators of a stack (N = k * planes * NT)
        flat = out.view(nnt, ncb, len(groups), -1)
        pl = torch.stack(planes)                                       # [P][T][cip][cop]
        for gi, stacks in enumerate(t2_stacks()):
            pos = 0
            for st in stacks:
                idx = torch.tensor([t for _, t in st], device=wg.device)
                sel = pl[:, idx]                                       # [P][k][cip][cop]
                sel = sel.reshape(len(planes), len(st), ncb, 2, 8, nnt, ntile)   # [P][k][cb][kc][8][nt][n]
                blk = sel.permute(5, 2, 3, 1, 0, 6, 4).reshape(nnt, ncb, -1)     # [nt][cb][kc][k][P][n][8]
                flat[:, :, gi, pos: pos + blk.shape[-1]] = blk
                pos += blk.shape[-1]
        return out.contiguous()
    if s2c4:
        # stride-2 conv over a COMPACT <= 4-channel input (csrc/tta_conv_tc.cu GEOM_S2C4): group = kd, entry = kh,
        # k-chunk 0 = [zeros | kw 0: 4 ch], k-chunk 1 = [kw 1: 4 ch | kw 2: 4 ch] (rows start at the even voxel 2w - 2)
        assert K == 3 and mode == 0 and stride == 2 and ci <= 4
        out = torch.zeros((nnt, 1, 3, 3, 2, len(planes), ntile, 8), dtype=_i16(), device=wg.device)
        for kd in range(3):
            for kh in range(3):
                for pi, plane in enumerate(planes):
                    for kw in range(3):
                        sel = plane[kd * 9 + kh * 3 + kw][:4].reshape(4, nnt, ntile)      # [ci 4][nt][n]
                        j = kw + 1
                        out[:, 0, kd, kh, j // 2, pi, :, (j % 2) * 4:(j % 2) * 4 + 4] = sel.permute(1, 2, 0)
        return out.contiguous()
    if K == 3 and lib.tta_conv_tc_s2pair(mode, K, stride, ci):
        # stride-2 conv over a one-chunk input: entry = a PAIR of taps of one parity class, tap a in
        # k-chunk 0 and tap b in k-chunk 1 (csrc/tta_conv_tc.cu, GEOM_S2 kPair); (1,1) stays single
        for gi in range(3):
            for e, (ta, tb) in enumerate(s2_pairs()):
                for pi, plane in enumerate(planes):
                    for kc, t in enumerate((ta, tb)):
                        if t is None:
                            continue
                        sel = plane[gi * 9 + t][:8].reshape(8, nnt, ntile)            # [ci 8][nt][n]
                        out[:, 0, gi, e, kc, pi] = sel.permute(1, 2, 0)
        return out.contiguous()
    for gi, taps in enumerate(groups):
        idx = torch.tensor(taps, device=wg.device)
        for pi, plane in enumerate(planes):
            sel = plane[idx]                                       # [E][cip][cop]
            sel = sel.reshape(len(taps), ncb, 2, 8, nnt, ntile)    # [E][cb][kc][8][nt][n]
            out[:, :, gi, : len(taps), :, pi] = sel.permute(4, 1, 0, 2, 5, 3)   # [nt][cb][E][kc][n][8]
    return out.contiguous()
