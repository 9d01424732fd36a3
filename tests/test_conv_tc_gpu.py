"""tcgen05 implicit-GEMM conv (tta_conv_tc) against torch fp32 conv on the CPU, every geometry
family the UNet uses (s1, s2, transposed s2, 1x1, and their input-gradient duals), ragged tile
edges (dims not multiples of the 16x8 tile), channel counts that need zero-padded k-chunks and
n-tiles, concat-slice views, and the accumulate epilogue.

Tolerance: operands are split 16-bit planes and the kernel drops the lo*lo term, so products
carry ~2^-22 (fp16 planes) / ~2^-16 (bf16 planes) relative error against a reference computed
from the SAME split-rounded inputs: rel-L2 <= 2e-5 (fp16 planes; measured 2e-6 .. 7e-6, growing
with K = 27*Cin because the tensor core's fp32 accumulator truncates) / 1e-4 (bf16 planes)."""
import pytest
import torch
import torch.nn.functional as F

from multimodal_tta_b200._lib import TTA_BF16, TTA_F16, TTA_F16_HI, check
from multimodal_tta_b200.layout import (from_chunked, join_planes, pack_bias, pack_weights_tc, split_planes,
                                        to_chunked, wg_dgrad, wg_forward)
from tests.util import planes_from, rel_l2, stream, wsplit

pytestmark = pytest.mark.gpu

GEOMS = [  # cin, cout, K, stride, transposed, dims
    (32, 32, 3, 1, False, (4, 16, 8)),      # exact tiles
    (32, 32, 3, 1, False, (5, 20, 12)),     # ragged h/w/d
    (4, 32, 3, 2, False, (8, 32, 16)),      # stem: cin 4 -> one chunk + OOB k-chunk
    (32, 64, 3, 2, False, (8, 16, 16)),
    (64, 128, 3, 2, False, (4, 12, 12)),
    (128, 256, 3, 1, False, (2, 6, 6)),     # n-tiles = 2, tiny spatial (two-plane tiles, ragged h/w)
    (256, 512, 3, 1, False, (8, 8, 8)),     # bottleneck: two d-planes per 128-row tile, split-K
    (64, 64, 3, 1, False, (5, 8, 16)),      # two-plane tiles, odd plane count, two w-tiles
    (48, 16, 3, 2, True, (3, 10, 6)),       # transposed s2, ragged
    (64, 3, 3, 2, True, (4, 16, 8)),        # head convT: cout 3 -> N=16 tile
    (256, 64, 3, 2, True, (4, 8, 8)),       # transposed s2 over small planes: two input planes per tile
    (40, 24, 3, 2, True, (3, 6, 5)),        # ... ragged, odd plane count
    (3, 3, 3, 1, False, (8, 16, 16)),       # head conv 3->3
    (256, 512, 1, 1, False, (4, 8, 8)),     # 1x1 shortcut
    (96, 40, 3, 1, False, (3, 9, 9)),       # odd chunk counts
]


def _tc(lib, hi, lo, ns, dt, N, c8i, idims, wp, bias, out, ons, c8o, odims, mode, K, s, acc=0, flags=0):
    check(lib.tta_conv_tc(hi.data_ptr(), lo.data_ptr(), ns, dt, N, c8i, *idims, wp.data_ptr(),
                          bias.data_ptr() if bias is not None else 0, out.data_ptr(), ons, c8o, *odims, mode, K, s,
                          acc, flags, 0, 0, stream()), "conv_tc")
    torch.cuda.synchronize()


@pytest.mark.parametrize("cin,cout,K,s,tr,dims", GEOMS)
def test_conv_tc_forward_and_dgrad(lib, cuda, cin, cout, K, s, tr, dims):
    torch.manual_seed(7)
    N = 2
    x = torch.randn(N, cin, *dims)
    w = torch.randn((cin, cout, K, K, K) if tr else (cout, cin, K, K, K)) * 0.1
    b = torch.randn(cout)
    pad = (K - 1) // 2
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    # reference on the rounded operands: weights also go through the fp16 hi/lo split
    whi, wlo = split_planes(w, TTA_F16)
    wv = join_planes(whi, wlo, TTA_F16)
    xr = xv.cpu().requires_grad_(True)
    ref = F.conv_transpose3d(xr, wv, b, stride=s, padding=pad, output_padding=s - 1) if tr else \
        F.conv3d(xr, wv, b, stride=s, padding=pad)
    odims = tuple(ref.shape[2:])
    c8i, c8o = (cin + 7) // 8, (cout + 7) // 8
    mode = 1 if tr else 0
    wp = pack_weights_tc(wg_forward(w.to(cuda), tr), mode, K, s, TTA_F16)
    out = torch.zeros((N, c8o, *odims, 8), device=cuda)
    _tc(lib, hi, lo, c8i * xv[0, 0].numel() * 8, TTA_F16, N, c8i, dims, wp, pack_bias(b.to(cuda)), out,
        c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s)
    got = from_chunked(out, cout).cpu()
    assert rel_l2(got, ref.detach()) < 2e-5, rel_l2(got, ref.detach())
    if K == 3:
        # resident weights (small-channel layers keep every weight blob in smem for the CTA's lifetime)
        # vs. per-stage weight copies (flags bit4): same MMAs in the same order -> identical result
        o_res = torch.zeros_like(out); o_str = torch.zeros_like(out)
        args = (c8i * xv[0, 0].numel() * 8, TTA_F16, N, c8i, dims, wp, pack_bias(b.to(cuda)))
        _tc(lib, hi, lo, *args, o_res, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2)
        _tc(lib, hi, lo, *args, o_str, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2 | 16)
        assert torch.equal(o_res, o_str)
    if K == 3 and dims[1] <= 8 and (s == 1 or tr):
        # small planes: two d-planes per 128-row tile (9 (kd, kh) groups) vs. one plane per tile (flags
        # bit5): the same products enter every accumulator in the same order -> identical result
        o_p2 = torch.zeros_like(out); o_p1 = torch.zeros_like(out)
        args = (c8i * xv[0, 0].numel() * 8, TTA_F16, N, c8i, dims, wp, pack_bias(b.to(cuda)))
        _tc(lib, hi, lo, *args, o_p2, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2)
        _tc(lib, hi, lo, *args, o_p1, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2 | 32)
        assert torch.equal(o_p2, o_p1)
    if K == 3:
        # CTA pairs (flags bit12: 2-CTA clusters, each CTA fetches half of every streamed weight blob, multicast
        # into both): same MMAs on the same bits -> identical result (no resident weights: bit4)
        o_one = torch.zeros_like(out); o_pair = torch.zeros_like(out)
        args = (c8i * xv[0, 0].numel() * 8, TTA_F16, N, c8i, dims, wp, pack_bias(b.to(cuda)))
        _tc(lib, hi, lo, *args, o_one, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2 | 16)
        _tc(lib, hi, lo, *args, o_pair, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2 | 16 | 4096)
        assert torch.equal(o_one, o_pair)
    if not tr and s == 2:
        # w-parity-split operand layout (flags bit3): same MMAs on the same bits -> identical result
        o_std = torch.zeros_like(out); o_ws = torch.zeros_like(out)
        args = (c8i * xv[0, 0].numel() * 8, TTA_F16, N, c8i, dims, wp, pack_bias(b.to(cuda)))
        _tc(lib, hi, lo, *args, o_std, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2)
        _tc(lib, wsplit(hi), wsplit(lo), *args, o_ws, c8o * ref[0, 0].numel() * 8, c8o, odims, mode, K, s, flags=2 | 8)
        assert torch.equal(o_ws, o_std)
    # pad channels of the last chunk stay exactly zero (+ bias pad 0)
    if cout % 8:
        assert float(out[:, -1, ..., cout % 8:].abs().max()) == 0.0
    # ---- dgrad in bf16 planes, written into a channel SLICE of a wider (concat) gradient buffer
    dy = torch.randn_like(ref)
    dhi, dlo, dyv = planes_from(dy.to(cuda), TTA_BF16)
    bhi, blo = split_planes(w, TTA_BF16)
    wvb = join_planes(bhi, blo, TTA_BF16)
    xr2 = xv.cpu().requires_grad_(True)
    ref2 = F.conv_transpose3d(xr2, wvb, None, stride=s, padding=pad, output_padding=s - 1) if tr else \
        F.conv3d(xr2, wvb, None, stride=s, padding=pad)
    (g2,) = torch.autograd.grad(ref2, xr2, dyv.cpu())
    wpd = pack_weights_tc(wg_dgrad(w.to(cuda), tr), 1 - mode, K, s, TTA_BF16)
    extra = 2
    gbuf = torch.full((N, c8i + extra, *dims, 8), 7.0, device=cuda)
    Vi = dims[0] * dims[1] * dims[2]
    gview_ptr_off = extra * Vi * 8
    gflat = gbuf.view(-1)
    sub = gflat[gview_ptr_off:]
    check(lib.tta_conv_tc(dhi.data_ptr(), dlo.data_ptr(), c8o * ref[0, 0].numel() * 8, TTA_BF16, N, c8o, *odims,
                          wpd.data_ptr(), 0, sub.data_ptr(), (c8i + extra) * Vi * 8, c8i, *dims, 1 - mode, K, s, 0, 0,
                          0, 0, stream()), "conv_tc dgrad")
    torch.cuda.synchronize()
    gx = from_chunked(gbuf[:, extra:], cin).cpu()
    assert rel_l2(gx, g2) < 1e-4, rel_l2(gx, g2)
    assert float((gbuf[:, :extra] - 7.0).abs().max()) == 0.0       # neighbouring slice untouched
    # accumulate epilogue
    check(lib.tta_conv_tc(dhi.data_ptr(), dlo.data_ptr(), c8o * ref[0, 0].numel() * 8, TTA_BF16, N, c8o, *odims,
                          wpd.data_ptr(), 0, sub.data_ptr(), (c8i + extra) * Vi * 8, c8i, *dims, 1 - mode, K, s, 1, 0,
                          0, 0, stream()), "conv_tc dgrad acc")
    torch.cuda.synchronize()
    assert rel_l2(from_chunked(gbuf[:, extra:], cin).cpu(), 2 * g2) < 1e-4
    # ---- dgrad with ONE fp16 plane (loss-scaled gradients): a single MMA per k-step
    hhi, _ = split_planes(to_chunked(dy.to(cuda)), TTA_F16_HI)
    dy16 = from_chunked(hhi.view(torch.float16).float(), cout).cpu()
    w16 = w.half().float()
    xr3 = xv.cpu().requires_grad_(True)
    ref3 = F.conv_transpose3d(xr3, w16, None, stride=s, padding=pad, output_padding=s - 1) if tr else \
        F.conv3d(xr3, w16, None, stride=s, padding=pad)
    (g3,) = torch.autograd.grad(ref3, xr3, dy16)
    wph = pack_weights_tc(wg_dgrad(w.to(cuda), tr), 1 - mode, K, s, TTA_F16_HI)
    gx3 = torch.zeros((N, c8i, *dims, 8), device=cuda)
    hhi = hhi.contiguous()
    check(lib.tta_conv_tc(hhi.data_ptr(), 0, c8o * ref[0, 0].numel() * 8, TTA_F16_HI, N, c8o, *odims,
                          wph.data_ptr(), 0, gx3.data_ptr(), c8i * Vi * 8, c8i, *dims, 1 - mode, K, s, 0, 0,
                          0, 0, stream()), "conv_tc dgrad fp16")
    torch.cuda.synchronize()
    assert rel_l2(from_chunked(gx3, cin).cpu(), g3) < 2e-5, rel_l2(from_chunked(gx3, cin).cpu(), g3)
    if tr and s == 2:
        # the dgrad of a transposed conv is a stride-2 conv: w-parity-split dY (flags bit3)
        g_std = torch.zeros_like(gx3); g_ws = torch.zeros_like(gx3)
        hws = wsplit(hhi)
        for buf, src, fl in ((g_std, hhi, 2), (g_ws, hws, 2 | 8)):
            check(lib.tta_conv_tc(src.data_ptr(), 0, c8o * ref[0, 0].numel() * 8, TTA_F16_HI, N, c8o, *odims,
                                  wph.data_ptr(), 0, buf.data_ptr(), c8i * Vi * 8, c8i, *dims, 1 - mode, K, s, 0, fl,
                                  0, 0, stream()), "conv_tc dgrad wsplit")
        torch.cuda.synchronize()
        assert torch.equal(g_ws, g_std)


def test_conv_tc_reads_concat_slice_view(lib, cuda):
    """Input is the SECOND half of a concat buffer (c8 pitch > view chunks)."""
    torch.manual_seed(8)
    N, ctot, c0, cin, cout, dims = 2, 48, 16, 32, 32, (4, 16, 8)
    xfull = torch.randn(N, ctot, *dims)
    w = torch.randn(cout, cin, 3, 3, 3) * 0.1
    hi, lo, xv = planes_from(xfull.to(cuda), TTA_F16)
    V = dims[0] * dims[1] * dims[2]
    whi, wlo = split_planes(w, TTA_F16)
    ref = F.conv3d(xv.cpu()[:, c0:c0 + cin], join_planes(whi, wlo, TTA_F16), None, padding=1)
    wp = pack_weights_tc(wg_forward(w.to(cuda), False), 0, 3, 1, TTA_F16)
    out = torch.zeros((N, cout // 8, *dims, 8), device=cuda)
    off = (c0 // 8) * V * 8
    check(lib.tta_conv_tc(hi.view(-1)[off:].data_ptr(), lo.view(-1)[off:].data_ptr(), (ctot // 8) * V * 8, TTA_F16,
                          N, cin // 8, *dims, wp.data_ptr(), 0, out.data_ptr(), (cout // 8) * V * 8, cout // 8, *dims,
                          0, 3, 1, 0, 0, 0, 0, stream()), "conv_tc view")
    torch.cuda.synchronize()
    assert rel_l2(from_chunked(out, cout).cpu(), ref) < 2e-5


@pytest.mark.parametrize("cin,cout,s,tr,dims,stats_c", [(32, 32, 1, False, (5, 20, 12), 32), (32, 64, 2, False, (8, 16, 16), 32),
                                                         (48, 16, 2, True, (3, 10, 6), 16), (64, 3, 2, True, (4, 16, 8), 3),
                                                         (128, 256, 1, False, (2, 6, 6), 256)])
def test_conv_tc_fused_norm_statistics(lib, cuda, cin, cout, s, tr, dims, stats_c):
    """The conv epilogue leaves per-CTA partial sums of y and y^2 over the leading chunks
    (tta_conv_tc stats_workspace); tta_norm_stats_finalize turns them into mean/rstd.  Checked against
    torch statistics of the kernel's own fp32 output: 1e-5 relative (fp32 partial sums, fp64 finalize),
    instance and batch mode, twice (the CTA slots are re-zeroed by every launch)."""
    import ctypes
    torch.manual_seed(9)
    N = 2
    x = torch.randn(N, cin, *dims) + 0.3
    w = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3)) * 0.1
    b = torch.randn(cout)
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    mode = 1 if tr else 0
    odims = tuple(d * s for d in dims) if tr else tuple((d - 1) // s + 1 for d in dims)
    c8i, c8o, sc8 = (cin + 7) // 8, (cout + 7) // 8, (stats_c + 7) // 8
    Vi, Vo = dims[0] * dims[1] * dims[2], odims[0] * odims[1] * odims[2]
    wp = pack_weights_tc(wg_forward(w.to(cuda), tr), mode, 3, s, TTA_F16)
    ks, grid = ctypes.c_int(0), ctypes.c_int(0)
    check(lib.tta_conv_tc_query(TTA_F16, N, c8i, *dims, c8o, *odims, mode, 3, s, 0, 2, ctypes.byref(ks),
                                ctypes.byref(grid), 0), "query")
    assert ks.value == 1 and 1 <= grid.value <= 148
    ws = torch.full((1024 + N * sc8 * grid.value * 16,), 123.0, device=cuda)   # garbage: the conv must zero its slots
    out = torch.zeros((N, c8o, *odims, 8), device=cuda)
    bias = pack_bias(b.to(cuda))
    for _ in range(2):
        check(lib.tta_conv_tc(hi.data_ptr(), lo.data_ptr(), c8i * Vi * 8, TTA_F16, N, c8i, *dims, wp.data_ptr(),
                              bias.data_ptr(), out.data_ptr(), c8o * Vo * 8, c8o, *odims, mode, 3, s, 0, 2,
                              ws.data_ptr(), sc8, stream()), "conv_tc stats")
    y = from_chunked(out, cout)[:, :stats_c]
    for batch_mode in (0, 1):
        mean = torch.zeros(N * sc8 * 8, device=cuda); rstd = torch.zeros_like(mean)
        check(lib.tta_norm_stats_finalize(ws.data_ptr(), N, sc8, grid.value, Vo, batch_mode, 1e-5, mean.data_ptr(),
                                          rstd.data_ptr(), stream()), "finalize")
        dimsr = (0, 2, 3, 4) if batch_mode else (2, 3, 4)
        mu = y.double().mean(dimsr, keepdim=True)
        var = ((y.double() - mu) ** 2).mean(dimsr, keepdim=True)
        mu_ref = mu.expand(N, stats_c, 1, 1, 1).reshape(N, stats_c)
        rs_ref = (1.0 / torch.sqrt(var + 1e-5)).expand(N, stats_c, 1, 1, 1).reshape(N, stats_c)
        got_mu = mean.view(N, sc8 * 8)[:, :stats_c].double()
        got_rs = rstd.view(N, sc8 * 8)[:, :stats_c].double()
        assert float((got_mu - mu_ref).abs().max()) < 1e-5 * float(y.abs().max())
        assert float(((got_rs - rs_ref) / rs_ref).abs().max()) < 1e-5


@pytest.mark.parametrize("cin,cout,s,tr,dims,acc", [(32, 32, 1, False, (5, 20, 12), 0), (32, 64, 2, False, (8, 16, 16), 1),
                                                     (16, 3, 2, True, (3, 10, 6), 0), (32, 32, 1, False, (4, 16, 8), 1)])
def test_conv_tc_fused_norm_backward_sums(lib, cuda, cin, cout, s, tr, dims, acc):
    """tta_conv_tc_bwd_norm: the dgrad epilogue also reduces dz = g*[gamma*xhat+beta > 0] and dz*xhat for
    up to two norm layers (two channel segments of its output, as in a skip concat);
    tta_norm_bwd_finalize turns the per-CTA partials into sums / dgamma / dbeta.  Reference: torch on
    the kernel's own g output (fp32 partial sums, fp64 finalize: 2e-5 relative)."""
    import ctypes

    class Seg(ctypes.Structure):
        _fields_ = [("c8_begin", ctypes.c_int), ("c8_count", ctypes.c_int), ("relu", ctypes.c_int), ("pad", ctypes.c_int),
                    ("y", ctypes.c_void_p), ("y_ns", ctypes.c_longlong), ("mean", ctypes.c_void_p),
                    ("rstd", ctypes.c_void_p), ("gamma", ctypes.c_void_p), ("beta", ctypes.c_void_p),
                    ("partial", ctypes.c_void_p)]
    torch.manual_seed(13)
    N = 2
    # forward conv cin -> cout; its dgrad maps dy [N, cout, odims] to g [N, cin, dims]
    w = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3)) * 0.1
    odims = tuple(d * s for d in dims) if tr else tuple((d - 1) // s + 1 for d in dims)
    dy = torch.randn(N, cout, *odims)
    mode = 1 if tr else 0
    c8i, c8o = (cin + 7) // 8, (cout + 7) // 8
    Vi, Vo = dims[0] * dims[1] * dims[2], odims[0] * odims[1] * odims[2]
    hhi, _ = split_planes(to_chunked(dy.to(cuda)), TTA_F16_HI)
    hhi = hhi.contiguous()
    wph = pack_weights_tc(wg_dgrad(w.to(cuda), tr), 1 - mode, 3, s, TTA_F16_HI)
    ks, grid = ctypes.c_int(0), ctypes.c_int(0)
    check(lib.tta_conv_tc_query(TTA_F16_HI, N, c8o, *odims, c8i, *dims, 1 - mode, 3, s, acc, 2, ctypes.byref(ks),
                                ctypes.byref(grid), 0), "query")
    assert ks.value == 1
    # two norm layers own the two halves of the gradient's channels (one when there is a single chunk)
    bounds = [(0, c8i)] if c8i == 1 else [(0, c8i // 2), (c8i // 2, c8i)]
    segs = (Seg * len(bounds))()
    keep, layers = [], []
    for i, (b0, b1) in enumerate(bounds):
        C = min(cin, b1 * 8) - b0 * 8                          # real channels of this segment
        Cp = (b1 - b0) * 8
        y = torch.randn(N, C, *dims) * 1.5 + 0.2
        gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.5
        mu = y.mean((2, 3, 4)); rs = 1.0 / torch.sqrt(y.var((2, 3, 4), unbiased=False) + 1e-5)
        ych = to_chunked(y.to(cuda))
        mean = torch.zeros(N, Cp, device=cuda); mean[:, :C] = mu.to(cuda)
        rstd = torch.zeros(N, Cp, device=cuda); rstd[:, :C] = rs.to(cuda)
        gp = torch.zeros(Cp, device=cuda); gp[:C] = gamma.to(cuda)
        bp = torch.zeros(Cp, device=cuda); bp[:C] = beta.to(cuda)
        part = torch.full((N * (b1 - b0) * grid.value * 16,), 55.0, device=cuda)   # garbage: kernel zeroes its slots
        keep += [ych, mean, rstd, gp, bp, part]
        segs[i] = Seg(b0, b1 - b0, 1, 0, ych.data_ptr(), (b1 - b0) * Vi * 8, mean.data_ptr(), rstd.data_ptr(),
                      gp.data_ptr(), bp.data_ptr(), part.data_ptr())
        layers.append((b0, b1, C, y, gamma, beta, mu, rs, part))
    g = torch.full((N, c8i, *dims, 8), 0.25 if acc else 0.0, device=cuda)
    g[..., :] = g[..., :] if cin % 8 == 0 else g
    for _ in range(2 if not acc else 1):
        check(lib.tta_conv_tc_bwd_norm(hhi.data_ptr(), 0, c8o * Vo * 8, TTA_F16_HI, N, c8o, *odims, wph.data_ptr(),
                                       g.data_ptr(), c8i * Vi * 8, c8i, *dims, 1 - mode, 3, s, acc, 2, segs, len(bounds),
                                       stream()), "conv_tc_bwd_norm")
    torch.cuda.synchronize()
    gout = from_chunked(g, c8i * 8).cpu()
    for (b0, b1, C, y, gamma, beta, mu, rs, part) in layers:
        gs = gout[:, b0 * 8:b0 * 8 + C]
        xh = (y - mu[:, :, None, None, None]) * rs[:, :, None, None, None]
        z = xh * gamma[None, :, None, None, None] + beta[None, :, None, None, None]
        dz = gs * (z > 0)
        s1, s2 = dz.sum((2, 3, 4)), (dz * xh).sum((2, 3, 4))
        Cp = (b1 - b0) * 8
        sums = torch.zeros(N * Cp * 2, device=cuda); dg = torch.zeros(Cp, device=cuda); db = torch.zeros(Cp, device=cuda)
        check(lib.tta_norm_bwd_finalize(part.data_ptr(), N, b1 - b0, C, grid.value, 0, sums.data_ptr(), dg.data_ptr(),
                                        db.data_ptr(), stream()), "bwd_finalize")
        sm = sums.view(N, Cp, 2).cpu()
        scale = float(s1.abs().max()) + float(s2.abs().max())
        assert float((sm[:, :C, 0] - s1).abs().max()) < 2e-5 * scale
        assert float((sm[:, :C, 1] - s2).abs().max()) < 2e-5 * scale
        assert float((db[:C].cpu() - s1.sum(0)).abs().max()) < 2e-5 * scale * N
        assert float((dg[:C].cpu() - s2.sum(0)).abs().max()) < 2e-5 * scale * N


@pytest.mark.parametrize("cin,cout,dims", [(64, 3, (4, 16, 8)), (64, 3, (5, 17, 9)), (32, 1, (3, 8, 8)),
                                            (16, 4, (2, 6, 5)), (64, 3, (20, 33, 20)), (48, 2, (9, 15, 7))])
def test_conv_tc_small_cout_transposed_gemm_col2im(lib, cuda, cin, cout, dims):
    """Transposed stride-2 conv with <= 4 output channels (the UNet head convT 64 -> 3) as a dense
    GEMM per input plane + shared-memory col2im (conv_t2s_kernel; flags bits 8..10 = Cout, weights from
    pack_weights_tc(t2s=True)): against torch conv_transpose3d on the same split-rounded operands
    (rel-L2 <= 2e-5 as for every fp16 hi/lo conv), with the fused norm statistics, twice (slots are
    re-zeroed by every launch), including tile edges (tiles advance by 15 x 7 input voxels) and
    several d-segments."""
    import ctypes
    torch.manual_seed(21)
    N = 2
    assert lib.tta_conv_tc_t2s(1, 3, 2, cin, cout, 1) == 1
    x = torch.randn(N, cin, *dims)
    w = torch.randn(cin, cout, 3, 3, 3) * 0.1
    b = torch.randn(cout)
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    whi, wlo = split_planes(w, TTA_F16)
    ref = F.conv_transpose3d(xv.cpu(), join_planes(whi, wlo, TTA_F16), b, stride=2, padding=1, output_padding=1)
    odims = tuple(ref.shape[2:])
    c8i = cin // 8
    Vi, Vo = dims[0] * dims[1] * dims[2], odims[0] * odims[1] * odims[2]
    wp = pack_weights_tc(wg_forward(w.to(cuda), True), 1, 3, 2, TTA_F16, t2s=True)
    flags = cout << 8
    ks, grid = ctypes.c_int(0), ctypes.c_int(0)
    check(lib.tta_conv_tc_query(TTA_F16, N, c8i, *dims, 1, *odims, 1, 3, 2, 0, flags, ctypes.byref(ks),
                                ctypes.byref(grid), 0), "query")
    assert ks.value == 1 and 1 <= grid.value <= 148
    ws = torch.full((1024 + N * grid.value * 16,), 55.0, device=cuda)
    out = torch.full((N, 1, *odims, 8), 9.0, device=cuda)
    bias = pack_bias(b.to(cuda))
    for _ in range(2):
        check(lib.tta_conv_tc(hi.data_ptr(), lo.data_ptr(), c8i * Vi * 8, TTA_F16, N, c8i, *dims, wp.data_ptr(),
                              bias.data_ptr(), out.data_ptr(), Vo * 8, 1, *odims, 1, 3, 2, 0, flags,
                              ws.data_ptr(), 1, stream()), "conv_tc t2s")
    torch.cuda.synchronize()
    got = from_chunked(out, cout).cpu()
    assert rel_l2(got, ref) < 2e-5, rel_l2(got, ref)
    assert float(out[..., cout:].abs().max()) == 0.0                     # pad channels exactly zero
    mean = torch.zeros(N * 8, device=cuda); rstd = torch.zeros_like(mean)
    check(lib.tta_norm_stats_finalize(ws.data_ptr(), N, 1, grid.value, Vo, 0, 1e-5, mean.data_ptr(), rstd.data_ptr(),
                                      stream()), "finalize")
    y = from_chunked(out, cout).double()
    mu = y.mean((2, 3, 4)); var = ((y - mu.view(N, cout, 1, 1, 1)) ** 2).mean((2, 3, 4))
    assert float((mean.view(N, 8)[:, :cout].double() - mu).abs().max()) < 1e-5 * float(y.abs().max())
    rs_ref = 1.0 / torch.sqrt(var + 1e-5)
    assert float(((rstd.view(N, 8)[:, :cout].double() - rs_ref) / rs_ref).abs().max()) < 1e-5
    # the generic parity-class path (no flag, generic packing) computes the same conv
    wpg = pack_weights_tc(wg_forward(w.to(cuda), True), 1, 3, 2, TTA_F16)
    out_g = torch.zeros_like(out)
    _tc(lib, hi, lo, c8i * Vi * 8, TTA_F16, N, c8i, dims, wpg, bias, out_g, Vo * 8, 1, odims, 1, 3, 2)
    assert rel_l2(out.cpu(), out_g.cpu()) < 2e-5


@pytest.mark.parametrize("cin,cout,dims,dt", [(4, 64, (8, 32, 16), TTA_F16), (3, 64, (6, 20, 12), TTA_F16_HI),
                                              (2, 32, (4, 36, 24), TTA_F16), (1, 16, (2, 16, 16), TTA_BF16)])
def test_stride2_conv_over_compact_input(lib, cuda, cin, cout, dims, dt):
    """GEOM_S2C4 (flags bit 15): stride-2 3x3x3 conv whose <= 4-channel input is stored [N][D][H][W][4] (network
    input; gradient of the head norm): against torch on the same rounded operands, and against the 8-channel chunk
    path (different MMA grouping -> same value up to fp32 summation order)."""
    torch.manual_seed(9)
    N = 2
    x = torch.randn(N, cin, *dims)
    w = torch.randn(cout, cin, 3, 3, 3) * 0.1
    b = torch.randn(cout)
    hi8, lo8, xv = planes_from(x.to(cuda), dt)
    whi, wlo = split_planes(w, dt)
    wv = join_planes(whi, wlo, dt)
    ref = F.conv3d(xv.cpu(), wv, b, stride=2, padding=1)
    odims = tuple(ref.shape[2:])
    V, Vo = dims[0] * dims[1] * dims[2], odims[0] * odims[1] * odims[2]
    c8o = (cout + 7) // 8
    hi4 = hi8[:, 0, ..., :4].contiguous(); lo4 = lo8[:, 0, ..., :4].contiguous()      # [N][D][H][W][4]
    wg = wg_forward(w.to(cuda), False)
    wp4 = pack_weights_tc(wg, 0, 3, 2, dt, s2c4=True)
    out4 = torch.zeros((N, c8o, *odims, 8), device=cuda)
    _tc(lib, hi4, lo4, V * 4, dt, N, 1, dims, wp4, pack_bias(b.to(cuda)), out4, c8o * Vo * 8, c8o, odims, 0, 3, 2,
        flags=32768)
    got = from_chunked(out4, cout).cpu()
    tol = {TTA_F16: 2e-5, TTA_F16_HI: 2e-5, TTA_BF16: 1e-4}[dt]
    assert rel_l2(got, ref) < tol, rel_l2(got, ref)
    out8 = torch.zeros_like(out4)
    _tc(lib, hi8, lo8, V * 8, dt, N, 1, dims, pack_weights_tc(wg, 0, 3, 2, dt), pack_bias(b.to(cuda)), out8,
        c8o * Vo * 8, c8o, odims, 0, 3, 2, flags=2)
    assert rel_l2(out4.cpu(), out8.cpu()) < 2e-6
    assert lib.tta_conv_tc_s2c4(0, 3, 2, cin) == 1 and lib.tta_conv_tc_s2c4(0, 3, 2, 8) == 0
    assert lib.tta_conv_tc_s2c4(0, 3, 1, 4) == 0 and lib.tta_conv_tc_s2c4(1, 3, 2, 4) == 0
