"""Op-level parity: every C-ABI kernel against the CPU oracle / plain torch fp32 on seeded inputs.
Tolerances are stated per test; integer work (Dice counts) is bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from multimodal_tta_b200._lib import TTA_BF16, TTA_F16, TTA_F16_HI, check
from multimodal_tta_b200.layout import (from_chunked, join_planes, pack_bias, pack_weights_simt, to_chunked,
                                        wg_dgrad, wg_forward)
from tests.util import planes_from, rel_l2, stream, wsplit

pytestmark = pytest.mark.gpu

GEOMS = [  # (cin, cout, K, stride, transposed, dims)
    (4, 32, 3, 2, False, (12, 10, 16)),
    (32, 32, 3, 1, False, (6, 9, 12)),
    (16, 24, 3, 1, False, (5, 5, 5)),
    (64, 16, 3, 2, True, (4, 5, 6)),
    (24, 3, 3, 2, True, (6, 6, 8)),
    (3, 3, 3, 1, False, (8, 8, 8)),
    (40, 48, 1, 1, False, (4, 6, 8)),
    (2, 32, 3, 2, False, (8, 8, 8)),
]


def _run_conv(lib, which, hi, lo, dt, N, cin, idims, wp, bias, cout, odims, mode, K, s, acc=None):
    dev = hi.device
    out = torch.zeros((N, (cout + 7) // 8, *odims, 8), device=dev) if acc is None else acc
    c8i, c8o = (cin + 7) // 8, (cout + 7) // 8
    ins = c8i * idims[0] * idims[1] * idims[2] * 8
    ons = c8o * odims[0] * odims[1] * odims[2] * 8
    args = [hi.data_ptr(), lo.data_ptr(), ins, dt, N, c8i, *idims, wp.data_ptr(),
            bias.data_ptr() if bias is not None else 0, out.data_ptr(), ons, c8o, *odims, mode, K, s,
            int(acc is not None)]
    if which == "simt":
        check(lib.tta_conv_simt(*args, stream()), "conv_simt")
    else:
        check(lib.tta_conv_tc(*args, 0, 0, 0, stream()), "conv_tc")
    return out


@pytest.mark.parametrize("cin,cout,K,s,tr,dims", GEOMS)
def test_conv_simt_forward_and_dgrad(lib, cuda, cin, cout, K, s, tr, dims):
    torch.manual_seed(1)
    N = 2
    x = torch.randn(N, cin, *dims)
    w = torch.randn((cin, cout, K, K, K) if tr else (cout, cin, K, K, K)) * 0.1
    b = torch.randn(cout)
    pad = (K - 1) // 2
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    xr = xv.cpu().requires_grad_(True)
    if tr:
        ref = F.conv_transpose3d(xr, w, b, stride=s, padding=pad, output_padding=s - 1)
    else:
        ref = F.conv3d(xr, w, b, stride=s, padding=pad)
    odims = tuple(ref.shape[2:])
    wp = pack_weights_simt(wg_forward(w.to(cuda), tr))
    out = _run_conv(lib, "simt", hi, lo, TTA_F16, N, cin, dims, wp, pack_bias(b.to(cuda)), cout, odims,
                    1 if tr else 0, K, s)
    got = from_chunked(out, cout).cpu()
    assert rel_l2(got, ref.detach()) < 2e-6          # fp32 FMA order only
    # dgrad: conv of dy with the dgrad weights must equal autograd's input gradient
    dy = torch.randn_like(ref)
    dhi, dlo, dyv = planes_from(dy.to(cuda), TTA_BF16)
    wpd = pack_weights_simt(wg_dgrad(w.to(cuda), tr))
    gx = _run_conv(lib, "simt", dhi, dlo, TTA_BF16, N, cout, odims, wpd, None, cin, dims, 0 if tr else 1, K, s)
    # reference computed on the bf16x2-rounded dy the kernel actually saw
    xr2 = xv.cpu().requires_grad_(True)
    ref2 = F.conv_transpose3d(xr2, w, b, stride=s, padding=pad, output_padding=s - 1) if tr else \
        F.conv3d(xr2, w, b, stride=s, padding=pad)
    (g2,) = torch.autograd.grad(ref2, xr2, dyv.cpu())
    assert rel_l2(from_chunked(gx, cin).cpu(), g2) < 2e-6
    # accumulate flag adds onto the existing output
    gx2 = _run_conv(lib, "simt", dhi, dlo, TTA_BF16, N, cout, odims, wpd, None, cin, dims, 0 if tr else 1, K, s,
                    acc=gx.clone())
    assert rel_l2(from_chunked(gx2, cin).cpu(), 2 * g2) < 2e-6


@pytest.mark.parametrize("cin,cout,dims", [(3, 3, (9, 10, 13)), (1, 1, (8, 8, 8)), (4, 2, (5, 6, 7)), (2, 2, (4, 4, 40))])
def test_conv_small_forward_and_dgrad(lib, cuda, cin, cout, dims):
    from multimodal_tta_b200.layout import pack_weights_small
    torch.manual_seed(9)
    N = 2
    x = torch.randn(N, cin, *dims)
    w = torch.randn(cout, cin, 3, 3, 3) * 0.2
    b = torch.randn(cout)
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    xr = xv.cpu().requires_grad_(True)
    ref = F.conv3d(xr, w, b, padding=1)
    V = dims[0] * dims[1] * dims[2]
    out = torch.zeros((N, 1, *dims, 8), device=cuda)
    wp = pack_weights_small(wg_forward(w.to(cuda), False), 0)
    bp = pack_bias(b.to(cuda))
    check(lib.tta_conv_small(hi.data_ptr(), lo.data_ptr(), V * 8, TTA_F16, N, cin, *dims, wp.data_ptr(), bp.data_ptr(),
                             out.data_ptr(), V * 8, cout, 0, stream()))
    assert rel_l2(from_chunked(out, cout).cpu(), ref.detach()) < 2e-6
    assert float(out[..., cout:].abs().max()) == 0.0 if cout < 8 else True
    dy = torch.randn_like(ref)
    dhi, _, dyv = planes_from(dy.to(cuda), TTA_F16_HI)
    (g,) = torch.autograd.grad(ref, xr, dyv.cpu())
    wpd = pack_weights_small(wg_dgrad(w.to(cuda), False), 1)
    gx = torch.zeros((N, 1, *dims, 8), device=cuda)
    check(lib.tta_conv_small(dhi.data_ptr(), 0, V * 8, TTA_F16_HI, N, cout, *dims, wpd.data_ptr(), 0, gx.data_ptr(),
                             V * 8, cin, 0, stream()))
    assert rel_l2(from_chunked(gx, cin).cpu(), g) < 2e-6


@pytest.mark.parametrize("batch_mode", [0, 1])
@pytest.mark.parametrize("C,dims", [(32, (8, 9, 10)), (3, (16, 16, 16)), (20, (5, 7, 3)), (16, (20, 24, 28))])
def test_norm_forward_backward(lib, cuda, batch_mode, C, dims):
    torch.manual_seed(2)
    N = 2
    V = dims[0] * dims[1] * dims[2]
    C8 = (C + 7) // 8
    y = (torch.randn(N, C, *dims) * 1.7 + 0.4)
    gamma = torch.rand(C) + 0.5
    beta = torch.randn(C) * 0.2
    resid = torch.randn(N, C, *dims)
    g_in = torch.randn(N, C, *dims)
    # oracle (CPU fp32 autograd)
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    if batch_mode:
        z = F.batch_norm(yr, None, None, gr, br, training=True, eps=1e-5)
    else:
        z = F.instance_norm(yr, weight=gr, bias=br, eps=1e-5)
    a = torch.relu(z) + resid
    a.backward(g_in)
    # device
    ych = to_chunked(y.to(cuda))
    mean = torch.zeros(N * C8 * 8, device=cuda); rstd = torch.zeros_like(mean)
    ws = torch.zeros(lib.tta_norm_workspace_floats(N, C8, V), device=cuda)
    gp = torch.zeros(C8 * 8, device=cuda); gp[:C] = gamma.to(cuda)
    bp = torch.zeros(C8 * 8, device=cuda); bp[:C] = beta.to(cuda)
    ns = C8 * V * 8
    check(lib.tta_norm_stats(ych.data_ptr(), ns, N, C8, V, batch_mode, 1e-5, mean.data_ptr(), rstd.data_ptr(),
                             ws.data_ptr(), 1, stream()))
    rch = to_chunked(resid.to(cuda))
    ohi = torch.zeros((N, C8, *dims, 8), dtype=torch.int16, device=cuda); olo = torch.zeros_like(ohi)
    check(lib.tta_norm_apply(ych.data_ptr(), ns, N, C8, V, mean.data_ptr(), rstd.data_ptr(), gp.data_ptr(),
                             bp.data_ptr(), 1, 1, rch.data_ptr(), 0, ns, ohi.data_ptr(), olo.data_ptr(), ns,
                             TTA_F16, 0, 0, batch_mode, 1e-5, 0, 0, 0, 0, stream()))
    got = from_chunked(join_planes(ohi, olo, TTA_F16), C).cpu()
    assert (got - a.detach()).abs().max() < 2e-5      # fp16x2 storage (22 bits) + fp32 stats
    # fused path: partial sums only, finalize inside the apply prologue -> identical planes
    mean2 = torch.zeros_like(mean); rstd2 = torch.zeros_like(rstd)
    check(lib.tta_norm_stats(ych.data_ptr(), ns, N, C8, V, batch_mode, 1e-5, 0, 0, ws.data_ptr(), 0, stream()))
    ohi2 = torch.zeros_like(ohi); olo2 = torch.zeros_like(ohi)
    check(lib.tta_norm_apply(ych.data_ptr(), ns, N, C8, V, mean2.data_ptr(), rstd2.data_ptr(), gp.data_ptr(),
                             bp.data_ptr(), 1, 1, rch.data_ptr(), 0, ns, ohi2.data_ptr(), olo2.data_ptr(), ns,
                             TTA_F16, ws.data_ptr(), 0, batch_mode, 1e-5, 0, 0, 0, 0, stream()))
    got2 = from_chunked(join_planes(ohi2, olo2, TTA_F16), C).cpu()
    assert (got2 - got).abs().max() < 2e-6            # same math, fp64 partial sums in another order
    if dims[2] % 2 == 0:
        # second copy in the w-parity-split layout (operand of a stride-2 tcgen05 conv): same bits,
        # rows permuted to [even w | odd w]
        whi = torch.zeros_like(ohi); wlo = torch.zeros_like(ohi)
        ohi3 = torch.zeros_like(ohi); olo3 = torch.zeros_like(ohi)
        check(lib.tta_norm_apply(ych.data_ptr(), ns, N, C8, V, mean.data_ptr(), rstd.data_ptr(), gp.data_ptr(),
                                 bp.data_ptr(), 1, 1, rch.data_ptr(), 0, ns, ohi3.data_ptr(), olo3.data_ptr(), ns,
                                 TTA_F16, 0, 0, batch_mode, 1e-5, whi.data_ptr(), wlo.data_ptr(), ns, dims[2],
                                 stream()))
        assert torch.equal(ohi3, ohi) and torch.equal(olo3, olo)
        assert torch.equal(whi, wsplit(ohi)) and torch.equal(wlo, wsplit(olo))
    assert torch.allclose(mean, mean2, rtol=1e-6, atol=1e-7) and torch.allclose(rstd, rstd2, rtol=1e-6)
    # backward
    gch = to_chunked(g_in.to(cuda))
    sums = torch.zeros(N * C8 * 8 * 2, device=cuda)
    dg = torch.zeros(C8 * 8, device=cuda); db = torch.zeros(C8 * 8, device=cuda)
    check(lib.tta_norm_bwd_reduce(gch.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, C8, C, V, mean.data_ptr(),
                                  rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode, sums.data_ptr(),
                                  dg.data_ptr(), db.data_ptr(), ws.data_ptr(), 1, 0, stream()))
    assert rel_l2(dg[:C].cpu(), gr.grad) < 1e-5
    assert rel_l2(db[:C].cpu(), br.grad) < 1e-5
    dhi = torch.zeros_like(ohi); dlo = torch.zeros_like(ohi); ahi = torch.zeros_like(ohi); alo = torch.zeros_like(ohi)
    check(lib.tta_norm_bwd_apply(gch.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, C8, V, mean.data_ptr(),
                                 rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode, sums.data_ptr(),
                                 dhi.data_ptr(), dlo.data_ptr(), ns, ahi.data_ptr(), alo.data_ptr(), ns, TTA_BF16,
                                 0, C, 0, 0, 0, stream()))
    if dims[2] % 2 == 0:
        dhw = torch.zeros_like(ohi); dlw = torch.zeros_like(ohi)
        check(lib.tta_norm_bwd_apply(gch.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, C8, V, mean.data_ptr(),
                                     rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode, sums.data_ptr(),
                                     dhw.data_ptr(), dlw.data_ptr(), ns, 0, 0, 0, TTA_BF16,
                                     0, C, 0, 0, dims[2], stream()))
        assert torch.equal(dhw, wsplit(dhi)) and torch.equal(dlw, wsplit(dlo))
    # fused path: finalize of the reductions inside the bwd-apply prologue
    dg2 = torch.zeros_like(dg); db2 = torch.zeros_like(db)
    dhi2 = torch.zeros_like(ohi); dlo2 = torch.zeros_like(ohi)
    check(lib.tta_norm_bwd_reduce(gch.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, C8, C, V, mean.data_ptr(),
                                  rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode, 0, 0, 0,
                                  ws.data_ptr(), 0, 0, stream()))
    check(lib.tta_norm_bwd_apply(gch.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, C8, V, mean.data_ptr(),
                                 rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode, 0,
                                 dhi2.data_ptr(), dlo2.data_ptr(), ns, 0, 0, 0, TTA_BF16, ws.data_ptr(), C,
                                 dg2.data_ptr(), db2.data_ptr(), 0, stream()))
    assert torch.allclose(dg, dg2, rtol=1e-6, atol=1e-7) and torch.allclose(db, db2, rtol=1e-6, atol=1e-7)
    dy2 = from_chunked(join_planes(dhi2, dlo2, TTA_BF16), C).cpu()
    assert rel_l2(dy2, yr.grad) < 3e-5
    dy = from_chunked(join_planes(dhi, dlo, TTA_BF16), C).cpu()
    assert rel_l2(dy, yr.grad) < 3e-5                 # bf16x2 storage (~16 bits)
    aux = from_chunked(join_planes(ahi, alo, TTA_BF16), C).cpu()
    assert rel_l2(aux, g_in) < 2e-5
    if lib.tta_norm_small_supported(N, V, batch_mode):
        # small-layer variants: statistics + apply (and reduction + apply) in ONE launch each; the
        # sums are reduced in another order (fp32 tree per CTA), everything else is the same code
        mean3 = torch.zeros_like(mean); rstd3 = torch.zeros_like(rstd)
        ohi4 = torch.zeros_like(ohi); olo4 = torch.zeros_like(ohi)
        check(lib.tta_norm_fwd_small(ych.data_ptr(), ns, N, C8, V, 1e-5, mean3.data_ptr(), rstd3.data_ptr(),
                                     gp.data_ptr(), bp.data_ptr(), 1, 1, rch.data_ptr(), 0, ns, ohi4.data_ptr(),
                                     olo4.data_ptr(), ns, TTA_F16, 0, 0, 0, 0, stream()))
        assert torch.allclose(mean3, mean, rtol=1e-5, atol=1e-6) and torch.allclose(rstd3, rstd, rtol=1e-5)
        got4 = from_chunked(join_planes(ohi4, olo4, TTA_F16), C).cpu()
        assert (got4 - a.detach()).abs().max() < 2e-5
        sums3 = torch.zeros_like(sums); dg3 = torch.zeros_like(dg); db3 = torch.zeros_like(db)
        dhi3 = torch.zeros_like(ohi); dlo3 = torch.zeros_like(ohi); ahi3 = torch.zeros_like(ohi); alo3 = torch.zeros_like(ohi)
        for _ in range(2):                           # twice: self-resetting block counters
            check(lib.tta_norm_bwd_small(gch.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, C8, C, V, mean.data_ptr(),
                                         rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, sums3.data_ptr(),
                                         dg3.data_ptr(), db3.data_ptr(), dhi3.data_ptr(), dlo3.data_ptr(), ns,
                                         ahi3.data_ptr(), alo3.data_ptr(), ns, TTA_BF16, 0, ws.data_ptr(), stream()))
        assert rel_l2(dg3[:C].cpu(), gr.grad) < 1e-5 and rel_l2(db3[:C].cpu(), br.grad) < 1e-5
        assert torch.allclose(sums3, sums, rtol=1e-4, atol=1e-5)
        assert rel_l2(from_chunked(join_planes(dhi3, dlo3, TTA_BF16), C).cpu(), yr.grad) < 3e-5
        assert torch.equal(ahi3, ahi) and torch.equal(alo3, alo)


@pytest.mark.parametrize("mode,R", [(1, 3), (1, 1), (0, 3), (0, 2)])
def test_head_entropy(lib, cuda, mode, R):
    from oracle.tent_oracle import entropy_loss
    torch.manual_seed(3)
    N, dims = 2, (6, 7, 9)
    V = dims[0] * dims[1] * dims[2]
    z = torch.randn(N, R, *dims) * 3
    zr = z.clone().requires_grad_(True)
    loss = entropy_loss(zr, "sigmoid" if mode == 1 else "softmax")
    loss.backward()
    ych = to_chunked(z.to(cuda))
    logits = torch.zeros(N, R, *dims, device=cuda)
    dhi = torch.zeros((N, 1, *dims, 8), dtype=torch.int16, device=cuda); dlo = torch.zeros_like(dhi)
    nb = lib.tta_head_entropy_blocks(N, V)
    part = torch.zeros(nb * N, device=cuda); lossd = torch.zeros(1, device=cuda)
    check(lib.tta_head_entropy(ych.data_ptr(), V * 8, N, R, V, mode, 1.0 / (N * V), 1.0, TTA_BF16, 0,
                               logits.data_ptr(), dhi.data_ptr(), dlo.data_ptr(), V * 8, part.data_ptr(),
                               lossd.data_ptr(), stream()))
    assert torch.equal(logits.cpu(), z)               # pure layout change: bit exact
    assert abs(float(lossd) - float(loss.detach())) < 1e-6 * max(1.0, abs(float(loss.detach())))
    dz = from_chunked(join_planes(dhi, dlo, TTA_BF16), R).cpu()
    assert rel_l2(dz, zr.grad) < 2e-5                 # bf16x2 storage
    assert float(join_planes(dhi, dlo, TTA_BF16)[..., R:].abs().max()) == 0.0 if R < 8 else True
    # loss-scaled single-plane fp16 gradient (the default backward operand format)
    S = 2.0 ** 12
    check(lib.tta_head_entropy(ych.data_ptr(), V * 8, N, R, V, mode, 1.0 / (N * V), S, TTA_F16_HI, 0,
                               logits.data_ptr(), dhi.data_ptr(), 0, V * 8, part.data_ptr(), lossd.data_ptr(),
                               stream()))
    dz16 = from_chunked(join_planes(dhi, dlo, TTA_F16_HI), R).cpu() / S
    assert rel_l2(dz16, zr.grad) < 6e-4              # 11-bit mantissa


def test_adam_matches_torch(lib, cuda):
    torch.manual_seed(4)
    n = 4870
    p0 = torch.randn(n)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    pd = p0.clone().to(cuda); m = torch.zeros(n, device=cuda); v = torch.zeros(n, device=cuda)
    step = torch.zeros(1, dtype=torch.int32, device=cuda)
    for it in range(5):
        g = torch.randn(n) * (10.0 ** -it)
        p.grad = g.clone()
        opt.step()
        gd = g.to(cuda)
        check(lib.tta_adam_step(pd.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8,
                                1.0, step.data_ptr(), stream()))
        assert (pd.cpu() - p.detach()).abs().max() < 5e-7, it
    assert int(step) == 5


@pytest.mark.parametrize("vw,roi", [(11, (6, 8, 6)), (11, (6, 8, 8)), (12, (6, 8, 8)), (16, (5, 7, 12))])
def test_gather_pack_windows_and_padding(lib, cuda, vw, roi):
    # roi W % 4 == 0 takes the four-voxels-per-thread kernel: 128-bit loads where the source run is
    # 16-byte aligned and inside the volume, per-element loads otherwise (vw = 11 / window w0 = 5, -3)
    torch.manual_seed(5)
    vol = torch.randn(2, 3, 9, 10, vw)
    wins = torch.tensor([[0, 0, 0, 0], [1, 3, 2, 5], [1, -2, -1, -3]], dtype=torch.int32)
    scale = torch.tensor([[1., 1., 1.], [1., 0., 1.], [0., 1., 1.]])
    hi = torch.zeros((3, 1, *roi, 8), dtype=torch.int16, device=cuda); lo = torch.zeros_like(hi)
    vd, wd, sd = vol.to(cuda), wins.to(cuda), scale.to(cuda)   # keep alive across the async launch
    check(lib.tta_gather_pack(vd.data_ptr(), 2, 3, 9, 10, vw, wd.data_ptr(),
                              sd.data_ptr(), 3, *roi, hi.data_ptr(), lo.data_ptr(),
                              roi[0] * roi[1] * roi[2] * 8, 1, 0, stream()))
    hw = torch.zeros_like(hi); lw = torch.zeros_like(hi)
    check(lib.tta_gather_pack(vd.data_ptr(), 2, 3, 9, 10, vw, wd.data_ptr(),
                              sd.data_ptr(), 3, *roi, hw.data_ptr(), lw.data_ptr(),
                              roi[0] * roi[1] * roi[2] * 8, 1, 1, stream()))
    assert torch.equal(hw, wsplit(hi)) and torch.equal(lw, wsplit(lo))   # w-parity-split variant
    # compact variant (wsplit = 2: [NB][D][H][W][4], what a GEOM_S2C4 stem reads) = the first four channels, and the
    # same from an fp16-staged volume (values exactly representable in fp16 here: identical planes)
    V = roi[0] * roi[1] * roi[2]
    hc = torch.full((3, *roi, 4), 7, dtype=torch.int16, device=cuda); lc = torch.full_like(hc, 7)
    check(lib.tta_gather_pack(vd.data_ptr(), 2, 3, 9, 10, vw, wd.data_ptr(), sd.data_ptr(), 3, *roi, hc.data_ptr(),
                              lc.data_ptr(), V * 4, 1, 2, stream()))
    assert torch.equal(hc, hi[:, 0, ..., :4]) and torch.equal(lc, lo[:, 0, ..., :4])
    v16 = vd.half()
    h16 = torch.zeros_like(hc); l16 = torch.zeros_like(hc)
    check(lib.tta_gather_pack_norm_f16(v16.data_ptr(), 2, 3, 9, 10, vw, wd.data_ptr(), sd.data_ptr(), 0, 3, *roi,
                                       h16.data_ptr(), l16.data_ptr(), V * 4, 1, 2, stream()))
    r16 = torch.zeros_like(hc); q16 = torch.zeros_like(hc)
    v32 = v16.float()
    check(lib.tta_gather_pack(v32.data_ptr(), 2, 3, 9, 10, vw, wd.data_ptr(), sd.data_ptr(), 3, *roi, r16.data_ptr(),
                              q16.data_ptr(), V * 4, 1, 2, stream()))
    assert torch.equal(h16, r16) and torch.equal(l16, q16) and int(l16.abs().max()) == 0
    got = from_chunked(join_planes(hi, lo, TTA_F16), 3).cpu()
    pv = F.pad(vol, (8, 8, 8, 8, 8, 8))
    for b, (vi, d0, h0, w0) in enumerate(wins.tolist()):
        ref = pv[vi, :, d0 + 8:d0 + 8 + roi[0], h0 + 8:h0 + 8 + roi[1], w0 + 8:w0 + 8 + roi[2]] * scale[b].view(3, 1, 1, 1)
        assert (got[b] - ref).abs().max() < 1e-6      # fp16x2 storage of O(1) values


def test_dice_counts_match_reference_golden(lib, cuda):
    import os
    from multimodal_tta_b200.evaluation import device_dice_counts, dice_iou_from_counts
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "dice_golden.npz"))
    for i in range(4):
        pred, gt = torch.from_numpy(gold[f"pred{i}"]), torch.from_numpy(gold[f"gt{i}"])
        logits = (pred.float() * 2 - 1) * 3.0          # sigmoid(+-3) on either side of 0.5
        counts = device_dice_counts(logits.to(cuda), gt.float().to(cuda), 0.5).cpu()
        dice, iou, valid = dice_iou_from_counts(counts)
        assert torch.equal(dice, torch.from_numpy(gold[f"dice{i}"]))     # bit exact
        assert torch.equal(iou, torch.from_numpy(gold[f"iou{i}"]))
        assert torch.equal(valid, torch.from_numpy(gold[f"valid{i}"]))


@pytest.mark.parametrize("shape", [(2, 3, 9, 10, 37), (1, 1, 5, 17, 6), (2, 3, 16, 16, 64), (1, 2, 33, 40, 48)])
def test_dice_counts_vector_and_scalar_paths(lib, cuda, shape):
    """tta_dice_counts: four voxels per 128-bit load when every row starts 16-byte aligned (V % 4 == 0), the scalar
    loop otherwise and for misaligned views; random logits around the threshold, thresholds 0.5 and 0.3
    (seg_eval.py:41-68, hecktor21.yaml:80); `out=` adds into a caller-owned zeroed slot."""
    from multimodal_tta_b200.evaluation import device_dice_counts
    torch.manual_seed(3)
    logits = torch.randn(shape) * 2
    labels = (torch.rand(shape) > 0.6).float()
    for thr in (0.5, 0.3):
        pred = (torch.sigmoid(logits) >= thr).long()
        gt = (labels > 0.5).long()
        ref = torch.stack([(pred * gt).flatten(2).sum(-1), pred.flatten(2).sum(-1), gt.flatten(2).sum(-1)], dim=-1)
        got = device_dice_counts(logits.to(cuda), labels.to(cuda), thr).cpu()
        assert torch.equal(got, ref)
        ring = torch.zeros((3, shape[0], shape[1], 3), dtype=torch.int64, device=cuda)
        device_dice_counts(logits.to(cuda), labels.to(cuda), thr, out=ring[1])
        assert torch.equal(ring[1].cpu(), ref) and int(ring[0].abs().sum()) == 0 and int(ring[2].abs().sum()) == 0
    # a view that starts 4 bytes into an allocation: rows are no longer 16-byte aligned
    flat = torch.zeros(logits.numel() + 1)
    flat[1:] = logits.flatten()
    lview = flat.to(cuda)[1:].view(shape)
    pred = (torch.sigmoid(logits) >= 0.5).long()
    ref = torch.stack([(pred * gt).flatten(2).sum(-1), pred.flatten(2).sum(-1), gt.flatten(2).sum(-1)], dim=-1)
    assert torch.equal(device_dice_counts(lview, labels.to(cuda), 0.5).cpu(), ref)


@pytest.mark.parametrize("cpv", [8, 4])
@pytest.mark.parametrize("mode,C,dims,batch_mode", [(1, 3, (9, 10, 37), 0), (0, 3, (8, 8, 32), 0), (1, 1, (5, 17, 6), 0),
                                                     (1, 2, (16, 8, 40), 1), (0, 4, (3, 9, 33), 0)])
def test_fused_head_forward_and_backward(lib, cuda, mode, C, dims, batch_mode, cpv):
    """tta_head_fused_{fwd,bwd} (norm apply + 3x3x3 conv + entropy, and its backward up to the
    norm-backward reduction) against torch fp32 autograd on the CPU; ragged tiles (dims are not
    multiples of the 8x8x32 tile).  fp32 FMA chains in a different order: 3e-6 relative.
    cpv = 8: y / dz in the 8-channel chunk layout; cpv = 4: the compact layout (y fp32 [N][V][4], dz fp16 [N][V][4],
    tta_norm_bwd_apply_c4) -- same math, half / a quarter of the bytes."""
    from multimodal_tta_b200.layout import pack_weights_small
    from oracle.tent_oracle import entropy_loss
    torch.manual_seed(11)
    N = 2
    V = dims[0] * dims[1] * dims[2]
    y = torch.randn(N, C, *dims) * 2 + 0.5
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.3
    w = torch.randn(C, C, 3, 3, 3) * 0.2
    b = torch.randn(C) * 0.1
    S = 2.0 ** 10                                   # loss scale carried by dlogits / dz / dgamma
    # ---- oracle
    yr = y.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    zn = F.batch_norm(yr, None, None, gr, br, training=True, eps=1e-5) if batch_mode else \
        F.instance_norm(yr, weight=gr, bias=br, eps=1e-5)
    a = torch.relu(zn)
    a.retain_grad()
    z = F.conv3d(a, w, b, padding=1)
    z.retain_grad()
    loss = entropy_loss(z, "sigmoid" if mode == 1 else "softmax")
    loss.backward()
    # ---- device
    ych = to_chunked(y.to(cuda))
    ns = V * 8
    mean = torch.zeros(N * 8, device=cuda); rstd = torch.zeros_like(mean)
    nws = max(lib.tta_norm_workspace_floats(N, 1, V), lib.tta_head_fused_workspace_floats(N, *dims))
    ws = torch.zeros(nws, device=cuda)
    gp = torch.zeros(8, device=cuda); gp[:C] = gamma.to(cuda)
    bp = torch.zeros(8, device=cuda); bp[:C] = beta.to(cuda)
    check(lib.tta_norm_stats(ych.data_ptr(), ns, N, 1, V, batch_mode, 1e-5, mean.data_ptr(), rstd.data_ptr(),
                             ws.data_ptr(), 1, stream()))
    ych8 = ych
    if cpv == 4:
        ych = ych8[:, 0, ..., :4].contiguous()                  # [N][D][H][W][4]
        ns_y = V * 4
    else:
        ns_y = ns
    wh = pack_weights_small(wg_forward(w, False), 0)            # HOST [27][8][8]
    bias = pack_bias(b.to(cuda))
    logits = torch.zeros(N, C, *dims, device=cuda); dlog = torch.zeros_like(logits)
    lossd = torch.zeros(1, device=cuda)
    check(lib.tta_head_fused_fwd(ych.data_ptr(), ns_y, cpv, N, C, *dims, mean.data_ptr(), rstd.data_ptr(), gp.data_ptr(),
                                 bp.data_ptr(), 1, wh.data_ptr(), bias.data_ptr(), mode, 1.0 / (N * V), S, 0,
                                 logits.data_ptr(), dlog.data_ptr(), ws.data_ptr(), lossd.data_ptr(), stream()))
    assert rel_l2(logits.cpu(), z.detach()) < 3e-6
    assert abs(float(lossd) - float(loss.detach())) < 2e-6 * max(1.0, abs(float(loss.detach())))
    assert rel_l2(dlog.cpu() / S, z.grad) < 1e-5
    # inference variant: no dlogits, same logits; a second call also proves the block counter reset
    logits2 = torch.zeros_like(logits); loss2 = torch.zeros(1, device=cuda)
    check(lib.tta_head_fused_fwd(ych.data_ptr(), ns_y, cpv, N, C, *dims, mean.data_ptr(), rstd.data_ptr(), gp.data_ptr(),
                                 bp.data_ptr(), 1, wh.data_ptr(), bias.data_ptr(), mode, 1.0 / (N * V), S, 0,
                                 logits2.data_ptr(), 0, ws.data_ptr(), loss2.data_ptr(), stream()))
    assert torch.equal(logits2, logits) and torch.equal(loss2, lossd)
    # ---- backward: masked gradient w.r.t. the norm output + reductions
    dz = torch.full((N, 1, *dims, 8), 7.0, device=cuda)
    dz4 = torch.full((N, *dims, 4), 7.0, dtype=torch.float16, device=cuda)
    sums = torch.zeros(N * 8 * 2, device=cuda)
    dg = torch.zeros(8, device=cuda); db = torch.zeros(8, device=cuda)
    for _ in range(2):                               # twice: self-resetting counters
        check(lib.tta_head_fused_bwd(dlog.data_ptr(), N, C, *dims, wh.data_ptr(), ych.data_ptr(), ns_y, cpv,
                                     mean.data_ptr(), rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode,
                                     (dz if cpv == 8 else dz4).data_ptr(), ns if cpv == 8 else V * 4, sums.data_ptr(),
                                     dg.data_ptr(), db.data_ptr(), ws.data_ptr(), stream()))
    dz_ref = a.grad * (zn.detach() > 0)
    if cpv == 8:
        assert rel_l2(from_chunked(dz, C).cpu() / S, dz_ref) < 1e-5
        assert float(dz[..., C:].abs().max()) == 0.0     # pad channels exactly zero
    else:
        got = dz4.float().permute(0, 4, 1, 2, 3)[:, :C].cpu() / S
        assert rel_l2(got, dz_ref) < 5e-4                # one fp16 rounding
        assert C == 4 or float(dz4[..., C:].abs().max()) == 0.0
    assert rel_l2(dg[:C].cpu() / S, gr.grad) < 2e-5
    assert rel_l2(db[:C].cpu() / S, br.grad) < 2e-5
    # the norm-backward apply consumes dz as its gradient source: dy = autograd's dL/dy
    dhi = torch.zeros((N, 1, *dims, 8), dtype=torch.int16, device=cuda); dlo = torch.zeros_like(dhi)
    if cpv == 8:
        check(lib.tta_norm_bwd_apply(dz.data_ptr(), ns, 0, 0, ych.data_ptr(), ns, N, 1, V, mean.data_ptr(),
                                     rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), 1, batch_mode, sums.data_ptr(),
                                     dhi.data_ptr(), dlo.data_ptr(), ns, 0, 0, 0, TTA_BF16, 0, C, 0, 0, 0, stream()))
    else:
        check(lib.tta_norm_bwd_apply_c4(dz4.data_ptr(), V * 4, ych.data_ptr(), ns_y, N, C, V, mean.data_ptr(),
                                        rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), batch_mode, sums.data_ptr(),
                                        dhi.data_ptr(), dlo.data_ptr(), ns, TTA_BF16, 0, 0, stream()))
        dy = from_chunked(join_planes(dhi, dlo, TTA_BF16), C).cpu() / S
        assert rel_l2(dy, yr.grad) < 6e-4                # dz went through fp16 (11 bits) first
        assert float(join_planes(dhi, dlo, TTA_BF16)[..., C:].abs().max()) == 0.0
        # compact gradient planes [N][V][4] (what a GEOM_S2C4 dgrad reads): the same values, 8 B per voxel
        chi = torch.zeros((N, *dims, 4), dtype=torch.int16, device=cuda); clo = torch.zeros_like(chi)
        check(lib.tta_norm_bwd_apply_c4(dz4.data_ptr(), V * 4, ych.data_ptr(), ns_y, N, C, V, mean.data_ptr(),
                                        rstd.data_ptr(), gp.data_ptr(), bp.data_ptr(), batch_mode, sums.data_ptr(),
                                        chi.data_ptr(), clo.data_ptr(), V * 4, TTA_BF16, 0, 1, stream()))
        assert torch.equal(chi, dhi[:, 0, ..., :4]) and torch.equal(clo, dlo[:, 0, ..., :4])
        return
    dy = from_chunked(join_planes(dhi, dlo, TTA_BF16), C).cpu() / S
    assert rel_l2(dy, yr.grad) < 5e-5                # bf16x2 storage (~16 bits)


@pytest.mark.parametrize("cin,cout,K,s,tr,dims,dyt", [
    (32, 32, 3, 1, False, (5, 9, 11), TTA_F16_HI), (4, 64, 3, 2, False, (8, 12, 16), TTA_F16_HI),
    (64, 128, 3, 2, False, (4, 8, 8), TTA_BF16), (48, 16, 3, 2, True, (3, 6, 5), TTA_F16_HI),
    (64, 3, 3, 2, True, (4, 8, 8), TTA_BF16), (3, 3, 3, 1, False, (6, 10, 12), TTA_F16_HI),
    (256, 512, 1, 1, False, (2, 4, 4), TTA_F16_HI), (96, 40, 3, 1, False, (3, 5, 7), TTA_BF16),
    (4, 4, 3, 1, False, (5, 18, 20), TTA_BF16), (2, 3, 3, 1, False, (9, 17, 33), TTA_F16_HI)])
def test_conv_weight_and_bias_gradients(lib, cuda, cin, cout, K, s, tr, dims, dyt):
    """tta_conv_wgrad / tta_bias_grad (supervised step) against torch autograd on the same rounded
    operands: Conv3d and ConvTranspose3d, strides 1 / 2, 1x1, ragged tiles, channel counts that are not multiples of
    the channel tiles, scaled single-plane fp16 and bf16x2 gradients, w-parity-split operands, split outputs."""
    torch.manual_seed(13)
    N, pad, S = 2, (K - 1) // 2, 64.0
    x = torch.randn(N, cin, *dims)
    w = (torch.randn((cin, cout, K, K, K) if tr else (cout, cin, K, K, K)) * 0.1).requires_grad_(True)
    b = torch.zeros(cout, requires_grad=True)
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    y = F.conv_transpose3d(xv.cpu(), w, b, stride=s, padding=pad, output_padding=s - 1) if tr else \
        F.conv3d(xv.cpu(), w, b, stride=s, padding=pad)
    dy = torch.randn_like(y)
    odims = tuple(y.shape[2:])
    if dyt == TTA_F16_HI:
        dch = to_chunked((dy * S).to(cuda))
        dhi = dch.half().view(torch.int16).contiguous(); dlo = dhi
        dyv = from_chunked(dhi.view(torch.float16).float(), cout).cpu() / S
        scale = 1.0 / S
    else:
        dhi, dlo, dv = planes_from(dy.to(cuda), TTA_BF16)
        dyv, scale = dv.cpu(), 1.0
    y.backward(dyv)
    c8i, c8o = (cin + 7) // 8, (cout + 7) // 8
    Vi, Vo = dims[0] * dims[1] * dims[2], odims[0] * odims[1] * odims[2]
    mode = 1 if tr else 0
    for ws in ((False, False), (True, True)):
        xw, dw_ = ws[0] and dims[2] % 2 == 0, ws[1] and odims[2] % 2 == 0
        xh, xl = (wsplit(hi), wsplit(lo)) if xw else (hi, lo)
        gh, gl = (wsplit(dhi), wsplit(dlo)) if dw_ else (dhi, dlo)
        gw = torch.zeros_like(w.detach()).to(cuda); gb = torch.zeros(cout, device=cuda)
        check(lib.tta_conv_wgrad(xh.data_ptr(), xl.data_ptr(), c8i * Vi * 8, *dims, int(xw), gh.data_ptr(), gl.data_ptr(),
                                 c8o * Vo * 8, dyt, *odims, int(dw_), N, mode, K, s, cin, cout, scale, gw.data_ptr(),
                                 1 if tr else 0, 0, 0, stream()), "wgrad")
        check(lib.tta_bias_grad(dhi.data_ptr(), dlo.data_ptr(), c8o * Vo * 8, dyt, N, cout, Vo, scale, gb.data_ptr(), 0, 0,
                                stream()), "bias_grad")
        assert rel_l2(gw.cpu(), w.grad) < 2e-5, rel_l2(gw.cpu(), w.grad)
        assert rel_l2(gb.cpu(), b.grad) < 2e-5
    if not tr and cout % 16 == 0:
        # split output (fused unit0 || shortcut conv): the two halves land in two Conv3d weight gradients
        h = cout // 2
        ga = torch.zeros((h, cin, K, K, K), device=cuda); gb2 = torch.zeros((cout - h, cin, K, K, K), device=cuda)
        check(lib.tta_conv_wgrad(hi.data_ptr(), lo.data_ptr(), c8i * Vi * 8, *dims, 0, dhi.data_ptr(), dlo.data_ptr(),
                                 c8o * Vo * 8, dyt, *odims, 0, N, mode, K, s, cin, cout, scale, ga.data_ptr(), 0, h,
                                 gb2.data_ptr(), stream()), "wgrad")
        assert rel_l2(torch.cat([ga, gb2]).cpu(), w.grad) < 2e-5


@pytest.mark.parametrize("cin,cout,s,tr,dims,flags", [
    (32, 32, 1, False, (5, 9, 11), 0), (32, 32, 1, False, (6, 16, 16), 1), (32, 32, 1, False, (5, 9, 11), 2),
    (64, 128, 2, False, (4, 8, 8), 0), (128, 256, 2, False, (4, 16, 8), 0), (4, 64, 2, False, (8, 12, 16), 0),
    (4, 64, 2, False, (8, 12, 16), 2), (32, 128, 2, False, (6, 32, 16), 0), (32, 128, 2, False, (6, 32, 16), 2),
    (48, 16, 2, True, (3, 6, 5), 0), (48, 16, 2, True, (3, 6, 5), 2), (64, 3, 2, True, (4, 8, 8), 0),
    (256, 64, 2, True, (3, 4, 4), 0), (128, 32, 2, True, (4, 16, 8), 0), (768, 128, 2, True, (2, 4, 4), 0),
    (3, 8, 1, False, (6, 10, 12), 0), (8, 3, 1, False, (6, 20, 12), 2), (24, 200, 1, False, (3, 5, 7), 0),
    (160, 40, 1, False, (3, 5, 7), 0), (64, 64, 1, False, (3, 18, 9), 0), (512, 512, 1, False, (2, 4, 4), 0),
    (128, 128, 1, False, (4, 20, 24), 0), (40, 24, 2, True, (2, 3, 4), 0)])
def test_conv_weight_gradients_tensor_core(lib, cuda, cin, cout, s, tr, dims, flags):
    """tta_conv_wgrad_tc (tcgen05, MN-major operands, M = 64 / 128 row tiles, N = 8 / 16 / 32 column tiles, parity
    sub-tiles of the finer tensor for stride 2) against torch autograd on the same rounded operands (x = fp16 hi + lo,
    dy = one scaled fp16 plane); flags = 1 uses the hi plane of x only and is compared on hi-rounded operands."""
    torch.manual_seed(17)
    N, K, pad, S = 2, 3, 1, 64.0
    x = torch.randn(N, cin, *dims)
    w = (torch.randn((cin, cout, K, K, K) if tr else (cout, cin, K, K, K)) * 0.1).requires_grad_(True)
    hi, lo, xv = planes_from(x.to(cuda), TTA_F16)
    if flags & 1:
        xv = from_chunked(hi.view(torch.float16).float(), cin)
    y = F.conv_transpose3d(xv.cpu(), w, None, stride=s, padding=pad, output_padding=s - 1) if tr else \
        F.conv3d(xv.cpu(), w, None, stride=s, padding=pad)
    dy = torch.randn_like(y)
    odims = tuple(y.shape[2:])
    dch = to_chunked((dy * S).to(cuda))
    dhi = dch.half().view(torch.int16).contiguous()
    dyv = from_chunked(dhi.view(torch.float16).float(), cout).cpu() / S
    y.backward(dyv)
    c8i, c8o = (cin + 7) // 8, (cout + 7) // 8
    Vi, Vo = dims[0] * dims[1] * dims[2], odims[0] * odims[1] * odims[2]
    mode = 1 if tr else 0
    assert lib.tta_conv_wgrad_tc_supported(mode, K, s, cin, cout, TTA_F16_HI) == 1
    xw, dw_ = (s == 2 and not tr), (s == 2 and tr)
    xh, xl = (wsplit(hi), wsplit(lo)) if xw else (hi, lo)
    gh = wsplit(dhi) if dw_ else dhi
    gw = torch.zeros_like(w.detach()).to(cuda)
    check(lib.tta_conv_wgrad_tc(xh.data_ptr(), xl.data_ptr(), c8i * Vi * 8, *dims, int(xw), gh.data_ptr(), c8o * Vo * 8,
                                *odims, int(dw_), N, mode, s, cin, cout, 1.0 / S, gw.data_ptr(), 1 if tr else 0, 0, 0,
                                flags, stream()), "wgrad_tc")
    torch.cuda.synchronize()
    assert rel_l2(gw.cpu(), w.grad) < 3e-5, rel_l2(gw.cpu(), w.grad)
    if not tr and cout % 16 == 0:
        h = cout // 2
        ga = torch.zeros((h, cin, K, K, K), device=cuda); gb2 = torch.zeros((cout - h, cin, K, K, K), device=cuda)
        check(lib.tta_conv_wgrad_tc(xh.data_ptr(), xl.data_ptr(), c8i * Vi * 8, *dims, int(xw), gh.data_ptr(),
                                    c8o * Vo * 8, *odims, int(dw_), N, mode, s, cin, cout, 1.0 / S, ga.data_ptr(), 0, h,
                                    gb2.data_ptr(), flags, stream()), "wgrad_tc")
        assert rel_l2(torch.cat([ga, gb2]).cpu(), w.grad) < 3e-5
