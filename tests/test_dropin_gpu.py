"""The drop-in call sites executed end to end on the GPU:

* ``TTASegmentationEvaluationStrategy.evaluate_epoch(model, loader, device)`` -- the ONLY place
  ``main.py method=tent_b200`` reaches the kernels from (/root/reference/src/core/experiment_manager.py:364-370,
  /root/reference/src/evaluation/seg_eval.py:238-479) -- against the CPU oracle run over the same loader and
  scored with oracle.dice_oracle (pinned to the reference's ``_binary_dice_iou``), full-volume and
  sliding-window modes, with the reference's ``domain`` batch formats;
* ``TentB200.run_step(batch)`` (trainer-step signature, seg_trainer.py:97-145);
* ``method.checkpoint`` (CheckpointHook format, hooks.py:53-70);
* the engine-state hazards a live trainer can hit: new weights after a plan was built, ``step`` and
  ``step_windows`` alternating on one plan, padded sliding-window tail batches on a single rank.
"""
import copy
from collections import defaultdict

import pytest
import torch

from multimodal_tta_b200 import SlidingWindowTTA, TentB200, UNetB200, create, get_evaluation_strategy
from multimodal_tta_b200.presets import BRATS_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, region_labels
from oracle.dice_oracle import evaluate_logits
from oracle.sliding_window_oracle import sliding_window_oracle
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from tests.util import make_pair, rel_l2

pytestmark = pytest.mark.gpu

DIMS = (32, 32, 32)


def _loader(domain_kind):
    """Four batches (B = 2, 2, 2, 1) with the `domain` entry in each of the formats the reference's
    _as_list_str accepts (seg_eval.py:20-38)."""
    batches = []
    for i, B in enumerate((2, 2, 2, 1)):
        dom = {"list": [f"site{(i + j) % 2}" for j in range(B)],
               "str": "siteA",
               "tensor": torch.tensor([(i + j) % 2 for j in range(B)]),
               "scalar": torch.tensor(i % 2),
               "none": None}[domain_kind]
        b = {"image": brats_volume(B, DIMS, seed=500 + i), "label": region_labels(B, 3, DIMS, seed=600 + i),
             "case_id": [f"c{i}_{j}" for j in range(B)], "index": torch.arange(B)}
        if dom is not None:
            b["domain"] = dom
        batches.append(b)
    return batches


def _domains_of(batch):
    d, B = batch.get("domain"), batch["image"].shape[0]
    if d is None:
        return [""] * B
    if isinstance(d, list):
        return [str(v) for v in d]
    if isinstance(d, str):
        return [d] * B
    return [str(int(d))] * B if d.ndim == 0 else [str(int(v)) for v in d]


def _oracle_metrics(to, batches, predict):
    """Reference aggregation (seg_eval.py:363-479) on oracle logits: overall + per-domain keys."""
    logits = [predict(b["image"]) for b in batches]
    labels = [b["label"] for b in batches]
    ref = evaluate_logits(logits, labels)
    per_dom = defaultdict(lambda: ([], []))
    for lg, b in zip(logits, batches):
        for i, dom in enumerate(_domains_of(b)):
            per_dom[dom][0].append(lg[i:i + 1]); per_dom[dom][1].append(b["label"][i:i + 1])
    for dom, (lgs, ys) in per_dom.items():
        m = evaluate_logits(lgs, ys)
        safe = dom if dom != "" else "unknown"
        for k in ("et_dc", "tc_dc", "wt_dc", "avg_dc", "miou"):
            ref[f"dom/{safe}/{k}"] = m[k]
    return ref


@pytest.mark.parametrize("domain_kind", ["list", "str", "tensor", "scalar", "none"])
def test_evaluate_epoch_matches_oracle(cuda, domain_kind):
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=61, device="cpu")   # manager moves it
    to = TentOracle(oracle, mode="sigmoid")
    batches = _loader(domain_kind)
    losses = []

    def predict(x):
        lg, loss = to.step(x)
        losses.append(loss * x.shape[0])
        return lg
    ref = _oracle_metrics(to, batches, predict)
    cfg = create({"evaluation": {"seg": {"threshold": 0.5, "region_order": ["ET", "TC", "WT"]}},
                  "method": {"name": "tent_b200", "entropy": "sigmoid", "lr": 1e-3}})
    strat = get_evaluation_strategy("tta_seg_eval")(cfg)
    got = strat.evaluate_epoch(prod, batches, cuda)
    assert all(isinstance(v, float) for v in got.values())
    assert set(ref) <= set(got), sorted(set(ref) - set(got))
    assert {k for k in got if k.startswith("dom/")} == {k for k in ref if k.startswith("dom/")}
    for k, v in ref.items():
        assert abs(got[k] - v) < 1e-3, (k, got[k], v)                   # north star: Dice identical to 1e-3
    assert abs(got["loss"] - sum(losses) / 7) < 1e-4
    assert got["jc"] == got["miou"]
    # the strategy adapted the LIVE model (non-episodic): parameters moved exactly as the oracle's
    perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
    assert float(perr.median()) < 1e-4
    assert strat.is_best_model({"avg_dc": 0.5}, {"avg_dc": 0.4}) and not strat.is_best_model({"avg_dc": 0.3}, {"avg_dc": 0.4})


def test_evaluate_epoch_sliding_window_mode(cuda):
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=62)
    to = TentOracle(oracle, mode="sigmoid")
    dims = (40, 48, 32)
    batches = [{"image": brats_volume(1, dims, seed=700 + i), "label": region_labels(1, 3, dims, seed=710 + i),
                "domain": ["d0"]} for i in range(2)]
    ref = _oracle_metrics(to, batches, lambda x: sliding_window_oracle(x, (32, 32, 32), 2, lambda w: to.step(w)[0],
                                                                       overlap=0.5))
    cfg = create({"evaluation": {"seg": {"threshold": 0.5}},
                  "method": {"sliding_window": {"enabled": True, "roi": [32, 32, 32], "sw_batch": 2, "overlap": 0.5}}})
    got = get_evaluation_strategy("tta_seg_eval")(cfg).evaluate_epoch(prod, batches, cuda)
    for k, v in ref.items():
        assert abs(got[k] - v) < 1e-3, (k, got[k], v)


def test_evaluate_epoch_rejects_bad_batches(cuda):
    _, prod = make_pair(BRATS_MODEL_CFG, seed=63)
    strat = get_evaluation_strategy("tta_seg_eval")(create({}))
    x = brats_volume(1, DIMS, seed=1)
    with pytest.raises(KeyError):
        strat.evaluate_epoch(prod, [{"image": x}], cuda)
    with pytest.raises(ValueError):
        strat.evaluate_epoch(prod, [{"image": x, "label": torch.zeros(1, 2, *DIMS)}], cuda)       # 2 regions != 3
    with pytest.raises(ValueError):
        strat.evaluate_epoch(prod, [{"image": x, "label": torch.zeros(*DIMS)}], cuda)             # 3-D label
    with pytest.raises(TypeError):
        strat.evaluate_epoch(torch.nn.Conv3d(4, 3, 3), [{"image": x, "label": torch.zeros(1, 3, *DIMS)}], cuda)
    # [R,D,H,W] labels are broadcast over the batch like the reference does (seg_eval.py:287-288)
    out = strat.evaluate_epoch(prod, [{"image": x, "label": region_labels(1, 3, DIMS, seed=2)[0]}], cuda)
    assert 0.0 <= out["avg_dc"] <= 1.0


def test_run_step_trainer_signature(cuda):
    oracle, prod = make_pair(BRATS_MODEL_CFG, seed=64)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    for s in (1, 2):
        batch = {"image": brats_volume(2, DIMS, seed=s), "label": region_labels(2, 3, DIMS, seed=s)}
        out = tp.run_step(batch)                                        # host batch: moved to the model's device
        _, loss = to.step(batch["image"])
        assert set(out) == {"loss"} and isinstance(out["loss"], float)
        assert abs(out["loss"] - loss) < 1e-4 * max(1.0, abs(loss))


def test_method_checkpoint_is_loaded_before_adaptation(cuda, tmp_path):
    """best_model.pth as the reference's CheckpointHook writes it (hooks.py:53-70), DataParallel prefix."""
    torch.manual_seed(65)
    from oracle.unet_oracle import OracleUNet
    source = OracleUNet.from_cfg(BRATS_MODEL_CFG)
    path = tmp_path / "best_model.pth"
    torch.save({"epoch": 3, "model_state_dict": {"module." + k: v for k, v in source.state_dict().items()},
                "optimizer_state_dict": {}, "best_metrics": {"avg_dc": 0.5}}, path)
    torch.manual_seed(66)
    prod = UNetB200(dict(BRATS_MODEL_CFG))                               # different random weights
    batches = _loader("list")[:2]
    to = TentOracle(source, mode="sigmoid")
    ref = _oracle_metrics(to, batches, lambda x: to.step(x)[0])
    cfg = create({"method": {"checkpoint": str(path)}})
    got = get_evaluation_strategy("tta_seg_eval")(cfg).evaluate_epoch(prod, batches, cuda)
    for k, v in ref.items():
        assert abs(got[k] - v) < 1e-3, (k, got[k], v)
    with pytest.raises(FileNotFoundError):
        get_evaluation_strategy("tta_seg_eval")(create({"method": {"checkpoint": str(tmp_path / "nope.pth")}})) \
            .evaluate_epoch(UNetB200(dict(BRATS_MODEL_CFG)), batches, cuda)


def test_new_weights_after_a_plan_was_built(cuda):
    """Packed conv weights, host kernel parameters and captured graphs all belong to the weights they were
    built from: loading another checkpoint must rebuild them (forward, load_state_dict, forward)."""
    oa, prod = make_pair(BRATS_MODEL_CFG, seed=71)
    torch.manual_seed(72)
    from oracle.unet_oracle import OracleUNet
    ob = OracleUNet.from_cfg(BRATS_MODEL_CFG)
    x = brats_volume(1, DIMS, seed=3)
    tp = TentB200(prod, {"cuda_graph": True})
    la = tp.step(x.cuda()).cpu()
    assert rel_l2(la, TentOracle(oa, mode="sigmoid").step(x)[0]) < 1e-3
    n_plans = len(prod.engine.plans)
    prod.load_state_dict(copy.deepcopy(ob.state_dict()), strict=False)   # norm affines keep their adapted values
    prod.to(cuda)
    tp2 = TentB200(prod, {"cuda_graph": True})
    tp2.reset()
    for nh in prod.norm_holders():                                       # source gamma/beta again
        nh.weight.data.fill_(1.0); nh.bias.data.zero_()
    lb = tp2.step(x.cuda()).cpu()
    ref_b = TentOracle(ob, mode="sigmoid").step(x)[0]
    assert rel_l2(lb, ref_b) < 1e-3, rel_l2(lb, ref_b)
    assert rel_l2(lb, la) > 1e-2                                         # really different weights
    assert n_plans == 1
    # an unchanged model keeps its plans (no repack on .to(same device))
    plan = next(iter(prod.engine.plans.values()))
    prod.to(cuda)
    tp2.step(x.cuda())
    assert next(iter(prod.engine.plans.values())) is plan


def test_step_and_step_windows_do_not_share_a_graph(cuda):
    """A batch whose shape equals (sw_batch, roi) after a window sweep must not replay the window-gather graph."""
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=73)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    vol = brats_volume(1, (32, 48, 32), seed=4)
    sw = SlidingWindowTTA(tp, (32, 32, 32), sw_batch=2, overlap=0.5)
    ref_sw = sliding_window_oracle(vol, (32, 32, 32), 2, lambda w: to.step(w)[0], overlap=0.5)
    assert rel_l2(sw(vol.cuda()).cpu(), ref_sw) < 1e-3
    x = brats_volume(2, (32, 32, 32), seed=5)                            # same plan key (2, 32, 32, 32)
    got = tp.step(x.cuda()).cpu()
    assert rel_l2(got, to.step(x)[0]) < 1e-3
    vol2 = brats_volume(1, (32, 48, 32), seed=6)
    ref_sw2 = sliding_window_oracle(vol2, (32, 32, 32), 2, lambda w: to.step(w)[0], overlap=0.5)
    assert rel_l2(sw(vol2.cuda()).cpu(), ref_sw2) < 1e-3


@pytest.mark.parametrize("use_graph", [True, False])
def test_single_rank_tail_batch_is_the_mean_over_real_windows(cuda, use_graph):
    """3 windows with sw_batch = 2: the second step holds one real and one zero-weight padding window; loss,
    gradient and update must equal the oracle's step on the single real window."""
    oracle, prod = make_pair(dict(BRATS_MODEL_CFG, deterministic=True), seed=74)
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": use_graph})
    vol = brats_volume(1, (64, 32, 32), seed=8)
    losses = []

    def pred(w):
        lg, loss = to.step(w)
        losses.append(loss)
        return lg
    ref = sliding_window_oracle(vol, (32, 32, 32), 2, pred, overlap=0.5)
    sw = SlidingWindowTTA(tp, (32, 32, 32), sw_batch=2, overlap=0.5)
    got = sw(vol.cuda()).cpu()
    assert sw.last_num_windows == 3 and sw.last_steps == 2
    assert rel_l2(got, ref) < 1e-3
    assert abs(float(tp.last_loss) - losses[-1]) < 1e-4 * max(1.0, abs(losses[-1]))
    g_o, g_p = to.last_grads, prod.engine.flat_grads().cpu() * 2.0       # flat_grads excludes the tail multiplier
    assert rel_l2(g_p, g_o) < 3e-3
    perr = (prod.engine.flat_params().cpu() - flat_gamma_beta(to.model)).abs()
    assert float(perr.median()) < 1e-4 and float((perr > 1e-4).float().mean()) < 0.05
