"""Host-side logic that needs no GPU: registry/config shims behave like the reference's, the model
holder tree reproduces MONAI's state-dict keys and torch's default init stream, the op graph has
the layer counts of SURVEY.md section 8a, tiling equals the oracle's, weight packers implement the
canonical gather semantics, and the product fails loudly without CUDA / without its library."""
import copy
import json
import os

import pytest
import torch
import torch.nn.functional as F

import multimodal_tta_b200 as pkg
from multimodal_tta_b200 import TentB200, UNetB200, _lib, compose_yaml, create, get_config, require_config
from multimodal_tta_b200.layout import (from_chunked, join_planes, pack_weights_simt, pack_weights_small, split_planes,
                                        tc_groups, to_chunked, wg_dgrad, wg_forward)
from multimodal_tta_b200.registry import Registry
from multimodal_tta_b200.sliding_window import gaussian_factors, plan_windows, scan_interval, shard_schedule, window_starts
from oracle import sliding_window_oracle as swo
from oracle.unet_oracle import BARE_DEFAULT_MODEL_CFG, BRATS_MODEL_CFG, HECKTOR_MODEL_CFG, OracleUNet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_registry_matches_reference_behaviour(capsys):
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "registry_golden.json")))
    r = Registry("models")
    r.register("a", int)
    r.register("a", float)
    assert capsys.readouterr().out == gold["dup_warning"]
    assert r.get("a").__name__ == gold["dup_result"] and r.list_all() == gold["list_all"]
    assert r.has("a") == gold["has_a"] and r.has("b") == gold["has_b"]
    with pytest.raises(KeyError) as e:
        r.get("zzz")
    assert str(e.value) == gold["missing_keyerror"]


def test_plugins_are_registered_under_reference_names():
    assert pkg.get_model("unet_b200") is UNetB200
    assert pkg.get_plugin("tent_b200") is TentB200
    assert pkg.get_evaluation_strategy("tta_seg_eval").__name__ == "TTASegmentationEvaluationStrategy"
    with pytest.raises(KeyError):
        pkg.get_model("unet")      # the reference's own class is not shadowed


def test_config_access_semantics():
    cfg = create({"a": {"b": 3, "c": None}, "l": [1, 2]})
    assert get_config(cfg, "a.b") == 3 and get_config(cfg, "a.c", 7) == 7 and get_config(cfg, "x.y", "d") == "d"
    assert require_config(cfg, "a.b", int) == 3
    with pytest.raises(ValueError):
        require_config(cfg, "a.c")
    with pytest.raises(TypeError):
        get_config({"a": 1}, "a")              # plain dict is rejected, like the reference
    with pytest.raises(TypeError):
        require_config(cfg, "a.b", str)


def test_method_yaml_composes_like_a_global_patch():
    cfg = compose_yaml(os.path.join(ROOT, "configs", "model", "unet_b200.yaml"),
                       os.path.join(ROOT, "configs", "method", "tent_b200.yaml"),
                       overrides={"model": dict(BRATS_MODEL_CFG)})
    assert cfg.model.name == "unet_b200" and cfg.task.eval_strategy == "tta_seg_eval"
    assert cfg.method.name == "tent_b200" and cfg.method.lr == 1e-3 and cfg.model.in_channels == 4
    model = pkg.get_model(cfg.model.name)(cfg.model)
    assert isinstance(model, UNetB200) and model.num_res_units == 2


@pytest.mark.parametrize("cfg", [BRATS_MODEL_CFG, HECKTOR_MODEL_CFG, BARE_DEFAULT_MODEL_CFG])
def test_state_dict_keys_and_default_init_equal_the_oracle(cfg):
    torch.manual_seed(0)
    m = UNetB200(dict(cfg))
    torch.manual_seed(0)
    o = OracleUNet.from_cfg(cfg)
    sm, so = m.state_dict(), o.state_dict()
    assert list(sm.keys()) == list(so.keys())
    assert all(torch.equal(sm[k], so[k]) for k in sm)


def test_load_reference_checkpoint_variants():
    o = OracleUNet.from_cfg(BRATS_MODEL_CFG)
    m = UNetB200(dict(BRATS_MODEL_CFG))
    m.load_state_dict({"module." + k: v for k, v in o.state_dict().items()})     # DataParallel prefix
    assert torch.equal(m.state_dict()["model.2.0.conv.weight"], o.state_dict()["model.2.0.conv.weight"])
    sd = copy.deepcopy(o.state_dict())
    sd["model.0.conv.unit0.adn.N.weight"] = torch.full((32,), 2.0)
    sd["model.0.conv.unit0.adn.N.bias"] = torch.full((32,), 0.5)
    m2 = UNetB200(dict(BRATS_MODEL_CFG))
    m2.load_state_dict(sd, strict=False)           # affine keys materialise gamma/beta
    assert float(m2.model[0].conv.unit0.adn.N.weight.detach()[0]) == 2.0
    with pytest.raises(RuntimeError):
        UNetB200(dict(BRATS_MODEL_CFG)).load_state_dict({"bogus": torch.zeros(1)})


def test_config_validation_errors():
    with pytest.raises(ValueError):
        UNetB200(dict(BRATS_MODEL_CFG, in_channels="auto"))
    with pytest.raises(ValueError):
        UNetB200(dict(BRATS_MODEL_CFG, norm="GROUP"))
    with pytest.raises(ValueError):
        UNetB200(dict(BRATS_MODEL_CFG, strides=[2, 2, 3, 2]))
    with pytest.raises(ValueError):
        UNetB200(dict(BRATS_MODEL_CFG, channels=[32]))
    with pytest.raises(ValueError):
        TentB200(UNetB200(dict(HECKTOR_MODEL_CFG)), {"entropy": "softmax"})   # degenerate for 1 channel
    with pytest.raises(TypeError):
        TentB200(OracleUNet.from_cfg(BRATS_MODEL_CFG))


def test_tent_configure_selects_only_norm_affine():
    m = UNetB200(dict(BRATS_MODEL_CFG))
    assert "model.0.conv.unit0.adn.N.weight" not in m.state_dict()       # InstanceNorm3d(affine=False)
    t = TentB200(m, {"cuda_graph": False})
    ps = t.adaptable_parameters()
    assert len(ps) == 34 and sum(p.numel() for p in ps) == 4870 == m.engine.n_adaptable
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 4870
    assert m.training and "model.0.conv.unit0.adn.N.weight" in m.state_dict()


@pytest.mark.parametrize("cfg,nconv,nnorm,ndgrad", [(dict(BRATS_MODEL_CFG, fuse_shortcut=False), 23, 17, 21),
                                                    (BRATS_MODEL_CFG, 19, 17, 18),   # 4 unit0||shortcut pairs fused
                                                    (dict(BARE_DEFAULT_MODEL_CFG, in_channels=4), 9, 8, 8)])
def test_op_graph_layer_counts(cfg, nconv, nnorm, ndgrad):
    m = UNetB200(dict(cfg))
    TentB200(m, {"cuda_graph": False})
    eng = m.engine
    eng._ensure_device(torch.device("cpu"), dry=True)
    plan = eng.build_plan(1, 32, 32, 32)
    # res-unit UNets end in norm -> 3x3x3 conv(R->R) -> entropy: that tail runs as the fused head
    # (norm apply + conv + loss finalize leave the forward list, dgrad + norm reduce merge)
    fh = int(plan.fused_head)
    assert fh == (cfg is not BARE_DEFAULT_MODEL_CFG and "num_res_units" in cfg and cfg["num_res_units"] > 0)
    assert len(plan.fwd) == nconv + nnorm - fh and len(plan.bwd) == ndgrad + nnorm
    # norms fed by a tcgen05 conv without split-K take their statistics from the conv epilogue
    assert 0 < plan.n_fused_stats <= nnorm
    assert plan.n_fused_bwd == 0         # opt-in (fuse_bwd_stats): measured slower than the streaming pass
    # layers with <= 4096 voxels per instance run statistics + apply as one launch (InstanceNorm only)
    assert plan.launches_fwd == 1 + nconv + 2 * nnorm + 2 - 3 * fh - plan.n_small_fwd
    assert set(plan.conv_backends.values()) <= {"tc", "small", "simt", "head"}
    if cfg is BRATS_MODEL_CFG:
        assert plan.conv_backends["model.2.1.conv.unit0.conv:fwd"] == "head"
        assert plan.conv_backends["model.0.conv.unit0.conv||residual:fwd"] == "tc"
    with pytest.raises(ValueError):
        eng.build_plan(1, 24, 32, 32)        # not divisible by 16


def test_product_fails_loudly_without_cuda_or_library(monkeypatch, tmp_path):
    m = UNetB200(dict(HECKTOR_MODEL_CFG))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 2, 16, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        TentB200(m, {}).step(torch.zeros(1, 2, 16, 16, 16))
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_no_product_module_imports_the_oracle():
    for fn in os.listdir(os.path.join(ROOT, "multimodal_tta_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "multimodal_tta_b200", fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_window_tiling_equals_oracle():
    for dims, roi, ov in [((155, 240, 240), (128, 128, 128), 0.5), ((40, 48, 36), (32, 32, 32), 0.5),
                          ((10, 50, 33), (16, 32, 32), 0.25), ((144, 144, 144), (96, 96, 96), 0.5)]:
        padded, pad_lo, starts = plan_windows(dims, roi, ov)
        psz = [max(a, b) for a, b in zip(dims, roi)]
        assert padded == psz
        assert starts == swo.window_starts(psz, roi, swo.scan_interval(psz, roi, ov))
        fs, wmin = gaussian_factors(roi)
        imp = (fs[0].view(-1, 1, 1) * fs[1].view(1, -1, 1) * fs[2].view(1, 1, -1)).clamp_(min=wmin)
        assert torch.allclose(imp, swo.gaussian_importance(roi), rtol=1e-6)


def test_shard_schedule_partitions_windows_exactly_once():
    for total, nb, world in [(18, 1, 4), (18, 2, 8), (48, 3, 8), (5, 2, 2), (7, 1, 1)]:
        per_rank = [list(shard_schedule(total, nb, world, r)) for r in range(world)]
        assert len({len(p) for p in per_rank}) == 1                       # every rank joins every all-reduce
        seen = [i for p in per_rank for idxs, _ in p for i in idxs if i is not None]
        assert sorted(seen) == list(range(total))
        for t in range(len(per_rank[0])):
            nv = per_rank[0][t][1]
            assert all(p[t][1] == nv for p in per_rank)
            assert nv == sum(1 for p in per_rank for i in p[t][0] if i is not None)


def test_layout_round_trips():
    torch.manual_seed(0)
    x = torch.randn(2, 11, 3, 4, 5)
    ch = to_chunked(x)
    assert ch.shape == (2, 2, 3, 4, 5, 8) and float(ch[:, 1, ..., 3:].abs().max()) == 0.0
    assert torch.equal(from_chunked(ch, 11), x)
    for tag, tol in [(0, 2 ** -21), (1, 2 ** -15), (2, 2 ** -10)]:
        hi, lo = split_planes(x, tag)
        assert float((join_planes(hi, lo, tag) - x).abs().max() / x.abs().max()) < tol


def _gather_conv(x, wg, mode, K, s):
    """Naive restatement of the canonical gather semantics (csrc/tta_conv_simt.cu header)."""
    N, ci, D, H, W = x.shape
    co, p = wg.shape[2], (K - 1) // 2
    od = ((D - 1) // s + 1, (H - 1) // s + 1, (W - 1) // s + 1) if mode == 0 else (D * s, H * s, W * s)
    out = torch.zeros(N, co, *od)
    for kd in range(K):
        for kh in range(K):
            for kw in range(K):
                w = wg[(kd * K + kh) * K + kw]              # [ci][co]
                for d in range(od[0]):
                    for h in range(od[1]):
                        for ww in range(od[2]):
                            if mode == 0:
                                i = (s * d - p + kd, s * h - p + kh, s * ww - p + kw)
                            else:
                                t = (d + p - kd, h + p - kh, ww + p - kw)
                                if any(v % s for v in t):
                                    continue
                                i = tuple(v // s for v in t)
                            if all(0 <= i[a] < (D, H, W)[a] for a in range(3)):
                                out[:, :, d, h, ww] += x[:, :, i[0], i[1], i[2]] @ w
    return out


@pytest.mark.parametrize("tr,s", [(False, 1), (False, 2), (True, 2)])
def test_weight_conventions_forward_and_dgrad(tr, s):
    torch.manual_seed(1)
    ci, co, K = 3, 2, 3
    x = torch.randn(1, ci, 4, 4, 4, requires_grad=True)
    w = torch.randn((ci, co, K, K, K) if tr else (co, ci, K, K, K))
    ref = F.conv_transpose3d(x, w, None, stride=s, padding=1, output_padding=s - 1) if tr else \
        F.conv3d(x, w, None, stride=s, padding=1)
    got = _gather_conv(x.detach(), wg_forward(w, tr), 1 if tr else 0, K, s)
    assert torch.allclose(got, ref.detach(), atol=1e-5)
    dy = torch.randn_like(ref)
    (g,) = torch.autograd.grad(ref, x, dy)
    gg = _gather_conv(dy, wg_dgrad(w, tr), 0 if tr else 1, K, s)
    assert torch.allclose(gg, g, atol=1e-5)
    wp = pack_weights_simt(wg_forward(w, tr))
    assert wp.shape == (27, 1, 1, 8, 8) and float(wp[:, 0, 0, ci:, :].abs().max()) == 0.0


def test_small_and_tc_pack_tables():
    wg = torch.arange(27 * 3 * 3, dtype=torch.float32).reshape(27, 3, 3)
    assert torch.equal(pack_weights_small(wg, 1)[:, :3, :3], wg.flip(0)) and pack_weights_small(wg, 0).device.type == "cpu"
    assert [len(g) for g in tc_groups(0, 3, 1)] == [9, 9, 9] and [len(g) for g in tc_groups(1, 3, 2)] == [18, 9]
    assert tc_groups(1, 3, 2)[1] == list(range(9)) and tc_groups(0, 1, 1) == [[0]]
    assert sorted(sum(tc_groups(1, 3, 2), [])) == list(range(27))


def test_domain_entry_formats_follow_the_reference():
    """seg_eval.py:20-38: None -> "", list/tuple -> str per item, str -> repeated, 0-d tensor -> int for the batch,
    B-element tensor -> one int per sample, anything else -> str(x) repeated."""
    from multimodal_tta_b200.evaluation import _as_list_str
    assert _as_list_str(None, 2) == ["", ""]
    assert _as_list_str(["a", 3], 2) == ["a", "3"] and _as_list_str(("x", "y"), 2) == ["x", "y"]
    assert _as_list_str("site", 3) == ["site"] * 3
    assert _as_list_str(torch.tensor(4), 2) == ["4", "4"]
    assert _as_list_str(torch.tensor([0, 1, 1]), 3) == ["0", "1", "1"]
    assert _as_list_str(torch.tensor([[2.0], [5.0]]), 2) == ["2", "5"]
    t = torch.tensor([0, 1, 2])
    assert _as_list_str(t, 2) == [str(t)] * 2              # wrong length: falls through, like the reference
    assert _as_list_str(7, 2) == ["7", "7"]


def test_presets_equal_the_oracle_configs():
    from multimodal_tta_b200 import presets
    assert presets.BRATS_MODEL_CFG == BRATS_MODEL_CFG and presets.HECKTOR_MODEL_CFG == HECKTOR_MODEL_CFG
    assert presets.BARE_DEFAULT_MODEL_CFG == BARE_DEFAULT_MODEL_CFG


def test_load_source_checkpoint_formats(tmp_path):
    """CheckpointHook format (hooks.py:53-70), with and without the DataParallel prefix, and a bare state dict."""
    from multimodal_tta_b200.evaluation import load_source_checkpoint
    torch.manual_seed(3)
    o = OracleUNet.from_cfg(BRATS_MODEL_CFG)
    k = "model.1.submodule.0.conv.unit1.conv.weight"
    for name, payload in (("hook.pth", {"epoch": 1, "model_state_dict": o.state_dict(), "optimizer_state_dict": {},
                                        "best_metrics": {"avg_dc": 0.1}}),
                          ("dp.pth", {"epoch": 1, "model_state_dict": {"module." + n: v for n, v in o.state_dict().items()}}),
                          ("bare.pth", o.state_dict())):
        torch.save(payload, tmp_path / name)
        m = UNetB200(dict(BRATS_MODEL_CFG))
        assert not torch.equal(m.state_dict()[k], o.state_dict()[k])
        load_source_checkpoint(m, str(tmp_path / name))
        assert all(torch.equal(m.state_dict()[n], v) for n, v in o.state_dict().items())
    with pytest.raises(FileNotFoundError):
        load_source_checkpoint(UNetB200(dict(BRATS_MODEL_CFG)), str(tmp_path / "missing.pth"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present (GPU box)")
def test_registration_lands_in_the_reference_registry():
    """Inside the reference checkout (``src.registry`` importable) the decorators must fill the reference's OWN
    registries -- the maps ExperimentManager.setup_model / setup_evaluation read
    (/root/reference/src/registry.py:60-66, src/core/experiment_manager.py:88-96,364-370)."""
    import subprocess
    import sys
    code = (
        "import multimodal_tta_b200 as pkg, multimodal_tta_b200.registry as r\n"
        "from src.registry import MODELS, EVALUATION_STRATEGIES, PLUGINS, get_model, get_evaluation_strategy\n"
        "assert r.USING_REFERENCE_REGISTRY and r.MODELS is MODELS and r.EVALUATION_STRATEGIES is EVALUATION_STRATEGIES\n"
        "assert MODELS.has('unet_b200') and EVALUATION_STRATEGIES.has('tta_seg_eval') and PLUGINS.has('tent_b200')\n"
        "assert get_model('unet_b200') is pkg.UNetB200\n"
        "assert get_evaluation_strategy('tta_seg_eval') is pkg.TTASegmentationEvaluationStrategy\n"
        "m = get_model('unet_b200')(pkg.create(dict(in_channels=4, num_classes=3, num_res_units=2, norm='INSTANCE')))\n"
        "assert sum(p.numel() for p in m.parameters()) == 19223961\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(["/root/reference", ROOT]))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr


def test_multimodal_model_keys_counts_and_plan():
    """holder tree of MultimodalUNetDeepFusion reproduces the oracle's (= the reference's) state-dict
    keys and shapes; every InstanceNorm becomes adaptable; the op graph builds without a GPU."""
    from multimodal_tta_b200 import MultimodalUNetB200
    from oracle.multimodal_oracle import MULTIMODAL_MODEL_CFG, OracleMultimodalUNet
    torch.manual_seed(0)
    o = OracleMultimodalUNet.from_cfg(MULTIMODAL_MODEL_CFG)
    m = MultimodalUNetB200(dict(MULTIMODAL_MODEL_CFG))
    so, sm = o.state_dict(), m.state_dict()
    assert list(so.keys()) == list(sm.keys()) and all(so[k].shape == sm[k].shape for k in so)
    assert sum(p.numel() for p in m.parameters()) == sum(p.numel() for p in o.parameters()) == 83075815
    assert "bottleneck_reduce.bias" not in sm and "decoder_stages.0.upsample.preconv.bias" in sm
    m.load_state_dict({"module." + k: v for k, v in so.items()})
    t = TentB200(m, {"cuda_graph": False})
    assert len(t.adaptable_parameters()) == 98 and m.engine.n_adaptable == 18816
    eng = m.engine
    eng._ensure_device(torch.device("cpu"), dry=True)
    plan = eng.build_plan(1, 32, 32, 32)
    assert set(plan.conv_backends.values()) == {"tc"} and not plan.fused_head and plan.x2 is not None
    # the four stems read the one packed input chunk through one-hot channel weights
    cl = eng.fused_layers[id(m.specific_encoders[2].layers[0].conv.unit0.conv)]
    assert cl.cin == 4 and float(cl.wg_fwd_host[:, [0, 1, 3]].abs().max()) == 0.0 and float(cl.wg_fwd_host[:, 2].abs().max()) > 0
    with pytest.raises(ValueError):
        MultimodalUNetB200(dict(MULTIMODAL_MODEL_CFG, norm="BATCH"))
    with pytest.raises(ValueError):
        eng.build_plan(1, 24, 32, 32)
