"""Generate tests/golden/dice_golden.npz by running the REFERENCE's own code in this container.

Run from the repo root (needs /root/reference, so it cannot run on the GPU box):
    python tests/golden/gen_dice_golden.py

``src/evaluation/seg_eval.py`` imports omegaconf / monai / tqdm at module level; none is
installed here, so empty stub modules are pre-seeded into ``sys.modules`` for the names the
import statements touch.  ``_binary_dice_iou`` (seg_eval.py:41-68) itself is pure torch, so the
stubs do not influence the numbers written here.  Also records the behaviour of the reference
``Registry`` (src/registry.py:10-56): duplicate -> overwrite with a printed warning,
missing -> KeyError.
"""
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _DictConfig(dict):
    pass


def main():
    sys.path.insert(0, REF)
    _stub("omegaconf", DictConfig=_DictConfig, OmegaConf=type("OmegaConf", (), {"select": staticmethod(lambda c, p: None)}))
    _stub("monai")
    _stub("monai.losses", DiceCELoss=object)
    _stub("tqdm", tqdm=lambda it, **kw: it)
    from src.evaluation.seg_eval import _binary_dice_iou  # noqa: E402
    from src.registry import Registry  # noqa: E402

    g = torch.Generator().manual_seed(42)
    cases = {}
    # case 0: random masks; case 1: one GT-empty region; case 2: all-empty pred; case 3: perfect
    shapes = [(2, 3, 6, 7, 8), (1, 3, 4, 4, 4), (2, 1, 5, 5, 5), (1, 3, 8, 8, 8)]
    for i, shp in enumerate(shapes):
        pred = (torch.rand(shp, generator=g) > 0.5).to(torch.uint8)
        gt = (torch.rand(shp, generator=g) > 0.6).to(torch.uint8)
        if i == 1:
            gt[:, 1] = 0
        if i == 2:
            pred[:] = 0
        if i == 3:
            pred = gt.clone()
        dice, iou, valid = _binary_dice_iou(pred, gt)
        cases[f"pred{i}"] = pred.numpy()
        cases[f"gt{i}"] = gt.numpy()
        cases[f"dice{i}"] = dice.numpy()
        cases[f"iou{i}"] = iou.numpy()
        cases[f"valid{i}"] = valid.numpy()
    np.savez_compressed(os.path.join(HERE, "dice_golden.npz"), **cases)

    reg = Registry("models")
    reg.register("a", int)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        reg.register("a", float)
    try:
        reg.get("zzz")
        missing = "no error"
    except KeyError as e:
        missing = str(e)
    rec = {"dup_warning": buf.getvalue(), "dup_result": reg.get("a").__name__, "missing_keyerror": missing,
           "list_all": reg.list_all(), "has_a": reg.has("a"), "has_b": reg.has("b")}
    with open(os.path.join(HERE, "registry_golden.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("wrote dice_golden.npz, registry_golden.json")


if __name__ == "__main__":
    main()
