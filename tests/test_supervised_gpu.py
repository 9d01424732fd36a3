"""the SUPERVISED step of the reference's trainer on the B200 kernels.  ``SegTrainer.run_step``
(/root/reference/src/core/trainers/seg_trainer.py:97-145: zero_grad -> logits = model(x) -> DiceCELoss ->
loss.backward() -> optimizer.step() -> float(loss)) runs UNCHANGED on a ``unet_b200`` built with ``trainable: true``:
the logits carry an autograd node whose backward is the CUDA backward (input gradients, norm backward, weight and
bias gradients), the loss and the optimizer stay the reference's own torch objects, and the packed operand blobs
follow the optimizer through a device-side repack.  Checked against the same step on the CPU oracle."""
import copy

import pytest
import torch

from multimodal_tta_b200 import UNetB200
from multimodal_tta_b200.presets import BARE_DEFAULT_MODEL_CFG, BRATS_MODEL_CFG, HECKTOR_MODEL_CFG
from multimodal_tta_b200.synthetic import brats_volume, region_labels
from oracle.dicece_oracle import dice_ce_loss, make_optimizer
from oracle.unet_oracle import OracleUNet
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def _pair(cfg, seed):
    torch.manual_seed(seed)
    oracle = OracleUNet.from_cfg(cfg)
    prod = UNetB200(dict(cfg, trainable=True, deterministic=True))
    prod.load_state_dict(copy.deepcopy(oracle.state_dict()))
    return oracle.train(), prod.to("cuda").train()


def _run_step(model, opt, x, y):
    """The reference's step, verbatim in shape (seg_trainer.py:105-145)."""
    opt.zero_grad()
    logits = model(x)
    loss = dice_ce_loss(logits, y)
    loss.backward()
    opt.step()
    return logits.detach(), float(loss.item())


# tolerances (measured, scripts/diag_supervised.py): one scaled fp16 gradient plane gives 1.4e-3 over all parameters at
# 64^3 (the fp32-exact CUDA-core conv backend: 5.7e-4), bf16 hi + lo planes less; 32^3 inputs normalise over 2^3 = 8
# voxels at the bottom level and are ill-conditioned (5e-3), as in the TENT parity tests
@pytest.mark.parametrize("cfg,C,R,dims,bwd,tol", [(BRATS_MODEL_CFG, 4, 3, (64, 64, 64), "fp16", 3e-3),
                                                  (BRATS_MODEL_CFG, 4, 3, (64, 48, 64), "bf16x2", 5e-3),
                                                  (HECKTOR_MODEL_CFG, 2, 1, (32, 32, 32), "fp16", 2e-2),
                                                  (dict(BARE_DEFAULT_MODEL_CFG, in_channels=4, num_classes=3), 4, 3,
                                                   (32, 32, 32), "fp16", 2e-2)])
def test_every_parameter_gradient_matches_autograd(cuda, cfg, C, R, dims, bwd, tol):
    oracle, prod = _pair(dict(cfg, bwd_precision=bwd), seed=81)
    x = brats_volume(2, dims, seed=5, channels=C)
    y = region_labels(2, R, dims, seed=6)
    lo = oracle(x)
    dice_ce_loss(lo, y).backward()
    lp = prod(x.cuda())
    assert lp.requires_grad and rel_l2(lp.detach().cpu(), lo.detach()) < 1e-4
    loss_p = dice_ce_loss(lp, y.cuda())
    loss_p.backward()
    po, pp = dict(oracle.named_parameters()), dict(prod.named_parameters())
    assert set(po) == set(pp)
    num = den = 0.0
    worst = (0.0, "")
    for n, p in po.items():
        g = pp[n].grad
        assert g is not None and g.shape == p.grad.shape, n
        e = rel_l2(g.cpu(), p.grad)
        if float(p.grad.norm()) > 1e-6:      # (a conv bias in front of a norm has an exactly-zero gradient: noise / noise)
            worst = max(worst, (e, n))
        num += float((g.cpu().double() - p.grad.double()).pow(2).sum()); den += float(p.grad.double().pow(2).sum())
    total = (num / den) ** 0.5
    print(f"[supervised {dims} {bwd}] all-parameter gradient rel-L2 {total:.2e}, worst tensor {worst[1]} {worst[0]:.2e}")
    assert total < tol and worst[0] < 20 * tol
    # a second backward on a fresh forward accumulates into .grad like autograd does
    dice_ce_loss(prod(x.cuda()), y.cuda()).backward()
    n0 = "model.0.conv.unit0.conv.weight" if "model.0.conv.unit0.conv.weight" in pp else next(iter(pp))
    assert rel_l2(pp[n0].grad.cpu(), 2 * po[n0].grad) < 20 * tol


def test_supervised_steps_follow_the_oracle_and_weights_are_repacked_on_device(cuda):
    oracle, prod = _pair(BRATS_MODEL_CFG, seed=82)
    oo, op = make_optimizer(oracle, lr=1e-3), make_optimizer(prod, lr=1e-3)     # larger lr: the weights really move
    w0 = prod.model[0].conv.unit0.conv.weight.detach().clone()
    for it in range(3):
        x = brats_volume(2, (32, 32, 32), seed=30 + it)
        y = region_labels(2, 3, (32, 32, 32), seed=40 + it)
        lo, loss_o = _run_step(oracle, oo, x, y)
        lp, loss_p = _run_step(prod, op, x.cuda(), y.cuda())
        print(f"[supervised step {it}] logits rel-L2 {rel_l2(lp.cpu(), lo):.1e}, loss {loss_p:.6f} vs {loss_o:.6f}")
        assert rel_l2(lp.cpu(), lo) < (1e-4 if it == 0 else 5e-2)     # Adam's sign(g) on noise-floor gradients (lr 1e-3)
        assert abs(loss_p - loss_o) < 1e-3 * max(1.0, abs(loss_o))
    assert float((prod.model[0].conv.unit0.conv.weight.detach() - w0).abs().max()) > 1e-4
    # the kernels run on the UPDATED weights: forward of the product == forward of a fresh oracle holding them
    torch.manual_seed(0)
    fresh = OracleUNet.from_cfg(BRATS_MODEL_CFG).train()
    fresh.load_state_dict({k: v.detach().cpu() for k, v in prod.state_dict().items()})
    x = brats_volume(1, (32, 32, 32), seed=50)
    with torch.no_grad():
        ref = fresh(x)
        got = prod(x.cuda()).cpu()
    assert rel_l2(got, ref) < 1e-4
    assert len(prod.engine.plans) == 2          # (2, 32^3) and (1, 32^3): repacking never dropped a plan


def test_trainable_false_keeps_the_inference_path(cuda):
    _, prod = _pair(BRATS_MODEL_CFG, seed=83)
    prod.trainable = False
    out = prod(brats_volume(1, (32, 32, 32), seed=1).cuda())
    assert not out.requires_grad
