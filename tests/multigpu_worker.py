"""torchrun worker (N >= 2 GPUs): N-rank TENT == 1-rank TENT on the concatenated batch.
Each rank adapts on its own volumes; one all-reduce of the flat [dgamma || dbeta] buffer per step
must leave every rank with the parameters the CPU oracle gets on the full batch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from multimodal_tta_b200 import SlidingWindowTTA, TentB200
from multimodal_tta_b200.synthetic import brats_volume
from oracle.sliding_window_oracle import sliding_window_oracle
from oracle.tent_oracle import TentOracle, flat_gamma_beta
from oracle.unet_oracle import BRATS_MODEL_CFG
from tests.util import make_pair, rel_l2


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    oracle, prod = make_pair(BRATS_MODEL_CFG, seed=11, device=f"cuda:{local}")
    to, tp = TentOracle(oracle, mode="sigmoid"), TentB200(prod, {"cuda_graph": True})
    xs = brats_volume(world, (64, 64, 64), seed=5)            # global batch = one volume per rank
    for it in range(2):
        lo, _ = to.step(xs)
        lp = tp.step(xs[rank:rank + 1].cuda()).cpu()
        assert rel_l2(lp, lo[rank:rank + 1]) < 1e-3, (rank, it, rel_l2(lp, lo[rank:rank + 1]))
    p = prod.engine.flat_params()
    gathered = [torch.empty_like(p) for _ in range(world)]
    dist.all_gather(gathered, p)
    for g in gathered:
        assert torch.equal(g, gathered[0])                     # bit-identical parameters on every rank
    perr = (p.cpu() - flat_gamma_beta(to.model)).abs()
    assert float(perr.median()) < 1e-5 and float((perr > 1e-4).float().mean()) < 0.03, (float(perr.median()),)
    # sliding window sharded over ranks == oracle with sw_batch * world windows per step
    oracle2, prod2 = make_pair(BRATS_MODEL_CFG, seed=12, device=f"cuda:{local}")
    to2, tp2 = TentOracle(oracle2, mode="sigmoid"), TentB200(prod2, {"cuda_graph": True})
    vol = brats_volume(1, (40, 48, 36), seed=9)
    ref = sliding_window_oracle(vol, (32, 32, 32), 1 * world, lambda w: to2.step(w)[0], overlap=0.5)
    got = SlidingWindowTTA(tp2, (32, 32, 32), sw_batch=1, overlap=0.5)(vol.cuda()).cpu()
    assert rel_l2(got, ref) < 2e-3, rel_l2(got, ref)
    if rank == 0:
        print("MULTIGPU_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
