"""world_size-2 gloo runs (CPU) of the N>1 host logic: the gradient all-reduce + 1/world Adam scale
keeps ranks identical and equals the single-rank mean; window sharding covers every window once."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multimodal_tta_b200 import TentB200, UNetB200
        from multimodal_tta_b200.sliding_window import shard_schedule
        from oracle.unet_oracle import HECKTOR_MODEL_CFG

        tent = TentB200(UNetB200(dict(HECKTOR_MODEL_CFG)), {"cuda_graph": False})
        assert tent.world_size == world

        class FakeEngine:            # the all-reduce only touches the flat gradient buffer
            dgb = torch.arange(6, dtype=torch.float32) * (rank + 1)
        eng = FakeEngine()
        scale = tent._allreduce_grads(eng)
        covered = [i for idxs, _ in shard_schedule(18, 2, world, rank) for i in idxs if i is not None]
        q.put((rank, eng.dgb.clone(), scale, covered))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_sharding():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = torch.arange(6, dtype=torch.float32)
    for rank, g, scale, _ in res:
        assert torch.equal(g, base * 3)               # sum over ranks: identical on every rank
        assert scale == 0.5                           # Adam sees the mean over ranks
    assert sorted(res[0][3] + res[1][3]) == list(range(18))
