"""CPU, world_size 2, gloo: the N>1 host logic of the batch-sharded MLP harness (tpp_mlir_b200/shard.py).
Row independence is what makes the path shard with no data-path collective: the oracle MLP on each rank's shard,
gathered, must equal the oracle MLP on the whole batch bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mlp_oracle(x, Ws, bs):
    import oracle

    a = x
    for W, b in zip(Ws, bs):
        y = np.empty((a.shape[0], W.shape[1]), np.uint16)
        oracle.fused_brgemm(2, a.shape[0], W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5, 4,
                            1, np.ascontiguousarray(a), W, y, b, 1)
        a = y
    return a


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from tpp_mlir_b200 import shard

    layers, batch = (64, 96, 64), 32
    if rank == 0:
        gen = oracle.TensorInit("normal", oracle.BF16, 123)
        Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
        bs = [gen.fill(k) for k in layers[1:]]
        x = gen.fill(batch, layers[0])
    else:  # other ranks start from garbage: the broadcast must overwrite it
        Ws = [np.full((c, k), 7, np.uint16) for c, k in zip(layers[:-1], layers[1:])]
        bs = [np.full((k,), 7, np.uint16) for k in layers[1:]]
        x = np.full((batch, layers[0]), 7, np.uint16)
    tens = [torch.from_numpy(a.view(np.int16)) for a in Ws + bs + [x]]
    shard.broadcast_parameters(tens, src=0)
    lo, hi = shard.shard_bounds(batch, rank, world, tile_m=8)
    y_local = _mlp_oracle(x[lo:hi], Ws, bs)
    full = shard.gather_rows(torch.from_numpy(y_local.view(np.int16)), dst=0)
    slow = shard.max_over_ranks(1.0 + rank)
    if rank == 0:
        want = _mlp_oracle(x, Ws, bs)
        ret["equal"] = bool(np.array_equal(full.numpy().view(np.uint16), want))
        ret["max"] = slow
        ret["bounds"] = (lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_batch_sharding_world2_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["equal"] is True
    assert ret["max"] == 2.0          # max over ranks, not rank 0's own time
    assert ret["bounds"] == (0, 16)


def test_shard_bounds_reject_unclean_splits():
    from tpp_mlir_b200 import shard

    assert shard.shard_bounds(2048, 3, 8, tile_m=128) == (768, 1024)
    with pytest.raises(ValueError):
        shard.shard_bounds(2048 + 128, 0, 8, tile_m=128)
    with pytest.raises(ValueError):
        shard.shard_bounds(256, 0, 3)
