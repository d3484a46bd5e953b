"""CPU, world_size 2 over gloo: the host-side data-parallel logic (sample sharding, flat-gradient
all-reduce, 1/world scaling) that bench.py / train.py run over NCCL on the GPUs."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    tu = importlib.import_module("n-body_pointcloudevolution_b200.train_utils")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    store = tu.ParamStore([3, 8, 3], device="cpu", seed=7)
    mine = list(tu.shard_samples(5, rank, world))
    # stand-in for backward: per-sample "gradient" = sample id + 1 everywhere
    store.zero_grad()
    for s in mine:
        store.flat_grad += float(s + 1)
    tu.allreduce_gradients(store, world)
    out[rank] = (mine, store.flat_grad.clone().numpy(), store.flat.clone().numpy())
    torch.distributed.destroy_process_group()


def test_shard_and_allreduce_two_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    s0, g0, p0 = out[0]
    s1, g1, p1 = out[1]
    assert sorted(s0 + s1) == [0, 1, 2, 3, 4] and abs(len(s0) - len(s1)) <= 1
    assert np.array_equal(g0, g1) and np.allclose(g0, 15.0)      # 1+2+3+4+5 on every rank
    assert np.array_equal(p0, p1)                                # same seed => identical replicas


def test_shard_samples_cover():
    import importlib
    tu = importlib.import_module("n-body_pointcloudevolution_b200.train_utils")
    for n, w in ((8, 8), (8, 3), (1, 4), (64, 8)):
        got = sorted(i for r in range(w) for i in tu.shard_samples(n, r, w))
        assert got == list(range(n))
