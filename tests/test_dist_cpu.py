"""Data-parallel host logic on CPU (gloo, world_size 2): bucketed gradient all-reduce and the DP semantics
the B200 trainer implements (local BatchNorm, averaged gradients, identical Adam on every rank), with the CPU
oracle standing in for the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sifnn_b200
    import sifnn_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1) bucketed all-reduce == plain sum, both buckets, any split
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(1000, generator=g)
        want = flat.clone()
        dist.all_reduce(want)
        ar = sifnn_b200.BucketedAllReduce(flat, 637)
        ar.start(0)
        ar.start(1)
        ar.finish()
        assert torch.equal(flat, want)
        # 2) DP step semantics on a 64x64 problem: each rank runs the oracle on ITS shard (local BatchNorm),
        #    gradients are summed with the product helper and scaled by 1/world, then one Adam step.
        sd = O.init_state_dict(1)
        lst, up, ndvi = O.synthetic_batch(4, seed=9, hr=64)
        sl = lambda t: sifnn_b200.shard_batch(t, rank, world)
        tr = O.Trainer(sd, "sr1", 0.99, -0.5, 1e-3)
        tr.loss_and_grads(sl(lst), sl(up), sl(ndvi))
        fg = tr.flat_grads().clone()
        ar = sifnn_b200.BucketedAllReduce(fg, 160096)
        ar.start(0)
        ar.start(1)
        ar.finish()
        fg /= world
        off = 0
        for k in tr.keys:
            n = tr.sd[k].numel()
            tr.sd[k].grad.copy_(fg[off:off + n].view_as(tr.sd[k]))
            off += n
        tr.opt.step()
        out[rank] = tr.flat_params().numpy()
        if rank == 0:
            # single-process emulation of the same sharding
            grads = []
            for r in range(world):
                e = O.Trainer(sd, "sr1", 0.99, -0.5, 1e-3)
                e.loss_and_grads(*(sifnn_b200.shard_batch(t, r, world) for t in (lst, up, ndvi)))
                grads.append(e.flat_grads())
            assert torch.allclose(fg, sum(grads) / world, rtol=1e-5, atol=1e-7)
    finally:
        dist.destroy_process_group()


def test_gloo_two_ranks_bucketed_allreduce_and_dp_semantics():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world
        assert np.array_equal(out[0], out[1])  # identical weights on every rank after the step
