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


def _fit_worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sifnn_b200
    import importlib
    F = importlib.import_module("sifnn_b200.fit")   # the package also exports a function called fit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank has its own shard: different loss sums and different batch counts (ragged last batch)
        sums = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64) * (rank + 1)
        qsum = torch.tensor([30.0, 0.5], dtype=torch.float64) * (rank + 1)
        n = 3 + rank
        s2, q2, n2 = F.reduce_epoch_sums(sums, qsum, n)
        # rank-local BatchNorm buffers diverge during an epoch; the checkpoint / evaluation must use rank 0's
        bn = torch.nn.BatchNorm2d(4)
        bn.running_mean.fill_(float(rank + 1))
        bn.running_var.fill_(2.0 * (rank + 1))
        bn.num_batches_tracked.fill_(10 + rank)
        F.sync_batchnorm_buffers(torch.nn.Sequential(bn), 0)
        # the early-stopping decision taken from the reduced validation loss is the same on every rank
        ck = F.model_checkpoint(10, patience=1)
        metrics = {"val_loss": []}
        decisions = []
        for epoch, local_val in enumerate([1.0 + rank, 0.9 - 0.5 * rank, 1.5 + 3 * rank, 2.0], start=1):   # per-rank values disagree on "improved"
            v, _, cnt = F.reduce_epoch_sums(torch.tensor([local_val], dtype=torch.float64), torch.zeros(0, dtype=torch.float64), 1)
            metrics["val_loss"].append(float(v[0]) / cnt)
            ck.test_update(torch.nn.Sequential(bn), metrics, "val_loss", epoch)
            decisions.append(ck.train_state)
        out[rank] = (s2.tolist(), q2.tolist(), n2, float(bn.running_mean[0]), float(bn.running_var[0]), int(bn.num_batches_tracked), decisions)
    finally:
        dist.destroy_process_group()


def test_gloo_fit_epoch_means_bn_buffers_and_early_stopping_agree_across_ranks():
    """fit() under data parallelism: epoch means reduced over ranks, BatchNorm buffers of rank 0 everywhere, identical
    break / continue decisions -- otherwise the ranks that continue hang in the next gradient all-reduce."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_fit_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0] == res[1]
    s, q, n, rm, rv, nbt, decisions = res[0]
    assert s == [3.0, 6.0, 9.0] and q == [90.0, 1.5] and n == 7
    assert (rm, rv, nbt) == (1.0, 2.0, 10)
    assert decisions[0] is None and decisions[1] == "continue" and decisions[-1] == "break"   # epoch 1 leaves train_state None, as the reference does
