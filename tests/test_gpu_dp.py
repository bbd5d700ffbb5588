"""Data-parallel training on real GPUs (-m gpu; needs >= 2 devices, otherwise skipped): N ranks over NCCL must
produce the single-process emulation of the same sharding -- oracle run per shard from identical weights (local
BatchNorm), gradients averaged, one Adam step -- and identical weights on every rank."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


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
    import torch.distributed as dist
    import sifnn_b200
    import sifnn_oracle as O
    import model as model_mod
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sd = O.init_state_dict(2)
        lst, up, ndvi = O.synthetic_batch(2 * world, seed=21)
        m = model_mod.ModelB_2(in_channels=2)
        m.load_state_dict(sd)
        m = m.cuda().train()
        tr = sifnn_b200.Trainer(m, "sr2", 0.5, -0.25, 1e-3)
        assert tr.world == world
        tr.broadcast_parameters(0)
        sl = lambda t: sifnn_b200.shard_batch(t, rank, world).cuda()
        losses = tr.step(sl(lst), sl(ndvi)).cpu().numpy()
        out[rank] = (torch.cat([p.detach().reshape(-1) for p in m.parameters()]).cpu().numpy(), losses)
        # the same step replayed from CUDA graphs (three segments around the two NCCL buckets) must reproduce eager
        ma, mb = (model_mod.ModelB_2(in_channels=2) for _ in range(2))
        for mm in (ma, mb):
            mm.load_state_dict(sd)
            mm.cuda().train()
        ta, tb = sifnn_b200.Trainer(ma, "sr2", 0.5, -0.25, 1e-3), sifnn_b200.Trainer(mb, "sr2", 0.5, -0.25, 1e-3)
        tb.capture(sl(lst), sl(ndvi))   # one graph per step, NCCL all-reduces captured; capture() preserves the training state
        for _ in range(2):
            la = ta.step(sl(lst), sl(ndvi))
            lb = tb.step_graph(sl(lst), sl(ndvi)).clone()
        pa = torch.cat([p.detach().reshape(-1) for p in ma.parameters()])
        pb = torch.cat([p.detach().reshape(-1) for p in mb.parameters()])
        assert torch.allclose(la, lb, rtol=1e-6), (la, lb)
        assert float((pa - pb).abs().max()) <= 1e-6 * float(pa.abs().max()), float((pa - pb).abs().max())
        assert getattr(tb, "_dp_single", False), "the data-parallel step should replay as ONE graph with the all-reduces captured"
        tb.release_graphs()   # captured collectives pin the NCCL communicator: drop the graphs before destroy_process_group()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_single_process_emulation():
    import torch.multiprocessing as mp
    import sifnn_b200
    import sifnn_oracle as O
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert np.array_equal(res[0][0], res[1][0])
    sd = O.init_state_dict(2)
    lst, up, ndvi = O.synthetic_batch(2 * world, seed=21)
    grads, losses = [], []
    for r in range(world):
        e = O.Trainer(sd, "sr2", 0.5, -0.25, 1e-3)
        _, sc, _ = e.loss_and_grads(*(sifnn_b200.shard_batch(t, r, world) for t in (lst, up, ndvi)))
        grads.append(e.flat_grads())
        losses.append(sc)
        assert np.allclose(res[r][1], sc, rtol=1e-4)
    ref = O.Trainer(sd, "sr2", 0.5, -0.25, 1e-3)
    fg = sum(grads) / world
    off = 0
    ref.opt.zero_grad()
    for k in ref.keys:
        n = ref.sd[k].numel()
        ref.sd[k].grad = fg[off:off + n].view_as(ref.sd[k]).clone()
        off += n
    ref.opt.step()
    want = ref.flat_params().numpy()
    d = np.abs(res[0][0] - want)
    assert (d <= 1e-4 * np.abs(want).max()).mean() > 0.99 and d.max() <= 2 * 1e-3 * 1.01
