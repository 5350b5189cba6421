"""N>1 host logic on CPU (gloo, world_size 2): batch sharding + MPJPE combination.  The forward itself
is replaced by the oracle here (the CUDA path cannot run without a GPU); the sharding code is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import _models as M
from cistgcn_b200.dist import shard_bounds, sharded_eval_mpjpe
from oracle import cistgcn_oracle as O


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model, sd, cfg = M.build(8, 22, "W2")
    x, tgt = O.synth_inputs(7, cfg)          # odd batch: ragged shards (4 + 3)

    def fwd(xs, ts):
        with torch.no_grad():
            p = O.forward(sd, cfg, xs)
        return p, O.mpjpe(p, ts, None).double().sum((0, 2))

    pred, (lo, hi), m_all, m_frames = sharded_eval_mpjpe(fwd, x, tgt)
    with torch.no_grad():
        full = O.forward(sd, cfg, x)
    ok = torch.allclose(pred, full[lo:hi], atol=1e-6)
    ok &= abs(m_all.item() - O.mpjpe(full, tgt).item()) < 1e-5
    ok &= torch.allclose(m_frames.float(), O.mpjpe(full, tgt, (0, 2)), atol=1e-5)
    out[rank] = (bool(ok), lo, hi)
    dist.destroy_process_group()


def test_sharded_mpjpe_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] == (True, 0, 4) and out[1] == (True, 4, 7)
