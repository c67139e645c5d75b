"""CPU, world_size 2 over gloo: the object sharding + single all-gather logic of genpose_b200.distributed
(the GPU path uses the same code with backend nccl)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genpose_b200 import distributed as D


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 65, 512):
        for w in (1, 2, 3, 8):
            b = D.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_objects, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    D.init_from_env(backend="gloo")

    def local_fn(lo, hi):
        ids = torch.arange(lo, hi, dtype=torch.float32)
        pose = ids[:, None, None] * 100 + torch.arange(5)[None, :, None] * 10 + torch.arange(9)[None, None, :]
        energy = ids[:, None, None] + torch.zeros(hi - lo, 5, 2)
        assert hi > lo, "local_fn must not be called for an empty shard"
        return {"pred_pose": pose.double() + 1e-9, "energy": energy}          # float64 poses (ODE sampler) must survive the gather

    out = D.run_sharded(local_fn, n_objects, keys=("pred_pose", "energy"))
    ref = local_fn(0, n_objects)
    ok = all(torch.equal(out[k], ref[k]) and out[k].dtype == ref[k].dtype for k in ref)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_objects,world", [(8, 2), (7, 2), (1, 2), (2, 3)])
def test_run_sharded_gloo(n_objects, world):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_objects, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
