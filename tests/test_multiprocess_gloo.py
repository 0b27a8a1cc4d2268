"""world_size-2 gloo test (CPU) of the N>1 path: pair sharding + the single pose all-gather."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from staticmapping_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = parallel.shard_pairs(n_pairs, rank, world)
    Ts, scores = [], []
    for i in mine:                       # stand-in for the per-pair alignment result
        T = np.eye(4); T[:3, 3] = [i, 2 * i, -i]
        Ts.append(T); scores.append(0.5 + i)
    table = parallel.allgather_poses(parallel.pack_poses(Ts, scores), n_pairs)
    q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_poses_world2():
    world, n_pairs = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(world):
        Ts, scores = parallel.unpack_poses(got[r])
        assert len(Ts) == n_pairs
        for i in range(n_pairs):
            assert np.array_equal(Ts[i][:3, 3], [i, 2 * i, -i]) and scores[i] == 0.5 + i
    assert np.array_equal(got[0], got[1])
