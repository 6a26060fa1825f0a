"""The N > 1 host path on the CPU (gloo, world_size 2): images are dealt round-robin to ranks with no data-path
collective, and the one thing that is exchanged - the secret key words rank 0 sampled - arrives bit-identical."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200ckks import synthetic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_images, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # key distribution as in bench.py: rank 0 owns the secret, everyone receives the same words
    sk = torch.zeros(4 * 4096, dtype=torch.int64)
    if rank == 0:
        sk = torch.from_numpy(np.random.default_rng(7).integers(0, 2 ** 62, 4 * 4096, dtype=np.int64))
    dist.broadcast(sk, 0)
    mine = synthetic.shard(n_images, rank, world)
    # every rank "infers" its own images; the only reduction is the max over ranks of the elapsed time
    t = torch.tensor([float(len(mine))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, int(sk.sum().item()), [float(synthetic.synthetic_image(i)[0]) for i in mine]))
    if rank == 0:
        out.put((gathered, float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_and_key_broadcast_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port, n_images = _free_port(), 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, out)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    all_items = sorted(i for g in gathered for i in g[0])
    assert all_items == list(range(n_images))                      # every image exactly once
    assert set(gathered[0][0]).isdisjoint(gathered[1][0])
    assert gathered[0][1] == gathered[1][1]                          # the same secret on both ranks
    assert tmax == 4.0                                               # ceil(7 / 2): the slowest rank sets the time
    # images are a function of the image id only, whichever rank draws them
    assert gathered[1][2][0] == float(synthetic.synthetic_image(1)[0])


def test_shard_edge_cases():
    assert synthetic.shard(0, 0, 4) == []
    assert synthetic.shard(3, 3, 4) == []
    assert synthetic.shard(50, 7, 8) == [7, 15, 23, 31, 39, 47]
    assert sorted(sum((synthetic.shard(50, r, 8) for r in range(8)), [])) == list(range(50))
