"""Host-side multi-process logic on CPU: env sharding and the episode-statistics reduction
(the path's only collective) with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from swarm_b200 import distributed as D
from swarm_b200._abi import STAT_NAMES


def test_shard_range_partitions_every_env_exactly_once():
    for total in (1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = D.shard_range(total, r, world)
                assert 0 <= lo <= hi <= total
                covered.extend(range(lo, hi))
                assert abs((hi - lo) - total / world) < 1.0 + 1e-9
            assert covered == list(range(total))


def test_seeds_depend_only_on_global_env_index():
    total = 1000
    one = D.global_env_seeds(5, 0, total)
    for world in (2, 4, 8):
        parts = [D.global_env_seeds(5, *D.shard_range(total, r, world)) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), one)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = {n: float((rank + 1) * (i + 1)) for i, n in enumerate(STAT_NAMES)}
    stats["return_sum"] = -1.5 * (rank + 1)
    out = D.all_reduce_stats(stats)
    lo, hi = D.shard_range(10, rank, world)
    t = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(t)
    q.put((rank, out, int(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_all_reduce_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, n_envs in results:
        assert n_envs == 10
        for i, n in enumerate(STAT_NAMES):
            if n == "return_sum":
                assert out[n] == -4.5
            else:
                assert out[n] == 3 * (i + 1)
    s = D.summarize(results[0][1])
    assert s["episodes_this_iter"] == 3 and s["episode_reward_mean"] == -1.5
