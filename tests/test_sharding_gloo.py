"""N > 1 host logic on the CPU: world_size-2 gloo process group on 127.0.0.1 -- env partition, shard-invariant
generation (global env ids), statistics all-reduce and max-over-ranks timing."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from antsrl_b200.generator import BatchedEnvironmentGenerator, CirclesGenerator, stack_states
from antsrl_b200.sharding import local_stats, max_over_ranks, reduce_stats, shard_envs, weak_scaling_envs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gen():
    return BatchedEnvironmentGenerator(48, 40, 12, 2, 0, CirclesGenerator(5, 3, 6), CirclesGenerator(4, 3, 6), 50,
                                       seed_base=1000)


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = shard_envs(total, rank, world)
    states = stack_states(_gen().generate_states(count, first))
    states["anthill_food"] = np.full(count, float(rank + 1))           # pretend deliveries
    states["holding"] = np.full((count, 12), 0.5)
    stats = reduce_stats(local_stats(states, last_reward=np.ones((count, 12))))
    tmax = max_over_ranks(10.0 + rank)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), first=first, count=count, x=states["x"], walls=states["walls"],
             stats=np.array([stats[k] for k in ("anthill_food", "carried_food", "reward_sum", "n_ants")]), tmax=tmax)
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_contiguous_and_balanced():
    for total in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_envs(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    assert weak_scaling_envs(512, 3, 8) == (1536, 512)


def test_two_rank_gloo_sharding(tmp_path):
    total, world = 5, 2
    port = _free_port()
    mp.start_processes(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True, start_method="fork")
    whole = stack_states(_gen().generate_states(total, 0))
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["first"]) for p in parts] == [0, 3] and [int(p["count"]) for p in parts] == [3, 2]
    # the union of the shards is the single-process generation, env by env (global env id -> seed)
    assert np.array_equal(np.concatenate([p["x"] for p in parts]), whole["x"])
    assert np.array_equal(np.concatenate([p["walls"] for p in parts]), whole["walls"])
    for p in parts:   # every rank sees the same reduced statistics
        assert np.allclose(p["stats"], [3 * 1.0 + 2 * 2.0, 5 * 12 * 0.5, 5 * 12, 5 * 12])
        assert float(p["tmax"]) == 11.0
