"""Multi-GPU layout of the step loop: environments are independent (no cross-env state in the reference:
environment/environment.py:21-47), so rank r of G owns the contiguous global env ids
[first, first + count) and never communicates during a step.  The only collective of the design is an optional
end-of-rollout reduction of a few statistics (anthill deliveries, carried food, reward sums) over NCCL (gloo in
CPU tests).  Per-env results do not depend on the partition: the Philox collision noise is keyed by the global env
id (AntsConfig.env_id_base)."""
import numpy as np


def shard_envs(total_envs, rank, world_size):
    """Contiguous, balanced partition: -> (first_env, n_envs) of `rank`."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d of %d" % (rank, world_size))
    base, rem = divmod(int(total_envs), int(world_size))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def weak_scaling_envs(envs_per_gpu, rank, world_size):
    """bench.py's weak-scaling layout: every rank holds `envs_per_gpu` envs; global ids are rank-major."""
    return rank * int(envs_per_gpu), int(envs_per_gpu)


STAT_KEYS = ("anthill_food", "carried_food", "reward_sum", "n_ants")


def local_stats(state, last_reward=None):
    """Per-rank statistics vector (float64) from an exported state dict."""
    v = np.zeros(len(STAT_KEYS), dtype=np.float64)
    v[0] = float(np.sum(state["anthill_food"]))
    v[1] = float(np.sum(state["holding"]))
    v[2] = 0.0 if last_reward is None else float(np.sum(last_reward))
    v[3] = float(np.asarray(state["holding"]).size)
    return v


def reduce_stats(vec, device=None):
    """Sum the statistics vector over all ranks (no-op without an initialised process group).  NCCL when `device` is
    a CUDA device, gloo on the CPU."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(vec, dtype=np.float64))
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dict(zip(STAT_KEYS, t.cpu().tolist()))


def max_over_ranks(value, device=None):
    """Timing rule of the benchmark: a multi-GPU time is the max over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
