"""
Sharding of a batch of independent envs over the GPUs of one box.

Each env's step touches only its own board, goals and state record (the torus
wrap is internal to a board), so the path shards with no exchange: rank g owns
the contiguous env range [lo, hi), the level pool is replicated, and every
per-env quantity that must not depend on the GPU count (level assignment, RNG
stream) is derived from the *global* env index.  The only collectives are
optional and off the step path: the gather of step outputs to every rank
(reward f32 + done u8 = 5 B/env) and the all-gather that reassembles the
observation batch for a learner that wants all of it on one GPU (NCCL over
NVLink / NVSwitch; 9 375 B/env for the training observation shape -- keeping the
observations sharded and the policy data-parallel avoids it altogether).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world):
    """Contiguous, balanced [lo, hi) of rank `rank` (first n_total % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_levels(levels, lo, hi):
    """First level of the envs [lo, hi): global env i starts on levels[i % len(levels)]."""
    return [levels[i % len(levels)] for i in range(lo, hi)]


def shard_rng_states(seed, n_total, lo, hi):
    """PCG64 words of the envs [lo, hi) out of n_total, independent of the sharding."""
    from .env import seed_rng_states
    return seed_rng_states(seed, n_total)[lo:hi]


def make_sharded_env(levels, n_total, seed=None, rank=None, world=None, **kwargs):
    """BatchedSafeLifeEnv over this rank's env range (rank/world default to torch.distributed)."""
    from .env import BatchedSafeLifeEnv
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n_total, rank, world)
    # the whole pool on every rank (a few MB); first level and level stride come from
    # the GLOBAL env index and count, so the level sequence of an env does not depend
    # on the number of ranks
    env = BatchedSafeLifeEnv(levels, hi - lo, env_offset=lo, total_envs=n_total, **kwargs)
    env.set_rng_state(shard_rng_states(seed, n_total, lo, hi))
    env.global_range = (lo, hi)
    return env


def gather_step_outputs(reward, done, n_total, group=None):
    """
    All-gather of (reward f32[n_local], done bool/u8[n_local]) into
    (reward f32[n_total], done u8[n_total]) in global env order.  Shards may
    differ by one env, so they are padded to the largest shard.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return reward.clone(), done.to(torch.uint8)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    packed = torch.zeros((width, 2), dtype=torch.float32, device=reward.device)
    packed[: hi - lo, 0] = reward
    packed[: hi - lo, 1] = done.to(torch.float32)
    flat = torch.empty((world * width, 2), dtype=torch.float32, device=reward.device)
    dist.all_gather_into_tensor(flat, packed, group=group)
    out = flat.view(world, width, 2)
    rewards = torch.cat([out[r, : h - l, 0] for r, (l, h) in enumerate(sizes)])
    dones = torch.cat([out[r, : h - l, 1] for r, (l, h) in enumerate(sizes)]).to(torch.uint8)
    return rewards, dones


def gather_observations(obs, n_total, group=None):
    """
    All-gather of the per-rank observation tensors [n_local, ...] into the whole
    batch [n_total, ...] in global env order, on every rank (one
    all_gather_into_tensor; shards padded to the largest one).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return obs.clone()
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if obs.shape[0] != hi - lo:
        raise ValueError("obs has %d envs, this rank owns %d" % (obs.shape[0], hi - lo))
    padded = obs
    if hi - lo < width:
        padded = torch.zeros((width,) + tuple(obs.shape[1:]), dtype=obs.dtype, device=obs.device)
        padded[: hi - lo] = obs
    flat = torch.empty((world * width,) + tuple(obs.shape[1:]), dtype=obs.dtype, device=obs.device)
    dist.all_gather_into_tensor(flat, padded.contiguous(), group=group)
    if all(h - l == width for l, h in sizes):
        return flat
    return torch.cat([flat[r * width: r * width + (h - l)] for r, (l, h) in enumerate(sizes)])
