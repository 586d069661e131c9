"""Multi-process (gloo, world_size 2, CPU) checks of the sharding logic."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from safelife_b200 import sharding


def test_shard_ranges_partition():
    for n in (0, 1, 7, 8, 4096, 65537):
        for world in (1, 2, 3, 8):
            ranges = [sharding.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def test_rng_and_levels_do_not_depend_on_world_size():
    whole = sharding.shard_rng_states(4, 10, 0, 10)
    parts = np.concatenate([sharding.shard_rng_states(4, 10, *sharding.shard_range(10, r, 3))
                            for r in range(3)])
    assert np.array_equal(whole, parts)
    levels = ["a", "b", "c"]
    joined = sum((sharding.shard_levels(levels, *sharding.shard_range(10, r, 4)) for r in range(4)), [])
    assert joined == [levels[i % 3] for i in range(10)]


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(n_total, rank, world)
        reward = torch.arange(lo, hi, dtype=torch.float32) * 0.5
        done = (torch.arange(lo, hi) % 3 == 0)
        r, d = sharding.gather_step_outputs(reward, done, n_total)
        want_r = torch.arange(n_total, dtype=torch.float32) * 0.5
        want_d = (torch.arange(n_total) % 3 == 0).to(torch.uint8)
        obs = (torch.arange(lo, hi, dtype=torch.int64)[:, None, None, None] +
               torch.arange(3 * 2 * 5).reshape(3, 2, 5)).to(torch.uint8)
        whole = sharding.gather_observations(obs, n_total)
        want_o = (torch.arange(n_total, dtype=torch.int64)[:, None, None, None] +
                  torch.arange(3 * 2 * 5).reshape(3, 2, 5)).to(torch.uint8)
        ret[rank] = bool(torch.equal(r, want_r) and torch.equal(d, want_d) and
                         torch.equal(whole, want_o) and whole.dtype == torch.uint8)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 9])
def test_gather_step_outputs_gloo(n_total):
    world = 2
    port = 29650 + n_total
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}
