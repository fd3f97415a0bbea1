"""N>1 host logic on CPU: world_size-2 gloo processes."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from farms_mujoco_b200.sharding import env_shard


def test_env_shard_partitions():
    for n, world in ((16384, 8), (10, 3), (7, 8), (65536, 4)):
        ranges = [env_shard(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        env_shard(4, 4, 4)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from farms_mujoco_b200 import models, mjcf_subset
    from farms_mujoco_b200.sharding import env_shard, synthetic_inputs, gather_env_statistics
    model = mjcf_subset.parse_mjcf(models.swimmer8().mjcf)
    n_envs = 11
    start, stop = env_shard(n_envs, rank, world)
    qpos, _, phase = synthetic_inputs(model, np.arange(start, stop))
    stats = torch.as_tensor(np.concatenate([qpos[:, 7:9], phase[:, None]], axis=1))
    full = gather_env_statistics(stats, world)
    ref_qpos, _, ref_phase = synthetic_inputs(model, np.arange(n_envs))
    ref = np.concatenate([ref_qpos[:, 7:9], ref_phase[:, None]], axis=1)
    ok = full.shape == (n_envs, 3) and np.array_equal(full.numpy(), ref)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == world
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_reproduce_single_process_inputs():
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}


def test_bind_host_to_device_is_harmless_without_a_gpu():
    """No NVML device here: the helper reports None and leaves the affinity alone."""
    import os
    from farms_mujoco_b200.sharding import bind_host_to_device
    before = os.sched_getaffinity(0)
    bound = bind_host_to_device(0)
    assert bound is None or set(bound) <= before
    if bound is None:
        assert os.sched_getaffinity(0) == before
