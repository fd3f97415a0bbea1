"""Environment sharding across ranks (one process per GPU).

Environments are independent: each rank steps a
contiguous range of environment ids with no data-path collective.  The only
collective is the optional end-of-rollout gather of per-environment statistics
(NCCL on GPUs; the CPU tests drive the same code over gloo).
"""

import numpy as np


def env_shard(n_envs, rank, world_size):
    """Contiguous env-id range ``[start, stop)`` of ``rank``; sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(int(n_envs), int(world_size))
    start = rank*base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def synthetic_inputs(model, env_ids, seed=1234):
    """Per-env initial joint angles U(-0.1, 0.1) rad and controller phase U(0, 2 pi),
    drawn in env-id order so that any shard reproduces the same environments
    (SURVEY.md section 8d)."""
    import torch
    env_ids = np.asarray(env_ids, dtype=np.int64)
    n_total = int(env_ids.max()) + 1 if len(env_ids) else 0
    gen = torch.Generator().manual_seed(seed)
    nj = model.nq - 7
    # one row per env id: row e only depends on (seed, e) because rows are drawn in order
    angles = (torch.rand((n_total, nj), generator=gen, dtype=torch.float64)*0.2 - 0.1).numpy()
    gen_phase = torch.Generator().manual_seed(seed + 1)
    phase = (torch.rand((n_total,), generator=gen_phase, dtype=torch.float64)*2*np.pi).numpy()
    qpos = np.tile(model.key_qpos, (len(env_ids), 1))
    qpos[:, 7:] += angles[env_ids]
    qvel = np.tile(model.key_qvel, (len(env_ids), 1))
    return qpos, qvel, phase[env_ids]


def bind_host_to_device(device):
    """Pin this process to the CPUs NVML lists as local to GPU ``device`` (its NUMA node), so that
    the pinned host buffers of ``step_host`` -- allocated after this call, first-touch -- and the
    host-side controller sit on the socket the GPU's PCIe root hangs off.  On a two-socket box with
    eight ranks the sensor download otherwise crosses the socket interconnect for half of them.
    Returns the CPU list it bound to, or None when NVML / the affinity call is unavailable (a
    single-socket host, a container without the permission): nothing else changes then."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63)//64)
        cpus = {64*w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:  # pylint: disable=broad-except
        return None


def gather_env_statistics(local, world_size):
    """all_gather of a per-env statistics tensor ``[n_local, k]`` -> ``[n_total, k]``.

    Shards may differ in length by one, so sizes are exchanged first."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return local
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world_size)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device))
    longest = int(max(s.item() for s in sizes))
    padded = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)])
