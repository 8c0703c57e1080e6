"""Population sharding over the GPUs of one box.

Trials are independent (the reference treats them as independent Pool tasks,
refinement/async_evaluatable.py:38-50), so the candidate axis is block-partitioned
over the ranks with no data-path collective; the only exchange is one all-gather of
the per-candidate residual rows f64[K_rank, 1+S] (NCCL over NVLink on GPUs, gloo in
the CPU tests).  A single pattern / trial is never split across GPUs.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def bind_process_to_device(device_index: int) -> bool:
    """Pin this process to the CPU cores next to GPU ``device_index`` (NVML's ideal affinity), so that pinned host
    buffers and the copy engines' traffic stay on the GPU's own NUMA node.  One process per GPU (the launch model
    of this package) otherwise lands anywhere; with 8 ranks streaming patterns to the host that costs PCIe
    bandwidth.  Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() \
                else device_index
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def shard_bounds(n_items: int, world_size: int) -> List[Tuple[int, int]]:
    """[lo, hi) of every rank: ceil(K / ws) candidates per rank, the tail ranks may be short/empty"""
    per = -(-int(n_items) // int(world_size)) if n_items > 0 else 0
    return [(min(r * per, n_items), min((r + 1) * per, n_items)) for r in range(world_size)]


def shard(items: Sequence, rank: int, world_size: int):
    lo, hi = shard_bounds(len(items), world_size)[rank]
    return items[lo:hi]


def gather_rows(local, n_total: int, world_size: int, group=None):
    """All-gather row blocks of unequal length: ``local`` is a torch tensor [k_rank, C] on the
    device the process group communicates on; returns [n_total, C] on every rank."""
    import torch
    import torch.distributed as dist
    per = -(-int(n_total) // int(world_size)) if n_total > 0 else 0
    cols = local.shape[1]
    padded = torch.zeros((per, cols), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world_size * per, cols), dtype=local.dtype, device=local.device)
    if world_size > 1:
        dist.all_gather_into_tensor(out, padded, group=group)
    else:
        out.copy_(padded)
    return out[:n_total]


def evaluate_population_sharded(mixtures: Sequence, evaluate, device=None, group=None) -> np.ndarray:
    """Each rank evaluates its block of the population with ``evaluate(list) -> f64[k, 1+S]``
    (on the GPU path: ``pyxrd_b200.batched.population_residuals``) and all ranks receive the full
    [K, 1+S] table, bit-identical to a single-rank evaluation of the same candidates."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard(mixtures, rank, world)
    n_spec = len(mixtures[0].specimens)
    res = np.asarray(evaluate(mine), dtype=np.float64).reshape(len(mine), 1 + n_spec) if len(mine) else \
        np.zeros((0, 1 + n_spec))
    t = torch.from_numpy(np.ascontiguousarray(res))
    if device is not None:
        t = t.to(device)
    return gather_rows(t, len(mixtures), world, group).cpu().numpy()


def evaluate_solutions_sharded(x, evaluate, device=None, group=None) -> np.ndarray:
    """The template path over several GPUs: every rank evaluates its block of the solution matrix ``x`` f64[K, d] with
    ``evaluate(x_block) -> f64[k] or f64[k, C]`` (on the GPU path ``PopulationTemplate.optimized_residuals`` /
    ``.residuals`` of the rank's own template) and all ranks receive the values of all K solutions in order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    lo, hi = shard_bounds(len(x), world)[rank]
    values = np.asarray(evaluate(x[lo:hi]), dtype=np.float64) if hi > lo else np.zeros(0)
    cols = int(values.shape[1]) if values.ndim == 2 else 1
    if world > 1:                                  # empty tail ranks learn the column count from rank 0
        c = torch.tensor([cols], dtype=torch.int64, device=device if device is not None else "cpu")
        dist.broadcast(c, src=0, group=group)
        cols = int(c.item())
    squeeze = values.ndim == 1 and cols == 1
    t = torch.from_numpy(np.ascontiguousarray(values.reshape(hi - lo, cols)))
    if device is not None:
        t = t.to(device)
    out = gather_rows(t, len(x), world, group).cpu().numpy()
    return out[:, 0] if squeeze else out
