"""Population sharding over the GPUs of one box.

Trials are independent (the reference treats them as independent Pool tasks,
refinement/async_evaluatable.py:38-50), so the candidate axis is block-partitioned
over the ranks with no data-path collective; the only exchange is one all-gather of
the per-candidate residual rows f64[K_rank, 1+S] (NCCL over NVLink on GPUs, gloo in
the CPU tests).  A single pattern / trial is never split across GPUs.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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


_pad_buffers = {}
_peer_gathers = {}


class PeerGather(object):
    """All-gather of residual rows by direct stores into peer memory (NVLink / NVSwitch) instead of a NCCL collective.

    Every rank owns two gathered tables f64[world * per, cols] in torch symmetric memory (mapped into every process of the
    box at rendezvous; two, used alternately, so that the stores of generation g + 1 never land in a table a peer is
    still copying to its host for generation g).  A gather is: the library's kernel ``pyxrd_put_rows_to_peers`` stores
    the rank's rows into the table of EVERY rank at the rank's row offset, then one cross-GPU barrier on the signal pads
    (``_SymmetricMemory.barrier``).  Measured inside the bench step (c2, 2048 candidates per GPU, 32 KB of rows per rank):
    2 GPUs 0.803 ms against 0.801 ms with the NCCL all-gather, 8 GPUs 0.825 against 0.808 -- the exchange is ~25 us of
    latency either way, so NCCL stays the default and this path is an option (``ctx=`` of ``evaluate_solutions_sharded``,
    ``bench.py --gather peer``); both give the single-GPU table bit for bit."""

    def __init__(self, ctx, per: int, cols: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.ctx, self.per, self.cols = ctx, int(per), int(cols)
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.tables, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty((self.world * self.per, self.cols), dtype=torch.float64, device=device)
            t.zero_()
            self.tables.append(t)
            self.handles.append(symm_mem.rendezvous(t, self.group))
        self.turn = 0
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    def __call__(self, local, n_total: int):
        t, h = self.tables[self.turn], self.handles[self.turn]
        self.turn ^= 1
        if local.shape[0]:
            src = local if local.is_contiguous() else local.contiguous()
            self.ctx.put_rows_to_peers(src.data_ptr(), src.shape[0], self.cols, list(h.buffer_ptrs), self.rank * self.per)
        if not self.ctx.stream_is_torch_current():
            self.ctx.synchronize()                 # the stores run on the context's stream, the barrier on torch's
        h.barrier(channel=0)
        return t[:n_total]


def peer_gather_for(ctx, per: int, cols: int, device, group=None):
    """cached ``PeerGather`` of this shape, or None where symmetric memory is not available (CPU groups, gloo, a torch
    without it, GPUs without peer access): the caller then uses the NCCL all-gather"""
    key = (id(ctx), int(per), int(cols), str(device), id(group))
    if key not in _peer_gathers:
        try:
            _peer_gathers[key] = PeerGather(ctx, per, cols, device, group)
        except Exception:
            _peer_gathers[key] = None
    return _peer_gathers[key]


def gather_rows(local, n_total: int, world_size: int, group=None):
    """All-gather row blocks of unequal length: ``local`` is a torch tensor [k_rank, C] on the
    device the process group communicates on (on the GPU path: the zero-copy view of the library's residual buffer,
    ``Context.residuals_tensor``); returns [n_total, C] on every rank.  Ranks hold ceil(n / ws) rows (the tail ranks
    fewer): blocks are padded to that length in a reused staging buffer, one ``all_gather_into_tensor``.  The result
    is a view of a reused buffer: copy it (``_to_host``) before the next gather."""
    import torch
    import torch.distributed as dist
    per = -(-int(n_total) // int(world_size)) if n_total > 0 else 0
    cols = local.shape[1]
    if world_size <= 1:
        return local[:n_total]
    key = (per, cols, local.dtype, str(local.device), world_size)
    bufs = _pad_buffers.get(key)
    if bufs is None:
        if len(_pad_buffers) > 16:
            _pad_buffers.clear()
        bufs = _pad_buffers[key] = (torch.zeros((per, cols), dtype=local.dtype, device=local.device),
                                    torch.empty((world_size * per, cols), dtype=local.dtype, device=local.device))
    padded, out = bufs
    if local.shape[0] == per:
        padded = local if local.is_contiguous() else local.contiguous()
    else:
        padded[: local.shape[0]].copy_(local)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_total]


_pinned = {}        # (shape, dtype) -> reused pinned host buffer of _to_host


def _to_host(t) -> np.ndarray:
    """host copy of a gathered table (never a view of the reused gather buffer).  Device tables go through a reused
    PINNED staging buffer (one asynchronous copy + one stream synchronisation: a pageable ``.cpu()`` costs an allocation
    and a staged copy per generation)"""
    if not t.is_cuda:
        return t.clone().numpy()
    import torch
    key = (tuple(t.shape), t.dtype)
    buf = _pinned.get(key)
    if buf is None:
        if len(_pinned) > 16:
            _pinned.clear()
        buf = _pinned[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return buf.numpy().copy()


def _as_rows(values, rows: int, device):
    """(tensor [rows, C], squeeze): the rank's block as a 2-D torch tensor on ``device``; a torch tensor that already
    lives on the device (zero-copy view of the library's result buffer) is used as it is"""
    import torch
    if isinstance(values, torch.Tensor):
        t = values if values.dim() == 2 else values.reshape(rows, -1)
        return (t if device is None or t.device == torch.device(device) else t.to(device)), False
    values = np.asarray(values, dtype=np.float64)
    squeeze = values.ndim == 1
    t = torch.from_numpy(np.ascontiguousarray(values.reshape(rows, -1) if rows else values.reshape(0, max(values.shape[-1] if values.ndim == 2 else 1, 1))))
    return (t.to(device) if device is not None else t), squeeze


def evaluate_population_sharded(mixtures: Sequence, evaluate, device=None, group=None) -> np.ndarray:
    """Each rank evaluates its block of the population with ``evaluate(list) -> f64[k, 1+S]``
    (on the GPU path: ``pyxrd_b200.batched.population_residuals``) and all ranks receive the full
    [K, 1+S] table, bit-identical to a single-rank evaluation of the same candidates.  ``evaluate`` may return a
    torch tensor on ``device``: it is gathered as it is, without a trip through host memory."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard(mixtures, rank, world)
    n_spec = len(mixtures[0].specimens)
    values = evaluate(mine) if len(mine) else np.zeros((0, 1 + n_spec))
    t, _ = _as_rows(values, len(mine), device)
    return _to_host(gather_rows(t, len(mixtures), world, group))


def evaluate_solutions_sharded(x, evaluate, device=None, group=None, columns: Optional[int] = None, ctx=None) -> np.ndarray:
    """The template path over several GPUs (what ``RefineAsyncHelper.do_async_evaluation`` does with Pool tasks,
    refinement/refine_async_helper.py:16-23, async_evaluatable.py:38-50): every rank evaluates its block of the solution
    matrix ``x`` f64[K, d] with ``evaluate(x_block) -> f64[k] | f64[k, C]`` and all ranks receive the values of all K
    solutions in order.  On the GPU path ``evaluate`` is ``lambda xb: template.residuals(xb, device=True)`` /
    ``template.optimized_residuals(xb, device=True)`` of the rank's own template: the rank's rows are gathered straight
    from the library's device buffer (NCCL over NVLink), one device -> host copy of the gathered table at the end.
    ``columns``: C when known (saves the broadcast that tells empty tail ranks the column count).  ``ctx``: the library
    context the rows come from -- with it (and ``columns``) the exchange is ``PeerGather``: the rank's rows are stored
    straight into every rank's gathered table over NVLink by the library's own kernel, one barrier, no NCCL collective."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    lo, hi = shard_bounds(len(x), world)[rank]
    values = evaluate(x[lo:hi]) if hi > lo else np.zeros(0)
    if isinstance(values, torch.Tensor):
        cols, squeeze = (int(values.shape[1]) if values.dim() == 2 else 1), values.dim() == 1
    else:
        values = np.asarray(values, dtype=np.float64)
        cols, squeeze = (int(values.shape[1]) if values.ndim == 2 else 1), values.ndim == 1
    if columns is not None:
        cols = int(columns)
    elif world > 1:                                # empty tail ranks learn the column count from rank 0
        c = torch.tensor([cols], dtype=torch.int64, device=device if device is not None else "cpu")
        dist.broadcast(c, src=0, group=group)
        cols = int(c.item())
    if isinstance(values, torch.Tensor):
        t = values.reshape(hi - lo, cols)
        if device is not None and t.device != torch.device(device):
            t = t.to(device)
    else:
        t = torch.from_numpy(np.ascontiguousarray(values.reshape(hi - lo, cols)))
        if device is not None:
            t = t.to(device)
    peer = None
    if ctx is not None and columns is not None and world > 1 and t.is_cuda:
        peer = peer_gather_for(ctx, -(-len(x) // world), cols, t.device, group)
    out = _to_host(peer(t, len(x)) if peer is not None else gather_rows(t, len(x), world, group))
    return out[:, 0] if (squeeze and cols == 1) else out
