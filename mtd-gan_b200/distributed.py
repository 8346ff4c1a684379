"""Data-parallel MTD-GAN training: one process per GPU, NCCL over NVLink.

The reference has no working multi-GPU path for PCGrad (nn.DataParallel hides `shared_parameters()`,
SURVEY §2.1), so the semantics are defined here: R ranks x B patches == one process on the concatenated
R*B batch.  Losses are batch means, so every gradient is averaged over ranks; because PCGrad is non-linear,
the per-task shared gradients are reduced BEFORE the projection:

    reduce-scatter g_0, g_1, g_2 (each rank keeps a 1/R shard)  ->  partial 3x3 Gram on the shard
    -> all-reduce 16 doubles -> identical coefficient solve on every rank (same `random.shuffle` stream)
    -> combine on the shard -> all-gather the merged gradient

which moves 2 x 114 MB less than three all-reduces.  Task-specific and generator gradients are plain
all-reduce(mean).  The collective choreography is independent of the arithmetic back-end (`ops` argument),
so world_size-2 gloo tests on CPU exercise it with torch stand-ins.

Overlap (CUDA): every collective of the discriminator step runs on ONE dedicated communication stream, forked from
and joined back into the compute stream with events (inside a captured step these are graph edges):

    compute : bwd(task-specific) | bwd(task 0) | bwd(task 1) | bwd(task 2) |wait| Gram, solve, combine |wait| AdamW
    comm    :                    AR(ts grads)  |     RS(g0)  |     RS(g1)  | RS(g2)        AR(gram)   AG(merged)

so only the last reduce-scatter, the 16-double Gram all-reduce and the all-gather are exposed.  A task's gradients
are gathered into the flat collective operand by ONE kernel (`mtd_segments_scale_copy`, 1/R folded in) instead of
`torch.zeros` + one `copy_` per parameter.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def init(backend: Optional[str] = None):
    """Initialise the default process group from the torchrun environment (idempotent)."""
    if dist.is_initialized():
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    dist.init_process_group(backend=backend, init_method="env://")


def active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world() -> int:
    return dist.get_world_size() if active() else 1


def rank() -> int:
    return dist.get_rank() if active() else 0


def shard_bounds(n: int, world_size: int, r: int):
    """Element range of rank r's shard of a length-n vector padded to a multiple of world_size."""
    per = (n + world_size - 1) // world_size
    return per, min(n, r * per), min(n, (r + 1) * per)


_comm_streams: dict = {}


def comm_stream(device) -> "torch.cuda.Stream":
    """The one communication stream of `device` (all collectives of a step are issued on it, in program order)."""
    key = torch.device(device).index
    st = _comm_streams.get(key)
    if st is None:
        st = _comm_streams[key] = torch.cuda.Stream(device=device)
    return st


def flatten(tensors: Sequence[torch.Tensor], pad_to: int = 1, scale: float = 1.0) -> torch.Tensor:
    """scale * cat(tensors), zero padded to a multiple of `pad_to`.  CUDA tensors: one kernel launch."""
    n = sum(t.numel() for t in tensors)
    total = (n + pad_to - 1) // pad_to * pad_to
    if tensors[0].is_cuda and tensors[0].dtype == torch.float32:
        from . import _ext
        from .weight_methods import _chunk_table, _float_bits
        dev = tensors[0].device
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        if total > n:
            flat[n:].zero_()
        rows, off, keep = [], 0, []
        sb = _float_bits(scale)
        for t in tensors:
            t = t.contiguous()
            keep.append(t)
            rows.append([t.data_ptr(), 0, 0, 0, flat.data_ptr() + 4 * off, t.numel(), sb, 0])
            off += t.numel()
        seg = _ext.device_table(rows, torch.int64, dev)
        chunks, n_chunks = _chunk_table([t.numel() for t in tensors], dev)
        _ext.call("mtd_segments_scale_copy", _ext.ptr(seg), _ext.ptr(chunks), n_chunks, _ext.stream())
        return flat
    flat = torch.zeros(total, dtype=tensors[0].dtype, device=tensors[0].device)
    off = 0
    for t in tensors:
        flat[off:off + t.numel()].copy_(t.reshape(-1))
        off += t.numel()
    if scale != 1.0:
        flat.mul_(scale)
    return flat


def unflatten(flat: torch.Tensor, like: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    out, off = [], 0
    for t in like:
        out.append(flat[off:off + t.numel()].view(t.shape))
        off += t.numel()
    return out


def allreduce_mean_grads(params: Sequence[torch.Tensor]):
    """p.grad <- mean over ranks, for every parameter that has a gradient (one flat all-reduce)."""
    if not active():
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    flat = flatten([p.grad for p in ps], scale=1.0 / dist.get_world_size())       # mean = sum of pre-scaled gradients
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    for p, g in zip(ps, unflatten(flat, [p.grad for p in ps])):
        p.grad = g


def allreduce_mean_list(tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    if not active():
        return list(tensors)
    flat = flatten(tensors, scale=1.0 / dist.get_world_size())
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return unflatten(flat, tensors)


class _Pending:
    """A collective issued on the communication stream; `wait()` joins it into the current stream."""

    def __init__(self, result, stream, keep):
        self.result, self.stream, self.keep = result, stream, keep

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.keep = None
        return self.result


def allreduce_mean_list_async(tensors: Sequence[torch.Tensor]) -> _Pending:
    """allreduce_mean_list whose all-reduce runs on the communication stream (CUDA) while the caller keeps computing."""
    tensors = list(tensors)
    if not active():
        return _Pending(tensors, None, None)
    flat = flatten(tensors, scale=1.0 / dist.get_world_size())
    views = unflatten(flat, tensors)
    if not flat.is_cuda:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return _Pending(views, None, None)
    comm = comm_stream(flat.device)
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return _Pending(views, comm, (flat, tensors))


class ShardedPCGrad:
    """Distributed PCGrad with the reduce-scatter of task k overlapping the backward pass of task k+1.

        pipe = ShardedPCGrad()
        for each task:  pipe.submit(task_grads)         # flatten (compute stream) + reduce-scatter (comm stream)
        merged = pipe.finish(orders, mean, gram_fn, solve_combine_fn)

    gram_fn / solve_combine_fn as in `pcgrad_sharded`.  Sums (not means) are reduced: the projection coefficients are
    invariant to a common scale and the 1/R of the batch mean is applied in the combine."""

    def __init__(self):
        self.shards, self.like, self.keep, self.comm = [], None, [], None
        self.R, self.r = dist.get_world_size(), dist.get_rank()

    def submit(self, task_grads: Sequence[torch.Tensor]):
        task_grads = list(task_grads)
        if self.like is None:
            self.like = task_grads
            self.n = sum(g.numel() for g in task_grads)
            self.per, _, _ = shard_bounds(self.n, self.R, self.r)
        flat = flatten(task_grads, pad_to=self.R)                     # zero padded to R * per
        if flat.is_cuda:
            self.comm = comm_stream(flat.device)
            self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                shard = torch.empty(self.per, dtype=flat.dtype, device=flat.device)
                dist.reduce_scatter_tensor(shard, flat, op=dist.ReduceOp.SUM)
            self.keep.append((flat, task_grads))                       # alive until the join in finish()
        else:
            shard = torch.empty(self.per, dtype=flat.dtype, device=flat.device)
            dist.reduce_scatter_tensor(shard, flat, op=dist.ReduceOp.SUM)
        self.shards.append(shard)

    def finish(self, orders, mean: bool, gram_fn, solve_combine_fn) -> List[torch.Tensor]:
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)        # all reduce-scatters (and anything queued before them)
        self.keep = []
        gram = gram_fn(self.shards)
        dist.all_reduce(gram, op=dist.ReduceOp.SUM)
        merged_shard = solve_combine_fn(self.shards, gram, orders, mean, 1.0 / self.R)
        full = torch.empty(self.R * self.per, dtype=merged_shard.dtype, device=merged_shard.device)
        dist.all_gather_into_tensor(full, merged_shard)
        return unflatten(full[:self.n], self.like)


def pcgrad_sharded(task_grads: Sequence[Sequence[torch.Tensor]], orders, mean: bool, gram_fn, solve_combine_fn):
    """Distributed PCGrad on per-task gradient lists (all tasks have a gradient for every parameter).

    gram_fn(shards: List[Tensor]) -> float64 tensor [16] with entry [a*4+b] (a <= b) = <shard_a, shard_b>
    solve_combine_fn(shards, gram16, orders, mean, scale) -> merged shard (Tensor)
    Returns the merged gradients as views of one flat buffer, identical on every rank.
    (All tasks at once; the training path submits task by task -- see ShardedPCGrad.)
    """
    pipe = ShardedPCGrad()
    for tg in task_grads:
        pipe.submit(tg)
    return pipe.finish(orders, mean, gram_fn, solve_combine_fn)
