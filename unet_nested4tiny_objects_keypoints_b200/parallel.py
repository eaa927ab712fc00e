"""Data-parallel host logic.  The reference's only multi-GPU mechanism is the
single-process ``torch.nn.DataParallel`` (trainer/trainer.py:283-285,336-338: scatter the batch,
replicate the weights every forward, reduce-add the gradients to GPU 0, BatchNorm statistics per
replica).  Here: one process per GPU, a full replica each, the batch sharded by rank, and exactly
one collective per step — ``all_reduce(SUM)`` of the flat fp32 gradient buffer (553 260 elements,
2.2 MB) — followed by the same optimizer step everywhere with ``grad_scale = 1/world``.

Everything in this module is backend-agnostic ``torch.distributed`` plumbing (NCCL on the GPUs,
gloo in the CPU tests); there is no data-path collective in inference.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def world_and_rank(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def shard_range(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of this rank's images.  Like the reference (``assert batch_size % device_count == 0``,
    trainer.py:285) the global batch must divide evenly."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by the world size {world} (trainer.py:285)")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    b, e = shard_range(t.shape[0], world, rank)
    return t[b:e]


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> float:
    """Sums the flat gradient buffer over the ranks in place and returns the scale (1/world) the
    optimizer must apply (the fused AdamW takes it as ``grad_scale``; no extra pass over the buffer)."""
    world, _ = world_and_rank(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_parameters(model, src: int = 0, group=None) -> None:
    """Initial weight/buffer sync of the replicas (DataParallel re-broadcasts every forward, trainer.py:338)."""
    world, _ = world_and_rank(group)
    if world == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)
