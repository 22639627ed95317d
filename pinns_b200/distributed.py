"""Data-parallel sharding of the collocation points (new capability; the reference has no
data parallelism -- SURVEY.md sections 2.3, 8e).

One process per GPU.  Collocation points are split into contiguous per-rank shards (or each
rank draws the counter range [first, first+n) of the job-wide Philox stream), theta / Adam
moments are replicated, and the packed vector [grad | dlambda | partial sums] is combined by
ONE sum-allreduce (NCCL over NVLink on GPUs; gloo in the CPU tests) per step.  The data term
lives on rank 0 only (data_weight 0 elsewhere) so the sum counts it once.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [first, first+count) of n_total points for `rank`; sizes differ by at most 1."""
    base, rem = divmod(int(n_total), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def data_weight(rank: int) -> float:
    return 1.0 if rank == 0 else 0.0


def allreduce_sum_(tensor, group=None):
    """In-place sum-allreduce of the packed vector when a process group is up and has >1 rank."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


class DataParallelStepper:
    """Drives one Engine per rank: loss+grad kernel -> one allreduce -> replicated Adam update."""

    def __init__(self, engine, rank: int = 0, world: int = 1, group=None):
        self.engine = engine
        self.rank, self.world, self.group = rank, world, group
        engine.set_data_weight(data_weight(rank))
        self._packed = engine.packed_tensor() if world > 1 else None
        self._l1 = None

    def loss_grad_device(self):
        eng = self.engine
        if self.world > 1 and eng_loss_is_l1(eng):
            import torch
            ptr = eng.l1_pass1()           # local sum|f| -> job-wide sum before the seeds are formed
            if self._l1 is None:
                self._l1 = eng.device_view(ptr, 1)
            allreduce_sum_(self._l1, self.group)
        eng.loss_grad_device()
        if self.world > 1:
            allreduce_sum_(self._packed, self.group)

    def adam_step(self):
        self.loss_grad_device()
        self.engine.adam_apply()


def eng_loss_is_l1(engine) -> bool:
    return getattr(engine, "loss_kind", "") in ("v3", "v3_l1sq")
