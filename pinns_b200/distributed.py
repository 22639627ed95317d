"""Data-parallel sharding of the collocation points (new capability; the reference has no
data parallelism -- SURVEY.md sections 2.3, 8e).

One process per GPU.  Collocation points are split into contiguous per-rank shards (or each
rank draws the counter range [first, first+n) of the job-wide Philox stream), theta / Adam
moments are replicated, and the packed vector [grad | dlambda | partial sums] is combined by
ONE sum over ranks per step.  On the fused path that sum happens INSIDE the reduction kernel:
every rank stores its partial vector into its peers' receive slots over NVLink (buffers shared
through CUDA IPC), waits for their flags and adds the slots in rank order, so a step is two
launches (residual+grad kernel, reduction+exchange+Adam) and no collective call
(`attach_peer_memory`).  Everywhere else it is one sum-allreduce (NCCL on GPUs, gloo in the CPU
tests).  The data term lives on rank 0 only (data_weight 0 elsewhere) so the sum counts it once.
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


def attach_peer_memory(engine, rank: int, world: int, group=None) -> bool:
    """Gathers the engines' CUDA IPC handles over the process group and joins them into one peer-memory exchange group.
    Returns False (and changes nothing) when the engine is not on the fused path, the job has one rank or more than
    eight, the L1^2 loss needs its scalar allreduce between two passes, or the loss is INF-L2's un-squared data norm
    (V1): its data gradient joins the packed vector AFTER the reduction kernel whenever rank 0's shard is too large to
    carry the data batches in warps of their own, i.e. after the in-kernel exchange -- the allreduce route sums the
    finished vector instead (the C ABI refuses the combination too: PINN_E_STATE)."""
    import torch.distributed as dist
    ok = (world > 1 and world <= 8 and engine.kernel_path == "fused" and not eng_loss_is_l1(engine)
          and getattr(engine, "loss_kind", "") not in ("v1", "v1_inf_l2")
          and dist.is_available() and dist.is_initialized())
    if not ok:
        return False
    # every rank must end up in the same mode: a rank that cannot export / map (IPC disabled in its container, no P2P
    # path) votes the whole group back to the allreduce
    try:
        mine = engine.comm_export()
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    good = all(h is not None for h in handles)
    if good:
        try:
            engine.comm_attach(rank, world, handles)
        except Exception:
            good = False
            try:
                engine.comm_detach()   # close whatever was mapped before the failure
            except Exception:
                pass
    votes = [None] * world
    dist.all_gather_object(votes, good, group=group)
    if not all(votes):
        if good:
            engine.comm_detach()
        return False
    dist.barrier(group=group)          # every rank has mapped its peers before anybody stores into them
    return True


def detach_peer_memory(engine, group=None) -> None:
    """Leaves the exchange group in the order CUDA IPC asks for: every rank unmaps its peers' buffers, and only after
    a barrier may anybody free (close) the buffer it exported."""
    import torch.distributed as dist
    if getattr(engine, "comm_attached", False):
        engine.comm_detach()
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)


class DataParallelStepper:
    """Drives one Engine per rank: loss+grad kernel -> sum over ranks -> replicated Adam update."""

    def __init__(self, engine, rank: int = 0, world: int = 1, group=None, peer_memory: bool = False):
        self.engine = engine
        self.rank, self.world, self.group = rank, world, group
        engine.set_data_weight(data_weight(rank))
        self._packed = engine.packed_tensor() if world > 1 else None
        self._l1 = None
        self.peer_memory = bool(peer_memory) and attach_peer_memory(engine, rank, world, group)

    def close(self):
        """Call on every rank before the engine is destroyed (see detach_peer_memory)."""
        if self.world > 1:
            detach_peer_memory(self.engine, self.group)
        self.peer_memory = False

    def loss_grad_device(self):
        eng = self.engine
        if self.world > 1 and eng_loss_is_l1(eng):
            import torch
            ptr = eng.l1_pass1()           # local sum|f| -> job-wide sum before the seeds are formed
            if self._l1 is None:
                self._l1 = eng.device_view(ptr, 1)
            allreduce_sum_(self._l1, self.group)
        eng.loss_grad_device()        # with peer memory attached the packed vector already holds the sum over ranks
        if self.world > 1 and not self.peer_memory:
            allreduce_sum_(self._packed, self.group)

    def adam_step(self):
        if self.peer_memory or self.world == 1:
            # residual+grad kernel, then ONE kernel: reduction (+ exchange over NVLink) + Adam -- the same two-launch
            # route on one GPU and on eight
            self.engine.adam_steps(1)
            return
        self.loss_grad_device()
        self.engine.adam_apply()


def eng_loss_is_l1(engine) -> bool:
    return getattr(engine, "loss_kind", "") in ("v3", "v3_l1sq")
