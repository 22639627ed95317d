"""world_size-2 gloo test of the data-parallel host logic: contiguous shards, the 1/N_f
factors of the WHOLE job, the data term on rank 0 only, ONE sum-allreduce of the packed vector.  The per-rank
"kernel" here is the CPU oracle evaluated on the rank's shard -- this checks the sharding arithmetic the GPU
path relies on (tests/test_multigpu_gpu.py repeats it with the CUDA engine when 2 GPUs are present)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tf_graph as tg
from pinns_b200.distributed import allreduce_sum_, shard_range
from tests.helpers import make_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, loss, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        c = make_case(tg.PDE_BURGERS, [2, 10, 10, 1], loss, 16, 101, seed=9)
        n_f = c["X_f"].shape[0]
        first, cnt = shard_range(n_f, rank, world)
        Xs = c["X_f"][first:first + cnt]
        prob = c["prob"]
        # residual term of the shard with the JOB-wide 1/N_f: evaluate with an empty data term and rescale
        zero_u = np.zeros((0, 2)), np.zeros((0, 1))
        th = torch.from_numpy(c["theta"].astype(np.float64)).requires_grad_(True)
        W, b = tg.unpack(th, prob.layers)
        x = tg.feed(Xs[:, 0:1]).requires_grad_(True); t = tg.feed(Xs[:, 1:2]).requires_grad_(True)
        f = tg.net_f_burgers(x, t, W, b, prob.lb, prob.ub, float(np.float32(prob.lam1)), float(np.float32(prob.lam2)))
        packed = torch.zeros(th.numel() + 2, dtype=torch.float64)
        if loss == tg.LOSS_V3:
            s_local = torch.tensor([float(f.detach().abs().sum())], dtype=torch.float64)
            allreduce_sum_(s_local)                                   # the extra scalar allreduce of the L1^2 loss
            res = (2.0 / n_f) * s_local[0] * f.abs().sum()            # d/dtheta of this equals the seeded backward
            packed[-1] = float(f.detach().abs().sum())
        else:
            res = (f * f).sum() / n_f
            packed[-2] = float(res.detach())
        g = torch.autograd.grad(res, th)[0]
        packed[:th.numel()] = g
        if rank == 0:                                                 # data term lives on rank 0 only
            th2 = torch.from_numpy(c["theta"].astype(np.float64)).requires_grad_(True)
            W2, b2 = tg.unpack(th2, prob.layers)
            up = tg.net_u(tg.feed(c["X_u"][:, 0:1]), tg.feed(c["X_u"][:, 1:2]), W2, b2, prob.lb, prob.ub)
            ld = ((tg.feed(c["u"]) - up) ** 2).sum() / c["X_u"].shape[0]
            packed[:th.numel()] += torch.autograd.grad(ld, th2)[0]
            packed[-2] += float(ld.detach())
        allreduce_sum_(packed)
        if rank == 0:
            out.put(packed.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("loss", [tg.LOSS_V4, tg.LOSS_V3])
def test_two_rank_allreduce_reproduces_single_process_gradient(loss):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, loss, out)) for r in range(2)]
    for p in procs:
        p.start()
    packed = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c = make_case(tg.PDE_BURGERS, [2, 10, 10, 1], loss, 16, 101, seed=9)
    ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"])
    P = ref.grad.size
    assert np.abs(packed[:P] - ref.grad).max() <= 1e-10 * np.abs(ref.grad).max()
    if loss == tg.LOSS_V4:
        assert abs(packed[-2] - ref.loss) <= 1e-12 * abs(ref.loss)
    else:
        n_f = c["X_f"].shape[0]
        assert abs(packed[-2] + packed[-1] ** 2 / n_f - ref.loss) <= 1e-12 * abs(ref.loss)


class _StubEngine:
    """Stands in for the CUDA engine in the host-side group logic: exports a handle unless told to fail."""

    def __init__(self, fail_export=False, path="fused", loss_kind="v4"):
        self.fail_export, self.kernel_path, self.loss_kind = fail_export, path, loss_kind
        self.attached = self.detached = False
        self.weight = None

    def set_data_weight(self, w):
        self.weight = w

    def packed_tensor(self):
        return torch.zeros(4)

    def comm_export(self):
        if self.fail_export:
            raise RuntimeError("IPC not available")
        return b"\0" * 64

    def comm_attach(self, rank, world, handles):
        assert len(handles) == world and all(len(h) == 64 for h in handles)
        self.attached = True

    def comm_detach(self):
        self.detached = True


def _vote_worker(rank, world, port, scenario, out):
    from pinns_b200.distributed import DataParallelStepper
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _StubEngine(fail_export=(scenario == "one-rank-fails" and rank == 1),
                          path=("generic" if scenario == "not-fused" else "fused"),
                          loss_kind=("v3" if scenario == "l1" else "v4"))
        st = DataParallelStepper(eng, rank, world, peer_memory=True)
        out.put((rank, st.peer_memory, eng.attached, eng.detached, eng.weight))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("scenario,expect", [("all-good", True), ("one-rank-fails", False), ("not-fused", False), ("l1", False)])
def test_peer_memory_group_is_all_or_nothing(scenario, expect):
    """attach_peer_memory: every rank ends in the same mode; one rank that cannot export sends the whole group back to the
    allreduce (nobody stays attached); paths / losses without the in-kernel exchange never try; data term on rank 0 only."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vote_worker, args=(r, 2, port, scenario, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, peer, attached, detached, weight in got:
        assert peer == expect
        assert weight == (1.0 if rank == 0 else 0.0)
        if expect:
            assert attached and not detached
        else:
            assert not attached or detached
