#!/usr/bin/env python
"""bench.py -- residual+grad collocation points/sec of the Burgers PINN hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic collocation points:
fused forward + Taylor derivatives + Burgers residual + MSE loss + full parameter gradient
(+ the N_u data term), one allreduce of the packed gradient when N>1, and the TF-1 Adam update.
Workload (BASELINE.json config 4): layers [2,20x8,1], nu = 0.01/pi, x~U[-1,1), t~U[0,0.99)
from the job-wide Philox stream (seed 1234), N_f = 16*2^20 points PER GPU (weak scaling; the
134 MB of inputs per GPU exceed the 126 MB L2, so no flush is needed between iterations).

`value`  : whole-job points/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the public class API with HOST buffers: each step feeds the
           step's collocation points from pinned host memory (pinn_feed_collocation: chunked H2D on a
           copy stream, the kernel starts on chunk k while chunk k+1 is on the bus) and reads the loss back.
`roofline`: FP32-FMA bound (>= 5700 FLOP/B, compute bound): algorithmic
           FLOPs (68 320 per point) / measured duration of the fused kernel, against an FFMA-only
           micro-kernel measured on this GPU in this run (MEASURED_PEAKS.json has no fp32 entry).
`cpu_baseline` / `--impl reference`: the oracle's torch-fp32 op-for-op restatement of the
           reference TF graph (+ TF-1 Adam) on the box's host cores (TensorFlow is not installed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LAYERS = [2] + [20] * 8 + [1]
LB = np.array([-1.0, 0.0])
UB = np.array([1.0, 0.99])
NU = 0.01 / np.pi
F_POINT = 68320.0            # algorithmic FLOPs per point: 6*S*P_w - 2*S*n0*n1, S=4, P_w=2860
N_U = 100
SEED = 1234
METRIC = "residual+grad collocation points/sec, Burgers PINN [2,20x8,1]"
UNIT = "points/s"


def workload_name(nf_per_gpu):
    return ("burgers-inference-sweep: layers [2,20x8,1], nu=0.01/pi, MSE residual loss + N_u=100 data term + Adam, "
            "N_f=%d per GPU (weak), synthetic U[-1,1)x[0,0.99) Philox seed 1234" % nf_per_gpu)


def make_theta():
    from pinns_b200.models import xavier_init_flat
    rng = np.random.default_rng(SEED)
    th = xavier_init_flat(LAYERS, rng)
    return th


def make_data():
    rng = np.random.default_rng(SEED + 1)
    X_u = LB + (UB - LB) * rng.random((N_U, 2))
    u = -np.sin(np.pi * X_u[:, 0:1])
    return X_u, u


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) during the timed region."""

    def __init__(self, device, period=0.1):
        super().__init__(daemon=True)
        self.device, self.period = device, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_reference_run(steps, warmup, n_sample, threads=None):
    """The reference's CPU path, restated: nested reverse sweeps + TF-1 Adam in torch fp32."""
    import torch
    from oracle import tf_graph as tg
    from oracle.optim import TF1Adam
    if threads is None:
        threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rng = np.random.default_rng(SEED)
    theta = make_theta().astype(np.float32)
    X_u, u = make_data()
    X_f = LB + (UB - LB) * rng.random((n_sample, 2))
    prob = tg.Problem(LAYERS, LB, UB, pde=tg.PDE_BURGERS, loss=tg.LOSS_V4, lam1=1.0, lam2=NU)
    opt = TF1Adam(theta.size, dtype=np.float32)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ev = tg.evaluate(theta, prob, X_u, u, X_f, dtype=torch.float32)
        theta = opt.step(theta, ev.grad.astype(np.float32))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = float(np.sum(times))
    return {"value": n_sample * len(times) / total, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d steps of %d points (torch %s fp32, reverse-over-reverse like tf.gradients, + TF-1 Adam)"
                      % (len(times), n_sample, torch.__version__),
            "ms_per_step": 1e3 * total / len(times)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    n_sample = args.ref_points
    res = cpu_reference_run(args.steps, args.warmup, n_sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.nf), "reference_sample_points_per_step": n_sample,
                   "note": "TensorFlow 1.x is not installed; the reference's CPU graph is timed as its op-for-op "
                           "torch-fp32 restatement (oracle/tf_graph.py) on this box's host cores"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pinns_b200 import Engine
    from pinns_b200.distributed import DataParallelStepper

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nf = args.nf
    nf_global = nf * world
    eng = Engine(LAYERS, LB, UB, pde="burgers", loss="v4", lambda1=1.0, lambda2=NU, device=local_rank)
    eng.use_torch_stream()
    eng.set_params(make_theta())
    X_u, u = make_data()
    eng.set_data(X_u, u)
    eng.adam_config(lr=1e-3)
    eng.sample_collocation(SEED, rank * nf, nf, nf_global)   # rank r owns counters [r*nf, (r+1)*nf)
    # N > 1: the sum over ranks happens inside the reduction kernel through peer memory (NVLink), no collective launch
    stepper = DataParallelStepper(eng, rank, world, peer_memory=(world > 1 and not args.nccl))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------- device-resident throughput ----------------
    for _ in range(max(args.warmup, 3)):
        stepper.adam_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.kernel_timing(True)
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        stepper.adam_step()
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count - l0
    k_ms, k_n = eng.kernel_time()
    eng.kernel_timing(False)
    sampler.stop()
    ms_step = ms_total / args.steps
    value = nf_global / (ms_step * 1e-3)
    loss_now = eng.loss_from_packed(eng.packed_tensor()[eng.num_params + 2:].tolist(), nf_global, "v4")

    # ---------------- end to end through the public class API with host buffers ----------------
    from pinns_b200.models import PhysicsInformedNN
    e2e_steps = max(3, min(args.steps, 10))
    host = torch.empty((nf, 2), dtype=torch.float32).pin_memory()
    host.copy_(torch.from_numpy(eng.get_collocation()))
    model = PhysicsInformedNN.__new__(PhysicsInformedNN)   # the reference class surface over the same engine
    model.engine, model.layers, model.lb, model.ub, model.nu = eng, LAYERS, LB, UB, NU

    def e2e_step():
        # H2D of this step's points from pinned memory, loss+grad(+allreduce)+Adam, D2H of the loss
        return model.train_step_from_host(host, nf_global=nf_global, stepper=stepper)

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1)) / e2e_steps
    e2e_value = nf_global / (e2e_ms * 1e-3)

    if stepper.peer_memory and eng.comm_status()[1]:
        raise RuntimeError("peer-memory exchange: a peer's flag did not arrive in time")
    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel ----------------
    peak_meas = Engine.measure_fma_peak(local_rank)
    sm_max = sampler.max_mhz or 1965
    peak_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = nf * F_POINT / (k_avg_ms * 1e-3) / 1e12
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel (per point x points per launch)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_fused_traffic.json")))
        traffic = float(tr["dram_bytes_per_point"]) * nf
    except Exception:
        pass
    roofline = {"bound": "fp32_fma", "achieved": achieved, "peak": peak_meas, "unit": "TFLOP/s",
                "frac": achieved / peak_meas if peak_meas > 0 else None,
                "peak_source": "FFMA-only micro-kernel measured on this GPU in this run (of measured); "
                               "MEASURED_PEAKS.json has no fp32 entry",
                "peak_nominal": peak_nominal, "frac_of_nominal": achieved / peak_nominal,
                "kernel": "pinn_fused_kernel<20,true>", "kernel_ms": k_avg_ms, "kernel_launches": k_n,
                "kernel_share_of_step": k_avg_ms / ms_step,
                "flops_per_point": F_POINT, "points_per_launch": nf, "traffic": traffic,
                "traffic_source": "profiles/r01_fused_traffic.json (ncu dram__bytes_read+write per point x points per launch)",
                "hbm_algorithmic_bytes_per_point": 8}
    cpu = None
    if world == 1:  # the CPU sample is timed on rank 0 at N = 1 only (bench contract); null in the scaling runs
        cpu = cpu_reference_run(steps=2, warmup=1, n_sample=args.cpu_points)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(nf), "layers": LAYERS, "nf_per_gpu": nf, "nf_global": nf_global,
                   "l2_policy": "inputs (%d MB per GPU) larger than the 126 MB L2; no flush" % (nf * 8 // 2 ** 20),
                   "parallelism": "dp%d" % world, "kernel_path": eng.kernel_path, "loss_after": loss_now,
                   "rank_sum": ("peer-memory exchange inside the reduction kernel" if stepper.peer_memory else
                                ("one NCCL allreduce of the packed vector" if world > 1 else "single GPU"))},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nf * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "api": "PhysicsInformedNN.train_step_from_host(pinned X_f): chunked H2D feed overlapped with the kernel + loss/grad(+allreduce) + Adam + loss D2H"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nf", type=int, default=16 * 2 ** 20, help="collocation points per GPU")
    ap.add_argument("--nccl", action="store_true", help="N > 1: combine the ranks with one NCCL allreduce instead of peer memory")
    ap.add_argument("--cpu-points", type=int, default=2 ** 20, help="points per step of the cpu_baseline sample")
    ap.add_argument("--ref-points", type=int, default=2 ** 18, help="points per step of --impl reference")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
