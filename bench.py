#!/usr/bin/env python
"""bench.py -- residual+grad collocation points/sec of the Burgers PINN hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over one batch of synthetic collocation points:
fused forward + Taylor derivatives + Burgers residual + MSE loss + full parameter gradient
(+ the N_u data term), one sum of the packed gradient over the ranks when N>1, and the TF-1 Adam update.
Workload (BASELINE.json config 4, north_star "near-linear scaling to 8xB200 on 64M collocation points"):
layers [2,20x8,1], nu = 0.01/pi, x~U[-1,1), t~U[0,0.99) from the job-wide Philox stream (seed 1234),
N_f = 64*2^20 points for the WHOLE JOB, sharded over the N GPUs (strong scaling: 64 Mi points on one GPU,
8 Mi per GPU at N = 8; every shard's inputs (>= 64 MB ... 512 MB) plus the kernel's 117 MB working set exceed the
126 MB L2, so no flush is needed between iterations).

`value`  : whole-job points/s, inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    : the same metric through the public class API with HOST buffers: each step feeds the
           step's collocation points from pinned host memory (pinn_feed_collocation: chunked H2D on a
           copy stream, the kernel starts on chunk k while chunk k+1 is on the bus) and reads the loss back.
`roofline`: FP32-FMA bound (>= 5700 FLOP/B, compute bound): algorithmic
           FLOPs (68 320 per point) / measured duration of the fused kernel, against an FFMA-only
           micro-kernel measured on this GPU in this run (MEASURED_PEAKS.json has no fp32 entry).
`sweep`  : BASELINE config 4's collocation sweep (1, 4, 16, 64 Mi points global) at this N.
`extra`  : the other BASELINE configs at the sizes the reference runs them (config 1: 10 456 + 100 points, V1 loss;
           config 3: Euler [2,200x5,3]; config 5: [2,128x8,1] on the tcgen05 path), N = 1 only -- and config 5 at
           16 Mi points global on every N.
`config.rank_sum_check` (N > 1): the N-rank summed packed vector of a fixed 1 Mi-point batch, through peer memory
           and through NCCL, against the single-GPU vector over the same Philox counters; the run FAILS above 5e-6.
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
F_POINT_W128 = 2759680.0     # [2,128x8,1]
F_POINT_EULER = 2895600.0    # [2,200x5,3], S = 3
N_U = 100
SEED = 1234
METRIC = "residual+grad collocation points/sec, Burgers PINN [2,20x8,1]"
UNIT = "points/s"
RANK_SUM_TOL = 5e-6


def workload_name(nf_global):
    return ("burgers-inference-sweep: layers [2,20x8,1], nu=0.01/pi, MSE residual loss + N_u=100 data term + Adam, "
            "N_f=%d for the whole job (strong scaling), synthetic U[-1,1)x[0,0.99) Philox seed 1234" % nf_global)


def xavier_flat(layers, rng):
    """initialize_NN / xavier_init (INF-L2:79-94) as a flat float32 vector; numpy only, so that the reference arm
    imports nothing of the product (same draws as pinns_b200.models.xavier_init_flat and oracle.tf_graph.xavier_init)."""
    parts = []
    for l in range(len(layers) - 1):
        n_in, n_out = int(layers[l]), int(layers[l + 1])
        w = rng.standard_normal(n_in * n_out)
        bad = np.abs(w) > 2.0
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2.0
        parts.append((w * np.sqrt(2 / (n_in + n_out))).astype(np.float32))
        parts.append(np.zeros(n_out, np.float32))
    return np.concatenate(parts)


def make_theta(layers=LAYERS):
    return xavier_flat(layers, np.random.default_rng(SEED))


def make_data(n_out=1, n_u=N_U):
    rng = np.random.default_rng(SEED + 1)
    X_u = LB + (UB - LB) * rng.random((n_u, 2))
    u = -np.sin(np.pi * X_u[:, 0:1])
    if n_out > 1:
        u = np.hstack([1.0 + 0.2 * u, u, 2.5 + 0.1 * u])[:, :n_out]
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
    """The reference's CPU path, restated: nested reverse sweeps + TF-1 Adam in torch fp32.  Imports the oracle only."""
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
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.nf_global), "reference_sample_points_per_step": n_sample,
                   "note": "TensorFlow 1.x is not installed; the reference's CPU graph is timed as its op-for-op "
                           "torch-fp32 restatement (oracle/tf_graph.py) on this box's host cores; each step is a bounded "
                           "sample of the workload's batch (points/s on the CPU is higher at the smaller batch)"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def timed_steps(torch, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(steps)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extra_config(torch, dev_index, name, layers, pde, loss, n_u, n_f, steps, fpp, peak_fp32, path="auto", ref=None):
    """One of the other BASELINE configs at the reference's own size: Adam steps on a fixed batch, device resident
    (the working set fits the L2 by design: the reference feeds the same N_f points every step)."""
    from pinns_b200 import Engine
    eng = Engine(layers, LB, UB, pde=pde, loss=loss, lambda1=1.0, lambda2=NU, rho=40.0, device=dev_index, path=path)
    eng.use_torch_stream()
    eng.set_params(make_theta(layers))
    X_u, u = make_data(layers[-1], n_u)
    eng.set_data(X_u, u)
    eng.adam_config(lr=1e-3)
    eng.sample_collocation(SEED, 0, n_f)
    if loss == "v5":
        eng.admm_init()
    eng.adam_steps(3)
    torch.cuda.synchronize()
    l0 = eng.launch_count
    ms = timed_steps(torch, eng.adam_steps, steps)      # the step as a user runs it (no events between the kernels: at these sizes
    launches = eng.launch_count - l0                    # they would serialise the programmatic dependent launches)
    eng.kernel_timing(True)                             # ... then the same steps again with CUDA events around the main kernel
    timed_steps(torch, eng.adam_steps, steps)
    k_ms, k_n = eng.kernel_time()
    eng.kernel_timing(False)
    kpath = eng.kernel_path
    small = kpath == "fused" and (n_f + 7) // 8 + (n_u + 7) // 8 <= 148 * 9   # one 8-point batch per warp (8 or 9 warps per SM): four lanes per point
    kname = {"fused": "pinn_fused_small_kernel<20,true>" if small else "pinn_fused_kernel<20,true>",
             "generic": "pinn_generic_kernel<%d>" % (4 if pde == "burgers" else 3),
             "tensor": "pinn_tc_kernel<%s> (tcgen05 + TMA, 3xTF32)" % ("4,1" if pde == "burgers" else "3,3")}[kpath]
    tflops = n_f * fpp / (ms * 1e-3) / 1e12
    out = {"workload": name, "layers": "[2,%dx%d,%d]" % (layers[1], len(layers) - 2, layers[-1]), "loss": loss, "n_f": n_f,
           "n_u": n_u, "ms_per_step": ms, "points_per_s": n_f / (ms * 1e-3), "steps": steps, "launches_per_step": launches / steps,
           "kernel": kname, "kernel_path": kpath, "kernel_ms": k_ms / max(k_n, 1), "algorithmic_tflops": tflops,
           "frac_fp32_peak": tflops / peak_fp32, "reference_lines": ref}
    if kpath == "tensor":
        peak_tf32 = tf32_peak()
        out["tensor_tflops_issued"] = 3.0 * tflops      # 3xTF32: hi*hi + hi*lo + lo*hi
        out["frac_tf32_peak"] = 3.0 * tflops / peak_tf32
        out["tf32_peak"] = peak_tf32
    eng.close()
    return out


def tf32_peak():
    """Dense TF32 tensor peak used as the tensor-path denominator: half of the driver-measured sustained bf16 figure
    (MEASURED_PEAKS.json), else half of the profiling guide's nominal 2250."""
    try:
        return 0.5 * float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        return 0.5 * 2250.0


def rank_sum_check(torch, dist, rank, world, local_rank, n_total=1 << 20):
    """Multi-GPU CORRECTNESS, on the hardware and in the driver-visible line: the packed vector [grad | dlambda | sums]
    of a fixed job-wide batch, summed over the ranks (a) inside the reduction kernel through NVLink peer memory and
    (b) by one NCCL allreduce, against (c) the single-GPU vector over the same Philox counters (computed on rank 0)."""
    from pinns_b200 import Engine
    from pinns_b200.distributed import DataParallelStepper, shard_range
    theta = make_theta()
    X_u, u = make_data()
    first, cnt = shard_range(n_total, rank, world)
    out = {"points": n_total, "tolerance": RANK_SUM_TOL}
    vecs = {}
    for mode in ("peer_memory", "nccl"):
        eng = Engine(LAYERS, LB, UB, pde="burgers", loss="v4", lambda1=1.0, lambda2=NU, device=local_rank)
        eng.use_torch_stream()
        eng.set_params(theta)
        eng.set_data(X_u, u)
        eng.sample_collocation(SEED, first, cnt, n_total)
        st = DataParallelStepper(eng, rank, world, peer_memory=(mode == "peer_memory"))
        st.loss_grad_device()
        torch.cuda.synchronize()
        v = eng.packed_tensor().cpu().numpy().copy()
        st.adam_step()
        st.adam_step()
        th = eng.get_params()
        gathered = [None] * world
        dist.all_gather_object(gathered, (v.tobytes(), th.tobytes()))
        vecs[mode] = v
        out[mode] = {"attached": bool(st.peer_memory), "bitwise_equal_across_ranks": all(g == gathered[0] for g in gathered)}
        torch.cuda.synchronize()
        st.close()          # unmap the peers' buffers everywhere, barrier, and only then free the exported buffer
        eng.close()
    if rank == 0:
        eng = Engine(LAYERS, LB, UB, pde="burgers", loss="v4", lambda1=1.0, lambda2=NU, device=local_rank)
        eng.use_torch_stream()
        eng.set_params(theta)
        eng.set_data(X_u, u)
        eng.sample_collocation(SEED, 0, n_total, n_total)
        eng.loss_grad_device()
        torch.cuda.synchronize()
        one = eng.packed_tensor().cpu().numpy().copy()
        eng.close()
        P = one.size - 10
        ok = True
        for mode in ("peer_memory", "nccl"):
            v = vecs[mode]
            eg = float(np.linalg.norm(v[:P].astype(np.float64) - one[:P]) / np.linalg.norm(one[:P].astype(np.float64)))
            es = float(np.max(np.abs(v[P + 2:P + 7].astype(np.float64) - one[P + 2:P + 7])
                              / np.maximum(np.abs(one[P + 2:P + 7].astype(np.float64)), 1e-30)))
            out[mode].update({"rel_err_grad": eg, "rel_err_sums": es})
            ok = ok and eg <= RANK_SUM_TOL and es <= RANK_SUM_TOL and out[mode]["bitwise_equal_across_ranks"]
        ok = ok and out["peer_memory"]["attached"]
        out["pass"] = bool(ok)
    flag = torch.tensor([1 if (rank != 0 or out.get("pass")) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["pass"] = bool(flag.item())
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pinns_b200 import Engine
    from pinns_b200.distributed import DataParallelStepper, shard_range

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    nf_global = args.nf_global
    first, nf = shard_range(nf_global, rank, world)

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

    check = None
    if world > 1:
        check = rank_sum_check(torch, dist, rank, world, local_rank)
        if not check["pass"]:
            if rank == 0:
                print(json.dumps({"error": "rank_sum_check failed", "rank_sum_check": check}), flush=True)
            raise SystemExit(3)

    eng = Engine(LAYERS, LB, UB, pde="burgers", loss="v4", lambda1=1.0, lambda2=NU, device=local_rank)
    eng.use_torch_stream()
    eng.set_params(make_theta())
    X_u, u = make_data()
    eng.set_data(X_u, u)
    eng.adam_config(lr=1e-3)
    eng.sample_collocation(SEED, first, nf, nf_global)   # rank r owns the counters of its contiguous shard
    # N > 1: the sum over ranks happens inside the reduction kernel through peer memory (NVLink), no collective launch
    stepper = DataParallelStepper(eng, rank, world, peer_memory=(world > 1 and not args.nccl))

    # ---------------- device-resident throughput ----------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
        stepper.adam_step()
    sampler = ClockSampler(local_rank)   # (NVML initialisation takes tens of ms and differs from rank to rank: before the barrier,
    eng.kernel_timing(True)              #  or the ranks enter the timed region skewed and the first exchange waits it out)
    barrier()
    sampler.start()
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
        # H2D of this step's points from pinned memory, loss+grad(+sum over ranks)+Adam, D2H of the loss
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
    del host

    # ---------------- BASELINE config 4's sweep at this N (device resident, same engine) ----------------
    sweep = []
    for mi in (1, 4, 16, 64):
        g = mi << 20
        if g > nf_global:
            continue
        f_s, n_s = shard_range(g, rank, world)
        eng.sample_collocation(SEED, f_s, n_s, g)
        for _ in range(3):
            stepper.adam_step()
        barrier()
        n_t = 5 if mi >= 16 else 20
        e0.record()
        for _ in range(n_t):
            stepper.adam_step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / n_t
        sweep.append({"nf_global": g, "nf_per_gpu": n_s, "ms_per_step": ms, "points_per_s": g / (ms * 1e-3), "steps": n_t})

    eng.synchronize()   # raises if a peer-memory exchange timed out anywhere above
    stepper_peer = stepper.peer_memory
    stepper.close()
    eng.close()

    # ---------------- BASELINE config 5 at 16 Mi points global on every N: tcgen05 path + one allreduce per step ----------------
    peak_meas = Engine.measure_fma_peak(local_rank)
    extra = {}
    if not args.no_extra:
        W128 = [2] + [128] * 8 + [1]
        g5 = args.nf_wide
        f5, n5 = shard_range(g5, rank, world)
        e5 = Engine(W128, LB, UB, pde="burgers", loss="v4", lambda1=1.0, lambda2=NU, device=local_rank)
        e5.use_torch_stream()
        e5.set_params(make_theta(W128))
        e5.set_data(X_u, u)
        e5.adam_config(lr=1e-3)
        e5.sample_collocation(SEED, f5, n5, g5)
        st5 = DataParallelStepper(e5, rank, world, peer_memory=False)
        for _ in range(2):
            st5.adam_step()
        barrier()
        e5.kernel_timing(True)
        e0.record()
        for _ in range(3):
            st5.adam_step()
        e1.record()
        barrier()
        ms5 = max_over_ranks(e0.elapsed_time(e1)) / 3
        k5_ms, k5_n = e5.kernel_time()
        tfl = g5 * F_POINT_W128 / (ms5 * 1e-3) / 1e12
        extra["config5_wide128_16Mi"] = {
            "workload": "BASELINE config 5: layers [2,128x8,1], N_f=%d global (strong), MSE loss + Adam" % g5,
            "n_gpus": world, "nf_per_gpu": n5, "ms_per_step": ms5, "points_per_s": g5 / (ms5 * 1e-3), "steps": 3,
            "kernel_path": e5.kernel_path, "kernel": "pinn_tc_kernel<4,1> (tcgen05 + TMA, 3xTF32)", "kernel_ms": k5_ms / max(k5_n, 1),
            "rank_sum": "one NCCL allreduce of the packed vector (464 KB)" if world > 1 else "single GPU",
            "algorithmic_tflops": tfl, "tensor_tflops_issued": 3 * tfl, "frac_tf32_peak": 3 * tfl / (tf32_peak() * world),
            "frac_fp32_peak": tfl / (peak_meas * world), "tf32_peak_per_gpu": tf32_peak()}
        e5.synchronize()
        e5.close()
        if world == 1:
            B200x8 = [2] + [200] * 8 + [1]
            EUL = [2] + [200] * 5 + [3]
            extra["config1_inference"] = extra_config(torch, local_rank, "BASELINE config 1: INF-L2 step, 10 456 + 100 points", LAYERS,
                                                      "burgers", "v1", 100, 10456, 200, F_POINT, peak_meas,
                                                      ref="Hwan_L2Regularization_Burgers.py:68-69,:126-141")
            extra["identification_batch_1000"] = extra_config(torch, local_rank, "AB-ADMM / ID-* batch: 1000 + 100 points, ADMM loss", LAYERS,
                                                              "burgers", "v5", 100, 1000, 200, F_POINT, peak_meas,
                                                              ref="Abgrall_ADMM.py:129-130,:200-252")
            extra["config3_euler_1000"] = extra_config(torch, local_rank, "BASELINE config 3: Euler ADMM step, 1000 + 200 points", EUL,
                                                       "euler", "v5", 200, 1000, 50, F_POINT_EULER, peak_meas,
                                                       ref="Euler_ADMM.py:128-133,:217-258")
            extra["identification_batch_5000"] = extra_config(torch, local_rank, "ID-ADMMb batch: 5000 + 100 points, ADMM loss", LAYERS,
                                                              "burgers", "v5", 100, 5000, 200, F_POINT, peak_meas,
                                                              ref="Burgers_ADMM_batch.py:118-119,:29-34")
            extra["config3_euler_65536"] = extra_config(torch, local_rank, "Euler [2,200x5,3] at 65 536 points (tensor path from 8192 points on)",
                                                        EUL, "euler", "v5", 200, 65536, 5, F_POINT_EULER, peak_meas, ref="Euler_ADMM.py:176-198")
            extra["config3_euler_65536_generic"] = extra_config(torch, local_rank, "the same on the FP32 generic kernel", EUL, "euler", "v5",
                                                                200, 65536, 3, F_POINT_EULER, peak_meas, path="generic", ref="Euler_ADMM.py:176-198")
            extra["abgrall_l2_wide200_1000"] = extra_config(torch, local_rank, "AB-L2 batch: [2,200x8,1], 1000 + 100 points", B200x8,
                                                            "burgers", "v4", 100, 1000, 50, 6731200.0, peak_meas,
                                                            ref="Abgrall_L2.py:59-60,:247")
            extra["abgrall_l2_wide200_262144"] = extra_config(torch, local_rank, "[2,200x8,1] at 262 144 points (tensor path)", B200x8,
                                                              "burgers", "v4", 100, 262144, 3, 6731200.0, peak_meas, ref="Abgrall_L2.py:247")
            extra["config5_wide128_262144"] = extra_config(torch, local_rank, "BASELINE config 5 net at 262 144 points", W128, "burgers",
                                                           "v4", 100, 262144, 5, F_POINT_W128, peak_meas, ref="north_star config 5")
    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel ----------------
    sm_max = sampler.max_mhz or 1965
    peak_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = nf * F_POINT / (k_avg_ms * 1e-3) / 1e12
    traffic, traffic_src = None, None
    for name in ("r02_fused_traffic.json", "r01_fused_traffic.json"):
        try:  # DRAM bytes per launch from the committed ncu --set full capture of this kernel (per point x points per launch)
            tr = json.load(open(os.path.join(ROOT, "profiles", name)))
            traffic = float(tr["dram_bytes_per_point"]) * nf
            traffic_src = "profiles/%s (ncu dram__bytes_read+write per point x points per launch)" % name
            break
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
                "traffic_source": traffic_src, "hbm_algorithmic_bytes_per_point": 8}
    cpu = None
    if world == 1:  # the CPU sample is timed on rank 0 at N = 1 only (bench contract); null in the scaling runs
        cpu = cpu_reference_run(steps=2, warmup=1, n_sample=args.cpu_points)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(nf_global), "layers": LAYERS, "nf_per_gpu": nf, "nf_global": nf_global,
                   "l2_policy": "inputs (%d MB per GPU) + the kernel's 117 MB stash/accumulator working set exceed the 126 MB L2; no flush"
                                % (nf * 8 // 2 ** 20),
                   "parallelism": "dp%d" % world, "kernel_path": "fused", "loss_after": loss_now,
                   "rank_sum": ("peer-memory exchange inside the reduction kernel" if stepper_peer else
                                ("one NCCL allreduce of the packed vector" if world > 1 else "single GPU")),
                   "rank_sum_check": check},
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nf * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms, "steps": e2e_steps,
                "api": "PhysicsInformedNN.train_step_from_host(pinned X_f): chunked H2D feed overlapped with the kernel + loss/grad(+sum over ranks) + Adam + loss D2H"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "sweep": sweep,
        "extra": extra,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nf-global", type=int, default=64 * 2 ** 20, help="collocation points of the whole job")
    ap.add_argument("--nf-wide", type=int, default=16 * 2 ** 20, help="collocation points (global) of the config-5 entry")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` configs")
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
