"""Single-box replacement of the reference's MPI sweep farm (reference:
Burgers/continuous_identification/scheduler.py:48-127,:157-169).

The reference runs one independent training script per free GPU, dispatching the Cartesian product of the
hyper-parameter lists from an MPI master that polls NVML every 30 s.  On one 8-GPU box none of that machinery is
needed: a queue of scenarios, one worker slot per GPU, each run a subprocess pinned to its GPU by the last positional
argument (the drivers' `gpu` argv, e.g. Abgrall_ADMM.py:414-420).  No tensors cross process boundaries, as in the
reference.  This is orchestration only; it is not on the data path.
"""
from __future__ import annotations

import itertools
import subprocess
import sys
import time
from typing import Dict, List, Sequence


def get_combinations(params: Dict[str, Sequence]) -> List[dict]:
    """Cartesian product of the hyper-parameter lists, in the reference's order (scheduler.py:48-68)."""
    keys = list(params)
    return [dict(zip(keys, vals)) for vals in itertools.product(*[params[k] for k in keys])]


def schedule_runs(command: Sequence[str], scenarios: List[dict], gpus: Sequence[int], arg_order: Sequence[str],
                  poll_s: float = 0.2, popen=subprocess.Popen) -> List[dict]:
    """Run `command + [scenario[k] for k in arg_order] + [gpu]` for every scenario, at most one per GPU at a time.
    Returns one record per scenario: {scenario, gpu, returncode}.  A crashed run is reported, not hidden
    (the reference's workers report RUN_FINISHED whatever happened, scheduler.py:165-169)."""
    pending = list(enumerate(scenarios))
    running = {}
    done: List[dict] = [None] * len(scenarios)
    free = list(gpus)
    while pending or running:
        while pending and free:
            idx, sc = pending.pop(0)
            gpu = free.pop(0)
            argv = list(command) + [str(sc[k]) for k in arg_order] + [str(gpu)]
            running[idx] = (popen(argv), gpu, sc)
        for idx in list(running):
            proc, gpu, sc = running[idx]
            rc = proc.poll()
            if rc is not None:
                done[idx] = {"scenario": sc, "gpu": gpu, "returncode": rc}
                free.append(gpu)
                del running[idx]
        if running:
            time.sleep(poll_s)
    return done


if __name__ == "__main__":  # python -m pinns_b200.sweep script.py  (hyper-parameters of scheduler.py:147-150)
    import torch
    script = sys.argv[1]
    grid = {"N_u": [100, 200], "N_f": [1000, 5000], "rho": [10.0, 40.0], "epochs": [100000]}
    res = schedule_runs([sys.executable, script], get_combinations(grid), list(range(max(torch.cuda.device_count(), 1))),
                        ["N_u", "N_f", "rho", "epochs"])
    for r in res:
        print(r)
