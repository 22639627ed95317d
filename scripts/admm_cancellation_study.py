"""How large is the gradient error of a float32 evaluation in the ADMM cancellation regime -- as a DISTRIBUTION, not one draw?

The reference evaluates the loss gradient right after its z/gamma update on the same batch (Burgers_ADMM_batch.py:204-210
then :118-119; Abgrall_ADMM.py:225-226): the seed rho (f - z) + gamma collapses to +-1/N_f and 3-4 digits of f cancel, so
rounding-level differences in f (1e-7) become 1e-5 ... 1e-4 of |g|.  The committed fixture is ONE such state.  Here: from the
fixture's parameters, fresh seeded batches; z/gamma updated from the float64 residuals (rounded to float32 like the
reference's variables); float64 gradient at that state = truth; against it (a) the oracle's float32 evaluation of the
reference graph (torch CPU), (b) the CUDA kernel, for each tanh variant in scripts/variants/ (and the in-tree build).

    python scripts/admm_cancellation_study.py            (parent: runs one child per library)
"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NB = 8


def states(name):
    import numpy as np, torch
    from oracle import tf_graph as tg
    from tests.helpers import load_ref_fixture, ref_problem
    fx = load_ref_fixture(name)
    p = ref_problem(name, fx)
    last = max(int(k[5:].split("_")[0]) for k in fx if k.startswith("stage") and k.endswith("_theta"))
    theta = np.float32(fx["stage%d_theta" % last])
    n_f = fx["vec_X_f"].shape[0]
    rng = np.random.default_rng(77)
    z, gamma = fx["vec_z"].astype(np.float64), fx["vec_gamma"].astype(np.float64)
    out = []
    for b in range(NB):
        X_f = p.lb + (p.ub - p.lb) * rng.random((n_f, 2))
        f = tg.evaluate(theta, p, fx["X_u"], fx["u_data"], X_f, z=z, gamma=gamma, want_grad=False).f
        z2, g2 = tg.admm_update(f, z, gamma, p.rho, n_f)
        z2, g2 = z2.astype(np.float32), g2.astype(np.float32)
        ref = tg.evaluate(theta, p, fx["X_u"], fx["u_data"], X_f, z=z2, gamma=g2)
        out.append((X_f, z2, g2, ref.grad))
    return fx, p, theta, out


def child(do_cpu):
    import numpy as np, torch
    from oracle import tf_graph as tg
    from pinns_b200 import Engine
    from tests.helpers import ENGINE_LOSS
    res = {}
    for name in ("ID-ADMMb", "AB-ADMM"):
        fx, p, theta, sts = states(name)
        eng = Engine(p.layers, p.lb, p.ub, pde=p.pde, loss=ENGINE_LOSS[p.loss], lambda1=p.lam1, lambda2=p.lam2, rho=p.rho)
        eng.set_params(theta)
        eng.set_data(fx["X_u"], fx["u_data"])
        e_gpu, e_cpu = [], []
        for X_f, z2, g2, gref in sts:
            eng.set_collocation(X_f)
            eng.admm_set_state(z2, g2)
            _, g = eng.loss_grad()
            e_gpu.append(float(np.linalg.norm(g[:eng.num_params] - gref) / np.linalg.norm(gref)))
            if do_cpu:
                g32 = tg.evaluate(theta, p, fx["X_u"], fx["u_data"], X_f, z=z2, gamma=g2, dtype=torch.float32).grad
                e_cpu.append(float(np.linalg.norm(g32 - gref) / np.linalg.norm(gref)))
        res[name] = {"gpu": e_gpu, "cpu_fp32_graph": e_cpu}
    print("RESULT " + json.dumps(res))


def fmt(v):
    import numpy as np
    v = np.asarray(v)
    return "median %.2e  mean %.2e  min %.2e  max %.2e" % (np.median(v), v.mean(), v.min(), v.max())


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(sys.argv[2] == "1")
    else:
        libs = [None] + sorted(glob.glob(os.path.join(ROOT, "scripts", "variants", "libpinn_t[0-9]*.so")))
        for k, lib in enumerate(libs):
            env = dict(os.environ)
            if lib:
                env["PINN_B200_LIB"] = lib
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "child", "1" if k == 0 else "0"], env=env,
                               capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if not line:
                print(lib, "FAILED", r.stderr[-2000:])
                continue
            res = json.loads(line[0][7:])
            for name, d in res.items():
                if d["cpu_fp32_graph"]:
                    print("%-10s %-22s gradient error / |g| over %d states: %s" % (name, "torch fp32 ref graph", NB, fmt(d["cpu_fp32_graph"])))
                print("%-10s %-22s gradient error / |g| over %d states: %s" % (name, os.path.basename(lib) if lib else "in-tree library", NB, fmt(d["gpu"])))
