"""A/B of the fused kernel's tanh variants on B200 (libraries built into scripts/variants/ by scripts/build_tanh_variants.sh):
gradient error against the reference-run fixtures (the ADMM ones cancel 3-4 digits of f), against the fp64 oracle on
random cases, and the 16 Mi-point step time.   python scripts/tanh_variants.py [child <lib>]"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import numpy as np, torch
    from oracle import tf_graph as tg
    from pinns_b200 import Engine
    from tests.helpers import ENGINE_LOSS, REF_RUNS, load_ref_fixture, make_case, make_engine, ref_problem, rel_err
    for name in REF_RUNS:
        fx = load_ref_fixture(name)
        p = ref_problem(name, fx)
        if len(p.layers) != 10 or p.layers[1] != 20:
            continue
        last = max(int(k[5:].split("_")[0]) for k in fx if k.startswith("stage") and k.endswith("_theta"))
        eng = Engine(p.layers, p.lb, p.ub, pde=p.pde, loss=ENGINE_LOSS[p.loss], lambda1=p.lam1, lambda2=p.lam2, rho=p.rho)
        eng.set_params(np.float32(fx["stage%d_theta" % last]))
        eng.set_data(fx["X_u"], fx["u_data"])
        eng.set_collocation(fx["vec_X_f"])
        if "vec_z" in fx:
            eng.admm_set_state(fx["vec_z"], fx["vec_gamma"])
        loss, grad = eng.loss_grad()
        P = eng.num_params
        _, f = eng.predict(fx["vec_X_f"])
        print("  ref %-9s grad err/|g| %.2e  loss rel %.2e  f max-rel %.2e" % (
            name, np.linalg.norm(grad[:P] - fx["vec_grad"]) / np.linalg.norm(fx["vec_grad"]),
            abs(loss - fx["vec_loss"]) / abs(fx["vec_loss"]), np.abs(f - fx["vec_f"]).max() / np.abs(fx["vec_f"]).max()))
    B20 = [2] + [20] * 8 + [1]
    for loss in (tg.LOSS_V4, tg.LOSS_V5):
        c = make_case(tg.PDE_BURGERS, B20, loss, 100, 10456, seed=2)
        ref = tg.evaluate(c["theta"], c["prob"], c["X_u"], c["u"], c["X_f"], c["z"], c["gamma"])
        eng = make_engine(c)
        l, g = eng.loss_grad()
        print("  oracle %-4s N_f=10456 grad L2-rel %.2e loss rel %.2e" % (loss, rel_err(g[:eng.num_params], ref.grad), abs(l - ref.loss) / abs(ref.loss)))
    eng = Engine(B20, [-1, 0], [1, 0.99], loss="v4", lambda2=0.01 / np.pi)
    eng.use_torch_stream()
    from bench import make_theta, make_data
    eng.set_params(make_theta()); eng.set_data(*make_data())
    eng.sample_collocation(1234, 0, 1 << 24)
    eng.adam_steps(3); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.adam_steps(5); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 5)
    print("  16 Mi points: %.3f ms/step  %.1f M points/s" % (best, (1 << 24) / best / 1e3))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for lib in sorted(glob.glob(os.path.join(ROOT, "scripts", "variants", "libpinn_t[0-9]*.so"))):
            print(os.path.basename(lib), flush=True)
            env = dict(os.environ, PINN_B200_LIB=lib)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
