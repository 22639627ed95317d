"""CPU study of the fused path's tanh (no GPU needed): the Taylor-forward + reverse schedule of the kernels evaluated in
float32 numpy on the ID-ADMMb reference fixture -- the case whose ADMM seed rho (f - z) + gamma cancels 3-4 digits of f --
with different tanh implementations, against the fixture's float64 gradient.  The hardware's ex2.approx / rcp.approx are
modelled as the correctly rounded function times (1 + e), |e| <= 2 ulp / 1 ulp uniformly random (their documented bounds).

    python scripts/tanh_study.py
"""
import sys

import numpy as np

sys.path.insert(0, '/root/repo')
from oracle import taylor as ty                                   # noqa: E402
from oracle import tf_graph as tg                                 # noqa: E402
from tests.helpers import load_ref_fixture, ref_problem          # noqa: E402

F = np.float32
rng = np.random.default_rng(0)


def ulp_noise(v, ulps):
    return (v * (1.0 + rng.uniform(-ulps, ulps, v.shape) * 2.0 ** -24)).astype(F)


def ex2_approx(y):
    return ulp_noise(np.exp2(y.astype(np.float64)), 2.0)


def tanh_fast(x):
    """1 - 2/(exp(2x)+1): the round-1 form."""
    y = np.minimum(x * F(2.885390081777927), F(60.0)).astype(F)
    e = ex2_approx(y)
    d = (e + F(1.0)).astype(F)
    r = (F(1.0) / d).astype(F)          # Newton-refined reciprocal: correctly rounded to ~1 ulp
    return (F(1.0) - F(2.0) * r).astype(F)


# odd minimax-like polynomial x * P(x^2) on |x| <= T (Taylor coefficients of tanh; the truncation error at T = 0.25 with
# five terms is 0.0088 T^10 ~ 8e-9 relative)
C = [F(-1.0 / 3.0), F(2.0 / 15.0), F(-17.0 / 315.0), F(62.0 / 2835.0), F(-1382.0 / 155925.0)]


def tanh_poly(x, nterms):
    s = (x * x).astype(F)
    p = C[nterms - 1]
    for k in range(nterms - 2, -1, -1):
        p = (p * s + C[k]).astype(F)
    return (x + x * (p * s).astype(F)).astype(F)


def tanh_hybrid(T, nterms):
    def f(x):
        return np.where(np.abs(x) < F(T), tanh_poly(x, nterms), tanh_fast(x)).astype(F)
    return f


def tanh_expm1(x):
    """(e - 1) / (e + 1) with e from ex2.approx: same exp error, no better at small x (kept for the record)."""
    y = np.minimum(x * F(2.885390081777927), F(60.0)).astype(F)
    e = ex2_approx(y)
    return ((e - F(1.0)) / (e + F(1.0))).astype(F)


def grad_f32(theta, prob, X_u, u_data, X_f, z, gamma, tanh):
    """oracle.taylor.evaluate in float32 with a pluggable tanh (H-stream reverse step like the kernels)."""
    theta = np.asarray(theta, F)
    Ws, bs = ty._unpack(theta, prob.layers)
    lb = prob.lb.astype(F)
    span = (prob.ub - prob.lb).astype(F)
    lam1, lam2, rho = F(prob.lam1), F(prob.lam2), F(prob.rho)
    L = len(Ws)

    def fwd(X, second):
        H = (F(2.0) * (X - lb) / span - F(1.0)).astype(F)
        Hx = np.zeros_like(H); Hx[:, 0] = F(2.0) / span[0]
        Ht = np.zeros_like(H); Ht[:, 1] = F(2.0) / span[1]
        Hxx = np.zeros_like(H)
        cache = []
        for l in range(L - 1):
            W, b = Ws[l], bs[l]
            Z = H @ W + b
            Zx, Zt, Zxx = Hx @ W, Ht @ W, Hxx @ W
            a = tanh(Z.astype(F))
            d1 = F(1.0) - a * a
            cache.append((H, Hx, Ht, Hxx))
            H, Hx, Ht, Hxx = a, d1 * Zx, d1 * Zt, d1 * (Zxx - F(2.0) * a * Zx * Zx)
            cache[-1] = cache[-1] + (H, Hx, Ht, Hxx)
        cache.append((H, Hx, Ht, Hxx))
        W, b = Ws[-1], bs[-1]
        return (H @ W + b, Hx @ W, Ht @ W, Hxx @ W), cache

    def rev(cache, Yb, Yxb, Ytb, Yxxb):
        gW, gb = [None] * L, [None] * L
        Hin, Hinx, Hint, Hinxx = cache[-1]
        W = Ws[-1]
        gW[-1] = Hin.T @ Yb + Hinx.T @ Yxb + Hint.T @ Ytb + Hinxx.T @ Yxxb
        gb[-1] = Yb.sum(0)
        Hb, Hxb, Htb, Hxxb = Yb @ W.T, Yxb @ W.T, Ytb @ W.T, Yxxb @ W.T
        for l in range(L - 2, -1, -1):
            Hin, Hinx, Hint, Hinxx, a, Hx, Ht, Hxx = cache[l]
            Zb, Zxb, Ztb, Zxxb = ty.reverse_step_hstream(a, Hx, Ht, Hxx, Hb, Hxb, Htb, Hxxb)
            gW[l] = Hin.T @ Zb + Hinx.T @ Zxb + Hint.T @ Ztb + Hinxx.T @ Zxxb
            gb[l] = Zb.sum(0)
            W = Ws[l]
            Hb, Hxb, Htb, Hxxb = Zb @ W.T, Zxb @ W.T, Ztb @ W.T, Zxxb @ W.T
        return np.concatenate([np.concatenate([gW[l].ravel(), gb[l].ravel()]) for l in range(L)]).astype(np.float64)

    Xu, ud, Xf = np.asarray(X_u, F), np.asarray(u_data, F), np.asarray(X_f, F)
    (Yu, _, _, _), cu = fwd(Xu, False)
    r = ud - Yu
    zero = np.zeros_like(Yu)
    g = rev(cu, (F(-2.0) * r / F(Xu.shape[0])).astype(F), zero, zero, zero)
    (Y, Yx, Yt, Yxx), cf = fwd(Xf, True)
    f = Yt + lam1 * Y * Yx - lam2 * Yxx
    fbar = (rho * (f - np.asarray(z, F)) + np.asarray(gamma, F)).astype(F)
    g = g + rev(cf, fbar * lam1 * Yx, fbar * lam1 * Y, fbar, -lam2 * fbar)
    return g, f


def main():
    for name in ("ID-ADMMb", "AB-ADMM"):
        fx = load_ref_fixture(name)
        p = ref_problem(name, fx)
        last = max(int(k[5:].split("_")[0]) for k in fx if k.startswith("stage") and k.endswith("_theta"))
        theta = np.float32(fx["stage%d_theta" % last])
        ref = fx["vec_grad"].astype(np.float64)
        print(name, "|g| = %.3e" % np.linalg.norm(ref))
        cands = [("np.tanh float32 (correctly rounded)", lambda x: np.tanh(x.astype(np.float64)).astype(F)),
                 ("fast form (round 1)", tanh_fast),
                 ("(e-1)/(e+1)", tanh_expm1)]
        for T in (0.125, 0.25, 0.5, 0.75):
            for nt in (3, 4, 5):
                cands.append(("hybrid poly |x|<%.3f, %d terms" % (T, nt), tanh_hybrid(T, nt)))
        for label, fn in cands:
            errs = []
            for rep in range(3):
                g, f = grad_f32(theta, p, fx["X_u"], fx["u_data"], fx["vec_X_f"], fx["vec_z"], fx["vec_gamma"], fn)
                errs.append(np.linalg.norm(g - ref) / np.linalg.norm(ref))
            print("  %-40s grad err / |g| = %.2e (max of 3 noise draws %.2e)" % (label, np.mean(errs), np.max(errs)))


if __name__ == "__main__":
    main()
