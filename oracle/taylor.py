"""Independent numpy fp64 restatement: Taylor-forward + one reverse sweep.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against runs of the reference's own
scripts over a TensorFlow-1 API stand-in (tests/test_reference_pin.py); TensorFlow itself unpinned.

This is the schedule the CUDA kernels use, written out
by hand with no autograd, so that it can be checked against ``oracle.tf_graph``
(which mirrors the reference's nested ``tf.gradients``: INF-L2:113-120,
AB-ADMM:170-180, EUL:176-198) and against finite differences.  Agreement of the
two restatements to ~1e-13 is asserted in tests/test_oracle.py.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from .tf_graph import (Evaluation, Problem, PDE_BURGERS, LOSS_V1, LOSS_V2, LOSS_V3, LOSS_V4, LOSS_V5,
                       LOSS_V6, LOSS_EULER_MSE)


def _unpack(theta, layers):
    Ws, bs, off = [], [], 0
    for l in range(len(layers) - 1):
        n_in, n_out = layers[l], layers[l + 1]
        Ws.append(theta[off:off + n_in * n_out].reshape(n_in, n_out))
        off += n_in * n_out
        bs.append(theta[off:off + n_out])
        off += n_out
    return Ws, bs


def _consts(prob):
    lb = prob.lb.astype(np.float32).astype(np.float64)
    span = (prob.ub - prob.lb).astype(np.float32).astype(np.float64)
    return lb, span


def forward(theta, prob: Problem, X, second_order: bool):
    """Propagate H, H_x, H_t (and H_xx) through the tanh MLP (INF-L2:96-107).
    Returns head outputs per stream and the per-layer cache for the reverse sweep."""
    Ws, bs = _unpack(theta, prob.layers)
    lb, span = _consts(prob)
    N = X.shape[0]
    H = 2.0 * (X - lb) / span - 1.0
    Hx = np.zeros_like(H); Hx[:, 0] = 2.0 / span[0]
    Ht = np.zeros_like(H); Ht[:, 1] = 2.0 / span[1]
    Hxx = np.zeros_like(H) if second_order else None
    cache = []
    L = len(Ws)
    for l in range(L - 1):
        W, b = Ws[l], bs[l]
        Z = H @ W + b
        Zx = Hx @ W
        Zt = Ht @ W
        Zxx = Hxx @ W if second_order else None
        a = np.tanh(Z)
        d1 = 1.0 - a * a
        d2 = -2.0 * a * d1
        cache.append((H, Hx, Ht, Hxx, a, Zx, Zt, Zxx))
        H, Hx, Ht = a, d1 * Zx, d1 * Zt
        if second_order:
            Hxx = d2 * Zx * Zx + d1 * Zxx
    W, b = Ws[-1], bs[-1]
    Y = H @ W + b
    Yx = Hx @ W
    Yt = Ht @ W
    Yxx = Hxx @ W if second_order else None
    cache.append((H, Hx, Ht, Hxx))
    return (Y, Yx, Yt, Yxx), cache


def reverse(theta, prob: Problem, cache, Ybar, Yxbar, Ytbar, Yxxbar):
    """One reverse sweep given the adjoints of the head outputs of every stream.
    Streams whose adjoint is None are skipped (data term: primal only)."""
    Ws, _ = _unpack(theta, prob.layers)
    L = len(Ws)
    gW = [None] * L
    gb = [None] * L
    Hin, Hinx, Hint, Hinxx = cache[-1]
    W = Ws[-1]

    def acc(*pairs):
        tot = 0.0
        for h, zb in pairs:
            if zb is not None and h is not None:
                tot = tot + h.T @ zb
        return tot

    gW[-1] = acc((Hin, Ybar), (Hinx, Yxbar), (Hint, Ytbar), (Hinxx, Yxxbar))
    gb[-1] = Ybar.sum(0)
    Hb = Ybar @ W.T
    Hxb = Yxbar @ W.T if Yxbar is not None else None
    Htb = Ytbar @ W.T if Ytbar is not None else None
    Hxxb = Yxxbar @ W.T if Yxxbar is not None else None
    for l in range(L - 2, -1, -1):
        Hin, Hinx, Hint, Hinxx, a, Zx, Zt, Zxx = cache[l]
        W = Ws[l]
        d1 = 1.0 - a * a
        d2 = -2.0 * a * d1
        d3 = -2.0 * d1 * (1.0 - 3.0 * a * a)
        Zb = d1 * Hb
        Zxb = Ztb = Zxxb = None
        if Hxb is not None:
            Zxb = d1 * Hxb
            Zb = Zb + d2 * Zx * Hxb
        if Htb is not None:
            Ztb = d1 * Htb
            Zb = Zb + d2 * Zt * Htb
        if Hxxb is not None:
            Zxxb = d1 * Hxxb
            Zxb = Zxb + 2.0 * d2 * Zx * Hxxb
            Zb = Zb + d2 * Zxx * Hxxb + d3 * Zx * Zx * Hxxb
        gW[l] = acc((Hin, Zb), (Hinx, Zxb), (Hint, Ztb), (Hinxx, Zxxb))
        gb[l] = Zb.sum(0)
        if l > 0:
            Hb = Zb @ W.T
            Hxb = Zxb @ W.T if Zxb is not None else None
            Htb = Ztb @ W.T if Ztb is not None else None
            Hxxb = Zxxb @ W.T if Zxxb is not None else None
    return np.concatenate([np.concatenate([gW[l].ravel(), gb[l].ravel()]) for l in range(L)])


def reverse_step_hstream(a, Hx, Ht, Hxx, Hb, Hxb, Htb, Hxxb):
    """The tanh reverse step of `reverse` restated in the OUTPUT streams (a, H_x, H_t, H_xx) of the layer instead of its
    Z streams: with H_x = d1 Z_x, H_t = d1 Z_t, H_xx = d1 Z_xx + d2 Z_x^2, d2 = -2 a d1, d3 = -2 d1 (1 - 3 a^2) the
    d3 Z_x^2 and d2 Z_xx terms collapse to -2 a H_xx - 2 H_x^2.  This is the form the CUDA kernels evaluate (their
    stash holds the H streams); tests/test_oracle.py checks it against the Z-stream form above."""
    d1 = 1.0 - a * a
    Zxxb = d1 * Hxxb
    Ztb = d1 * Htb
    Zxb = d1 * Hxb - 4.0 * a * Hx * Hxxb
    Zb = d1 * Hb - 2.0 * a * (Hx * Hxb + Ht * Htb + Hxx * Hxxb) - 2.0 * Hx * Hx * Hxxb
    return Zb, Zxb, Ztb, Zxxb


def burgers_residual(Y, Yx, Yt, Yxx, lam1, lam2):
    return Yt + lam1 * Y * Yx - lam2 * Yxx


def euler_residual(Y, Yx, Yt):
    """EUL:176-198 expanded by the product rule."""
    k = 0.4
    rho, u, E = Y[:, 0:1], Y[:, 1:2], Y[:, 2:3]
    rx, ux, Ex = Yx[:, 0:1], Yx[:, 1:2], Yx[:, 2:3]
    rt, ut, Et = Yt[:, 0:1], Yt[:, 1:2], Yt[:, 2:3]
    p = k * (E - 0.5 * rho * u * u)
    px = k * (Ex - 0.5 * rx * u * u - rho * u * ux)
    f1 = rt + rx * u + rho * ux
    f2 = rt * u + rho * ut + rx * u * u + 2 * rho * u * ux + px
    f3 = Et + ux * E + u * Ex + ux * p + u * px
    return np.hstack([f1, f2, f3]), p, px


def euler_adjoints(Y, Yx, Yt, p, px, B):
    """Adjoints of (rho,u,E) and their x/t derivatives given B = dL/d(f1,f2,f3) [N,3]."""
    k = 0.4
    rho, u, E = Y[:, 0:1], Y[:, 1:2], Y[:, 2:3]
    rx, ux, Ex = Yx[:, 0:1], Yx[:, 1:2], Yx[:, 2:3]
    rt, ut, Et = Yt[:, 0:1], Yt[:, 1:2], Yt[:, 2:3]
    b1, b2, b3 = B[:, 0:1], B[:, 1:2], B[:, 2:3]
    p_r, p_u, p_E = -0.5 * k * u * u, -k * rho * u, k
    px_r, px_u = -k * u * ux, -k * (rx * u + rho * ux)
    px_rx, px_ux, px_Ex = -0.5 * k * u * u, -k * rho * u, k
    rb = b1 * ux + b2 * (ut + 2 * u * ux + px_r) + b3 * (ux * p_r + u * px_r)
    ub = b1 * rx + b2 * (rt + 2 * rx * u + 2 * rho * ux + px_u) + b3 * (Ex + ux * p_u + px + u * px_u)
    Eb = b3 * (ux + ux * p_E)
    rxb = b1 * u + b2 * (u * u + px_rx) + b3 * u * px_rx
    uxb = b1 * rho + b2 * (2 * rho * u + px_ux) + b3 * (E + p + u * px_ux)
    Exb = b2 * px_Ex + b3 * (u + u * px_Ex)
    rtb = b1 + b2 * u
    utb = b2 * rho
    Etb = b3 * np.ones_like(u)
    return np.hstack([rb, ub, Eb]), np.hstack([rxb, uxb, Exb]), np.hstack([rtb, utb, Etb])


def evaluate(theta, prob: Problem, X_u, u_data, X_f, z: Optional[np.ndarray] = None,
             gamma: Optional[np.ndarray] = None) -> Evaluation:
    theta = np.asarray(theta, np.float64).astype(np.float32).astype(np.float64)
    Xu = np.asarray(X_u, np.float64).astype(np.float32).astype(np.float64)
    ud = np.asarray(u_data, np.float64).astype(np.float32).astype(np.float64)
    Xf = np.asarray(X_f, np.float64).astype(np.float32).astype(np.float64)
    n_u, n_f = Xu.shape[0], Xf.shape[0]
    lam1 = float(np.float32(prob.lam1)); lam2 = float(np.float32(prob.lam2))
    rho = float(np.float32(prob.rho))
    burgers = prob.pde == PDE_BURGERS
    if z is not None:
        z = np.asarray(z, np.float32).astype(np.float64)
        gamma = np.asarray(gamma, np.float32).astype(np.float64)

    # data term: primal stream only
    (Yu, _, _, _), cache_u = forward(theta, prob, Xu, second_order=False)
    r = ud - Yu
    L = prob.loss
    if L == LOSS_V1:
        nr = np.sqrt((r * r).sum())
        loss_data = nr
        Yu_bar = -r / nr
    else:
        loss_data = (r * r).sum() / n_u
        Yu_bar = -2.0 * r / n_u
    grad = reverse(theta, prob, cache_u, Yu_bar, None, None, None)

    # residual term
    (Y, Yx, Yt, Yxx), cache_f = forward(theta, prob, Xf, second_order=burgers)
    dlam = np.zeros(2)
    if burgers:
        f = burgers_residual(Y, Yx, Yt, Yxx, lam1, lam2)
    else:
        f, p, px = euler_residual(Y, Yx, Yt)
    if L in (LOSS_V1,):
        loss_res = (f * f).mean(); fbar = 2.0 * f / n_f
    elif L in (LOSS_V4, LOSS_EULER_MSE):
        loss_res = (f * f).sum() / n_f; fbar = 2.0 * f / n_f
    elif L == LOSS_V3:
        s = np.abs(f).sum()
        loss_res = s * s / n_f; fbar = 2.0 / n_f * s * np.sign(f)
    elif L == LOSS_V2:
        loss_res = (gamma * f).sum() + rho / 2 * ((f - z + gamma / rho) ** 2).sum()
        fbar = 2.0 * gamma + rho * (f - z)
    elif L in (LOSS_V5, LOSS_V6):
        loss_res = rho / 2 * ((f - z + gamma / rho) ** 2).sum()
        fbar = rho * (f - z) + gamma
    else:
        raise ValueError(L)
    if burgers:
        Yb = fbar * lam1 * Yx
        Yxb = fbar * lam1 * Y
        Ytb = fbar
        Yxxb = -lam2 * fbar
        dlam[0] = (fbar * Y * Yx).sum()
        dlam[1] = -(fbar * Yxx).sum()
        grad = grad + reverse(theta, prob, cache_f, Yb, Yxb, Ytb, Yxxb)
    else:
        Yb, Yxb, Ytb = euler_adjoints(Y, Yx, Yt, p, px, fbar)
        grad = grad + reverse(theta, prob, cache_f, Yb, Yxb, Ytb, None)
    return Evaluation(loss=float(loss_data + loss_res), grad=grad, dlam=dlam, u_pred=Yu, f=f)
