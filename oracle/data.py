"""Data preparation of the reference drivers, restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows INF-L2:160-190 (meshgrid, X_star, lb/ub, IC/BC stacking, LHS collocation
points with the IC/BC points appended, seeded choice of N_u points),
AB-ADMM:264-309 and EUL:274-333.  ``sol`` is a dict with the arrays of the
reference's .mat fixtures (keys x, t, usol [, rhosol, Enersol]); the copies
under tests/golden/data/*.npz are made by tests/golden/make_fixtures.py.
pyDOE is not installed: ``lhs`` restates its default (non-centred, no
criterion) algorithm on numpy's legacy global RNG.
"""
from __future__ import annotations

import numpy as np


def lhs(n, samples, rng=np.random):
    """pyDOE.lhs(n, samples) default: one uniform draw inside each of `samples`
    equal strata per dimension, then an independent permutation per column."""
    cut = np.linspace(0, 1, samples + 1)
    u = rng.rand(samples, n)
    a, b = cut[:samples], cut[1:samples + 1]
    rdpoints = np.zeros_like(u)
    for j in range(n):
        rdpoints[:, j] = u[:, j] * (b - a) + a
    H = np.zeros_like(rdpoints)
    for j in range(n):
        order = rng.permutation(range(samples))
        H[:, j] = rdpoints[order, j]
    return H


def burgers_grid(sol):
    """INF-L2:162-177 / AB-ADMM:273-300."""
    t = sol['t'].flatten()[:, None]
    x = sol['x'].flatten()[:, None]
    Exact = np.real(sol['usol']).T
    X, T = np.meshgrid(x, t)
    X_star = np.hstack((X.flatten()[:, None], T.flatten()[:, None]))
    u_star = Exact.flatten()[:, None]
    lb = X_star.min(0)
    ub = X_star.max(0)
    xx1 = np.hstack((X[0:1, :].T, T[0:1, :].T)); uu1 = Exact[0:1, :].T
    xx2 = np.hstack((X[:, 0:1], T[:, 0:1])); uu2 = Exact[:, 0:1]
    xx3 = np.hstack((X[:, -1:], T[:, -1:])); uu3 = Exact[:, -1:]
    X_u_all = np.vstack([xx1, xx2, xx3])
    u_all = np.vstack([uu1, uu2, uu3])
    return dict(x=x, t=t, Exact=Exact, X=X, T=T, X_star=X_star, u_star=u_star, lb=lb, ub=ub,
                X_u_all=X_u_all, u_all=u_all)


def burgers_inference_inputs(sol, N_u=100, N_f=10000, seed=1234):
    """INF-L2:179-190: X_f = LHS points UNION all IC/BC points; N_u of the IC/BC points by seeded choice."""
    g = burgers_grid(sol)
    np.random.seed(seed)
    X_f = g['lb'] + (g['ub'] - g['lb']) * lhs(2, N_f)
    X_f = np.vstack((X_f, g['X_u_all']))
    idx = np.random.choice(g['X_u_all'].shape[0], N_u, replace=False)
    g.update(X_u=g['X_u_all'][idx, :], u=g['u_all'][idx, :], X_f=X_f)
    return g


def burgers_identification_inputs(sol, N_u=100, N_f=1000, seed=1234):
    """AB-ADMM:302-309 (seeded choice) and :91-93 (first uniform collocation batch)."""
    g = burgers_grid(sol)
    np.random.seed(seed)
    idx = np.random.choice(g['X_u_all'].shape[0], N_u, replace=False)
    x_phys = np.random.uniform(g['lb'][0], g['ub'][0], [N_f, 1])
    t_phys = np.random.uniform(g['lb'][1], g['ub'][1], [N_f, 1])
    g.update(X_u=g['X_u_all'][idx, :], u=g['u_all'][idx, :], X_f=np.hstack([x_phys, t_phys]))
    return g


def euler_inputs(sol, N_data=200, N_f=1000, seed=1234):
    """EUL:281-333 and :85-87."""
    t = sol['t'].flatten()[:, None]
    x = sol['x'].flatten()[:, None]
    Er = np.real(sol['rhosol']).T; Eu = np.real(sol['usol']).T; EE = np.real(sol['Enersol']).T
    X, T = np.meshgrid(x, t)
    X_star = np.hstack((X.flatten()[:, None], T.flatten()[:, None]))
    lb = X_star.min(0); ub = X_star.max(0)
    dom = np.vstack([np.hstack((X[0:1, :].T, T[0:1, :].T)), np.hstack((X[:, 0:1], T[:, 0:1])),
                     np.hstack((X[:, -1:], T[:, -1:]))])
    def stack(E):
        return np.vstack([E[0:1, :].T, E[:, 0:1], E[:, -1:]])
    np.random.seed(seed)
    idx = np.random.choice(dom.shape[0], N_data, replace=False)
    x_phys = np.random.uniform(lb[0], ub[0], [N_f, 1])
    t_phys = np.random.uniform(lb[1], ub[1], [N_f, 1])
    return dict(x=x, t=t, X_star=X_star, lb=lb, ub=ub,
                rho_star=Er.flatten()[:, None], u_star=Eu.flatten()[:, None], E_star=EE.flatten()[:, None],
                X_u=dom[idx, :], u=np.hstack([stack(Er)[idx, :], stack(Eu)[idx, :], stack(EE)[idx, :]]),
                X_f=np.hstack([x_phys, t_phys]))
