"""numpy restatement of the device collocation sampler (Philox4x32-10, Salmon et al. 2011).

TEST INFRASTRUCTURE ONLY.  The reference draws `np.random.uniform(lb, ub, [N_f,1])` twice per step on the
host (AB-ADMM:220-221, EUL:232-233); the device replacement is a counter-based stream so that point i of the
job is the same whatever the number of GPUs.  This file pins the kernel's bit pattern.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, seed: int, word2: int = 0):
    """counter: uint64 array (point index); word2: third counter word (stream selector); returns four uint32 arrays."""
    c = np.asarray(counter, np.uint64)
    c0 = (c & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (c >> np.uint64(32)).astype(np.uint32)
    c2 = np.full_like(c0, word2)
    c3 = np.zeros_like(c0)
    k0 = np.uint32(seed & 0xFFFFFFFF)
    k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def sample_collocation(seed: int, first_index: int, n: int, lb, ub) -> np.ndarray:
    """float32 [n,2]: x = fma(span_x, U0, lb_x), t = fma(span_t, U1, lb_t), U = (bits >> 8) * 2^-24."""
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    o0, o1, _, _ = philox4x32_10(idx, seed)
    u0 = (o0 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    u1 = (o1 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    lbf = np.asarray(lb, np.float64).astype(np.float32)
    span = (np.asarray(ub, np.float64) - np.asarray(lb, np.float64)).astype(np.float32)
    # fmaf: one rounding; emulate in float64 (exact product of two float32 fits) then round once
    x = (span[0].astype(np.float64) * u0.astype(np.float64) + lbf[0].astype(np.float64)).astype(np.float32)
    t = (span[1].astype(np.float64) * u1.astype(np.float64) + lbf[1].astype(np.float64)).astype(np.float32)
    return np.stack([x, t], axis=1)


def _lhs_round(r, rnd, dim, k0, k1):
    """round function of the stratum permutation (uint32 arithmetic, wraps)"""
    with np.errstate(over="ignore"):
        h = r ^ np.uint32((k0 + rnd * 0x9E3779B9) & 0xFFFFFFFF)
        h = h * np.uint32(0x85EBCA6B)
        h = h ^ (h >> np.uint32(13))
        h = h + np.uint32(k1 ^ (((dim + 1) * 0xC2B2AE35) & 0xFFFFFFFF))
        h = h * np.uint32(0xC2B2AE35)
        h = h ^ (h >> np.uint32(16))
    return h


def lhs_half_bits(n_design: int) -> int:
    bits = 2
    while (1 << bits) < n_design:
        bits += 1
    return (bits + 1) // 2


def lhs_perm(idx: np.ndarray, n_design: int, dim: int, seed: int) -> np.ndarray:
    """pi_dim(i): six-round balanced Feistel network on 2*half_bits bits, cycle-walked into [0, n_design)."""
    hb = lhs_half_bits(n_design)
    mask = np.uint32((1 << hb) - 1)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    out = np.asarray(idx, np.uint64).copy()
    todo = np.ones(out.shape, bool)
    while todo.any():
        i = out[todo]
        L = ((i >> np.uint64(hb)).astype(np.uint32)) & mask
        R = i.astype(np.uint32) & mask
        for rnd in range(6):
            F = _lhs_round(R, rnd, dim, k0, k1) & mask
            L, R = R, L ^ F
        i = (L.astype(np.uint64) << np.uint64(hb)) | R.astype(np.uint64)
        out[todo] = i
        todo[todo] = i >= np.uint64(n_design)
    return out


def sample_lhs(seed: int, first_index: int, n: int, n_design: int, lb, ub) -> np.ndarray:
    """float32 [n,2]: points [first_index, first_index + n) of the device Latin hypercube design (pinn_sample_lhs):
    x_d = lb_d + (ub_d - lb_d) * ((pi_d(i) + U_d(i)) / N) in float64, rounded once to float32.  Same construction as
    pyDOE.lhs (INF-L2:183; oracle.data.lhs): one uniform draw per stratum and axis, strata permuted per axis."""
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    o0, o1, _, _ = philox4x32_10(idx, seed, word2=1)
    lb = np.asarray(lb, np.float64)
    w = np.asarray(ub, np.float64) - lb
    cols = []
    for d, o in enumerate((o0, o1)):
        u = (o >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
        s = lhs_perm(idx, n_design, d, seed).astype(np.float64)
        cols.append((lb[d] + w[d] * ((s + u) / np.float64(n_design))).astype(np.float32))
    return np.stack(cols, axis=1)
