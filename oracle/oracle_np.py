"""Second, independent restatement of the hot path in numpy (cross-checks oracle/arianna_oracle.c).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  PARITY UNPINNED (no Julia here).

numpy float64 element-wise arithmetic is IEEE-754 binary64 with one rounding per ufunc call and no FMA
contraction, so each line below is one operation of SURVEY.md Appendix A.  Citations: path:line under
/root/reference.
"""
from __future__ import annotations

import math

import numpy as np

TWO_PI = 6.283185307179586  # Julia's 2π in binary64 (particle_1d.jl:53)


# ---------------------------------------------------------------------------------------------------------
# build_schedule, src/simulation.jl:95-117 (pure Python, exact integer arithmetic)
# ---------------------------------------------------------------------------------------------------------
def _unique(seq):
    seen, out = set(), []
    for v in seq:
        if v not in seen:
            seen.add(v)
            out.append(v)
    return out


def build_schedule(steps: int, burn: int, spec):
    if isinstance(spec, (int, np.integer)) and not isinstance(spec, bool):
        # collect(burn:Δt:steps) ∪ [steps]   (:95-97)
        return _unique(list(range(burn, steps + 1, int(spec))) + [steps])
    if isinstance(spec, float):
        # unique(vcat([burn], [burn + Int(base^n) for n in 0:floor(Int, log(base, steps-burn))], [steps])) (:104-106)
        nmax = math.floor(math.log(steps - burn, spec))
        mid = []
        for n in range(0, nmax + 1):
            v = spec ** n
            if v != math.floor(v):
                raise ValueError(f"InexactError: Int({v})")  # Julia's Int(::Float64) on a non-integer
            mid.append(burn + int(v))
        return _unique([burn] + mid + [steps])
    block = list(spec)
    # blocks = [block .+ burn .+ (m-1)*block[end] for m in 1:nblock]; filter(≤ steps, unique(vcat(blocks..., [steps]))) (:113-117)
    nblock = (steps - burn) // block[-1]
    out = []
    for m in range(1, nblock + 1):
        out.extend(b + burn + (m - 1) * block[-1] for b in block)
    return [t for t in _unique(out + [steps]) if t <= steps]


# ---------------------------------------------------------------------------------------------------------
# Philox4x32-10 in numpy (vectorised over counters)
# ---------------------------------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for v in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & mask
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & mask
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


# ---------------------------------------------------------------------------------------------------------
# A.1: replay sweep (metropolis.jl:176-212 + particle_1d.jl:20-59, potential(x) = x^2)
# ---------------------------------------------------------------------------------------------------------
def _logq(delta, sigma, lognorm):
    s2 = sigma * sigma
    t1 = (-(delta * delta)) / (2.0 * s2)
    return t1 - lognorm


def lognorm(sigma):
    sigma = np.asarray(sigma, dtype=np.float64)
    return np.log(TWO_PI * (sigma * sigma)) / 2.0


def categorical(weight, u):
    """Inverse-CDF scan of Distributions.Categorical [EXT] -- 0-based index per chain."""
    n = len(weight)
    k = np.zeros(u.shape, dtype=np.int64)
    cp = np.full(u.shape, weight[0], dtype=np.float64)
    for _ in range(n - 1):
        go = (cp <= u) & (k < n - 1)
        k = np.where(go, k + 1, k)
        cp = np.where(go, cp + np.asarray(weight)[np.minimum(k, n - 1)], cp)
    return k


def sweep_replay(x, e, beta, sigma, weight, u_cat, z, u_acc, acc, tot, ln=None):
    """In-place K-step replay over all chains; returns decisions [K][M]."""
    sigma = np.asarray(sigma, dtype=np.float64)
    ln = lognorm(sigma) if ln is None else np.asarray(ln)
    K, M = z.shape
    dec = np.zeros((K, M), dtype=np.uint8)
    idx = np.arange(M)
    with np.errstate(over="ignore", invalid="ignore"):
        for s in range(K):
            k = categorical(weight, u_cat[s]) if u_cat is not None else np.zeros(M, dtype=np.int64)
            sg, lnk = sigma[k], ln[k]
            delta = 0.0 + (sg * z[s])
            lqf = _logq(delta, sg, lnk)
            e1 = e.copy()
            x += delta
            e[:] = x * x
            dlogp = ((-e) * beta) - ((-e1) * beta)
            delta = -delta
            lqb = _logq(delta, sg, lnk)
            ex = np.exp((dlogp + lqb) - lqf)
            alpha = np.where(ex > 1.0, 1.0, ex)
            a = alpha > u_acc[s]
            tot[k, idx] += 1
            acc[k, idx] += a
            xr = x + delta
            x[:] = np.where(a, x, xr)
            e[:] = np.where(a, e, xr * xr)
            dec[s] = a
    return dec


# ---------------------------------------------------------------------------------------------------------
# A.2: PGMC estimator record sums (gradients.jl:93-121, estimator.jl:111-134)
# ---------------------------------------------------------------------------------------------------------
def pgmc_replay(x, e, beta, sigma, learn_ids, z):
    sigma = np.asarray(sigma, dtype=np.float64)
    ln = lognorm(sigma)
    n_learn, q_batch, M = z.shape
    out = np.zeros((n_learn, 5))
    for l, k in enumerate(learn_ids):
        sg = sigma[k]
        for b in range(q_batch):
            delta = 0.0 + sg * z[l, b]
            gf = (delta * delta) / (sg * sg * sg) - 1.0 / sg
            lqf = _logq(delta, sg, ln[k])
            e1 = e.copy()
            x += delta
            e[:] = x * x
            dlogp = ((-e) * beta) - ((-e1) * beta)
            r = delta * delta
            delta = -delta
            gb = (delta * delta) / (sg * sg * sg) - 1.0 / sg
            lqb = _logq(delta, sg, ln[k])
            x += delta
            e[:] = x * x
            ex = np.exp((dlogp + lqb) - lqf)
            alpha = np.where(ex > 1.0, 1.0, ex)
            j = r * alpha
            dj = j * np.where(alpha == 1.0, gf, gb)
            out[l] += [j.sum(), dj.sum(), gf.sum(), (gf * gf).sum(), float(M)]
    return out


# ---------------------------------------------------------------------------------------------------------
# learning_step!, src/PolicyGuided/learning.jl -- general P (numpy linear algebra, like the reference)
# gd = averaged record: j (scalar), dj (P,), glq (P,), g (P,P)
# ---------------------------------------------------------------------------------------------------------
def learning_step(kind, hp, theta, j, dj, glq, g):
    theta = np.asarray(theta, dtype=np.float64)
    dj, glq, g = np.atleast_1d(dj), np.atleast_1d(glq), np.atleast_2d(g)
    I = np.eye(theta.size)
    if kind == "VPG":
        return theta + hp[0] * dj
    if kind == "BLPG":
        return theta + hp[0] * (dj - j * glq)
    if kind == "BLAPG":
        eta = math.sqrt(2 * hp[0] / (dj @ dj + hp[1]))
        return theta + eta * (dj - j * glq)
    if kind == "NPG":
        return theta + hp[0] * np.linalg.inv(g + hp[1] * I) @ dj
    if kind == "ANPG":
        Finv = np.linalg.inv(g + hp[1] * I)
        eta = math.sqrt(2 * hp[0] / (dj @ (Finv @ dj)))
        return theta + eta * Finv @ dj
    if kind == "BLANPG":
        Finv = np.linalg.inv(g + hp[1] * I)
        bj = dj - j * glq
        eta = math.sqrt(2 * hp[0] / (bj @ (Finv @ bj)))
        return theta + eta * Finv @ bj
    return theta
