"""GPU accuracy tests of the device math layer (csrc/math64.cuh) through arianna_debug_math: the DEVICE code paths
(MUFU.RSQ64H seed, MUFU.EX2 filter, constant-bank polynomials) against long-double references."""
import numpy as np
import pytest

import montecarlo_b200 as mb

pytestmark = pytest.mark.gpu
LD = np.longdouble


def ulp_err(got, ref):
    u = np.spacing(np.abs(ref.astype(np.float64))).astype(LD)
    return np.abs((got.astype(LD) - ref) / u).astype(np.float64)


@pytest.fixture(scope="module")
def eng():
    with mb.CudaEnsemble(16, 2.0, [0.1]) as e:
        yield e


def test_device_exp(eng):
    rng = np.random.default_rng(1)
    x = np.concatenate([-rng.random(300000) * 2, -rng.random(300000) * 50, -rng.random(100000) * 700,
                        [0.0, -0.0, -1e-300, -707.99, -1e-17, 3.0, 1e300]])
    out = eng.debug_math(0, a=x)
    assert ulp_err(out, np.exp(np.minimum(x, 0).astype(LD))).max() <= 1.5
    y = np.array([5.0, 1e300, -709.0, -1e10, np.nan, -np.inf, np.inf])
    assert list(eng.debug_math(0, a=y)) == [1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0]


def test_device_neg2log_sqrt_sincos(eng):
    rng = np.random.default_rng(2)
    k = np.concatenate([rng.integers(1, 2 ** 53, size=500000, dtype=np.uint64) | np.uint64(1),
                        np.array([1, 3, 2 ** 53 - 1, 2 ** 52 + 1, 2 ** 52 - 1], dtype=np.uint64),
                        rng.integers(1, 2 ** 20, size=100000, dtype=np.uint64) | np.uint64(1)])
    w = eng.debug_math(1, b=k)
    assert ulp_err(w, -2 * np.log(k.astype(LD) * LD(2) ** -53)).max() <= 2.5 and w.min() > 0
    k52 = np.concatenate([rng.integers(1, 2 ** 52, size=500000, dtype=np.uint64) | np.uint64(1),
                          np.array([1, 3, 2 ** 52 - 1, 2 ** 51 + 1, 2 ** 51 - 1], dtype=np.uint64),
                          rng.integers(1, 2 ** 20, size=100000, dtype=np.uint64) | np.uint64(1)])
    w52 = eng.debug_math(9, b=k52)                                   # the Box-Muller radius form (one DADD)
    assert ulp_err(w52, -2 * np.log(k52.astype(LD) * LD(2) ** -52)).max() <= 2.5 and w52.min() > 0
    v = np.concatenate([rng.random(300000) * 75, 10.0 ** rng.uniform(-16, 2, size=300000)])
    assert ulp_err(eng.debug_math(2, a=v), np.sqrt(v.astype(LD))).max() <= 1.0     # Newton from the MUFU.RSQ64H seed
    kk = np.concatenate([rng.integers(0, 2 ** 53, size=500000, dtype=np.uint64),
                         np.array([0, 1, 2 ** 50, 2 ** 51, 2 ** 52, 2 ** 53 - 1], dtype=np.uint64)])
    sc = eng.debug_math(3, b=kk).reshape(-1, 2)
    ang = (LD(2) * np.arctan(LD(1)) * 4) * (kk.astype(LD) * LD(2) ** -53)
    assert np.abs(sc[:, 0] - np.sin(ang).astype(np.float64)).max() <= 2.3e-16
    assert np.abs(sc[:, 1] - np.cos(ang).astype(np.float64)).max() <= 2.3e-16


def test_device_box_muller(eng):
    rng = np.random.default_rng(5)
    b0 = rng.integers(0, 2 ** 64, size=400000, dtype=np.uint64)
    b1 = rng.integers(0, 2 ** 64, size=400000, dtype=np.uint64)
    z = eng.debug_math(4, b=b0, c=b1).reshape(-1, 2)
    u1 = ((b0 >> np.uint64(12)) | np.uint64(1)).astype(LD) * LD(2) ** -52
    u2 = (b1 >> np.uint64(11)).astype(LD) * LD(2) ** -53
    r = np.sqrt(-2 * np.log(u1))
    twopi = LD(2) * np.arctan(LD(1)) * 4
    assert np.abs(z[:, 0] - (r * np.cos(twopi * u2)).astype(np.float64)).max() < 4e-15
    assert np.abs(z[:, 1] - (r * np.sin(twopi * u2)).astype(np.float64)).max() < 4e-15


def test_device_philox_kat(eng):
    """Both device forms of the Philox4x32-10 block (general, and the per-chain hoisted PhiloxChain that the sweep
    uses) against the oracle's Philox, which is itself pinned by the Random123 known-answer vectors."""
    from oracle import oracle_np as N
    rng = np.random.default_rng(11)
    n = 20000
    sid = np.concatenate([rng.integers(0, 2 ** 64, size=n - 4, dtype=np.uint64),
                          np.array([0, 1, 2 ** 32 - 1, 2 ** 64 - 1], dtype=np.uint64)])
    p = np.concatenate([rng.integers(0, 2 ** 32, size=n - 4, dtype=np.uint64),
                        np.array([0, 1, 2 ** 31, 2 ** 32 - 1], dtype=np.uint64)])
    lo, hi = sid & np.uint64(0xffffffff), sid >> np.uint64(32)
    for kind, sub in ((6, 0), (7, 0), (7, 1), (7, 2)):
        got = eng.debug_math(kind, a=np.full(n, float(sub)), b=sid, c=p).reshape(-1, 4)
        want = np.stack(N.philox4x32_10(lo, p, hi, np.full(n, sub, dtype=np.uint64), 1, 0x41524941), axis=1)
        assert np.array_equal(got.astype(np.uint64), want), (kind, sub)
    # general form with a block index beyond 32 bits: p_hi lands in bits 8.. of the last counter word
    pbig = p + (np.uint64(5) << np.uint64(32))
    got = eng.debug_math(7, a=np.full(n, 2.0), b=sid, c=pbig).reshape(-1, 4)
    want = np.stack(N.philox4x32_10(lo, p, hi, np.full(n, 2 | (5 << 8), dtype=np.uint64), 1, 0x41524941), axis=1)
    assert np.array_equal(got.astype(np.uint64), want)


@pytest.mark.parametrize("pbits,kind", [(11, 5), (12, 8), (11, 10), (12, 11)])
def test_device_fp32_filter_never_changes_a_decision(eng, pbits, kind):
    """Prefix filter (11-bit prefix: odd step of a pair, 12-bit: even step) on the DEVICE (MUFU.EX2, FFMA.RM) against
    the plain FP64 decision, half of the cases adversarial near-ties.  Kinds 10 / 11: the headline sweep's form (argument
    in binary-log units, prefix as the filter's addend bits)."""
    rng = np.random.default_rng(7 + pbits)
    n = 4_000_000
    x = -rng.random(n) * rng.choice([0.01, 1.0, 3.0, 30.0, 300.0], size=n)
    w = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
    r = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
    h = n // 2
    tie = np.exp(x[:h]) * (1 + rng.normal(size=h) * 2.0 ** -rng.integers(18, 40, size=h))
    k = np.clip(tie * 2.0 ** 53, 0, 2 ** 53 - 1).astype(np.uint64)
    rb = 53 - pbits
    w[:h] = (w[:h] & ~np.uint64(2 ** pbits - 1)) | (k >> np.uint64(rb))
    r[:h] = (k & np.uint64(2 ** rb - 1)) << np.uint64(64 - rb)
    xin = x * 1.4426950408889634 if kind >= 10 else x
    d = eng.debug_math(kind, a=xin, b=w, c=r).reshape(-1, 2)
    assert np.array_equal(d[:, 0], d[:, 1])
    u = np.concatenate([k.astype(np.float64) * 2.0 ** -53])
    with np.errstate(all="ignore"):
        truth = np.minimum(LD(1), np.exp(x[:h].astype(LD))) > u.astype(LD)
    assert (truth != d[:h, 1].astype(bool)).sum() <= 40          # only ulp-level ties of exp() itself
