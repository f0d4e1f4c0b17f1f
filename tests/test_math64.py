"""CPU accuracy tests of montecarlo_b200/csrc/math64.cuh (compiled for the host with -DARIANNA_MATH_HOST) against
80-bit long-double references: the FP64 transcendentals of the fused sweep must stay within ≈2 ulp, and the FP32
accept filter must NEVER change a Metropolis decision relative to the plain FP64 evaluation."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LD = np.longdouble


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("m64") / "libm64.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-DARIANNA_MATH_HOST", "-shared", "-fPIC",
                           "-o", str(so), os.path.join(ROOT, "tests", "math64_host.cpp")])
    return C.CDLL(str(so))


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def ulp_err(got, ref):
    u = np.spacing(np.abs(ref.astype(np.float64))).astype(LD)
    return np.abs((got.astype(LD) - ref) / u).astype(np.float64)


def test_exp_core_accuracy(lib):
    rng = np.random.default_rng(1)
    x = np.concatenate([-rng.random(200000) * 2, -rng.random(200000) * 50, -rng.random(100000) * 700,
                        [0.0, -0.0, -1e-300, -707.99, -1e-17, 3.0, 1e300]])
    out = np.empty_like(x)
    lib.m64_exp(P(x), P(out), C.c_long(x.size))
    ref = np.exp(np.minimum(x, 0).astype(LD))
    assert ulp_err(out, ref).max() <= 1.5
    # min(1, exp(x)) classification: x > 0 -> 1, x < -708 / NaN / inf -> 0
    y = np.array([5.0, 1e300, -709.0, -1e10, np.nan, -np.inf, np.inf])
    o = np.empty_like(y)
    lib.m64_exp(P(y), P(o), C.c_long(y.size))
    assert list(o) == [1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0]


def test_neg2log_accuracy(lib):
    rng = np.random.default_rng(2)
    k = np.concatenate([rng.integers(1, 2 ** 53, size=400000, dtype=np.uint64) | np.uint64(1),
                        np.array([1, 3, 2 ** 53 - 1, 2 ** 52 + 1, 2 ** 52 - 1], dtype=np.uint64),
                        (np.uint64(2 ** 53) - rng.integers(1, 2 ** 20, size=100000, dtype=np.uint64)) | np.uint64(1),
                        rng.integers(1, 2 ** 20, size=100000, dtype=np.uint64) | np.uint64(1)])
    out = np.empty(k.size)
    lib.m64_neg2log(P(k), P(out), C.c_long(k.size))
    ref = -2 * np.log(k.astype(LD) * LD(2) ** -53)
    assert ulp_err(out, ref).max() <= 2.5
    assert out.min() > 0                         # odd lattice: u1 < 1, the radius never collapses to 0
    out2 = np.empty(k.size)
    lib.m64_neg2log_words(P(k), P(out2), C.c_long(k.size))
    assert np.array_equal(out, out2)             # the DADD-normalised variant is the same function, bit for bit
    # the 52-bit form the Box-Muller radius uses: −2 ln(k 2^-52), one exact DADD
    k52 = np.concatenate([rng.integers(1, 2 ** 52, size=400000, dtype=np.uint64) | np.uint64(1),
                          np.array([1, 3, 2 ** 52 - 1, 2 ** 51 + 1, 2 ** 51 - 1], dtype=np.uint64),
                          (np.uint64(2 ** 52) - rng.integers(1, 2 ** 20, size=100000, dtype=np.uint64)) | np.uint64(1),
                          rng.integers(1, 2 ** 20, size=100000, dtype=np.uint64) | np.uint64(1)])
    out3 = np.empty(k52.size)
    lib.m64_neg2log_k52(P(k52), P(out3), C.c_long(k52.size))
    assert ulp_err(out3, -2 * np.log(k52.astype(LD) * LD(2) ** -52)).max() <= 2.5 and out3.min() > 0


def test_sqrt_is_correctly_rounded(lib):
    rng = np.random.default_rng(3)
    w = np.concatenate([rng.random(300000) * 75, 10.0 ** rng.uniform(-16, 2, size=300000)])
    out = np.empty_like(w)
    lib.m64_sqrt(P(w), P(out), C.c_long(w.size))
    assert ulp_err(out, np.sqrt(w.astype(LD))).max() <= 0.5001


def test_sincos_turn_accuracy(lib):
    rng = np.random.default_rng(4)
    k = np.concatenate([rng.integers(0, 2 ** 53, size=500000, dtype=np.uint64),
                        np.array([0, 1, 2 ** 50, 2 ** 50 - 1, 2 ** 51, 2 ** 52, 2 ** 53 - 1, 3 * 2 ** 50], dtype=np.uint64)])
    s, c = np.empty(k.size), np.empty(k.size)
    lib.m64_sincos(P(k), P(s), P(c), C.c_long(k.size))
    ang = (LD(2) * np.arctan(LD(1)) * 4) * (k.astype(LD) * LD(2) ** -53)
    assert np.abs(s - np.sin(ang).astype(np.float64)).max() <= 2.3e-16
    assert np.abs(c - np.cos(ang).astype(np.float64)).max() <= 2.3e-16
    assert (s[-8], c[-8]) == (0.0, 1.0) and (s[-4], c[-4]) == (1.0, 0.0)      # exact at the quadrant points
    assert np.abs(s * s + c * c - 1).max() < 5e-16


def test_sincos_table_form_agrees_with_polynomial_form(lib):
    """Two independent evaluations (1024-direction table + short Taylor vs quadrant reduction + fdlibm kernels)."""
    rng = np.random.default_rng(14)
    k = rng.integers(0, 2 ** 53, size=500000, dtype=np.uint64)
    s, c, s2, c2 = (np.empty(k.size) for _ in range(4))
    lib.m64_sincos(P(k), P(s), P(c), C.c_long(k.size))
    lib.m64_sincos_poly(P(k), P(s2), P(c2), C.c_long(k.size))
    assert np.abs(s - s2).max() <= 3.4e-16 and np.abs(c - c2).max() <= 3.4e-16


def test_box_muller_matches_oracle_definition(lib):
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    b0 = rng.integers(0, 2 ** 64, size=100000, dtype=np.uint64)
    b1 = rng.integers(0, 2 ** 64, size=100000, dtype=np.uint64)
    z0, z1 = np.empty(b0.size), np.empty(b0.size)
    lib.m64_box_muller(P(b0), P(b1), P(z0), P(z1), C.c_long(b0.size))
    u1 = ((b0 >> np.uint64(12)) | np.uint64(1)).astype(LD) * LD(2) ** -52
    u2 = (b1 >> np.uint64(11)).astype(LD) * LD(2) ** -53
    r = np.sqrt(-2 * np.log(u1))
    twopi = LD(2) * np.arctan(LD(1)) * 4
    assert np.abs(z0 - (r * np.cos(twopi * u2)).astype(np.float64)).max() < 4e-15
    assert np.abs(z1 - (r * np.sin(twopi * u2)).astype(np.float64)).max() < 4e-15
    zz = np.concatenate([z0, z1])
    assert abs(zz.mean()) < 4 / np.sqrt(zz.size) and abs(zz.std() - 1) < 4 / np.sqrt(2 * zz.size)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_fp32_filter_never_changes_a_decision(lib, mode):
    """mode 0: 23-bit cell of a 53-bit word (XOSHIRO path); mode 1 / 3: 11- / 12-bit prefix + lazy 42- / 41-bit
    refinement (native, odd / even step of a pair); mode 2: directed-rounding float cell of an arbitrary double u
    (replay / EXACT paths); mode 4 / 5: the headline sweep's form of 1 / 3 (argument in binary-log units, prefix as the
    filter's addend bits)."""
    rng = np.random.default_rng(6 + mode)
    n = 2_000_000
    x = -rng.random(n) * rng.choice([0.01, 1.0, 3.0, 30.0, 300.0], size=n)
    w = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
    r = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
    # adversarial half: u within 2^-18 .. 2^-40 (relative) of exp(x)
    h = n // 2
    tie = np.exp(x[:h]) * (1 + rng.normal(size=h) * 2.0 ** -rng.integers(18, 40, size=h))
    k = np.clip(tie * 2.0 ** 53, 0, 2 ** 53 - 1).astype(np.uint64)
    if mode in (0, 2):
        w[:h] = (k << np.uint64(11)) | (w[:h] & np.uint64(0x7ff))
    elif mode in (1, 4):
        w[:h] = (w[:h] & ~np.uint64(0x7ff)) | (k >> np.uint64(42))
        r[:h] = (k & np.uint64(2 ** 42 - 1)) << np.uint64(22)
    else:
        w[:h] = (w[:h] & ~np.uint64(0xfff)) | (k >> np.uint64(41))
        r[:h] = (k & np.uint64(2 ** 41 - 1)) << np.uint64(23)
    x = np.concatenate([x, [0.0, -0.0, 1e-300, 5.0, -708.0, -709.0, -1e10, -np.inf, np.nan, 1e308, -1e-320]])
    w = np.concatenate([w, rng.integers(0, 2 ** 64, size=11, dtype=np.uint64)])
    r = np.concatenate([r, rng.integers(0, 2 ** 64, size=11, dtype=np.uint64)])
    f, ref, u = np.empty(x.size, np.uint8), np.empty(x.size, np.uint8), np.empty(x.size)
    xin = x
    if mode >= 4:
        with np.errstate(all="ignore"):
            xin = x * 1.4426950408889634                # the sweep multiplies by β·log2e instead of β
            x = xin * 0.6931471805599453                 # what the exact path evaluates: RN(y·ln2)
    lib.m64_accept(P(xin), P(w), P(r), C.c_int(mode), P(f), P(ref), P(u), C.c_long(x.size))
    assert np.array_equal(f, ref)
    assert np.array_equal(u[:h], k.astype(np.float64) * 2.0 ** -53)             # bit-assembled uniform is exact
    with np.errstate(all="ignore"):
        truth = np.minimum(LD(1), np.exp(x.astype(LD))) > u.astype(LD)
    # the FP64 decision itself differs from the infinitely precise one only on ulp-level ties of exp()
    assert (truth != ref.astype(bool)).sum() <= 20
    assert list(ref[-11:]) == [1, 1, 1, 1, 0, 0, 0, 0, 0, 1, 1]


def test_exact_div(lib):
    """exact_div(n, d, RN(1/d)) == n / d bit for bit: random and adversarial mantissas (all ones, all zeros, few bits)
    over the exponent range the fast path covers, and the fall-back outside it (zeros, infinities, NaN, subnormals)."""
    rng = np.random.default_rng(7)
    N = 4_000_000

    def rnd(emin, emax, n):
        m = rng.integers(0, 2 ** 52, size=n, dtype=np.uint64)
        e = rng.integers(emin + 1023, emax + 1024, size=n, dtype=np.uint64)
        kind = rng.integers(0, 8, size=n)
        low = rng.integers(0, 256, size=n, dtype=np.uint64)
        m = np.where(kind == 0, np.uint64(2 ** 52 - 1) ^ low, m)          # mantissa of (almost) all ones
        m = np.where(kind == 1, low, m)                                   # (almost) all zeros
        m = np.where(kind == 2, m & ~np.uint64(2 ** 40 - 1), m)           # few significant bits
        return ((e << np.uint64(52)) | m).view(np.float64)

    n = -rnd(-400, 399, N)
    d = rnd(-300, 299, N)
    sig = rng.random(1000) * 3 + 1e-3                                     # d = 2σ² of real pools, δ = σ z
    n[:100000] = -((sig[rng.integers(0, 1000, 100000)] * rng.standard_normal(100000)) ** 2)
    d[:100000] = 2 * sig[rng.integers(0, 1000, 100000)] ** 2
    out = np.empty(N)
    lib.m64_exact_div(P(n), P(d), P(out), C.c_long(N))
    assert np.array_equal(out, n / d)
    # outside the fast range: the IEEE division itself
    with np.errstate(all="ignore"):
        n2 = np.array([0.0, -0.0, -np.inf, np.nan, -1e-320, -1e300, -1e-200, -4.0, -1.0, -1e308])
        d2 = np.array([0.02, 0.02, 0.02, 0.02, 0.02, 0.02, 0.02, 1e-320, 1e305, 3.0])
        o2 = np.empty(n2.size)
        lib.m64_exact_div(P(n2), P(d2), P(o2), C.c_long(n2.size))
        assert np.array_equal(o2, n2 / d2, equal_nan=True) and np.array_equal(np.signbit(o2), np.signbit(n2 / d2))
