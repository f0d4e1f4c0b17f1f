"""CPU tests of the oracle itself: public known-answer vectors of its RNG building blocks, the C restatement
against the independent numpy restatement, the committed golden fixtures, and the analytic / reference-test
anchors.  The reference has no golden vectors for this path: PARITY UNPINNED."""
import math
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import oracle_np as N


# ---- RNG building blocks: public known-answer vectors ------------------------------------------------------
@pytest.mark.parametrize("ctr,key,out", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_kat(ctr, key, out):
    """Random123 kat_vectors for philox4x32-10."""
    assert [int(v) for v in O.philox4x32_10(ctr, key)] == out
    got = N.philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1])
    assert [int(v) for v in got] == out


def test_xoshiro_kat():
    """xoshiro256++ with state (1,2,3,4): first output rotl(1+4,23)+1 = 41943041."""
    r, s = O.xoshiro_next([1, 2, 3, 4])
    assert r == 41943041
    assert [int(v) for v in s] == [7, 0, 262146, 211106232532992]


def test_ziggurat_tables_shape():
    ki, wi, fi = O.ziggurat_tables()
    assert ki[1] == 0 and fi[0] == 1.0
    assert np.all(np.diff(fi[1:]) < 0)           # f decreasing with x
    assert np.all(np.diff(wi[1:]) > 0)           # x_i increasing
    assert abs(wi[255] * 2.0 ** 51 - 3.6541528853610088) < 1e-15


def test_julia_like_normal_moments():
    ens = O.Ensemble(np.zeros(512), 2.0, [0.1])
    ens.seed_xoshiro(7)
    uc, z, ua = ens.draws_xoshiro(4000)
    n = z.size
    assert abs(z.mean()) < 4 / math.sqrt(n)
    assert abs(z.std() - 1) < 4 / math.sqrt(2 * n)
    assert abs((z ** 4).mean() - 3) < 4 * math.sqrt(96 / n)
    assert 0 <= uc.min() and uc.max() < 1 and abs(uc.mean() - 0.5) < 4 / math.sqrt(12 * n)
    assert abs(ua.mean() - 0.5) < 4 / math.sqrt(12 * n)


def test_philox_draws_layout_and_moments():
    M, K = 64, 1001
    uc, z, ua = O.draws_philox(42, 5, M, 0, K)
    # chunk invariance incl. an odd split: draws are a pure function of (seed + chain, step)
    uc2, z2, ua2 = O.draws_philox(42, 5, M, 333, K - 333)
    assert np.array_equal(z[333:], z2) and np.array_equal(ua[333:], ua2) and np.array_equal(uc[333:], uc2)
    # shard invariance: chain offset shifts the stream id
    uc3, z3, ua3 = O.draws_philox(42, 5 + 10, M - 10, 0, 50)
    assert np.array_equal(z[:50, 10:], z3)
    # seed + c − 1 semantics (metropolis.jl:262): (seed, chain c+1) == (seed+1, chain c)
    uc4, z4, ua4 = O.draws_philox(43, 5, M - 1, 0, 50)
    assert np.array_equal(z[:50, 1:], z4)
    n = z.size
    assert abs(z.mean()) < 4 / math.sqrt(n) and abs(z.std() - 1) < 4 / math.sqrt(2 * n)
    # numpy restatement of the counter layout for chain 3, pair 0 (lazy-refinement layout):
    # even step: 12-bit prefix + 41 refinement bits, odd step: 11 + 42; the radius uniform takes the top 52 bits of A0
    sid = 42 + 5 + 3
    # layout v2: ctr = (sid_lo, p, sid_hi, sub), key = (tag, 'ARIA'); pair p = 0, sub-blocks 0 and 1
    o0 = N.philox4x32_10(sid & 0xffffffff, 0, sid >> 32, 0, 1, 0x41524941)
    o1 = N.philox4x32_10(sid & 0xffffffff, 0, sid >> 32, 1, 1, 0x41524941)
    A0 = int(o0[0]) | (int(o0[1]) << 32)
    B0 = int(o0[2]) | (int(o0[3]) << 32)
    A1 = int(o1[0]) | (int(o1[1]) << 32)
    B1 = int(o1[2]) | (int(o1[3]) << 32)
    assert ua[0, 3] == (((A0 & 0xfff) << 41) | (A1 >> 23)) * 2.0 ** -53
    assert ua[1, 3] == (((B0 & 0x7ff) << 42) | (B1 >> 22)) * 2.0 ** -53
    u1, u2 = ((A0 >> 12) | 1) * 2.0 ** -52, (B0 >> 11) * 2.0 ** -53
    r = math.sqrt(-2.0 * math.log(u1))
    assert abs(z[0, 3] - r * math.cos(2 * math.pi * u2)) < 1e-14 and abs(z[1, 3] - r * math.sin(2 * math.pi * u2)) < 1e-14


def test_philox_stream_statistics():
    """Distributional checks of the native stream (layout v3: 52-bit radius uniform, 12/11-bit prefix + lazy
    refinement): normal draws, accept and categorical uniforms are correctly distributed, mutually uncorrelated,
    and uncorrelated along the step axis and across neighbouring chains (stream ids differ by one)."""
    from scipy import stats
    M, K = 512, 2000
    uc, z, ua = O.draws_philox(1234, 77, M, 0, K)
    n = z.size
    assert stats.kstest(z.ravel(), "norm").pvalue > 1e-3
    assert stats.kstest(ua.ravel(), "uniform").pvalue > 1e-3
    assert stats.kstest(uc.ravel(), "uniform").pvalue > 1e-3
    # even and odd steps use different bit fields of the block (cos/sin half, 12/11-bit prefix): check both
    for par in (0, 1):
        assert stats.kstest(z[par::2].ravel(), "norm").pvalue > 1e-3
        assert stats.kstest(ua[par::2].ravel(), "uniform").pvalue > 1e-3
    lim = 4.5 / math.sqrt(n)

    def corr(a, b):
        a, b = a.ravel() - a.mean(), b.ravel() - b.mean()
        return float((a * b).mean() / (a.std() * b.std()))
    assert abs(corr(z, ua)) < lim and abs(corr(z, uc)) < lim and abs(corr(ua, uc)) < lim
    assert abs(corr(z[:-1], z[1:])) < lim and abs(corr(z[::2], z[1::2])) < 1.5 * lim      # lag 1, and within a pair
    assert abs(corr(z[:-1] ** 2, z[1:] ** 2)) < lim                                        # shared radius must not show
    assert abs(corr(ua[:-1], ua[1:])) < lim and abs(corr(z[:, :-1], z[:, 1:])) < lim      # neighbouring chains
    # tails: P(|z| > 4) = 6.33e-5
    tail = (np.abs(z) > 4).sum()
    assert abs(tail - 6.334e-5 * n) < 5 * math.sqrt(6.334e-5 * n)


# ---- the sweep: C restatement == numpy restatement, bit for bit ---------------------------------------------
@pytest.mark.parametrize("sigma,weight", [([0.1], [1.0]), ([0.2] * 7, [0.4] + [0.1] * 6), ([1.5, 0.01], [0.3, 0.7])])
def test_c_matches_numpy_replay(sigma, weight):
    M, K, beta = 257, 40, 2.0
    x0 = O.init_synthetic(3, 0, M)
    uc, z, ua = O.draws_philox(3, 0, M, 0, K)
    ens = O.Ensemble(x0, beta, sigma, weight)
    dec, mov, _ = ens.sweep_replay(uc, z, ua, want_decisions=True)
    x, e = x0.copy(), x0 * x0
    acc = np.zeros((len(sigma), M), dtype=np.int64)
    tot = np.zeros_like(acc)
    dec_np = N.sweep_replay(x, e, beta, sigma, weight, uc, z, ua, acc, tot)
    assert np.array_equal(dec, dec_np)
    assert np.array_equal(ens.x, x) and np.array_equal(ens.e, e)
    assert np.array_equal(ens.acc, acc) and np.array_equal(ens.tot, tot)
    assert np.array_equal(ens.e, ens.x * ens.x)          # e is always potential(x) (particle_1d.jl:33)


def test_reject_path_is_not_a_restore():
    """x ← fl(fl(x+δ)−δ) on reject (metropolis.jl:187 re-applies the negated move)."""
    x0 = np.array([0.1 + 2.0 ** -54 * 3])
    ens = O.Ensemble(x0, 2.0, [1.0])
    z = np.array([[3.0]])
    ens.sweep_replay(None, z, np.array([[0.999999]]))     # α = exp(-2(x'^2 - x^2)) ≈ 0 → reject
    assert ens.acc[0, 0] == 0
    assert ens.x[0] == (x0[0] + 3.0) - 3.0
    assert ens.e[0] == ens.x[0] * ens.x[0]


def test_strict_inequality_and_always_drawn_uniforms():
    ens = O.Ensemble(np.array([0.0]), 2.0, [0.1])
    dec, _, alp = ens.sweep_replay(None, np.array([[0.0]]), np.array([[1.0 - 2.0 ** -53]]), want_decisions=True,
                                   want_alpha=True)
    assert alp[0, 0] == 1.0 and dec[0, 0] == 1            # δ = 0 → α = 1 > u
    ens = O.Ensemble(np.array([5.0]), 2.0, [0.1])
    dec, _, alp = ens.sweep_replay(None, np.array([[-50.0]]), np.array([[0.0]]), want_decisions=True,
                                   want_alpha=True)
    assert alp[0, 0] == 1.0 and dec[0, 0] == 1            # downhill move always accepted


def test_nan_state_rejects():
    ens = O.Ensemble(np.array([np.inf]), 2.0, [0.1])
    with np.errstate(all="ignore"):
        dec, _, alp = ens.sweep_replay(None, np.array([[1.0]]), np.array([[0.5]]), want_decisions=True, want_alpha=True)
    assert np.isnan(alp[0, 0]) and dec[0, 0] == 0         # min(1, NaN) = NaN; NaN > u is false


def test_categorical_scan():
    w = [0.4] + [0.1] * 6
    u = np.array([0.0, 0.39999, 0.4, 0.45, 0.5, 0.95, 0.999999999])
    assert list(N.categorical(w, u)) == [0, 0, 1, 1, 2, 6, 6]
    # C side through moves_out
    ens = O.Ensemble(np.zeros(u.size), 2.0, [0.2] * 7, w)
    _, mov, _ = ens.sweep_replay(u[None, :], np.zeros((1, u.size)), np.full((1, u.size), 0.5), want_decisions=True)
    assert list(mov[0]) == [0, 0, 1, 1, 2, 6, 6]


# ---- golden fixtures -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["replay_single.npz", "replay_multi.npz", "replay_doublewell.npz"])
def test_golden_replay(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    ens = O.Ensemble(g["x0"], float(g["beta"]), g["sigma"], g["weight"], potential=int(g["pot"]))
    dec, mov, _ = ens.sweep_replay(g["u_cat"], g["z"], g["u_acc"], want_decisions=True)
    assert np.array_equal(ens.x, g["x"]) and np.array_equal(ens.e, g["e"])
    assert np.array_equal(ens.acc, g["acc"]) and np.array_equal(ens.tot, g["tot"])
    assert np.array_equal(np.packbits(dec), g["decisions"]) and np.array_equal(mov, g["moves"])
    assert ens.callback_energy() == float(g["energy"])
    assert np.array_equal(ens.callback_acceptance(), g["acceptance"], equal_nan=True)  # NaN: some chain never tried the move
    if int(g["pot"]) == O.POT_HARMONIC:  # independent numpy restatement reproduces the fixture as well
        x, e = g["x0"].copy(), g["x0"] * g["x0"]
        acc, tot = np.zeros_like(g["acc"]), np.zeros_like(g["tot"])
        N.sweep_replay(x, e, float(g["beta"]), g["sigma"], g["weight"], g["u_cat"], g["z"], g["u_acc"], acc, tot)
        assert np.array_equal(x, g["x"]) and np.array_equal(acc, g["acc"])


def test_golden_philox(golden_dir):
    g = np.load(os.path.join(golden_dir, "philox_native.npz"))
    seed, off, K = int(g["seed"]), int(g["offset"]), int(g["K"])
    M = g["x0"].size
    assert np.array_equal(O.init_synthetic(seed, off, M), g["x0"])
    uc, z, ua = O.draws_philox(seed, off, M, 0, K)
    assert np.array_equal(uc[:4], g["u_cat0"]) and np.array_equal(z[:4], g["z0"]) and np.array_equal(ua[:4], g["u_acc0"])
    ens = O.Ensemble(g["x0"], float(g["beta"]), g["sigma"], g["weight"])
    ens.sweep_replay(uc, z, ua)
    assert np.array_equal(ens.x, g["x"]) and np.array_equal(ens.acc, g["acc"]) and np.array_equal(ens.tot, g["tot"])


def test_golden_pgmc(golden_dir):
    g = np.load(os.path.join(golden_dir, "pgmc.npz"))
    ens = O.Ensemble(g["x0"], float(g["beta"]), g["sigma"], [1.0 / g["sigma"].size] * g["sigma"].size)
    gd = ens.pgmc_replay(int(g["q_batch"]), g["learn_ids"], g["z"])
    assert np.array_equal(gd, g["gd"]) and np.array_equal(ens.x, g["x"])
    x, e = g["x0"].copy(), g["x0"] * g["x0"]
    gd_np = N.pgmc_replay(x, e, float(g["beta"]), g["sigma"], list(g["learn_ids"]), g["z"])
    np.testing.assert_allclose(gd_np, gd, rtol=1e-13)     # numpy sums pairwise, C sums sequentially
    assert np.array_equal(x, g["x"])


# ---- the reference's own (statistical) assertions, run on the oracle -------------------------------------------
@pytest.mark.parametrize("beta", [2.0, 2.5, 3.0])
def test_distribution_anchor(beta):
    """test/distribution_test.jl:31-37 at larger M / fewer steps: mean ≈ 0, std ≈ 1/√(2β); ⟨E⟩ = 1/(2β)."""
    M, burn, K = 8192, 1000, 200
    x0 = O.init_synthetic(42, 0, M)
    ens = O.Ensemble(x0, beta, [0.1])
    ens.seed_xoshiro(42)
    ens.sweep_xoshiro(burn)
    xs, es = [], []
    for _ in range(20):
        ens.sweep_xoshiro(K)
        xs.append(ens.x.copy())
        es.append(ens.callback_energy())
    xs = np.concatenate(xs)
    s = 1 / math.sqrt(2 * beta)
    # chains are independent; successive samples K=200 steps apart at σ=0.1 are still correlated → conservative n_eff = M
    assert abs(xs.mean()) < 4 * s / math.sqrt(M)
    assert abs(xs.std() - s) < 4 * s / math.sqrt(2 * M)
    assert abs(np.mean(es) - 1 / (2 * beta)) < 4 * math.sqrt(1 / (2 * beta ** 2) / M)
    # stationary acceptance of a Gaussian random walk on a Gaussian target: (2/π) atan(2s/σ)  
    acc = ens.callback_acceptance()[0]
    assert abs(acc - 2 / math.pi * math.atan(2 * s / 0.1)) < 5e-3


def test_ad_backends_anchor():
    """test/ad_backends_test.jl:11-32: δ = 0, σ = 0.2 → logq = −log(2π·0.04)/2, ∂σ logq = −1/σ = −5 (atol 1e-10)."""
    assert abs(O.log_proposal_density(0.0, 0.2) - (-math.log(2 * math.pi * 0.04) / 2)) < 1e-10
    assert abs(O.dlogq_dsigma(0.0, 0.2) - (-5.0)) < 1e-10
    # analytic gradient == central finite difference of the restated log density elsewhere
    for d, s in [(0.3, 0.2), (-1.1, 0.7), (2.0, 1.2)]:
        h = 1e-6
        fd = (O.log_proposal_density(d, s + h) - O.log_proposal_density(d, s - h)) / (2 * h)
        assert abs(O.dlogq_dsigma(d, s) - fd) < 1e-7


def test_pgmc_record_definition():
    """One sample by hand (gradients.jl:93-109)."""
    x0, beta, sig, z = 0.3, 2.0, 0.5, 0.8
    ens = O.Ensemble(np.array([x0]), beta, [sig])
    gd = ens.pgmc_replay(1, [0], np.array([[[z]]]))[0]
    d = sig * z
    alpha = min(1.0, math.exp(-beta * ((x0 + d) ** 2 - x0 ** 2)))
    gf = d * d / sig ** 3 - 1 / sig
    np.testing.assert_allclose(gd, [d * d * alpha, d * d * alpha * gf, gf, gf * gf, 1.0], rtol=1e-14)
    assert ens.x[0] == (x0 + d) - d


KINDS = [("VPG", O.OPT_VPG, (1e-3, 0.0)), ("BLPG", O.OPT_BLPG, (1e-3, 0.0)), ("BLAPG", O.OPT_BLAPG, (1e-6, 1e-6)),
         ("NPG", O.OPT_NPG, (1e-2, 1e-6)), ("ANPG", O.OPT_ANPG, (1e-6, 1e-6)), ("BLANPG", O.OPT_BLANPG, (1e-6, 1e-6))]


@pytest.mark.parametrize("name,kind,hp", KINDS)
def test_learning_rules_c_vs_numpy(name, kind, hp):
    gd = [0.031, 0.012, -0.4, 2.1]
    th_c = O.learning_step(kind, hp[0], hp[1], gd, 0.2)
    th_n = N.learning_step(name, hp, [0.2], gd[0], [gd[1]], [gd[2]], [[gd[3]]])[0]
    assert abs(th_c - th_n) < 1e-15
    assert th_c != 0.2
    assert O.learning_step(O.OPT_STATIC, 0.0, 0.0, gd, 0.2) == 0.2


def test_pgmc_learns_sigma_on_oracle():
    """test/pgmc_test.jl:47-51 on the oracle at larger M / fewer steps: every learner drives σ towards ≈1.2."""
    M, beta = 2048, 2.0
    sigma = np.full(7, 0.2)
    weight = [0.4] + [0.1] * 6
    x0 = O.init_synthetic(42, 0, M)
    ens = O.Ensemble(x0, beta, sigma, weight)
    ens.seed_xoshiro(42)
    ens.sweep_xoshiro(300)
    # faster learning rates than the reference test (it runs 5·10^4 updates; here 400)
    opts = [(O.OPT_STATIC, 0, 0), (O.OPT_VPG, 0.05, 0), (O.OPT_BLPG, 0.05, 0), (O.OPT_BLAPG, 2e-3, 1e-6),
            (O.OPT_NPG, 0.5, 1e-6), (O.OPT_ANPG, 2e-3, 1e-6), (O.OPT_BLANPG, 2e-3, 1e-6)]
    learn = list(range(1, 7))
    for it in range(400):
        ens.sweep_xoshiro(1)
        gd = ens.pgmc_xoshiro(4, learn)
        for l, k in enumerate(learn):
            avg = gd[l, :4] / gd[l, 4]
            ens.sigma[k] = O.learning_step(opts[k][0], opts[k][1], opts[k][2], avg, ens.sigma[k])
    assert ens.sigma[0] == 0.2
    assert np.all(np.abs(ens.sigma[1:] - 1.2) < 0.2), ens.sigma
    assert abs(ens.callback_energy() - 0.25) < 5e-2       # pgmc_test.jl:45


# ---- build_schedule (simulation.jl:95-117) -------------------------------------------------------------------
def test_build_schedule_restatement():
    assert N.build_schedule(100, 10, 30) == [10, 40, 70, 100]
    assert N.build_schedule(100, 10, 45) == [10, 55, 100]
    s = N.build_schedule(10 ** 5, 1000, [0, 10])          # MC_harmonic_oscillator.jl:18-19
    assert s[0] == 1000 and s[1] == 1010 and s[-1] == 10 ** 5 and len(s) == 9901
    assert N.build_schedule(1000, 10, 2.0) == [10, 11, 12, 14, 18, 26, 42, 74, 138, 266, 522, 1000]
    assert N.build_schedule(100, 0, [0, 3, 10])[:5] == [0, 3, 10, 13, 20]
    with pytest.raises(ValueError):
        N.build_schedule(1000, 10, 1.5)


def test_sha256_and_julia_xoshiro_seeding():
    """Xoshiro(n) of Julia 1.7-1.10 [EXT; pinned by test_julia_rng_known_answers]: SHA-256 of n's 32-bit little-endian limbs, digest
    read as four little-endian UInt64 (metropolis.jl:262-263 builds Xoshiro(seed + c - 1) per chain).  The oracle's C
    SHA-256, hashlib and the product's vectorised numpy hash (montecarlo_b200/julia_rng.py) agree; SHA-256 itself is
    pinned by the FIPS 180-4 known answers."""
    import hashlib
    import struct
    from montecarlo_b200 import julia_rng as J
    kat = {b"abc": "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad",
           b"": "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
           b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq":
               "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"}
    for m, h in kat.items():
        assert O.sha256(m).hex() == h == hashlib.sha256(m).hexdigest()
    for m in (b"a" * 55, b"a" * 56, b"a" * 64, b"a" * 119, bytes(range(200))):      # padding edge cases
        assert O.sha256(m) == hashlib.sha256(m).digest()
    assert J.make_seed(42) == [42] and J.make_seed(2 ** 32) == [0, 1] and J.make_seed(0) == [0]
    seeds = np.array([0, 1, 42, 43, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 17, 2 ** 63 - 1], dtype=np.uint64)
    st = J.xoshiro_states(seeds)
    for sd, row in zip(seeds, st):
        limbs = J.make_seed(int(sd))
        d = hashlib.sha256(struct.pack("<%dI" % len(limbs), *limbs)).digest()
        assert np.array_equal(np.frombuffer(d, dtype="<u8"), row)
        assert np.array_equal(O.xoshiro_seed_julia(int(sd)), row)
    assert np.array_equal(J.xoshiro_state(2 ** 70 + 3), np.frombuffer(
        hashlib.sha256(struct.pack("<3I", 3, 0, 64)).digest(), dtype="<u8"))           # three limbs: hashlib route
    # per-chain seeds seed + c - 1: the oracle's chain seeding == the product's
    ens = O.Ensemble(np.zeros(100), 2.0, [0.1])
    ens.seed_xoshiro(42, chain_offset=7, julia=True)
    assert np.array_equal(ens.states, J.xoshiro_states(np.arange(49, 149)))
    with pytest.raises(ValueError):
        J.make_seed(-1)


# ---- Julia's Random stdlib [EXT]: known answers printed in the Julia manual --------------------------------------
# The reference draws EVERYTHING from it (rngs = [Xoshiro(seed + c - 1) ...] src/metropolis.jl:262-263; rand at
# :184/:206 via Distributions.Categorical, randn at example/particle_1d/particle_1d.jl:57 via Distributions.Normal).
# The docstrings of `Xoshiro` and `randn` (Julia 1.7 - 1.10 manual, Random chapter) print:
#     julia> rng = Xoshiro(1234);  julia> x1 = rand(rng, 2)      ->  0.32597672886359486, 0.5490511363155669
#     julia> rng = Xoshiro(123);   julia> randn(rng, ComplexF64) ->  -0.45660053706486897 - 1.0346749725929225im
#     julia> randn(rng, ComplexF32, (2, 3))  ->  -1.14806-0.153912im  0.056538+1.0954im   0.419454-0.543347im
#                                                 0.34807+0.693657im  -0.948661+0.291442im  -0.0538589-0.463085im
# (randn(rng, Complex{T}) = Complex(SQRT_HALF randn(rng, T), SQRT_HALF randn(rng, T)); arrays fill column-major;
#  randn(rng, Float32) = Float32(randn(rng)) in those versions.)
JULIA_RAND_1234 = [0.32597672886359486, 0.5490511363155669]
JULIA_RANDN_123_C64 = [-0.45660053706486897, -1.0346749725929225]
JULIA_RANDN_123_C32 = [-1.14806, -0.153912, 0.34807, 0.693657, 0.056538, 1.0954, -0.948661, 0.291442,
                       0.419454, -0.543347, -0.0538589, -0.463085]
# leading entries of the literal tables in normal.jl (1-based ki[1], ki[3], wi[1..3], fi[2])
JULIA_KI = {0: 0x0007799ec012f7b2, 1: 0, 2: 0x0006045f4c7de363}
JULIA_WI = {0: 1.7367254121602630e-15, 1: 9.5586603514556339e-17, 2: 1.2708704834810623e-16}
JULIA_FI = {0: 1.0, 1: 9.7710170126767082e-01}


def _julia_sig(v, digits=6):
    """a number the way Julia's 6-significant-digit array display prints it, as a float"""
    return float("%.*g" % (digits, v))


def test_julia_rng_known_answers():
    """Pins the oracle's restatement of Julia's generator bit for bit: SHA-256 seeding, xoshiro256++, rand(Float64) =
    (next >>> 11) 2^-53, and randn's ziggurat (literal tables + fast path) reproduce the manual's printed values."""
    from montecarlo_b200 import julia_rng as J
    st = O.xoshiro_seed_julia(1234)
    assert np.array_equal(st, J.xoshiro_state(1234))
    got = []
    for _ in range(2):
        r, st = O.xoshiro_next(st)
        got.append((r >> 11) * 2.0 ** -53)
    assert got == JULIA_RAND_1234
    z = O.xoshiro_randn_stream(O.xoshiro_seed_julia(123), 14)
    sqrt_half = 0.7071067811865476                      # Float64(SQRT_HALF)
    assert [float(sqrt_half * v) for v in z[:2]] == JULIA_RANDN_123_C64
    c32 = np.float32(0.70710677) * z[2:].astype(np.float32)
    assert [_julia_sig(v) for v in c32] == JULIA_RANDN_123_C32
    ki, wi, fi = O.ziggurat_tables()
    assert all(int(ki[i]) == v for i, v in JULIA_KI.items())
    assert all(float(wi[i]) == v for i, v in JULIA_WI.items())
    assert all(float(fi[i]) == v for i, v in JULIA_FI.items())


def test_ziggurat_tables_are_the_generated_ones():
    """oracle/zig_tables_julia.h and the product's csrc/zig_tables_julia.inc are the same generated data; when mpmath is
    importable the exact-arithmetic recursion is re-run and must give the committed files byte for byte."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(root, "oracle", "zig_tables_julia.h")).read()
    b = open(os.path.join(root, "montecarlo_b200", "csrc", "zig_tables_julia.inc")).read()
    assert a == b and a.count("0x") == 3 * 256
    pytest.importorskip("mpmath")
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "make_zig_tables.py"), "--check"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr


def test_julia_prediction_fixture_is_current(golden_dir):
    """tests/golden/julia_prediction_config1.json -- the falsifiable prediction of the reference's own output for its
    example script (config 1) under Julia 1.7 - 1.10 -- is what the oracle computes today (make_julia_prediction.py)."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_julia_prediction", os.path.join(golden_dir, "make_julia_prediction.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(golden_dir, "julia_prediction_config1.json")))
    got = json.loads(json.dumps(mod.record(*mod.predict())))
    assert got == want
    # the initial condition is 4 rand(Xoshiro(42)) - 2, ten consecutive draws of ONE generator (MC_harmonic_oscillator.jl:10-13)
    assert want["x0"][0] == "0.5173804925704357" and len(want["x0"]) == 10
    assert want["records"] == 9902 and want["energy_head"][0].startswith("0 ") and want["acceptance_head"][0] == "0 [NaN]"
