/*
 * arianna_oracle.c -- CPU restatement of the Arianna.jl multi-chain Metropolis hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (montecarlo_b200/, libarianna_cuda.so) may
 * include, link, import or call this file.  Its only legitimate callers are tests/, the checker in
 * __graft_entry__.smoke(), and bench.py's `cpu_baseline` / `--impl reference` legs.
 *
 * PARITY UNPINNED: the reference (pure Julia) cannot be executed in this environment and its test-suite
 * holds no golden vectors / known-answer values for this path (only statistical assertions, see
 * test/distribution_test.jl:36-37, test/pgmc_test.jl:45,50, test/ad_backends_test.jl:31-32).  This oracle
 * is a line-by-line restatement of the reference sources cited at each function; the statistical
 * assertions of the reference tests and the public known-answer vectors of the RNG building blocks
 * are what pin it: Philox4x32-10 (Random123 kat_vectors), xoshiro256++, and -- for the third-party stream
 * the reference actually draws from, Julia's Random stdlib [EXT] -- the known answers printed in the Julia
 * manual (Xoshiro(1234) -> rand; Xoshiro(123) -> fourteen consecutive randn), which pin the SHA-256
 * seeding, the generator, rand(Float64) and the ziggurat's literal tables + fast path bit for bit.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (see oracle/Makefile).  -ffp-contract=off is
 * REQUIRED: every arithmetic statement below is one IEEE-754 binary64 operation in the reference's order.
 *
 * All citations are path:line under /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define AO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Public algorithm; known-answer vectors from   */
/* Random123's kat_vectors are checked in tests/test_oracle.py.                                      */
/* ------------------------------------------------------------------------------------------------ */
#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

AO_API void ao_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0; k1 += PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Stream layout shared with the CUDA engine's native-Philox mode (DESIGN.md "RNG stream layout").
 *   key  = (tag, 0x41524941)                     tag: 0 = initial condition, 1 = Metropolis, 2 = estimator
 *   ctr  = (sid_lo, p_lo, sid_hi, sub | p_hi<<8) sid = seed + global 0-based chain index  (mirrors the
 *                                                reference's per-chain seed  seed + c - 1, metropolis.jl:262);
 *                                                p = block index of the stream, sub = sub-block of p (0..255)
 *   out  = 4 words -> A = out[0] | out[1] << 32 ; B = out[2] | out[3] << 32
 *   u53(w) = (w >> 11) * 2^-53 in [0,1)          (same map as Julia's rand(Float64) [EXT])               */
#define AO_KEY1 0x41524941u
enum { AO_TAG_INIT = 0, AO_TAG_METROPOLIS = 1, AO_TAG_ESTIMATOR = 2 };

static inline void philox_block(uint64_t sid, uint64_t p, uint32_t sub, uint32_t tag, uint64_t *A, uint64_t *B)
{
    uint32_t ctr[4] = {(uint32_t)sid, (uint32_t)p, (uint32_t)(sid >> 32), sub | ((uint32_t)(p >> 32) << 8)};
    uint32_t key[2] = {tag, AO_KEY1};
    uint32_t o[4];
    ao_philox4x32_10(ctr, key, o);
    *A = (uint64_t)o[0] | ((uint64_t)o[1] << 32);
    *B = (uint64_t)o[2] | ((uint64_t)o[3] << 32);
}

static inline double u53(uint64_t w) { return (double)(w >> 11) * 0x1.0p-53; }
/* (0,1) variant for the Box-Muller radius: odd 52-bit lattice (top 52 bits of the word), so log() never sees 0 and
 * never returns 0; the low 12 bits of the word are left for the accept-uniform prefix of the pair's even step */
static inline double u52_open0(uint64_t w) { return (double)((w >> 12) | 1) * 0x1.0p-52; }

#define AO_TWO_PI 6.283185307179586 /* binary64 nearest of 2pi == Julia's 2π (particle_1d.jl:53) */

/* Box-Muller pair from the two B words of blocks (b0, b1). */
static inline void box_muller(uint64_t B0, uint64_t B1, double *z0, double *z1)
{
    double u1 = u52_open0(B0);
    double u2 = u53(B1);
    double r = sqrt(-2.0 * log(u1));
    double a = AO_TWO_PI * u2;
    *z0 = r * cos(a);
    *z1 = r * sin(a);
}

/* x0 = 4u - 2 : mirrors `System(4rand(rng) - 2, β)` (example/.../MC_harmonic_oscillator.jl:13) with the
 * engine's own counter-based stream (tag 0, block (0, 0), word A). */
AO_API void ao_init_synthetic(int64_t seed, int64_t chain_offset, int64_t M, double *x)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        uint64_t A, B;
        philox_block((uint64_t)(seed + chain_offset + c), 0, 0, AO_TAG_INIT, &A, &B);
        x[c] = 4.0 * u53(A) - 2.0;
    }
}

/* Native-mode Metropolis draws for MC steps t0 .. t0+K-1 of chains [chain_offset, chain_offset+M), written
 * as step-major [K][M] arrays so that native mode == replay of these arrays.
 *   pair p = t >> 1:  sub-block 0: words (A, B): A>>12 -> Box-Muller u1 (|1, 52 bits), B>>11 -> Box-Muller u2;
 *                                 A & 0xfff -> 12-bit PREFIX of u_acc(2p), B & 0x7ff -> 11-bit prefix of u_acc(2p+1)
 *                     sub-block 1: A>>23 -> 41 refinement bits of u_acc(2p), B>>22 -> 42 of u_acc(2p+1)
 *                                 u_acc = ((prefix << (53 - bits)) | refinement) * 2^-53   (the engine only generates
 *                                 this block when its FP32 filter cannot decide from the prefix alone)
 *   quad q = t >> 2:  block (q, sub-block 2): its four 32-bit output words w0..w3 are the categorical uniforms of
 *                                 steps 4q .. 4q+3:  u_cat(t) = w[t & 3] * 2^-32   (only consumed when n_moves > 1; one
 *                                 block serves four steps -- layout v4)
 *   z(2p) = r cos(2π u2), z(2p+1) = r sin(2π u2).
 * u_cat may be NULL (single-move pools do not consume it in native mode). */
AO_API void ao_draws_philox(int64_t seed, int64_t chain_offset, int64_t M, int64_t t0, int64_t K,
                            double *u_cat, double *z, double *u_acc)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        uint64_t sid = (uint64_t)(seed + chain_offset + c);
        for (int64_t s = 0; s < K; ++s) {
            uint64_t t = (uint64_t)(t0 + s);
            uint64_t p = t >> 1;
            uint64_t A0, B0, A1, B1;
            philox_block(sid, p, 0, AO_TAG_METROPOLIS, &A0, &B0);
            philox_block(sid, p, 1, AO_TAG_METROPOLIS, &A1, &B1);
            double z0, z1;
            box_muller(A0, B0, &z0, &z1);
            int odd = (int)(t & 1);
            z[s * M + c] = odd ? z1 : z0;
            /* even step: 12-bit prefix (A0's low 12 bits) + 41 refinement bits; odd step: 11 + 42 */
            uint64_t prefix = odd ? (B0 & 0x7ffu) : (A0 & 0xfffu);
            uint64_t refine = odd ? (B1 >> 22) : (A1 >> 23);
            u_acc[s * M + c] = (double)((prefix << (odd ? 42 : 41)) | refine) * 0x1.0p-53;
            if (u_cat) {
                uint64_t A2, B2;
                philox_block(sid, t >> 2, 2, AO_TAG_METROPOLIS, &A2, &B2);
                uint64_t half = (t & 2) ? B2 : A2;                       /* words (w0, w1) = A, (w2, w3) = B */
                uint32_t w = (t & 1) ? (uint32_t)(half >> 32) : (uint32_t)half;
                u_cat[s * M + c] = (double)w * 0x1.0p-32;
            }
        }
    }
}

/* Estimator normals: sample index q (per chain, 0-based, counted since creation), pair p = q >> 1,
 * block (p, sub 0) with tag 2: A -> u1, B -> u2.  z laid out [n][M] for samples q0 .. q0+n-1. */
AO_API void ao_draws_pgmc_philox(int64_t seed, int64_t chain_offset, int64_t M, int64_t q0, int64_t n, double *z)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        uint64_t sid = (uint64_t)(seed + chain_offset + c);
        for (int64_t s = 0; s < n; ++s) {
            uint64_t q = (uint64_t)(q0 + s);
            uint64_t A, B;
            philox_block(sid, q >> 1, 0, AO_TAG_ESTIMATOR, &A, &B);
            double z0, z1;
            box_muller(A, B, &z0, &z1);
            z[s * M + c] = (q & 1) ? z1 : z0;
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* The Metropolis step: src/metropolis.jl:176-190 (mc_step!) + :203-212 (mc_sweep!) specialised to   */
/* example/particle_1d/particle_1d.jl with potential(x) = x^2 (MC_harmonic_oscillator.jl:4).         */
/* Operation order: SURVEY.md Appendix A.1.                                                          */
/* ------------------------------------------------------------------------------------------------ */

/* potential ids shared with include/arianna_cuda.h */
enum { AO_POT_HARMONIC = 0, AO_POT_QUARTIC = 1, AO_POT_DOUBLE_WELL = 2 };

static inline double potential(int pot, double x)
{
    switch (pot) {
    default:
    case AO_POT_HARMONIC: return x * x;                       /* potential(x) = x^2 */
    case AO_POT_QUARTIC: { double x2 = x * x; return x2 * x2; } /* x^4 = (x^2)^2 (Base.literal_pow) */
    case AO_POT_DOUBLE_WELL: { double w = x * x - 1.0; return w * w; } /* (x^2 - 1)^2 */
    }
}

/* log_proposal_density, particle_1d.jl:52-54:
 *   -(δ)^2 / (2σ^2) - log(2π*σ^2)/2   ==  ((-(δ*δ)) / (2*(σ*σ))) - lognorm
 * lognorm = log((2π)*(σ*σ))/2 is a per-move constant passed in by the caller (host-computed so that the
 * CUDA engine and this oracle consume the identical binary64 value). */
static inline double log_proposal_density(double delta, double sigma, double lognorm)
{
    double s2 = sigma * sigma;
    double t1 = (-(delta * delta)) / (2.0 * s2);
    return t1 - lognorm;
}

AO_API double ao_lognorm(double sigma)
{
    double s2 = sigma * sigma;
    return log(AO_TWO_PI * s2) / 2.0;
}

/* Distributions.Categorical inverse-CDF scan consuming one uniform [EXT], SURVEY.md A.1. 0-based result. */
static inline int categorical(int n, const double *w, double u)
{
    int k = 0;
    double cp = w[0];
    while (cp <= u && k < n - 1) {
        ++k;
        cp = cp + w[k];
    }
    return k;
}

/* One mc_sweep! step (A.1).  Returns the decision; *k_out the chosen move; *alpha_out = α. */
static inline int mc_step_exact(double *x, double *e, double beta, int pot, int n_moves, const double *sigma,
                                const double *weight, const double *lognorm, double u_cat, double z,
                                double u_acc, int *k_out, double *alpha_out)
{
    int k = categorical(n_moves, weight, u_cat);           /* metropolis.jl:206 */
    double delta = 0.0 + (sigma[k] * z);                   /* particle_1d.jl:57  rand(rng, Normal(0, σ)) [EXT] */
    double lqf = log_proposal_density(delta, sigma[k], lognorm[k]); /* metropolis.jl:178 */
    double e1 = *e;                                        /* particle_1d.jl:31 */
    *x = *x + delta;                                       /* :32 */
    *e = potential(pot, *x);                               /* :33 */
    double dlogp = ((-(*e)) * beta) - ((-e1) * beta);      /* metropolis.jl:98, particle_1d.jl:21 */
    delta = -delta;                                        /* particle_1d.jl:38 */
    double lqb = log_proposal_density(delta, sigma[k], lognorm[k]); /* metropolis.jl:182 */
    double arg = (dlogp + lqb) - lqf;                      /* metropolis.jl:183, left-assoc */
    double ex = exp(arg);
    double alpha = (ex > 1.0) ? 1.0 : ex;                  /* min(one(T), ·); NaN propagates as in Julia */
    *k_out = k;
    if (alpha_out) *alpha_out = alpha;
    if (alpha > u_acc) {                                   /* metropolis.jl:184 strict > */
        return 1;
    } else {
        *x = *x + delta;                                   /* perform_action_cached! == perform_action! (:119) */
        *e = potential(pot, *x);                           /* with the NEGATED δ: not an exact restore */
        return 0;
    }
}

/* Replay sweep: K steps for M chains from step-major draw arrays [K][M].
 * acc/tot: [n_moves][M] int64 cumulative counters (Move.accepted_calls / total_calls, metropolis.jl:208-209).
 * decisions/moves (optional): [K][M] uint8.  alpha_out (optional): [K][M]. */
AO_API void ao_sweep_replay(int64_t M, int64_t K, double *x, double *e, double beta, int pot, int n_moves,
                            const double *sigma, const double *weight, const double *lognorm,
                            const double *u_cat, const double *z, const double *u_acc, int64_t *acc,
                            int64_t *tot, uint8_t *decisions, uint8_t *moves, double *alpha_out)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        double xc = x[c], ec = e[c];
        for (int64_t s = 0; s < K; ++s) {
            int k;
            double alpha;
            double uc = u_cat ? u_cat[s * M + c] : 0.0;
            int d = mc_step_exact(&xc, &ec, beta, pot, n_moves, sigma, weight, lognorm, uc, z[s * M + c],
                                  u_acc[s * M + c], &k, &alpha);
            acc[(int64_t)k * M + c] += d;
            tot[(int64_t)k * M + c] += 1;
            if (decisions) decisions[s * M + c] = (uint8_t)d;
            if (moves) moves[s * M + c] = (uint8_t)k;
            if (alpha_out) alpha_out[s * M + c] = alpha;
        }
        x[c] = xc;
        e[c] = ec;
    }
}

/* Replay sweep with per-chain β (parallel-tempering style ensembles; same A.1 arithmetic). */
AO_API void ao_sweep_replay_betas(int64_t M, int64_t K, double *x, double *e, const double *betas, int pot,
                                  int n_moves, const double *sigma, const double *weight, const double *lognorm,
                                  const double *u_cat, const double *z, const double *u_acc, int64_t *acc,
                                  int64_t *tot, uint8_t *decisions)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        double xc = x[c], ec = e[c];
        for (int64_t s = 0; s < K; ++s) {
            int k;
            double uc = u_cat ? u_cat[s * M + c] : 0.0;
            int d = mc_step_exact(&xc, &ec, betas[c], pot, n_moves, sigma, weight, lognorm, uc, z[s * M + c],
                                  u_acc[s * M + c], &k, NULL);
            acc[(int64_t)k * M + c] += d;
            tot[(int64_t)k * M + c] += 1;
            if (decisions) decisions[s * M + c] = (uint8_t)d;
        }
        x[c] = xc;
        e[c] = ec;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Float32 ensembles: Particle{Float32}, Displacement{Float32}, ComponentArray(σ = 0.1f0)             */
/* (example/particle_1d/particle_1d.jl:9-16,26-28: the system is generic in T<:AbstractFloat; mc_step!'s  */
/* T is the PARAMETERS' element type, metropolis.jl:176).  Julia's promotion rules decide which        */
/* operation runs in which precision [EXT Base/Distributions semantics, restated]:                    */
/*   δ   = zero(δ) + σ*randn(rng, Float32)            Float32;  randn(rng, Float32) = Float32(randn(rng))  */
/*   lq  = -(δ)^2 / (2σ^2)  -  log(2π*σ^2)/2          Float32 quotient  -  Float64 (2π is Float64)  => Float64 */
/*   x  += δ;  e = x^2;  Δ = (-e)*β - (-e1)*β         Float32                                            */
/*   α   = min(one(Float32), exp((Δ + lq_b) - lq_f))  Float64 (Float32 + Float64);  α > rand(rng) in Float64 */
/* One IEEE operation per statement; compile with -ffp-contract=off (float expressions stay float:    */
/* FLT_EVAL_METHOD == 0 on x86-64).                                                                   */
/* ------------------------------------------------------------------------------------------------ */
static inline float potential_f32(int pot, float x)
{
    switch (pot) {
    default:
    case AO_POT_HARMONIC: return x * x;
    case AO_POT_QUARTIC: { float x2 = x * x; return x2 * x2; }
    case AO_POT_DOUBLE_WELL: { float w = x * x - 1.0f; return w * w; }
    }
}

/* log(2π·σ²)/2 for a Float32 σ: 2π*σ^2 = Float64(2π) * Float64(σ*σ)  (the Float32 square is promoted) */
AO_API double ao_lognorm_f32(float sigma)
{
    float s2 = sigma * sigma;
    return log(AO_TWO_PI * (double)s2) / 2.0;
}

static inline int mc_step_exact_f32(float *x, float *e, float beta, int pot, float sigma, double lognorm, double z64,
                                    double u_acc, double *alpha_out)
{
    float z = (float)z64;                                  /* randn(rng, Float32) = Float32(randn(rng)) [EXT] */
    float delta = 0.0f + (sigma * z);                      /* particle_1d.jl:57 */
    float s2 = sigma * sigma;
    float t1 = (-(delta * delta)) / (2.0f * s2);           /* :53, Float32 */
    double lqf = (double)t1 - lognorm;                     /* Float32 - Float64 */
    float e1 = *e;                                         /* :31 */
    *x = *x + delta;                                       /* :32 */
    *e = potential_f32(pot, *x);                           /* :33 */
    float dlogp = ((-(*e)) * beta) - ((-e1) * beta);       /* metropolis.jl:98, Float32 */
    delta = -delta;                                        /* particle_1d.jl:38 */
    float t1b = (-(delta * delta)) / (2.0f * s2);
    double lqb = (double)t1b - lognorm;                    /* metropolis.jl:182 */
    double arg = ((double)dlogp + lqb) - lqf;              /* :183 */
    double ex = exp(arg);
    double alpha = (ex > 1.0) ? 1.0 : ex;                  /* min(one(Float32), ·) promotes to Float64 */
    if (alpha_out) *alpha_out = alpha;
    if (alpha > u_acc) return 1;                           /* :184 */
    *x = *x + delta;                                       /* :187 */
    *e = potential_f32(pot, *x);
    return 0;
}

/* Replay sweep of a Float32 ensemble (single-move pool): draws are the SAME Float64 arrays as for Float64 ensembles
 * (z is rounded to Float32 inside, as randn(rng, Float32) does). */
AO_API void ao_sweep_replay_f32(int64_t M, int64_t K, float *x, float *e, float beta, const float *betas, int pot,
                                float sigma, const double *z, const double *u_acc, int64_t *acc, uint8_t *decisions)
{
    double lognorm = ao_lognorm_f32(sigma);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        float xc = x[c], ec = e[c];
        float b = betas ? betas[c] : beta;
        for (int64_t s = 0; s < K; ++s) {
            int d = mc_step_exact_f32(&xc, &ec, b, pot, sigma, lognorm, z[s * M + c], u_acc[s * M + c], NULL);
            acc[c] += d;
            if (decisions) decisions[s * M + c] = (uint8_t)d;
        }
        x[c] = xc;
        e[c] = ec;
    }
}

/* The engine's native Float32 stream (csrc/kernels_f32.cuh): tag 3, ONE block per pair of steps p = t >> 1, words
 * (w0, w1, w2, w3): u1 = ((w0 >> 8) + 1/2) 2^-24, angle = w1 2^-32 turns, z(2p) = r cos, z(2p+1) = r sin with
 * r = sqrt(-2 ln u1); accept uniforms w2 2^-32 and w3 2^-32.  The device evaluates the Box-Muller in FP32 on the MUFU
 * pipe; here it is evaluated in Float64 libm and rounded: the two agree to ~1e-6 relative, so this only follows the
 * device loosely (native-mode parity is statistical). */
AO_API void ao_draws_philox_f32(int64_t seed, int64_t chain_offset, int64_t M, int64_t t0, int64_t K, double *z,
                                double *u_acc)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        uint64_t sid = (uint64_t)(seed + chain_offset + c);
        for (int64_t s = 0; s < K; ++s) {
            uint64_t t = (uint64_t)(t0 + s), A, B;
            philox_block(sid, t >> 1, 0, 3, &A, &B);
            uint32_t w0 = (uint32_t)A, w1 = (uint32_t)(A >> 32), w2 = (uint32_t)B, w3 = (uint32_t)(B >> 32);
            double u1 = ((double)(w0 >> 8) + 0.5) * 0x1.0p-24;
            double r = sqrt(-2.0 * log(u1));
            double a = AO_TWO_PI * ((double)w1 * 0x1.0p-32);
            z[s * M + c] = (double)(float)((t & 1) ? r * sin(a) : r * cos(a));
            u_acc[s * M + c] = (double)((t & 1) ? w3 : w2) * 0x1.0p-32;
        }
    }
}

/* callback_energy for Float32 chains: mean(system.e for system in chains) accumulates in Float32 [EXT Statistics]. */
AO_API float ao_callback_energy_f32(int64_t M, const float *e)
{
    float s = 0.0f;
    for (int64_t c = 0; c < M; ++c) s = s + e[c];
    return s / (float)M;
}

/* ------------------------------------------------------------------------------------------------ */
/* Callbacks                                                                                         */
/* ------------------------------------------------------------------------------------------------ */

/* callback_energy, particle_1d.jl:68-70: mean(system.e for system in chains) -- sequential sum [EXT]. */
AO_API double ao_callback_energy(int64_t M, const double *e)
{
    double s = 0.0;
    for (int64_t c = 0; c < M; ++c) s = s + e[c];
    return s / (double)M;
}

/* callback_acceptance, metropolis.jl:319-321: per-move mean over chains of accepted_calls/total_calls;
 * 0/0 = NaN propagates (the t = 0 store_first record). */
AO_API void ao_callback_acceptance(int64_t M, int n_moves, const int64_t *acc, const int64_t *tot, double *out)
{
    for (int k = 0; k < n_moves; ++k) {
        double s = 0.0;
        for (int64_t c = 0; c < M; ++c)
            s = s + (double)acc[(int64_t)k * M + c] / (double)tot[(int64_t)k * M + c];
        out[k] = s / (double)M;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* PGMC estimator: src/PolicyGuided/gradients.jl:93-121 (pgmc_estimate / sample_gradient_data) and    */
/* estimator.jl:111-134 (fold over chains x q_batch per learnable move).  SURVEY.md A.2.             */
/* Analytic ∂σ log q = δ²/σ³ − 1/σ replaces the AD backends (gradients.jl:28-33, ext/ZygoteExt.jl, ext/EnzymeExt.jl).          */
/* ------------------------------------------------------------------------------------------------ */
AO_API double ao_dlogq_dsigma(double delta, double sigma)
{
    return (delta * delta) / (sigma * sigma * sigma) - 1.0 / sigma;
}

AO_API double ao_log_proposal_density(double delta, double sigma)
{
    return log_proposal_density(delta, sigma, ao_lognorm(sigma));
}

/* z: [n_learn][q_batch][M].  gd_out: [n_learn][5] = sums of (j, ∇j, ∇logq_forward, g, n) in the fold order of
 * estimator.jl:113-129 (chains outer, batch inner).  Chain state drifts by the perform/undo rounding. */
AO_API void ao_pgmc_replay(int64_t M, int q_batch, int n_learn, const int *learn_ids, double *x, double *e,
                           double beta, int pot, const double *sigma, const double *lognorm, const double *z,
                           double *gd_out)
{
    for (int l = 0; l < n_learn; ++l) {                     /* estimator.jl:112 */
        int k = learn_ids[l];
        double sj = 0.0, sdj = 0.0, sgf = 0.0, sg = 0.0, sn = 0.0;
        for (int64_t c = 0; c < M; ++c) {
            for (int b = 0; b < q_batch; ++b) {
                double zz = z[((int64_t)l * q_batch + b) * M + c];
                double delta = 0.0 + (sigma[k] * zz);       /* sample_action! gradients.jl:119 */
                double gf = ao_dlogq_dsigma(delta, sigma[k]); /* ∇logq_forward :97 (analytic) */
                double lqf = log_proposal_density(delta, sigma[k], lognorm[k]);
                double e1 = e[c];                           /* perform_action! :98 */
                x[c] = x[c] + delta;
                e[c] = potential(pot, x[c]);
                double dlogp = ((-e[c]) * beta) - ((-e1) * beta); /* :99 */
                double r = delta * delta;                   /* reward :100, particle_1d.jl:42-44 */
                delta = -delta;                             /* :101 */
                double gb = ao_dlogq_dsigma(delta, sigma[k]); /* :102 */
                double lqb = log_proposal_density(delta, sigma[k], lognorm[k]);
                x[c] = x[c] + delta;                        /* perform_action_cached! :103 (undo, with drift) */
                e[c] = potential(pot, x[c]);
                double ex = exp((dlogp + lqb) - lqf);       /* :104 */
                double alpha = (ex > 1.0) ? 1.0 : ex;
                double j = r * alpha;                       /* :105 */
                double dj = j * ((alpha == 1.0) ? gf : gb); /* :106 */
                double g = gf * gf;                         /* :107 */
                sj = sj + j; sdj = sdj + dj; sgf = sgf + gf; sg = sg + g; sn = sn + 1.0; /* :68-76 */
            }
        }
        gd_out[l * 5 + 0] = sj; gd_out[l * 5 + 1] = sdj; gd_out[l * 5 + 2] = sgf;
        gd_out[l * 5 + 3] = sg; gd_out[l * 5 + 4] = sn;
    }
}

/* learning_step!, src/PolicyGuided/learning.jl, P = 1.  gd = AVERAGED GradientData (gradients.jl:83-85):
 * gd[0] = j, gd[1] = ∇j, gd[2] = ∇logq_forward, gd[3] = g.  Returns the new θ. */
enum { AO_OPT_STATIC = 0, AO_OPT_VPG = 1, AO_OPT_BLPG = 2, AO_OPT_BLAPG = 3, AO_OPT_NPG = 4, AO_OPT_ANPG = 5,
       AO_OPT_BLANPG = 6 };

AO_API double ao_learning_step(int kind, double p1, double p2, const double *gd, double theta)
{
    double j = gd[0], dj = gd[1], glq = gd[2], g = gd[3];
    switch (kind) {
    case AO_OPT_VPG: return theta + p1 * dj;                               /* learning.jl:32-34 */
    case AO_OPT_BLPG: return theta + p1 * (dj - j * glq);                   /* :50-52 */
    case AO_OPT_BLAPG: {                                                    /* :76-79 */
        double eta = sqrt(2 * p1 / (dj * dj + p2));
        return theta + eta * (dj - j * glq);
    }
    case AO_OPT_NPG: return theta + p1 * (1.0 / (g + p2 * 1.0)) * dj;       /* :103-105 */
    case AO_OPT_ANPG: {                                                     /* :130-134 */
        double Finv = 1.0 / (g + p2 * 1.0);
        double eta = sqrt(2 * p1 / (dj * (Finv * dj)));
        return theta + eta * Finv * dj;
    }
    case AO_OPT_BLANPG: {                                                   /* :159-164 */
        double Finv = 1.0 / (g + p2 * 1.0);
        double bj = dj - j * glq;
        double eta = sqrt(2 * p1 / (bj * (Finv * bj)));
        return theta + eta * Finv * bj;
    }
    default: return theta;                                                  /* Static: skipped, estimator.jl:72 */
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* Julia's RNG front-end [EXT: stdlib Random, Julia 1.7 - 1.10], restated from its published algorithm */
/* and PINNED by the known answers printed in the Julia manual (Xoshiro(1234) -> rand, Xoshiro(123) -> */
/* randn; tests/test_oracle.py::test_julia_rng_known_answers): SHA-256 seeding, xoshiro256++, rand,    */
/* the ziggurat's tables and fast path.  The wedge / tail slow paths (1 %) are restated, not pinned.   */
/* xoshiro256++ (Blackman & Vigna); self-check: state (1,2,3,4) -> first output 41943041.             */
/* rand(Float64) = (next >> 11) * 2^-53;  randn = 256-layer ziggurat (Marsaglia-Tsang / Doornik ZIGNOR */
/* as in Julia's Random/normal.jl, with Julia's literal tables (zig_tables_julia.h).                   */
/* ------------------------------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }

AO_API uint64_t ao_xoshiro_next(uint64_t s[4])
{
    uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
    uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl64(s[3], 45);
    return r;
}

/* Default seeding when the host does not upload real Xoshiro states: splitmix64 of the per-chain seed
 * seed + c - 1 (metropolis.jl:262).  NOT Julia's seeding (SHA-256 based, version dependent) [EXT]. */
AO_API void ao_xoshiro_seed(uint64_t seed, uint64_t s[4])
{
    uint64_t zz = seed;
    for (int i = 0; i < 4; ++i) {
        zz += 0x9E3779B97F4A7C15ull;
        uint64_t v = zz;
        v = (v ^ (v >> 30)) * 0xBF58476D1CE4E5B9ull;
        v = (v ^ (v >> 27)) * 0x94D049BB133111EBull;
        s[i] = v ^ (v >> 31);
    }
}

/* Julia's own seeding of `Xoshiro(n::Integer)` [EXT Random stdlib, Julia 1.7 - 1.10; PINNED by the Julia manual's known answers]:
 *   seed!(rng, n) = seed!(rng, make_seed(n));  make_seed(n) = the 32-bit limbs of n, least significant first;
 *   seed!(rng, v::Vector{UInt32}): s0..s3 = reinterpret(UInt64, sha256(reinterpret(UInt8, v)))
 * i.e. the SHA-256 digest of the limbs' little-endian bytes read as four little-endian 64-bit words.  This is what
 * `rngs = [Xoshiro(seed + c - 1) ...]` (metropolis.jl:262-263) evaluates to.  SHA-256 per FIPS 180-4 (KATs in
 * tests/test_oracle.py).  Julia 1.11 changed the scheme. */
static const uint32_t sha_k[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t rotr32(uint32_t v, int k) { return (v >> k) | (v << (32 - k)); }

static void sha256_block(uint32_t h[8], const uint8_t blk[64])
{
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t t1 = hh + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + sha_k[i] + w[i];
        uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

/* SHA-256 of an arbitrary message -> 32 digest bytes. */
AO_API void ao_sha256(const uint8_t *msg, int64_t len, uint8_t out[32])
{
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    int64_t off = 0;
    for (; off + 64 <= len; off += 64) sha256_block(h, msg + off);
    uint8_t blk[128];
    int64_t rem = len - off;
    memset(blk, 0, sizeof blk);
    memcpy(blk, msg + off, (size_t)rem);
    blk[rem] = 0x80;
    int nb = rem + 9 <= 64 ? 1 : 2;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; ++i) blk[64 * nb - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (int i = 0; i < nb; ++i) sha256_block(h, blk + 64 * i);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
}

AO_API void ao_xoshiro_seed_julia(uint64_t seed, uint64_t s[4])
{
    uint8_t msg[8], dig[32];
    int len = (seed >> 32) ? 8 : 4;                         /* make_seed: one limb unless n >= 2^32 */
    for (int i = 0; i < len; ++i) msg[i] = (uint8_t)(seed >> (8 * i));
    ao_sha256(msg, len, dig);
    for (int k = 0; k < 4; ++k) {
        uint64_t v = 0;
        for (int i = 0; i < 8; ++i) v |= (uint64_t)dig[8 * k + i] << (8 * i);
        s[k] = v;
    }
}

/* states: [M][4]; chain c (0-based) gets Xoshiro(seed + chain_offset + c) == the reference's seed + c - 1 (1-based) */
AO_API void ao_xoshiro_seed_chains_julia(int64_t seed, int64_t chain_offset, int64_t M, uint64_t *states)
{
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) ao_xoshiro_seed_julia((uint64_t)(seed + chain_offset + c), states + 4 * c);
}

#define ZIG_R 3.6541528853610088       /* ziggurat_nor_r (normal.jl) rounded to binary64 */
#define ZIG_INV_R 0.27366123732975828 /* inv(ziggurat_nor_r) in binary64 */
/* Julia's literal tables ki / wi / fi (normal.jl [EXT]): the randmtzig recursion evaluated exactly and rounded   */
/* once (scripts/make_zig_tables.py), NOT randmtzig.c's double-precision recipe (that one is 3e-12..3e-10 off and  */
/* moves every draw in its 12th digit).  Pinned by the Julia manual's known answers: Xoshiro(123);                */
/* randn(rng, ComplexF64) = -0.45660053706486897 - 1.0346749725929225im (tests/test_oracle.py).                   */
#include "zig_tables_julia.h"
#define zig_ki kZigKiJulia
#define zig_wi kZigWiJulia
#define zig_fi kZigFiJulia
static void zig_init(void) {}

/* Expose the tables so the CUDA engine's XOSHIRO mode can be handed the *identical* binary64 tables. */
AO_API void ao_ziggurat_tables(uint64_t *ki, double *wi, double *fi)
{
    zig_init();
    memcpy(ki, zig_ki, sizeof zig_ki);
    memcpy(wi, zig_wi, sizeof zig_wi);
    memcpy(fi, zig_fi, sizeof zig_fi);
}

static inline double xo_rand(uint64_t s[4]) { return (double)(ao_xoshiro_next(s) >> 11) * 0x1.0p-53; }

static double xo_randn(uint64_t s[4])
{
    for (;;) {
        uint64_t r = ao_xoshiro_next(s) >> 12; /* 52 random bits */
        int64_t rabs = (int64_t)(r >> 1);
        int idx = (int)(rabs & 0xFF);
        double x = (double)((r & 1) ? -rabs : rabs) * zig_wi[idx];
        if ((uint64_t)rabs < zig_ki[idx]) return x; /* ~99% */
        if (idx == 0) {
            for (;;) {
                double xx = -ZIG_INV_R * log(xo_rand(s));
                double yy = -log(xo_rand(s));
                if (yy + yy > xx * xx) return ((rabs >> 8) & 1) ? -ZIG_R - xx : ZIG_R + xx;
            }
        } else if ((zig_fi[idx - 1] - zig_fi[idx]) * xo_rand(s) + zig_fi[idx] < exp(-0.5 * x * x)) {
            return x;
        }
        /* else: retry (tail-recursive randn(rng) in Julia) */
    }
}

AO_API double ao_xoshiro_rand(uint64_t s[4]) { return xo_rand(s); }
AO_API double ao_xoshiro_randn(uint64_t s[4])
{
    zig_init();
    return xo_randn(s);
}

/* states: [M][4].  Default per-chain seeding seed + c (0-based c) == reference seed + c - 1 (1-based). */
AO_API void ao_xoshiro_seed_chains(int64_t seed, int64_t chain_offset, int64_t M, uint64_t *states)
{
    for (int64_t c = 0; c < M; ++c) ao_xoshiro_seed((uint64_t)(seed + chain_offset + c), states + 4 * c);
}

/* Reference draw order per step (metropolis.jl:206 then particle_1d.jl:57 then metropolis.jl:184):
 * u_cat = rand, z = randn, u_acc = rand.  Arrays are step-major [K][M]. */
AO_API void ao_draws_xoshiro(int64_t M, int64_t K, uint64_t *states, double *u_cat, double *z, double *u_acc)
{
    zig_init();
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        uint64_t s[4];
        memcpy(s, states + 4 * c, sizeof s);
        for (int64_t t = 0; t < K; ++t) {
            u_cat[t * M + c] = xo_rand(s);
            z[t * M + c] = xo_randn(s);
            u_acc[t * M + c] = xo_rand(s);
        }
        memcpy(states + 4 * c, s, sizeof s);
    }
}

/* The CPU baseline: the reference's Metropolis.make_step! with parallel=true (Transducers.tcollect over
 * chains, metropolis.jl:265,303-307) == OpenMP over chains; each chain runs K x A.1 with its own generator.
 * Returns nothing; x/e/acc/tot/states updated in place. */
AO_API void ao_sweep_xoshiro(int64_t M, int64_t K, double *x, double *e, double beta, int pot, int n_moves,
                             const double *sigma, const double *weight, const double *lognorm, uint64_t *states,
                             int64_t *acc, int64_t *tot)
{
    zig_init();
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < M; ++c) {
        uint64_t s[4];
        memcpy(s, states + 4 * c, sizeof s);
        double xc = x[c], ec = e[c];
        for (int64_t t = 0; t < K; ++t) {
            double uc = xo_rand(s);
            double zz = xo_randn(s);
            double ua = xo_rand(s);
            int k;
            int d = mc_step_exact(&xc, &ec, beta, pot, n_moves, sigma, weight, lognorm, uc, zz, ua, &k, NULL);
            acc[(int64_t)k * M + c] += d;
            tot[(int64_t)k * M + c] += 1;
        }
        x[c] = xc;
        e[c] = ec;
        memcpy(states + 4 * c, s, sizeof s);
    }
}

/* CPU baseline for the estimator (estimator.jl:111-134 with foldxt == OpenMP reduction over chains).
 * Uses the per-chain generator for the normals; returns the summed record per learnable move. */
AO_API void ao_pgmc_xoshiro(int64_t M, int q_batch, int n_learn, const int *learn_ids, double *x, double *e,
                            double beta, int pot, const double *sigma, const double *lognorm, uint64_t *states,
                            double *gd_out)
{
    zig_init();
    for (int l = 0; l < n_learn; ++l) {
        int k = learn_ids[l];
        double sj = 0.0, sdj = 0.0, sgf = 0.0, sg = 0.0, sn = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : sj, sdj, sgf, sg, sn)
        for (int64_t c = 0; c < M; ++c) {
            uint64_t s[4];
            memcpy(s, states + 4 * c, sizeof s);
            for (int b = 0; b < q_batch; ++b) {
                double zz = xo_randn(s);
                double gd1[5];
                int one = 0;
                double xs = x[c], es = e[c];
                ao_pgmc_replay(1, 1, 1, &one, &xs, &es, beta, pot, sigma + k, lognorm + k, &zz, gd1);
                x[c] = xs; e[c] = es;
                sj += gd1[0]; sdj += gd1[1]; sgf += gd1[2]; sg += gd1[3]; sn += gd1[4];
            }
            memcpy(states + 4 * c, s, sizeof s);
        }
        gd_out[l * 5 + 0] = sj; gd_out[l * 5 + 1] = sdj; gd_out[l * 5 + 2] = sgf;
        gd_out[l * 5 + 3] = sg; gd_out[l * 5 + 4] = sn;
    }
}

/* Team size of the OpenMP loops above; launchers such as torchrun export OMP_NUM_THREADS=1, which the CPU arm of
 * bench.py must override to time the reference path on ALL host cores. */
AO_API void ao_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

AO_API int ao_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
