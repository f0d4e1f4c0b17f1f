// kernels_f32.cuh -- Float32 ensembles: Particle{Float32}, Displacement{Float32}, ComponentArray(σ = 0.1f0).
//
// The reference's system is generic in T <: AbstractFloat (example/particle_1d/particle_1d.jl:9-16, 26-28) and
// mc_step! is generic in the parameters' element type (src/metropolis.jl:176).  Julia's promotion rules decide the
// precision of every operation of the Float32 instantiation [EXT Base / Distributions, restated in
// oracle/arianna_oracle.c:mc_step_exact_f32]:
//     δ  = zero(δ) + σ·randn(rng, Float32)                       Float32   (randn(rng, Float32) = Float32(randn(rng)))
//     lq = −(δ)²/(2σ²) − log(2π·σ²)/2                            Float32 quotient − Float64 constant  =>  Float64
//     x += δ;  e = potential(x);  Δ = (−e)·β − (−e₁)·β           Float32
//     α  = min(one(Float32), exp((Δ + lq_b) − lq_f)) > rand(rng) Float64
// HBM layout: x[M] f32 (4 B per chain), acc[M] u32; single-move pools, native Philox stream or replay.
//
//  * replay (sweep_replay_f32_kernel): the caller's Float64 draws, EXACT operation order -> decisions and Float32
//    positions bit-identical to the oracle.
//  * native (sweep_f32_kernel): a stream of its own (tag 3): ONE Philox block per pair of steps, words
//    (w0, w1, w2, w3): Box-Muller in FP32 on the MUFU pipe -- u1 = ((w0 >> 8) + ½)·2^-24, angle = w1·2^-32 turns,
//    r = √(−2 ln u1) through LG2 / SQRT, sin / cos through MUFU.SIN / MUFU.COS -- and the accept uniforms of the two
//    steps are w2·2^-32 and w3·2^-32 (their top 12 bits feed the integer-domain accept filter, the exact 32-bit value
//    the rare FP64 fallback: the decision is exactly `exp_f64(Float64(Δ)) > u`, as the reference evaluates it).
//    FP32 lanes are twice as many as FP64 ones and nothing here touches the FP64 pipe outside the rare fallback and
//    the energy sum of the fused reduction (accumulated in FP64 on purpose: the reference's Float32 `mean` loses
//    digits at 2^27 chains).  Parity of the native stream is statistical (3σ), as the north star asks; the oracle also
//    follows it loosely (same words, libm Float64 Box-Muller rounded to Float32).
#pragma once

namespace arianna {

constexpr uint32_t kTagMetropolisF32 = 3;

struct F32Params {
    float *x;
    uint32_t *acc;
    const float *betas;
    float beta;
    int64_t M;
    int64_t K;
    int64_t t0;
    uint64_t sid0;
    float sigma;
    double lognorm;         // log(2π·Float64(σ·σ))/2 in Float64 (particle_1d.jl:53 under promotion)
    int reduce;
    double *partials;
    unsigned int *ticket;
    double *sums;
    const m64::MathTables *tables;
    const double *z;        // replay: [K][M] Float64 draws (rounded to Float32 inside, like randn(rng, Float32))
    const double *u_acc;    // replay: [K][M]
    uint8_t *decisions;     // replay: [K][M] or nullptr
};

template <int POT>
__device__ __forceinline__ float potential_f32(float x)
{
    if constexpr (POT == POT_HARMONIC) {
        return __fmul_rn(x, x);
    } else if constexpr (POT == POT_QUARTIC) {
        const float x2 = __fmul_rn(x, x);
        return __fmul_rn(x2, x2);
    } else {
        const float w = __fsub_rn(__fmul_rn(x, x), 1.0f);
        return __fmul_rn(w, w);
    }
}

// EXACT: one IEEE operation per statement in the reference's order and precision (never contracted).
// `decide(arg)` evaluates  min(1, exp(arg)) > u  for the Float64 arg.
template <int POT, class Decide>
__device__ __forceinline__ int mc_step_f32_exact(float &x, float &e, float beta, float sigma, double lognorm, float z,
                                                 Decide decide)
{
    float delta = __fadd_rn(0.0f, __fmul_rn(sigma, z));                              // particle_1d.jl:57
    const float s2 = __fmul_rn(sigma, sigma);
    const float t1 = __fdiv_rn(-__fmul_rn(delta, delta), __fmul_rn(2.0f, s2));       // :53 (Float32)
    const double lqf = __dsub_rn((double)t1, lognorm);                               // Float32 − Float64
    const float e1 = e;
    x = __fadd_rn(x, delta);                                                         // :32
    e = potential_f32<POT>(x);                                                       // :33
    const float dlogp = __fsub_rn(__fmul_rn(-e, beta), __fmul_rn(-e1, beta));        // metropolis.jl:98
    delta = -delta;                                                                  // particle_1d.jl:38
    const double lqb = lqf;                                                          // only δ·δ enters: bitwise equal
    const double arg = __dsub_rn(__dadd_rn((double)dlogp, lqb), lqf);                // :183
    if (decide(arg)) return 1;
    x = __fadd_rn(x, delta);                                                         // :187
    e = potential_f32<POT>(x);
    return 0;
}

// The integer-domain accept filter of math64.cuh for a Float32 argument (no rounding of the argument: the bound of
// exp_accept_prefix holds a fortiori); ambiguous cases evaluate exp in FP64 against the exact 32-bit uniform.
template <int PBITS>
__device__ __forceinline__ bool exp_accept_prefix_f32(float a, uint32_t f, uint32_t w, m64::Tab tb)
{
    const float Es = m64::ex2_approx(fmaf(a, 1.44269504f, (float)PBITS));
    const float v = m64::fma_floor_offset(Es, f);               // floor(Es(1 − 2^-13)) − f − 1.5·2^23 (math64.cuh)
    bool acc = (a >= 0.0f) || (v > -m64::kFloorMagic);
    const bool rej = v < -(m64::kFloorMagic + 1.0f);
    if (!(acc || rej)) {
        const double x = (double)a;
        const uint32_t t = (uint32_t)__double2hiint(x) - 0x7ff00000u;
        const bool core = (t - 0x00100000u) < (0x40962000u - 0x00100000u);
        const bool tiny_pos = t >= 0x80100000u;
        acc = tiny_pos || (core && (m64::exp_core(x, tb) > __dmul_rn((double)w, 0x1p-32)));
    }
    return acc;
}

// FP32 Box-Muller of the native stream (see the header).  Fast-math intrinsics on purpose: MUFU.LG2 / SQRT / SIN / COS.
__device__ __forceinline__ void box_muller_f32(uint32_t w0, uint32_t w1, float &z0, float &z1)
{
    const float u1 = fmaf((float)(w0 >> 8), 0x1p-24f, 0x1p-25f);        // ((w0 >> 8) + ½)·2^-24 ∈ (0, 1), exact
    const float r = __fsqrt_rn(-1.38629436f * __log2f(u1));             // √(−2 ln 2 · log2 u1)
    const float ang = fmaf((float)w1, 1.46291808e-09f, -3.14159274f);   // 2π·w1·2^-32 − π ∈ [−π, π]
    float s, c;
    __sincosf(ang, &s, &c);
    z0 = -r * c;                                                        // cos(θ) = −cos(θ − π)
    z1 = -r * s;
}

template <int POT, int ARITH>
__global__ void __launch_bounds__(kBlock, 4) sweep_f32_kernel(const F32Params p)
{
    __shared__ m64::MathTables s_T;
    load_tables(&s_T, p.tables);
    __syncthreads();
    const m64::Tab tb = shared_tab(&s_T);
    const int64_t tend = p.t0 + p.K;
    double sum_e = 0.0;
    unsigned long long sum_acc = 0ull;
    uint32_t cnt = 0;
    const float sigma = p.sigma;
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < p.M; c += (int64_t)gridDim.x * kBlock) {
        float x = p.x[c];
        float e = potential_f32<POT>(x);
        uint32_t acc = p.acc[c];
        const float beta = p.betas ? p.betas[c] : p.beta;
        const PhiloxChain<kTagMetropolisF32, 0> ph(p.sid0 + (uint64_t)c);
        auto step = [&](float z, uint32_t w) {
            const uint32_t f = w >> 20;                                  // 12-bit prefix of u = w·2^-32
            if constexpr (ARITH == ARITH_EXACT) {
                acc += mc_step_f32_exact<POT>(x, e, beta, sigma, p.lognorm, z, [&](double arg) {
                    return m64::exp_accept_prefix<12>(arg, f, [&]() { return __dmul_rn((double)w, 0x1p-32); }, tb);
                });
            } else {
                // FAST: symmetric proposal -> the log q terms cancel; accept iff exp_f64(Float64(β(e − e'))) > u
                const float e0 = potential_f32<POT>(x);
                const float xn = fmaf(sigma, z, x);
                const float en = potential_f32<POT>(xn);
                const bool a = exp_accept_prefix_f32<12>(beta * (e0 - en), f, w, tb);
                x = a ? xn : x;
                count_if(acc, a ? 1u : 0u);
            }
        };
        auto pair = [&](uint32_t pr, bool do0, bool do1) {
            const U64Pair b = ph.block(pr);
            float z0, z1;
            box_muller_f32(b.a_lo, b.a_hi, z0, z1);
            if (do0) step(z0, b.b_lo);
            if (do1) step(z1, b.b_hi);
        };
        const uint32_t ta = (uint32_t)p.t0, te = (uint32_t)tend;
        if (te > ta) {
            uint32_t pr = ta >> 1;
            if (ta & 1u) { pair(pr, false, true); ++pr; }
            const uint32_t pr_end = te >> 1;
#pragma unroll 1
            for (; pr < pr_end; ++pr) pair(pr, true, true);
            if (te & 1u) pair(pr, true, false);
        }
        p.x[c] = x;
        p.acc[c] = acc;
        if (p.reduce) {
            sum_e += (double)potential_f32<POT>(x);
            sum_acc += acc;
            ++cnt;
        }
    }
    if (p.reduce) {
        double vals[3] = {sum_e, (double)sum_acc / (double)tend, (double)cnt};
        block_reduce_and_finish<3>(vals, 3, p.partials, p.ticket, p.sums, false);
    }
}

// Replay of caller-supplied Float64 draws through the Float32 step (EXACT order, always).
template <int POT>
__global__ void __launch_bounds__(kBlock) sweep_replay_f32_kernel(const F32Params p)
{
    __shared__ m64::MathTables s_T;
    load_tables(&s_T, p.tables);
    __syncthreads();
    const m64::Tab tb = shared_tab(&s_T);
    constexpr int PF = 4;
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < p.M; c += (int64_t)gridDim.x * kBlock) {
        float x = p.x[c];
        float e = potential_f32<POT>(x);
        uint32_t acc = p.acc[c];
        const float beta = p.betas ? p.betas[c] : p.beta;
        for (int64_t s0 = 0; s0 < p.K; s0 += PF) {
            double zz[PF], ua[PF];
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                if (s0 + i < p.K) {
                    const size_t o = (size_t)(s0 + i) * p.M + c;
                    zz[i] = __ldcs(p.z + o);
                    ua[i] = __ldcs(p.u_acc + o);
                }
            }
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                if (s0 + i < p.K) {
                    const double u = ua[i];
                    const int d = mc_step_f32_exact<POT>(x, e, beta, p.sigma, p.lognorm, __double2float_rn(zz[i]),
                                                         [&](double arg) {
                        float ulo, uhi;
                        m64::ucell_from_double(u, ulo, uhi);
                        return m64::exp_accept(arg, ulo, uhi, [&]() { return u; }, tb);
                    });
                    acc += d;
                    if (p.decisions) __stcs(p.decisions + (size_t)(s0 + i) * p.M + c, (uint8_t)d);
                }
            }
        }
        p.x[c] = x;
        p.acc[c] = acc;
    }
}

// x0 = Float32(4u − 2) with the SAME uniform as the Float64 ensembles (stream tag 0)
__global__ void __launch_bounds__(kBlock) init_kernel_f32(float *x, int64_t M, uint64_t sid0)
{
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < M; c += (int64_t)gridDim.x * kBlock) {
        const U64Pair b = philox_block<kTagInit>(sid0 + (uint64_t)c, 0, 0);
        x[c] = __double2float_rn(__dsub_rn(__dmul_rn(4.0, u53(b.a_lo, b.a_hi)), 2.0));
    }
}

// Conversions between the Float32 state and the Float64 views of the generic entry points, and e = potential(x).
__global__ void __launch_bounds__(kBlock) f32_from_f64_kernel(float *dst, const double *src, int64_t M)
{
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < M; c += (int64_t)gridDim.x * kBlock)
        dst[c] = __double2float_rn(src[c]);
}
__global__ void __launch_bounds__(kBlock) f64_from_f32_kernel(double *dst, const float *src, int64_t M, int pot, int energy)
{
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < M; c += (int64_t)gridDim.x * kBlock) {
        float v = src[c];
        if (energy)
            v = pot == POT_HARMONIC ? potential_f32<POT_HARMONIC>(v)
                                    : pot == POT_QUARTIC ? potential_f32<POT_QUARTIC>(v) : potential_f32<POT_DOUBLE_WELL>(v);
        dst[c] = (double)v;
    }
}
__global__ void __launch_bounds__(kBlock) energy_kernel_f32(const float *x, float *e, int64_t M, int pot)
{
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < M; c += (int64_t)gridDim.x * kBlock) {
        const float v = x[c];
        e[c] = pot == POT_HARMONIC ? potential_f32<POT_HARMONIC>(v)
                                   : pot == POT_QUARTIC ? potential_f32<POT_QUARTIC>(v) : potential_f32<POT_DOUBLE_WELL>(v);
    }
}

// Standalone callback sums of a Float32 ensemble: [Σ Float64(e), Σ acc/tot, count]
__global__ void __launch_bounds__(kBlock) callback_reduce_f32_kernel(const float *x, const uint32_t *acc, int64_t M,
                                                                     int64_t steps_done, int pot, double *partials,
                                                                     unsigned int *ticket, double *sums)
{
    double vals[3] = {0.0, 0.0, 0.0};
    for (int64_t c = (int64_t)blockIdx.x * kBlock + threadIdx.x; c < M; c += (int64_t)gridDim.x * kBlock) {
        const float v = x[c];
        const float e = pot == POT_HARMONIC ? potential_f32<POT_HARMONIC>(v)
                                            : pot == POT_QUARTIC ? potential_f32<POT_QUARTIC>(v) : potential_f32<POT_DOUBLE_WELL>(v);
        vals[0] += (double)e;
        vals[1] += (double)acc[c] / (double)steps_done;
        vals[2] += 1.0;
    }
    block_reduce_and_finish<3>(vals, 3, partials, ticket, sums, false);
}

}  // namespace arianna
