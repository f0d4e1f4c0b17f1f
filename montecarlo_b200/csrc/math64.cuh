// math64.cuh -- FP64 transcendental kernels of the sweep, written for the B200 FP64 pipe (64 DFMA/clk/SM).
//
// The fused sweep is instruction-issue bound (DESIGN.md "Roofline"): CUDA's libm exp/log/sincospi/sqrt cost ≈150
// FP64 instructions per chain-step.  These replacements are specialised to what the sweep actually needs and take
// their arguments straight from the Philox INTEGER words (no int->fp64 conversion instructions):
//
//   exp_core / exp_accept / exp_nonpos   exp for the accept test / PGMC α      11 FP64, 1 table load, integer range checks
//   exp_accept_prefix<P> the accept test against a P-bit prefix of u          FP32 + MUFU.EX2 + one FFMA.RM; FP64 only when ambiguous
//   neg2log_k52(k)       −2·ln(k·2^-52), k ∈ [1, 2^52)  (Box-Muller radius²)   10 FP64 (incl. the exact int->fp64 DADD), 3 table loads
//   neg2log_u53(n)       −2·ln(n·2^-53), n ∈ [1, 2^53)  (reference form, clz)   9 FP64, 3 table loads
//   sqrt_pos(w)          √w, w > 0 normal                                       6 FP64 + 1 MUFU.RSQ64H
//   sincos_turn53_tab    sin/cos(2π·k·2^-53), k ∈ [0, 2^53)                    12 FP64, one 16-byte table load (1024 directions)
//
// Accuracy target: ≤ 2 ulp (validated against long-double references on the host, tests/test_math64.py, and on
// the device through arianna_debug_math, tests/test_gpu_math.py).  Every function is also compilable for the host
// (ARIANNA_MATH_HOST) so the accuracy tests run without a GPU.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDA_ARCH__) || (defined(__CUDACC__) && !defined(ARIANNA_MATH_HOST))
#define AM_DEV 1
#define AM_FN __device__ __forceinline__
#else
#define AM_DEV 0
#define AM_FN static inline
#include <cmath>
#include <cstring>
#endif

namespace arianna {
namespace m64 {

// ---- bit helpers ---------------------------------------------------------------------------------------
AM_FN double hilo2double(uint32_t hi, uint32_t lo)
{
#if AM_DEV
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    std::memcpy(&d, &b, 8);
    return d;
#endif
}
AM_FN uint32_t double2hi(double d)
{
#if AM_DEV
    return (uint32_t)__double2hiint(d);
#else
    uint64_t b;
    std::memcpy(&b, &d, 8);
    return (uint32_t)(b >> 32);
#endif
}
AM_FN uint32_t double2lo(double d)
{
#if AM_DEV
    return (uint32_t)__double2loint(d);
#else
    uint64_t b;
    std::memcpy(&b, &d, 8);
    return (uint32_t)b;
#endif
}
AM_FN double ll2double_bits(int64_t b)
{
#if AM_DEV
    return __longlong_as_double(b);
#else
    double d;
    std::memcpy(&d, &b, 8);
    return d;
#endif
}
AM_FN int clz64(uint64_t v)
{
#if AM_DEV
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
AM_FN double fma64(double a, double b, double c)
{
#if AM_DEV
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
// MUFU.RSQ64H: ≈22-bit reciprocal square root seed (PTX rsqrt.approx.ftz.f64).  The host emulation degrades a
// correctly rounded value to 22 bits so that the Newton steps are exercised from a realistic seed.
AM_FN double rsqrt_seed(double w)
{
#if AM_DEV
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
    return y;
#else
    double y = 1.0 / std::sqrt(w);
    uint64_t b;
    std::memcpy(&b, &y, 8);
    b &= ~((uint64_t(1) << 30) - 1);  // keep 22 mantissa bits
    std::memcpy(&y, &b, 8);
    return y;
#endif
}

// ---- polynomial coefficients -----------------------------------------------------------------------------
// On the device these live in the constant bank so that DFMA reads them as c[bank][offset] operands: a full 64-bit
// immediate otherwise costs two UMOV/MOV issue slots per use, and the sweep is issue-bound.
#if AM_DEV
#define AM_CONST __constant__
#else
#define AM_CONST static const
#endif
AM_CONST double kExpK[8] = {
    0x1.71547652b82fep+5,   // [0] 32/ln2
    0x1.62e42fefa0000p-6,   // [1] ln2/32, top 37 bits (n·hi exact for |n| < 2^16)
    0x1.cf79abc9e3b3ap-45,  // [2] ln2/32 − hi
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,  // [3..6] expm1 Taylor, |r| ≤ ln2/64
    6755399441055744.0};    // [7] 1.5·2^52
AM_CONST double kLogK[4] = {-2.0 / 7.0, 1.0 / 3.0, -0.4, -2.0 / 3.0};  // −2·log1p(r) = r(−2 + r(1 + k3 r + ½r² + k2 r³ + k1 r⁴ + k0 r⁵))
// High-order coefficients whose terms are below 5·10^-11 of the result only need 20 mantissa bits: as doubles with a
// zero low word they are encoded as 32-bit IMMEDIATES of DFMA/DMUL (no constant load, no register pair).  Truncation
// errors: k2 r⁴·3.6e-7 < 2·10^-17, 1/120·φ⁴·6e-8 < 10^-19, 1/12·φ⁴/2·2.4e-7 < 10^-18 (|r| ≤ 2^-8, |φ| ≤ π/1024).
constexpr double kLogK10t = 0x1.3cf3cp-1;   // 1/3 + 2/7 = 13/21: k0 r + k1 = k0 (m rc) + (k1 − k0), with k0 rc from the table
constexpr double kLogK2t = -0x1.99999p-2;   // −2/5
constexpr double kTrig120t = 0x1.11111p-7;  // 1/120
constexpr double kTrigM12t = -0x1.55555p-4;  // −1/12
AM_CONST double kSinK[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,
                            -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01};
AM_CONST double kCosK[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,
                            2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02};
AM_CONST double kTrigK[3] = {1.0 / 120.0, -1.0 / 6.0, 1.0 / 24.0};
AM_CONST double kTurnK[2] = {0x1.921fb54442d18p-51,  // 2π·2^-53
                             6755399441055744.0};    // 1.5·2^52

// ---- tables (filled on the host once, copied to shared memory by each CTA) ---------------------------------
constexpr int kTrigTab = 1024; // directions of the full circle held as (sin, cos) pairs
constexpr int kLogTab = 128;  // mantissa intervals of [√½, √2)
constexpr int kExpTab = 32;   // 2^(j/32)
constexpr int kETab = 64;     // −2·E·ln2 for E = 0 … −53
constexpr uint32_t kHxBase = 0x3fe6a09eu;  // high word of √½ (fdlibm's log reduction constant)

struct alignas(16) MathTables {
    double sincos[kTrigTab][2]; // (sin, cos)(2π·i/1024), correctly rounded; exact 0 / ±1 on the axes
    // one 32-byte record per mantissa interval i, so that ONE index serves an LDS.128 and an LDS.64:
    //   [0] rc_i = 1 / c_i (c_i: centre of the interval; exactly 1.0 for the interval holding 1)
    //   [1] k0·rc_i, k0 = −2/7: the first Horner step k0 r + k1 of the log1p polynomial becomes ONE fma on m with a
    //       single immediate, fma(m, k0 rc, k1 − k0) (two constants in one DFMA cost two MOVs per pair of steps)
    //   [2] −2·ln(c_i) == +2·ln(rc_i), computed from the ROUNDED rc_i        [3] unused
    double log_rec[kLogTab][4];
    double e_m2ln2[kETab];     // −2·E·ln2 at index n = 53 + E (E = −53 … 0): rising with the exponent field, so the
                               // byte offset is (high word >> 17) & 0x7ff8 plus a constant -- no negation (IADD3) per pair
    double exp2_j[kExpTab];    // 2^(j/32)
};

// Host-side table construction (long double where available).
static inline void build_math_tables(MathTables &T)
{
    for (int i = 0; i < kLogTab; ++i) {
        const uint32_t h0 = kHxBase + ((uint32_t)i << 13), h1 = h0 + (1u << 13);
        uint64_t b0 = (uint64_t)h0 << 32, b1 = (uint64_t)h1 << 32;
        double a, b;
        __builtin_memcpy(&a, &b0, 8);
        __builtin_memcpy(&b, &b1, 8);
        double c = 0.5 * (a + b);
        if (a <= 1.0 && 1.0 < b) c = 1.0;
        const double rc = (double)(1.0L / (long double)c);
        T.log_rec[i][0] = rc;
        T.log_rec[i][1] = (double)((long double)rc * (-2.0L / 7.0L));
        T.log_rec[i][2] = (c == 1.0) ? 0.0 : (double)(2.0L * __builtin_logl((long double)rc));
        T.log_rec[i][3] = 0.0;
    }
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int j = 0; j < kTrigTab / 4; ++j) {
        // first quadrant from the library, the other three by exact rotation (so the axes hold exact 0 / ±1)
        const double sj = j == 0 ? 0.0 : (double)__builtin_sinl(two_pi * j / kTrigTab);
        const double cj = j == 0 ? 1.0 : (double)__builtin_cosl(two_pi * j / kTrigTab);
        const int q = kTrigTab / 4;
        T.sincos[j][0] = sj;          T.sincos[j][1] = cj;
        T.sincos[j + q][0] = cj;      T.sincos[j + q][1] = -sj;
        T.sincos[j + 2 * q][0] = -sj; T.sincos[j + 2 * q][1] = -cj;
        T.sincos[j + 3 * q][0] = -cj; T.sincos[j + 3 * q][1] = sj;
    }
    for (int n = 0; n < kETab; ++n) T.e_m2ln2[n] = n <= 53 ? (double)(2.0L * (53 - n) * 0.693147180559945309417232121458176568L) : 0.0;
    for (int j = 0; j < kExpTab; ++j) T.exp2_j[j] = (double)__builtin_exp2l((long double)j / 32.0L);
}

// ---- table handle ------------------------------------------------------------------------------------------
// On the device a MathTables copy lives in shared memory and is addressed through its 32-bit shared-window address
// with explicit ld.shared: passing C++ pointers makes ptxas re-derive the window base (S2UR SR_CgaCtaId + UMOV + ULEA)
// next to the uses, i.e. inside the issue-bound loops; one pinned register holds it instead (kernels.cuh:shared_tab).
// On the host (accuracy tests) the handle is a plain pointer.
#if AM_DEV
struct Tab { uint32_t s; };
AM_FN double tab_ld(Tab t, uint32_t byte_off)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(t.s + byte_off));
    return v;
}
AM_FN void tab_ld2(Tab t, uint32_t byte_off, double &a, double &b)
{
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(t.s + byte_off));
}
#else
struct Tab { const MathTables *p; };
AM_FN double tab_ld(Tab t, uint32_t byte_off)
{
    double v;
    std::memcpy(&v, reinterpret_cast<const char *>(t.p) + byte_off, 8);
    return v;
}
AM_FN void tab_ld2(Tab t, uint32_t byte_off, double &a, double &b)
{
    a = tab_ld(t, byte_off);
    b = tab_ld(t, byte_off + 8);
}
#endif
constexpr uint32_t kOffSinCos = (uint32_t)offsetof(MathTables, sincos);
constexpr uint32_t kOffLogRec = (uint32_t)offsetof(MathTables, log_rec);
constexpr uint32_t kOffEM2ln2 = (uint32_t)offsetof(MathTables, e_m2ln2);
constexpr uint32_t kOffExp2 = (uint32_t)offsetof(MathTables, exp2_j);

// ---- exp ------------------------------------------------------------------------------------------------
// exp(x) for x ∈ [−708, 0] (also correct for small positive x): n = round(x·32/ln2) = 32m + j,
// r = x − n·ln2/32 (|r| ≤ ln2/64), exp(x) = 2^m · 2^(j/32) · (1 + expm1(r)).  11 FP64 instructions, no range checks:
// outside the interval the result is garbage and the CALLER must not use it (see exp_class / exp_nonpos).
AM_FN double exp_core(double x, Tab tb)
{
    const double nf = fma64(x, kExpK[0], kExpK[7]);
    const int32_t n = (int32_t)double2lo(nf);
    const double nd = nf - kExpK[7];
    double r = fma64(nd, -kExpK[1], x);
    r = fma64(nd, -kExpK[2], r);
    double q = kExpK[3];
    q = fma64(q, r, kExpK[4]);
    q = fma64(q, r, kExpK[5]);
    q = fma64(q, r, kExpK[6]);
    q = fma64(q, r, 0.5);
    const double p = r * fma64(q, r, 1.0);                  // expm1(r) = r·(1 + r·q)
    const double t = tab_ld(tb, kOffExp2 + 8u * (uint32_t)(n & 31));
    const double res = fma64(t, p, t);
    return hilo2double(double2hi(res) + ((uint32_t)(n >> 5) << 20), double2lo(res));
}

// Integer classification of x from its high word (ALU pipe instead of two FP64 compares + selects):
//   kExpOne   x ≥ +0 finite                -> min(1, exp(x)) = 1
//   kExpCore  x ∈ [−708, −0]               -> exp_core(x)
//   kExpZero  x < −708, ±inf, NaN          -> 0 (underflow; NaN/inf states reject like min(1, NaN) > u does)
enum { kExpZero = 0, kExpCore = 1, kExpOne = 2 };
AM_FN int exp_class(double x)
{
    const uint32_t t = double2hi(x) - 0x7ff00000u;
    if (t >= 0x80100000u) return kExpOne;
    return (t - 0x00100000u) < (0x40962000u - 0x00100000u) ? kExpCore : kExpZero;
}

// min(1, exp(x)) as a value (PGMC's α).
AM_FN double exp_nonpos(double x, Tab tb)
{
    const int c = exp_class(x);
    const double v = exp_core(x, tb);
    return c == kExpCore ? v : (c == kExpOne ? 1.0 : 0.0);
}

// u53 of a raw 64-bit word given as two 32-bit halves: (w >> 11)·2^-53 ∈ [0,1), exactly, without an int->fp64
// conversion instruction (two exact DADDs on bit-assembled doubles).
AM_FN double u53_words(uint32_t lo, uint32_t hi)
{
    const double dh = hilo2double(0x41E00000u, hi >> 11);                       // 2^31 + k_hi·2^-21
    const double dl = hilo2double(0x3FE00000u, (hi << 21) | (lo >> 11));         // 2^-1 + k_lo·2^-53
    return (dh - 2147483648.5) + dl;
}

AM_FN float ex2_approx(float y)
{
#if AM_DEV
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y));
    return r;
#else
    return exp2f(y);
#endif
}
AM_FN float uint_as_float(uint32_t b)
{
#if AM_DEV
    return __uint_as_float(b);
#else
    float f;
    std::memcpy(&f, &b, 4);
    return f;
#endif
}

// The Metropolis accept test  min(1, exp(x)) > u,  u ∈ [0, 1), decided EXACTLY as the FP64 evaluation would, but
// through an FP32 filter:  E ≈ exp(x) from MUFU.EX2 with a rigorous relative error bound ε, and u known only through
// a coarse float cell [u_lo, u_hi] (its leading bits).  E(1−ε) ≥ u_hi accepts, E(1+ε) < u_lo rejects; only the
// ambiguous sliver evaluates exp_core and asks `exact_u()` for the full 53-bit uniform on the FP64 pipe.
// Why: on B200 the FP64 pipe is shared with IMAD.WIDE/IMAD.HI (Philox) and is THE bound of the sweep
// (profiles/microbench/pipes.cu); the FP32, XU (F2F, MUFU) and ALU pipes run beside it for free.  Letting the
// filter see only the leading bits of u is what allows the native stream to spend ONE Philox block per pair of
// steps (lazy uniform refinement, DESIGN.md "RNG stream layout").
// Error budget: a = RN32(x) (2^-24), y = a·log2e (2·2^-24), EX2 (2^-22) -> |Ê/E − 1| ≤ 2^-22 + 3·2^-24·|x|;
// ε = 2^-21·(1 + |a|) over-covers it by > 2.5x.
template <class ExactU>
AM_FN bool exp_accept(double x, float ulo, float uhi, ExactU exact_u, Tab tb)
{
    const float a = (float)x;
    const float E = ex2_approx(a * 1.44269504f);
    const float eps = fmaf(fabsf(a), 4.76837158e-07f, 4.76837158e-07f);      // 2^-21·(|a| + 1)
    const float Elo = fmaf(-E, eps, E), Ehi = fmaf(E, eps, E);
    // x ≥ 0 (incl. −0 and +inf): min(1, exp(x)) = 1 > u.  Otherwise u ∈ [ulo, uhi] and the ε bound is strict.
    bool acc = (a >= 0.0f) || (Elo >= uhi);
    // strict: when E underflowed to 0 in FP32 the relative bound is void, but then α < 2^-125 < any non-zero u_lo;
    // with u_lo == 0 the comparison is false and the exact path decides (α > 0 = u accepts).
    const bool rej = Ehi < ulo;
    if (!(acc || rej)) {
        // rare: full FP64 decision.  NaN lands here too (every FP32 comparison is false) and rejects via `core`.
        const uint32_t t = double2hi(x) - 0x7ff00000u;
        const bool core = (t - 0x00100000u) < (0x40962000u - 0x00100000u);  // x ∈ [−708, −0]
        const bool tiny_pos = t >= 0x80100000u;                             // 0 ≤ x, finite (float(x) rounded to −0… never)
        acc = tiny_pos || (core && (exp_core(x, tb) > exact_u()));
    }
    return acc;
}

// The same decision for the native stream, where u is known through its P-bit PREFIX f (u ∈ [f, f+1)·2^-P): the
// comparison needs no float cell assembled from f (5 ALU instructions per step) and its error bound is a constant:
//   Es ≈ 2^P·exp(x) from MUFU.EX2(a·log2e + P), a = RN32(x).  Valid for −127 ≤ a < 0:
//   |Es/(2^P e^x) − 1| ≤ 2^-24|a| (a = RN32(x)) + 2^-25|a| (log2e) + 2^-17·ln2 (rounding of the fma at |y'| < 256)
//   + 2^-22 (EX2) =: δ < 2^-15.6.
// ONE fused multiply-add rounded toward −∞ does the scaling by c = 1 − 2^-13, the subtraction of f AND the floor:
//   v = fma_rd(Es, c, −(1.5·2^23 + f)) = floor(Es·c) − f − 1.5·2^23   exactly
// because the exact value lies in (−2^24, −2^23), where binary32 has ulp 1, and the addend is an integer; the addend
// is assembled by bit arithmetic (bits(−1.5·2^23) + f: the mantissa of a float of that binade counts in units of 1).
// With g = floor(Es·c) − f (an integer, read off v):
//   g ≥  1  accepts:  Es·c ≥ f+1  ⇒  2^P e^x ≥ Es(1−δ) ≥ (f+1)(1−δ)/(1−2^-13) > f+1  ⇒  exp(x) > (f+1)2^-P > u;
//   g ≤ −2  rejects:  Es·c < f−1  ⇒  2^P e^x ≤ Es(1+δ) < (f−1)(1+δ)/(1−2^-13) < f−1 + 2^P·1.17·2^-13 ≤ f−0.41 (P ≤ 12)
//                     ⇒  exp(x) < f 2^-P ≤ u;
//   g ∈ {−1, 0} is decided by the FP64 exp against the exact uniform (lazy refinement).
// a < −127: Es = 0 (the EX2 argument is below −172, flushed), g = −f rejects every f ≥ 2 — correct because
// exp(x) < 2^-183 < 2^-P ≤ u — and sends f ≤ 1 to the exact path.  a ≥ 0 (incl. −0, +inf) accepts: α = 1 > u.
// NaN: v is NaN, both comparisons are false, the exact path rejects -- min(1, NaN) > u is false in the reference too.
// Three instructions (FFMA.RM + two FSETP) replace two FMUL, two F2I and two ISETP of the two-sided integer form.
constexpr float kFloorMagic = 12582912.0f;                  // 1.5·2^23
AM_FN float fma_floor_offset(float Es, uint32_t f)
{
    const float m = uint_as_float(0xCB400000u + f);         // −(1.5·2^23 + f), f < 2^12
#if AM_DEV
    return __fmaf_rd(Es, 0.9998779296875f, m);              // c = 1 − 2^-13
#else
    if (Es != Es) return Es;
    if (Es > 3.0e38f) return Es;                            // +inf stays +inf
    return (float)(floor((double)Es * (double)0.9998779296875f) + (double)m);     // exact: integer-valued, |·| < 2^24
#endif
}
// PBITS = length of the prefix (11 for the odd step of a pair, 12 for the even one: DESIGN.md "RNG stream layout");
// the bound above is for PBITS + log2e·127 < 256 and 2^PBITS·1.17·2^-13 < 1, i.e. PBITS ≤ 12.
// MAGIC = false: `fm` = f and `x` = the argument of exp.
// MAGIC = true (the headline sweep): `fm` = the pre-assembled bits 0xCB400000 | f of the addend (built by the one LOP3
// that extracts the prefix, exp_prefix_bits below) and `x` = the argument ALREADY IN BINARY-LOG UNITS, y = x·log2(e)
// (the caller multiplies by β·log2e instead of β): the scaling FFMA becomes an FADD and its constant leaves the loop.
// The bound above holds a fortiori (a = RN32(y): 2^-24|y|·ln2 = 2^-24|x|; no log2e rounding term); the exact path
// recovers x = y·ln2 (one more rounding of the argument, 2^-53 relative: the FP64 decision is as sharp as before).
constexpr uint32_t kFloorMagicBits = 0xCB400000u;           // bits(−1.5·2^23)
template <int PBITS, bool MAGIC = false, class ExactU>
AM_FN bool exp_accept_prefix(double x, uint32_t fm, ExactU exact_u, Tab tb)
{
    static_assert(PBITS <= 12, "the one-sided floor form needs 2^PBITS * 1.17 * 2^-13 < 1");
    const float a = (float)x;
    const float Es = MAGIC ? ex2_approx(a + (float)PBITS) : ex2_approx(fmaf(a, 1.44269504f, (float)PBITS));
    float v;
    if constexpr (MAGIC) {
#if AM_DEV
        v = __fmaf_rd(Es, 0.9998779296875f, uint_as_float(fm));
#else
        v = fma_floor_offset(Es, fm & 0xfffu);
#endif
    } else {
        v = fma_floor_offset(Es, fm);
    }
    // MAGIC: no test of a ≥ 0 (one FSETP per step).  For y ≥ 0, Es ≥ 2^P(1 − 2^-22) gives floor(Es·c) ≥ 2^P − 1 ≥ f,
    // i.e. g ≥ 0: accepted here unless f = 2^P − 1, which the exact path accepts (x ≥ 0: tiny_pos, or exp(−0) = 1 > u).
    bool acc = MAGIC ? (v > -kFloorMagic) : ((a >= 0.0f) || (v > -kFloorMagic));            // g ≥ 1
    const bool rej = v < -(kFloorMagic + 1.0f);              // g ≤ −2
    if (!(acc || rej)) {
        if constexpr (MAGIC) x = x * 0x1.62e42fefa39efp-1;   // y·ln2: only this rare exact path needs the argument itself
        const uint32_t t = double2hi(x) - 0x7ff00000u;
        const bool core = (t - 0x00100000u) < (0x40962000u - 0x00100000u);  // x ∈ [−708, −0]
        const bool tiny_pos = t >= 0x80100000u;                             // 0 ≤ x, finite
        acc = tiny_pos || (core && (exp_core(x, tb) > exact_u()));
    }
    return acc;
}

// (w & (2^PBITS − 1)) | magic in ONE LOP3: `magic` must be a REGISTER holding kFloorMagicBits (an immediate would make
// it two instructions: a LOP3 has one immediate slot).
template <int PBITS>
AM_FN uint32_t exp_prefix_bits(uint32_t w, uint32_t magic)
{
#if AM_DEV
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(w), "n"((1u << PBITS) - 1u), "r"(magic));   // (a & b) | c
    return r;
#else
    return (w & ((1u << PBITS) - 1u)) | magic;
#endif
}

// Filter cell from the top 23 bits of a raw 64-bit word whose u is (w >> 11)·2^-53 (XOSHIRO mode).
AM_FN float ulo_from_word23(uint32_t w_hi) { return uint_as_float(0x3f800000u | (w_hi >> 9)) - 1.0f; }
// Exact native-mode uniform: u = ((f << (53 − PBITS)) | r)·2^-53 with r = word >> (11 + PBITS), the 53 − PBITS
// refinement bits (PBITS = 11: 42 bits, PBITS = 12: 41 bits).
template <int PBITS>
AM_FN double u53_prefix_refine(uint32_t f, uint32_t r_lo, uint32_t r_hi)
{
    // k = (f << (53 − PBITS)) | (word >> (11 + PBITS)): k_hi = top 21 bits, k_lo = low 32 bits
    const uint32_t k_hi = (f << (21 - PBITS)) | (r_hi >> (11 + PBITS));
    const uint32_t k_lo = (r_hi << (21 - PBITS)) | (r_lo >> (11 + PBITS));
    const double dh = hilo2double(0x41E00000u, k_hi);
    const double dl = hilo2double(0x3FE00000u, k_lo);
    return (dh - 2147483648.5) + dl;
}

// Float cell of an arbitrary double u ∈ [0,1) (replay / EXACT paths): directed roundings on the XU pipe.
AM_FN void ucell_from_double(double u, float &ulo, float &uhi)
{
#if AM_DEV
    ulo = __double2float_rd(u);
    uhi = __double2float_ru(u);
#else
    ulo = (float)u;
    if ((double)ulo > u) ulo = nextafterf(ulo, -1.0f);
    uhi = (float)u;
    if ((double)uhi < u) uhi = nextafterf(uhi, 2.0f);
#endif
}

// Reference decision (no filter) -- used by the accuracy tests to prove the filter never changes a decision.
AM_FN bool exp_accept_ref(double x, double u, Tab tb)
{
    const int c = exp_class(x);
    return c == kExpOne || (c == kExpCore && exp_core(x, tb) > u);
}

// ---- exact division by a per-move constant --------------------------------------------------------------------
// n / d, correctly rounded, for a divisor whose correctly rounded reciprocal y = RN(1/d) is known (d = 2σ² of
// log_proposal_density, particle_1d.jl:53).  q0 = RN(n·y) is within 2 ulp of n/d; one residual step makes it faithful
// (< 1 ulp), and by Markstein's theorem a second residual step from a faithful quotient with y = RN(1/d) returns
// RN(n/d) exactly -- as long as nothing over/underflows on the way: |n| ∈ [2^-400, 2^400] (checked here on the
// exponent field) and d ∈ [2^-300, 2^300] (checked by the host, which passes y = 0 otherwise).  5 FP64 instructions
// and no branch in the common case instead of the ~20 + slow-path call of an IEEE division; bit-identical results
// (tests/test_math64.py::test_exact_div; the replay tests compare every decision and position with the oracle's `/`).
AM_FN double exact_div(double n, double d, double y)
{
    const uint32_t ex = (double2hi(n) >> 20) & 0x7ffu;
    if (y != 0.0 && ex - 623u <= 800u) {
#if AM_DEV
        double q = __dmul_rn(n, y);
#else
        double q = n * y;
#endif
        q = fma64(fma64(-d, q, n), y, q);
        return fma64(fma64(-d, q, n), y, q);
    }
#if AM_DEV
    return __ddiv_rn(n, d);
#else
    return n / d;
#endif
}

// ---- −2·ln(n·2^-53) --------------------------------------------------------------------------------------
// Core: u = m·2^E with m ∈ [√½, √2) given by its words (hx, lx) after fdlibm's fold; i = mantissa interval.
AM_FN double neg2log_core(uint32_t hx_folded, uint32_t lx, uint32_t n8, Tab tb)     // n8 = 8·(53 + E)
{
    const double m = hilo2double(hx_folded, lx);
    const uint32_t i32 = ((hx_folded - kHxBase) >> 13) * 32u;   // byte offset of the record of mantissa interval i
    double rc, rck0;
    tab_ld2(tb, kOffLogRec + i32, rc, rck0);
    const double r = fma64(m, rc, -1.0);    // |r| ≤ 2^-8
    // −2·log1p(r) = r·(−2 + r·(1 − (2/3)r + (1/2)r² − (2/5)r³ + (1/3)r⁴ − (2/7)r⁵))
    double q = fma64(m, rck0, kLogK10t);    // = k0 r + k1
    q = fma64(q, r, kLogK2t);
    q = fma64(q, r, 0.5);
    q = fma64(q, r, kLogK[3]);
    q = fma64(q, r, 1.0);
    const double t = r * fma64(q, r, -2.0);
    return (tab_ld(tb, kOffEM2ln2 + n8) + tab_ld(tb, kOffLogRec + 16u + i32)) + t;
}

// From the integer n ∈ [1, 2^53) (clz normalisation; reference formulation used by the accuracy tests).
AM_FN double neg2log_u53(uint64_t n, Tab tb)
{
    const int lz = clz64(n);                // lz ∈ [11, 63]
    const uint64_t nm = n << lz;            // bit 63 set; at most 53 significant bits
    int E = 10 - lz;                        // n·2^-53 = (nm/2^63)·2^E
    uint32_t hx = 0x3ff00000u | (uint32_t)((nm >> 43) & 0xfffffu);
    const uint32_t lx = (uint32_t)(nm >> 11);
    hx += 0x3ff00000u - kHxBase;            // fdlibm: fold the mantissa's top bit into the exponent
    E += (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + kHxBase;
    return neg2log_core(hx, lx, 8u * (uint32_t)(53 + E), tb);
}

// Same value, from the two 32-bit halves of k = n (k_hi: top 21 bits, k_lo: low 32 bits): u = k·2^-53 is first
// assembled EXACTLY with two DADDs (no clz / 64-bit normalising shifts: the FP64 adder does the normalisation and
// 12 ALU-pipe instructions disappear from the hot loop), then split into exponent and mantissa words.
AM_FN double neg2log_words(uint32_t k_hi, uint32_t k_lo, Tab tb)
{
    const double dh = hilo2double(0x41E00000u, k_hi);   // 2^31 + k_hi·2^-21
    const double dl = hilo2double(0x3FE00000u, k_lo);   // 2^-1 + k_lo·2^-53
    const double u = (dh - 2147483648.5) + dl;          // exact
    const uint32_t hx = double2hi(u) + (0x3ff00000u - kHxBase);
    const uint32_t n8 = ((hx >> 17) & 0x7ff8u) - 8u * 0x3cau;    // 8·(53 + E), E = (hx >> 20) − 0x3ff ∈ [−53, 0]
    return neg2log_core((hx & 0x000fffffu) + kHxBase, double2lo(u), n8, tb);
}

// −2·ln(k·2^-52) for a 52-bit k given as (k_hi: top 20 bits, k_lo: low 32 bits), k ≥ 1: the Box-Muller radius of the
// native stream.  k fits the mantissa field of one double: bits(0x433 | k) = 2^52 + k, so ONE exact DADD with an
// immediate recovers k as a normalised double (the 53-bit form above needs two DADDs and three materialised
// constants); the 2^-52 is folded into the exponent bookkeeping.
AM_FN double neg2log_k52(uint32_t k_hi, uint32_t k_lo, Tab tb)
{
    const double kd = hilo2double(0x43300000u | k_hi, k_lo) - 4503599627370496.0;   // exact
    const uint32_t hx = double2hi(kd) + (0x3ff00000u - kHxBase);
    const uint32_t n8 = ((hx >> 17) & 0x7ff8u) - 8u * 0x3feu;    // 8·(53 + E), E = (hx >> 20) − (0x3ff + 52) ∈ [−52, 0]
    return neg2log_core((hx & 0x000fffffu) + kHxBase, double2lo(kd), n8, tb);
}

// ---- √w, w > 0 normal ---------------------------------------------------------------------------------------
AM_FN double sqrt_pos(double w)
{
    // y = (1/√w)(1+ε), |ε| ≲ 2^-22.  One coupled Newton step: g = √w(1 − 1.5ε² + …) (≈ 2^-43 relative); the final
    // correction g + (w − g²)·h only needs h = 1/(2√w) to the seed's 22 bits: the result is √w(1 + O(ε³)) before the
    // last rounding, so h is NOT refined (6 FP64 instructions instead of 7; ≤ 1 ulp in tests/test_math64.py).
    const double y = rsqrt_seed(w);
    double g = w * y;
    const double h = 0.5 * y;
    const double r = fma64(-h, g, 0.5);
    g = fma64(g, r, g);
    const double d = fma64(-g, g, w);
    return fma64(d, h, g);
}

// ---- sin/cos(2π·k·2^-53) ------------------------------------------------------------------------------------
// Quadrant reduction in the integer domain: q = round(4t) mod 4, φ = 2π(t − q/4) ∈ [−π/4, π/4); fdlibm kernels.
AM_FN void sincos_turn53(uint64_t k, double &sn, double &cs)
{
    const uint64_t kk = k + (uint64_t(1) << 50);
    const uint32_t q = (uint32_t)(kk >> 51) & 3u;
    const int64_t rem = (int64_t)(kk & ((uint64_t(1) << 51) - 1)) - (int64_t(1) << 50);
    const double d = ll2double_bits(0x4338000000000000LL + rem) - kTurnK[1];  // exact int -> fp64
    const double x = d * kTurnK[0];
    const double z = x * x;
    double ps = kSinK[0];
    ps = fma64(ps, z, kSinK[1]);
    ps = fma64(ps, z, kSinK[2]);
    ps = fma64(ps, z, kSinK[3]);
    ps = fma64(ps, z, kSinK[4]);
    ps = fma64(ps, z, kSinK[5]);
    const double s = fma64(z * x, ps, x);
    double pc = kCosK[0];
    pc = fma64(pc, z, kCosK[1]);
    pc = fma64(pc, z, kCosK[2]);
    pc = fma64(pc, z, kCosK[3]);
    pc = fma64(pc, z, kCosK[4]);
    pc = fma64(pc, z, kCosK[5]);
    const double c = fma64(z * z, pc, fma64(-0.5, z, 1.0));
    // rotate by q quarter turns: (cos, sin)(φ + qπ/2)
    const bool swap = q & 1u;
    double cc = swap ? s : c, ss = swap ? c : s;
    const uint32_t neg_c = ((q + 1u) >> 1) & 1u;  // q = 1, 2
    const uint32_t neg_s = q >> 1;                // q = 2, 3
    cs = hilo2double(double2hi(cc) ^ (neg_c << 31), double2lo(cc));
    sn = hilo2double(double2hi(ss) ^ (neg_s << 31), double2lo(ss));
}

// Table form used by the sweep: the circle is cut into 1024 directions a_i = 2π·i/1024 held as correctly rounded
// (sin, cos) pairs in shared memory; k = i·2^43 + rem with |rem| ≤ 2^42, φ = 2π·rem·2^-53, |φ| ≤ π/1024, so
//   sin φ = φ + φ³(−1/6 + φ²/120)              (truncation φ⁶/5040 ≈ 2·10^-19 relative)
//   cos φ − 1 = φ²(−1/2 + φ²/24)               (truncation φ⁶/720 ≈ 10^-18)
//   sin(a+φ) = fma(C, sin φ, fma(S, cos φ − 1, S)),  cos(a+φ) = fma(−S, sin φ, fma(C, cos φ − 1, C)).
// 12 FP64 instructions + one 16-byte shared load instead of 18 FP64 + ~20 integer/select instructions of the
// quadrant-reduced polynomial form above (no quadrant swap, no sign fix-up, four polynomial constants instead of
// twelve).  Error ≤ 0.5 ulp (table) + 0.5 ulp (inner fma) + 0.5 ulp (outer fma); exact on the axes (S or C = 0, ±1).
AM_FN void sincos_turn53_tab(uint32_t k_hi, uint32_t k_lo, Tab tb, double &sn, double &cs)
{
    const uint32_t kk = k_hi + (1u << 10);                 // k + 2^42: round to the nearest direction
    const uint32_t i = (kk >> 11) & (uint32_t)(kTrigTab - 1);
    // rem = ((kk & 0x7ff) − 0x400)·2^32 + k_lo, converted exactly: bits(1.5·2^52) + rem, minus 1.5·2^52
    const double d = hilo2double(0x43380000u - 0x400u + (kk & 0x7ffu), k_lo) - 6755399441055744.0;   // 1.5·2^52
    const double phi = d * kTurnK[0];
    const double z = phi * phi;
    const double ps = fma64(z, kTrig120t, kTrigK[1]);      // 1/120, −1/6
    const double sphi = fma64(z * phi, ps, phi);
    const double hz = z * -0.5;
    const double cm1 = fma64(hz, z * kTrigM12t, hz);       // cos φ − 1 = −z/2·(1 − z/12): one immediate per instruction
    double S, C;
    tab_ld2(tb, kOffSinCos + 16u * i, S, C);
    sn = fma64(C, sphi, fma64(S, cm1, S));
    cs = fma64(-S, sphi, fma64(C, cm1, C));
}

// Box-Muller from two raw 64-bit Philox words (B0 -> radius, B1 -> angle); same definition as the oracle:
//   u1 = ((B0 >> 12) | 1)·2^-52 ∈ (0,1) (odd lattice: never 0 or 1, so −2 ln u1 > 0 without a special case),
//   u2 = (B1 >> 11)·2^-53, z0 = √(−2 ln u1)·cos(2π u2), z1 = …·sin(2π u2).
AM_FN void box_muller_u64(uint64_t B0, uint64_t B1, Tab tb, double &z0, double &z1)
{
    const uint32_t a_lo = (uint32_t)B0, a_hi = (uint32_t)(B0 >> 32);
    const double w = neg2log_k52(a_hi >> 12, ((a_hi << 20) | (a_lo >> 12)) | 1u, tb);
    const double r = sqrt_pos(w);
    double s, c;
    const uint32_t b_lo = (uint32_t)B1, b_hi = (uint32_t)(B1 >> 32);
    sincos_turn53_tab(b_hi >> 11, (b_hi << 21) | (b_lo >> 11), tb, s, c);
    z0 = r * c;
    z1 = r * s;
}

}  // namespace m64
}  // namespace arianna
