// Host build of montecarlo_b200/csrc/math64.cuh for the CPU accuracy tests (tests/test_math64.py).
// g++ -O2 -std=c++17 -ffp-contract=off -DARIANNA_MATH_HOST -shared -fPIC
#include "../montecarlo_b200/csrc/math64.cuh"

using namespace arianna::m64;
static MathTables T;
static bool ready = false;
static void init() { if (!ready) { build_math_tables(T); ready = true; } }

extern "C" {
void m64_exp(const double *x, double *out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = exp_nonpos(x[i], Tab{&T}); }
void m64_neg2log(const uint64_t *k, double *out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = neg2log_u53(k[i], Tab{&T}); }
void m64_neg2log_words(const uint64_t *k, double *out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = neg2log_words((uint32_t)(k[i] >> 32), (uint32_t)k[i], Tab{&T}); }
void m64_neg2log_k52(const uint64_t *k, double *out, long n) { init(); for (long i = 0; i < n; ++i) out[i] = neg2log_k52((uint32_t)(k[i] >> 32), (uint32_t)k[i], Tab{&T}); }
void m64_sqrt(const double *x, double *out, long n) { for (long i = 0; i < n; ++i) out[i] = sqrt_pos(x[i]); }
void m64_sincos(const uint64_t *k, double *s, double *c, long n) { init(); for (long i = 0; i < n; ++i) sincos_turn53_tab((uint32_t)(k[i] >> 32), (uint32_t)k[i], Tab{&T}, s[i], c[i]); }
void m64_sincos_poly(const uint64_t *k, double *s, double *c, long n) { for (long i = 0; i < n; ++i) sincos_turn53(k[i], s[i], c[i]); }
void m64_box_muller(const uint64_t *b0, const uint64_t *b1, double *z0, double *z1, long n) { init(); for (long i = 0; i < n; ++i) box_muller_u64(b0[i], b1[i], Tab{&T}, z0[i], z1[i]); }
// mode 0: XOSHIRO-style word (23-bit cell, u = (w >> 11) 2^-53); mode 1 / 3: native (11- / 12-bit prefix = the low
// bits of w, refinement word r given separately); mode 2: directed-rounding float cell of an arbitrary double
void m64_accept(const double *x, const uint64_t *w, const uint64_t *r, int mode, unsigned char *filt,
                unsigned char *ref, double *u_out, long n)
{
    init();
    for (long i = 0; i < n; ++i) {
        const uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
        if (mode == 0 || mode == 2) {
            const double u = u53_words(lo, hi);
            if (mode == 0) filt[i] = exp_accept(x[i], ulo_from_word23(hi), ulo_from_word23(hi) + 1.1920929e-07f, [&] { return u; }, Tab{&T});
            else { float a, b; ucell_from_double(u, a, b); filt[i] = exp_accept(x[i], a, b, [&] { return u; }, Tab{&T}); }
            ref[i] = exp_accept_ref(x[i], u, Tab{&T});
            u_out[i] = u;
        } else if (mode == 1) {
            const uint32_t f = lo & 0x7ffu;
            const double u = u53_prefix_refine<11>(f, (uint32_t)r[i], (uint32_t)(r[i] >> 32));
            filt[i] = exp_accept_prefix<11>(x[i], f, [&] { return u; }, Tab{&T});
            ref[i] = exp_accept_ref(x[i], u, Tab{&T});
            u_out[i] = u;
        } else if (mode == 3) {
            const uint32_t f = lo & 0xfffu;
            const double u = u53_prefix_refine<12>(f, (uint32_t)r[i], (uint32_t)(r[i] >> 32));
            filt[i] = exp_accept_prefix<12>(x[i], f, [&] { return u; }, Tab{&T});
            ref[i] = exp_accept_ref(x[i], u, Tab{&T});
            u_out[i] = u;
        } else {
            // modes 4 / 5: the headline sweep's form -- x[i] is the argument in binary-log units (y = x log2e), the prefix
            // arrives as the filter's addend bits; reference = the FP64 decision on RN(y ln2)
            const double xr = x[i] * 0x1.62e42fefa39efp-1;
            if (mode == 4) {
                const uint32_t fm = exp_prefix_bits<11>(lo, kFloorMagicBits);
                const double u = u53_prefix_refine<11>(fm & 0x7ffu, (uint32_t)r[i], (uint32_t)(r[i] >> 32));
                filt[i] = exp_accept_prefix<11, true>(x[i], fm, [&] { return u; }, Tab{&T});
                ref[i] = exp_accept_ref(xr, u, Tab{&T});
                u_out[i] = u;
            } else {
                const uint32_t fm = exp_prefix_bits<12>(lo, kFloorMagicBits);
                const double u = u53_prefix_refine<12>(fm & 0xfffu, (uint32_t)r[i], (uint32_t)(r[i] >> 32));
                filt[i] = exp_accept_prefix<12, true>(x[i], fm, [&] { return u; }, Tab{&T});
                ref[i] = exp_accept_ref(xr, u, Tab{&T});
                u_out[i] = u;
            }
        }
    }
}
void m64_exact_div(const double *n, const double *d, double *out, long cnt)
{
    for (long i = 0; i < cnt; ++i) {
        const double y = (d[i] >= 0x1p-300 && d[i] <= 0x1p300) ? 1.0 / d[i] : 0.0;    // host_inv2s2 of arianna_cuda.cu
        out[i] = exact_div(n[i], d[i], y);
    }
}
void m64_u53(const uint64_t *w, double *out, long n) { for (long i = 0; i < n; ++i) out[i] = u53_words((uint32_t)w[i], (uint32_t)(w[i] >> 32)); }
void m64_tables(double *out) { init(); __builtin_memcpy(out, &T, sizeof T); }
}
