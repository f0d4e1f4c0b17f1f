/*
 * harmonic_oscillator.c -- the reference's example script
 * example/particle_1d/harmonic_oscillator/MC_harmonic_oscillator.jl (beta = 2, sigma = 0.1, burn 1000, energy and
 * acceptance stored every 10 steps) written against the C ABI alone: no Python, no torch, no Julia.  It is the
 * smallest complete host of libarianna_cuda.so and shows what a binding in any language has to call.
 *
 *   gcc -O2 -I include examples/harmonic_oscillator.c -L montecarlo_b200 -larianna_cuda \
 *       -Wl,-rpath,$PWD/montecarlo_b200 -o harmonic_oscillator
 *   ./harmonic_oscillator [n_chains] [steps]          # prints "t energy [acceptance]" like energy.dat / acceptance.dat
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "arianna_cuda.h"

#define CHECK(h, call)                                                                        \
    do {                                                                                      \
        int32_t rc_ = (call);                                                                 \
        if (rc_ != ARIANNA_OK) {                                                              \
            fprintf(stderr, "%s failed (%d): %s\n", #call, (int)rc_, arianna_last_error(h)); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

int main(int argc, char **argv)
{
    const int64_t M = argc > 1 ? atoll(argv[1]) : 1 << 20;
    const int64_t steps = argc > 2 ? atoll(argv[2]) : 2000, burn = 1000, every = 10;
    if (M < 1 || steps < burn) { fprintf(stderr, "usage: %s [n_chains >= 1] [steps >= %lld]\n", argv[0], (long long)burn); return 1; }

    arianna_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg;
    cfg.device = -1;
    cfg.n_chains = M;
    cfg.seed = 42;
    cfg.beta = 2.0;
    cfg.potential = ARIANNA_POT_HARMONIC;
    cfg.n_moves = 1;                                   /* pool = (Move(Displacement, StandardGaussian, sigma = 0.1, 1.0),) */
    cfg.sigma[0] = 0.1;
    cfg.weight[0] = 1.0;
    cfg.rng_mode = ARIANNA_RNG_PHILOX;
    cfg.arith_mode = ARIANNA_ARITH_FAST;

    arianna_handle *h = NULL;
    CHECK(NULL, arianna_create(&cfg, &h));
    CHECK(h, arianna_init_synthetic(h, cfg.seed));     /* chains = [System(4rand(rng) - 2, beta) for _ in 1:M] */

    /* sampletimes = build_schedule(steps, burn, 10) = burn:10:steps  ->  store intervals K = [burn, 10, 10, ...] */
    const int32_t n = (int32_t)((steps - burn) / every) + 1;
    int64_t *K = malloc(sizeof *K * n);
    double *rec = malloc(sizeof *rec * 3 * n);
    if (!K || !rec) return 3;
    K[0] = burn;
    for (int32_t i = 1; i < n; ++i) K[i] = every;

    /* run!(simulation): Metropolis + StoreCallbacks(callback_energy, callback_acceptance) in ONE call */
    CHECK(h, arianna_sweep_series(h, n, K, rec));

    double ms = 0.0, e_sum = 0.0;
    CHECK(h, arianna_timing(h, &ms, NULL));
    int64_t t = 0;
    for (int32_t i = 0; i < n; ++i) {
        t += K[i];
        const double energy = rec[3 * i] / rec[3 * i + 2], acceptance = rec[3 * i + 1] / rec[3 * i + 2];
        if (i < 3 || i == n - 1) printf("%lld %.16g [%.16g]\n", (long long)t, energy, acceptance);
        else if (i == 3) printf("...\n");
        e_sum += energy;
    }
    printf("# %lld chains x %lld steps in %.3f ms of device time = %.3e chain-steps/s; <E> over the stores = %.6f "
           "(analytic 1/(2 beta) = 0.25)\n", (long long)M, (long long)t, ms, (double)M * (double)t / (ms * 1e-3), e_sum / n);

    free(K);
    free(rec);
    CHECK(h, arianna_destroy(h));
    return 0;
}
