// kernels_replay.cuh -- K6t: the replay sweep fed by bulk asynchronous copies (1-D TMA, cp.async.bulk + mbarrier).
//
// Replay mode consumes caller-supplied draws z[K][M], u_acc[K][M] (and u_cat[K][M] for multi-move pools), step-major,
// always in EXACT arithmetic: 16 B (24 B) of draws per chain-step against ~60 instructions,
// i.e. HBM-bound -- and each step is one long dependent FP64 chain, i.e. latency-bound per warp.  The per-thread `__ldcs`
// loads of sweep_replay_kernel keep at most 64 B per thread in flight; here the draws of a CTA's 512 chains travel as
// [kReplayKT steps][512 chains] tiles: a PRODUCER WARP issues one bulk copy per step row and array (4 KB contiguous in
// the step-major arrays) into a ring of kReplayStages shared-memory stages, each guarded by a `full` mbarrier that
// completes on the byte count (complete_tx) and released through an `empty` mbarrier on which every consumer warp
// arrives (no CTA-wide barrier anywhere in the loop).  Two stages (64 KB per CTA, two CTAs per SM) are always in flight
// while the third is consumed, independent of register pressure and occupancy; the copies carry an L2 evict-first
// policy (the draws are read exactly once).  Every consumer thread runs TWO chains with their steps interleaved, which
// doubles the independent work per warp.  The ring runs across the CTA's chain blocks without draining.  Results are
// bit-identical to sweep_replay_kernel (same mc_step_exact).
//
// Requirements of cp.async.bulk: 16-byte aligned addresses and sizes -> M even and 16-byte aligned arrays; the host
// falls back to sweep_replay_kernel otherwise (arianna_cuda.cu).
#pragma once

namespace arianna {

constexpr int kReplayKT = 4;        // steps per tile
constexpr int kReplayStages = 3;    // ring depth
constexpr int kReplayIlp = 2;       // chains per consumer thread (independent dependency chains, interleaved)
constexpr int kReplayChains = kBlock * kReplayIlp;      // chains per tile row
constexpr int kReplayThreads = kBlock + 32;             // 8 consumer warps + 1 producer warp

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (TMA, 1-D); completion is signalled on `mbar` by byte count
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "l"(policy) : "memory");
}

__host__ __device__ constexpr size_t replay_tma_smem_bytes(int n_moves)
{
    const int na = n_moves > 1 ? 3 : 2;
    return 1024 /* exp table + barriers + pool parameters */ + (size_t)kReplayStages * na * kReplayKT * kReplayChains * 8 +
           (n_moves > 1 ? (size_t)n_moves * 8 * kReplayChains : 0);
}

// N chains of one thread through one EXACT Metropolis step, phase by phase: the arithmetic of every chain is the
// straight-line common case of mc_step_exact (exact_div's fast path, the FP32 accept filter deciding), written so that
// the N dependency chains sit in ONE basic block and the scheduler interleaves them (a single chain leaves the FP64
// pipe waiting on its own previous result: `wait` was the top stall).  A chain whose divisor / dividend is outside
// exact_div's range or whose filter is ambiguous is flagged and redone by the scalar mc_step_exact from its ORIGINAL
// state -- same functions, same operation order, so the results are bit-identical to the scalar kernel.
template <int POT, int N>
__device__ __forceinline__ void mc_step_exact_ilp(double (&x)[N], double (&e)[N], const double (&beta)[N],
                                                  const double (&sigma)[N], const double (&lognorm)[N],
                                                  const double (&inv)[N], const double (&z)[N], const double (&ua)[N],
                                                  int (&dec)[N], m64::Tab tb)
{
    double xn[N], en[N], ndelta[N];
    bool acc[N], slow[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const double delta = __dadd_rn(0.0, __dmul_rn(sigma[q], z[q]));            // particle_1d.jl:57
        const double dd = __dmul_rn(2.0, __dmul_rn(sigma[q], sigma[q]));
        const double n = -__dmul_rn(delta, delta);
        const uint32_t ex = ((uint32_t)__double2hiint(n) >> 20) & 0x7ffu;
        const bool div_ok = inv[q] != 0.0 && ex - 623u <= 800u;                      // m64::exact_div's fast range
        double t1 = __dmul_rn(n, inv[q]);
        t1 = __fma_rn(__fma_rn(-dd, t1, n), inv[q], t1);
        t1 = __fma_rn(__fma_rn(-dd, t1, n), inv[q], t1);                             // == n / dd (Markstein)
        const double lqf = __dsub_rn(t1, lognorm[q]);                                // particle_1d.jl:53
        xn[q] = __dadd_rn(x[q], delta);                                              // :32
        en[q] = potential<POT, ARITH_EXACT>(xn[q]);                                  // :33
        const double dlogp = __dsub_rn(__dmul_rn(-en[q], beta[q]), __dmul_rn(-e[q], beta[q]));   // metropolis.jl:98
        ndelta[q] = -delta;                                                          // particle_1d.jl:38
        const double arg = __dsub_rn(__dadd_rn(dlogp, lqf), lqf);                    // metropolis.jl:183 (lqb == lqf bitwise)
        // the FP32 filter of m64::exp_accept, verbatim
        float ulo, uhi;
        m64::ucell_from_double(ua[q], ulo, uhi);
        const float a = (float)arg;
        const float E = m64::ex2_approx(a * 1.44269504f);
        const float eps = fmaf(fabsf(a), 4.76837158e-07f, 4.76837158e-07f);
        const float Elo = fmaf(-E, eps, E), Ehi = fmaf(E, eps, E);
        acc[q] = (a >= 0.0f) || (Elo >= uhi);
        slow[q] = !div_ok || !(acc[q] || (Ehi < ulo));
    }
#pragma unroll
    for (int q = 0; q < N; ++q) {
        if (slow[q]) {                                       // rare: the scalar step, from the untouched state
            dec[q] = mc_step_exact<POT>(x[q], e[q], beta[q], sigma[q], lognorm[q], inv[q], z[q], ua[q], tb);
        } else {
            const double xr = __dadd_rn(xn[q], ndelta[q]);   // reject: re-applied negated move (metropolis.jl:187)
            const double er = potential<POT, ARITH_EXACT>(xr);
            x[q] = acc[q] ? xn[q] : xr;
            e[q] = acc[q] ? en[q] : er;
            dec[q] = acc[q] ? 1 : 0;
        }
    }
}

template <int POT, bool MULTI>
__global__ void __launch_bounds__(kReplayThreads, 2) sweep_replay_tma_kernel(const ReplayParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NA = MULTI ? 3 : 2;
    constexpr uint32_t kRowBytes = kReplayChains * 8;
    constexpr uint32_t kArrBytes = kReplayKT * kRowBytes;           // one array of one stage
    constexpr uint32_t kStageBytes = NA * kArrBytes;
    // header (1 KB): exp2 table (256 B) | full barriers (32 B) | empty barriers (32 B) | sigma, weight, lognorm, 1/(2σ²) (4 x 128 B)
    constexpr uint32_t kOffFull = 256, kOffEmpty = 288, kOffSigma = 320, kOffWeight = 448, kOffLognorm = 576, kOffInv = 704, kOffStages = 1024;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const int nm = p.pool.n_moves;
    {
        // only the 2^(j/32) table of the math layer is touched on the replay path (exp_accept -> exp_core)
        const double *src = reinterpret_cast<const double *>(reinterpret_cast<const char *>(p.tables) + m64::kOffExp2);
        if (threadIdx.x < m64::kExpTab) reinterpret_cast<double *>(smem_raw)[threadIdx.x] = src[threadIdx.x];
        if (threadIdx.x < kMaxMoves) {
            reinterpret_cast<double *>(smem_raw + kOffSigma)[threadIdx.x] = p.pool.sigma[threadIdx.x];
            reinterpret_cast<double *>(smem_raw + kOffWeight)[threadIdx.x] = p.pool.weight[threadIdx.x];
            reinterpret_cast<double *>(smem_raw + kOffLognorm)[threadIdx.x] = p.pool.lognorm[threadIdx.x];
            reinterpret_cast<double *>(smem_raw + kOffInv)[threadIdx.x] = p.pool.inv2s2[threadIdx.x];
        }
        if (threadIdx.x == 0) {
            for (int s = 0; s < kReplayStages; ++s) {
                mbar_init(base + kOffFull + 8u * s, 1u);                  // the producer's arrive.expect_tx
                mbar_init(base + kOffEmpty + 8u * s, kWarpsPerBlock);     // one arrival per consumer warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const int T = (int)((p.K + kReplayKT - 1) / kReplayKT);           // tiles per chain block
    const int64_t nblocks = (p.M + kReplayChains - 1) / kReplayChains;
    const int64_t my_blocks = (int64_t)blockIdx.x < nblocks ? (nblocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_items = my_blocks * T;                            // (chain block, tile) pairs of this CTA, in order

    if (threadIdx.x >= kBlock) {
        // ---- producer warp: one lane feeds the ring, waiting for a stage to be released before refilling it ----
        if (threadIdx.x == kBlock) {
            const uint64_t policy = l2_evict_first_policy();
            int t = 0;
            int64_t j = 0;
            for (int64_t i = 0; i < n_items; ++i) {
                const uint32_t st = (uint32_t)(i % kReplayStages);
                if (i >= kReplayStages) mbar_wait(base + kOffEmpty + 8u * st, (uint32_t)((i / kReplayStages - 1) & 1));
                const int64_t c0 = ((int64_t)blockIdx.x + j * gridDim.x) * kReplayChains;
                const uint32_t nvalid = (uint32_t)(p.M - c0 < kReplayChains ? p.M - c0 : kReplayChains);
                const int rows = (int)(p.K - (int64_t)t * kReplayKT < kReplayKT ? p.K - (int64_t)t * kReplayKT : kReplayKT);
                const uint32_t bar = base + kOffFull + 8u * st, dst0 = base + kOffStages + st * kStageBytes;
                const uint32_t row_b = nvalid * 8u;
                mbar_expect_tx(bar, (uint32_t)rows * row_b * NA);
                size_t off = (size_t)t * kReplayKT * p.M + c0;
                for (int r = 0; r < rows; ++r, off += p.M) {
                    bulk_g2s(dst0 + r * kRowBytes, p.z + off, row_b, bar, policy);
                    bulk_g2s(dst0 + kArrBytes + r * kRowBytes, p.u_acc + off, row_b, bar, policy);
                    if constexpr (MULTI) bulk_g2s(dst0 + 2 * kArrBytes + r * kRowBytes, p.u_cat + off, row_b, bar, policy);
                }
                if (++t == T) { t = 0; ++j; }
            }
        }
        return;
    }

    // ---- consumer warps: kReplayIlp chains per thread, their serial Metropolis steps interleaved ----
    const m64::Tab tb{base - m64::kOffExp2};                          // table handle: MathTables-relative offsets
    const double *s_sigma = reinterpret_cast<const double *>(smem_raw + kOffSigma);
    const double *s_weight = reinterpret_cast<const double *>(smem_raw + kOffWeight);
    const double *s_lognorm = reinterpret_cast<const double *>(smem_raw + kOffLognorm);
    const double *s_inv = reinterpret_cast<const double *>(smem_raw + kOffInv);
    uint32_t *s_acc = reinterpret_cast<uint32_t *>(smem_raw + kOffStages + kReplayStages * kStageBytes);
    uint32_t *s_tot = s_acc + (MULTI ? nm * kReplayChains : 0);
    const double sigma0 = p.pool.sigma[0], lognorm0 = p.pool.lognorm[0], inv0 = p.pool.inv2s2[0];

    double x[kReplayIlp], e[kReplayIlp], beta[kReplayIlp];
    uint32_t acc[kReplayIlp];
    int64_t c[kReplayIlp];
    bool live[kReplayIlp];
#pragma unroll
    for (int q = 0; q < kReplayIlp; ++q) { x[q] = 0.0; e[q] = 0.0; beta[q] = p.beta; acc[q] = 0; c[q] = 0; live[q] = false; }
    int t = 0;
    int64_t j = 0;
    for (int64_t i = 0; i < n_items; ++i) {
        const uint32_t st = (uint32_t)(i % kReplayStages), par = (uint32_t)((i / kReplayStages) & 1);
        if (t == 0) {                                                  // first tile of a chain block: load the chains
#pragma unroll
            for (int q = 0; q < kReplayIlp; ++q) {
                const int lc = q * kBlock + threadIdx.x;               // column of the tile row
                c[q] = ((int64_t)blockIdx.x + j * gridDim.x) * kReplayChains + lc;
                live[q] = c[q] < p.M;
                if (live[q]) {
                    x[q] = p.x[c[q]];
                    e[q] = potential<POT, ARITH_EXACT>(x[q]);
                    beta[q] = p.betas ? p.betas[c[q]] : p.beta;
                    if constexpr (MULTI) {
                        for (int k = 0; k < nm; ++k) {
                            s_acc[k * kReplayChains + lc] = p.acc[(size_t)k * p.M + c[q]];
                            s_tot[k * kReplayChains + lc] = p.tot[(size_t)k * p.M + c[q]];
                        }
                    } else {
                        acc[q] = p.acc[c[q]];
                    }
                }
            }
        }
        mbar_wait(base + kOffFull + 8u * st, par);                     // the tile's bytes have landed
        const int rows = (int)(p.K - (int64_t)t * kReplayKT < kReplayKT ? p.K - (int64_t)t * kReplayKT : kReplayKT);
        const uint32_t a0 = base + kOffStages + st * kStageBytes + 8u * threadIdx.x;
        const size_t o0 = (size_t)t * kReplayKT * p.M;
#pragma unroll 1
        for (int r = 0; r < rows; ++r) {
            // both chains of the thread run unconditionally (a chain past the end of the ensemble computes on stale
            // shared memory and is never stored): straight-line code that the scheduler interleaves
            int d[kReplayIlp];
            double zz[kReplayIlp], uu[kReplayIlp], sg[kReplayIlp], ln[kReplayIlp], iv[kReplayIlp];
            int kk[kReplayIlp];
#pragma unroll
            for (int q = 0; q < kReplayIlp; ++q) {
                const uint32_t a = a0 + r * kRowBytes + q * (kBlock * 8u);
                zz[q] = lds_f64(a);
                uu[q] = lds_f64(a + kArrBytes);
                if constexpr (MULTI) {
                    kk[q] = categorical(nm, s_weight, lds_f64(a + 2 * kArrBytes));
                    sg[q] = s_sigma[kk[q]]; ln[q] = s_lognorm[kk[q]]; iv[q] = s_inv[kk[q]];
                } else {
                    kk[q] = 0;
                    sg[q] = sigma0; ln[q] = lognorm0; iv[q] = inv0;
                }
            }
            mc_step_exact_ilp<POT, kReplayIlp>(x, e, beta, sg, ln, iv, zz, uu, d, tb);
#pragma unroll
            for (int q = 0; q < kReplayIlp; ++q) {
                if constexpr (MULTI) {
                    const int lc = q * kBlock + threadIdx.x;
                    s_acc[kk[q] * kReplayChains + lc] += d[q];
                    s_tot[kk[q] * kReplayChains + lc] += 1;
                } else {
                    acc[q] += d[q];
                }
            }
            if (p.decisions) {
#pragma unroll
                for (int q = 0; q < kReplayIlp; ++q)
                    if (live[q]) __stcs(p.decisions + o0 + (size_t)r * p.M + c[q], (uint8_t)d[q]);
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(base + kOffEmpty + 8u * st);   // this warp is done with the stage
        if (t == T - 1) {                                              // last tile of the block: write the chains back
#pragma unroll
            for (int q = 0; q < kReplayIlp; ++q) {
                if (live[q]) {
                    p.x[c[q]] = x[q];
                    if constexpr (MULTI) {
                        const int lc = q * kBlock + threadIdx.x;
                        for (int k = 0; k < nm; ++k) {
                            p.acc[(size_t)k * p.M + c[q]] = s_acc[k * kReplayChains + lc];
                            p.tot[(size_t)k * p.M + c[q]] = s_tot[k * kReplayChains + lc];
                        }
                    } else {
                        p.acc[c[q]] = acc[q];
                    }
                }
            }
        }
        if (++t == T) { t = 0; ++j; }
    }
}

}  // namespace arianna
