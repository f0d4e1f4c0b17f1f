// pipes.cu -- which B200 (sm_100a) pipes overlap?  Measures cycles per warp-instruction per SMSP for pure and
// mixed instruction streams (independent dependency chains, 8 warps/SMSP, no memory traffic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(double *out, uint32_t seed, double a, double b)
{
    double d0 = threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
    uint32_t i0 = threadIdx.x + seed, i1 = i0 * 3u, i2 = i0 * 5u, i3 = i0 * 7u;
    uint32_t l0 = i0 ^ 0x1234, l1 = i1 ^ 0x777, l2 = i2, l3 = i3;
    uint64_t w0 = i0, w1 = i1 + 11, w2 = i2 + 13, w3 = i3 + 17;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE & 1) {  // DFMA x4
                d0 = fma(d0, a, b); d1 = fma(d1, a, b); d2 = fma(d2, a, b); d3 = fma(d3, a, b);
            }
            if (MODE & 2) {  // IMAD.WIDE-like: mul.hi + mul.lo (Philox round multiply) x4 -> compiler emits IMAD.WIDE.U32
                uint64_t p0 = (uint64_t)i0 * 0xD2511F53u, p1 = (uint64_t)i1 * 0xCD9E8D57u;
                uint64_t p2 = (uint64_t)i2 * 0xD2511F53u, p3 = (uint64_t)i3 * 0xCD9E8D57u;
                i0 = (uint32_t)(p0 >> 32) ^ (uint32_t)p1; i1 = (uint32_t)(p1 >> 32) ^ (uint32_t)p0;
                i2 = (uint32_t)(p2 >> 32) ^ (uint32_t)p3; i3 = (uint32_t)(p3 >> 32) ^ (uint32_t)p2;
            }
            if (MODE & 4) {  // LOP3 x4 (3-input xor chains)
                l0 = l0 ^ l1 ^ 0x9E3779B9u; l1 = l1 ^ l2 ^ 0xBB67AE85u; l2 = l2 ^ l3 ^ 0x1234567u; l3 = l3 ^ l0 ^ 0x7654321u;
            }
            if (MODE & 8) {  // plain 32-bit IMAD (lo) x4
                i0 = i0 * i0 + 1u; i1 = i1 * i1 + 3u; i2 = i2 * i2 + 5u; i3 = i3 * i3 + 7u;
            }
            if (MODE & 32) {  // F2F.F32.F64 + back (2 conversions per chain) x4
                d0 = (double)((float)d0) + a; d1 = (double)((float)d1) + a; d2 = (double)((float)d2) + a; d3 = (double)((float)d3) + a;
            }
            if (MODE & 64) {  // MUFU.EX2 x4
                float f0 = __uint_as_float(l0), f1 = __uint_as_float(l1), f2 = __uint_as_float(l2), f3 = __uint_as_float(l3);
                asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(f0)); asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(f1));
                asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(f2)); asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(f3));
                l0 = __float_as_uint(f0); l1 = __float_as_uint(f1); l2 = __float_as_uint(f2); l3 = __float_as_uint(f3);
            }
            if (MODE & 128) {  // F2I.U64.F64.CEIL x4 (result folded back with integer ops)
                unsigned long long q0 = __double2ull_ru(d0), q1 = __double2ull_ru(d1), q2 = __double2ull_ru(d2), q3 = __double2ull_ru(d3);
                d0 = __longlong_as_double((__double_as_longlong(d0) ^ (q0 & 1))); d1 = __longlong_as_double((__double_as_longlong(d1) ^ (q1 & 1)));
                d2 = __longlong_as_double((__double_as_longlong(d2) ^ (q2 & 1))); d3 = __longlong_as_double((__double_as_longlong(d3) ^ (q3 & 1)));
            }
            if (MODE & 256) {  // IMAD.HI.U32 x4
                i0 = __umulhi(i0, 0xD2511F53u) + 1u; i1 = __umulhi(i1, 0xCD9E8D57u) + 3u; i2 = __umulhi(i2, 0xD2511F53u) + 5u; i3 = __umulhi(i3, 0xCD9E8D57u) + 7u;
            }
            if (MODE & 512) {  // pure IMAD.WIDE.U32 with 64-bit accumulate: one instruction per chain step
                w0 = (uint64_t)(uint32_t)w0 * 0xD2511F53u + w0; w1 = (uint64_t)(uint32_t)w1 * 0xCD9E8D57u + w1;
                w2 = (uint64_t)(uint32_t)w2 * 0xD2511F53u + w2; w3 = (uint64_t)(uint32_t)w3 * 0xCD9E8D57u + w3;
            }
            if (MODE & 1024) {  // DADD x4
                d0 = d0 + a; d1 = d1 + a; d2 = d2 + a; d3 = d3 + a;
            }
            if (MODE & 16) {  // FFMA x4 (fp32)
                float f0 = __uint_as_float(l0), f1 = __uint_as_float(l1), f2 = __uint_as_float(l2), f3 = __uint_as_float(l3);
                f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f);
                l0 = __float_as_uint(f0); l1 = __float_as_uint(f1); l2 = __float_as_uint(f2); l3 = __float_as_uint(f3);
            }
        }
    }
    out[blockIdx.x * 256 + threadIdx.x] = d0 + d1 + d2 + d3 + (double)(i0 ^ i1 ^ i2 ^ i3 ^ l0 ^ l1 ^ l2 ^ l3) + (double)(w0 ^ w1 ^ w2 ^ w3);
}

template <int MODE>
void run(const char *name, int ninst_per_u, double *out, int sms, double ghz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 4;  // 4 CTAs x 8 warps = 32 warps/SM = 8 warps/SMSP
    k<MODE><<<grid, 256>>>(out, 1, 0.999999, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(out, 1, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst_per_smsp = 8.0 * ITERS * 8.0 * ninst_per_u;  // 8 warps/SMSP
    const double cycles = ms * 1e-3 * ghz * 1e9;
    printf("%-30s %8.3f ms  %7.3f SMSP-cycles per group per warp (group = 4 independent chains x 1 op of each listed kind)\n",
           name, ms, cycles / (8.0 * ITERS * 8.0));
    (void)warp_inst_per_smsp;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double *out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 4 * 256);
    const double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    run<1>("DFMA", 4, out, p.multiProcessorCount, ghz);
    run<2>("IMAD.WIDE(+xor)", 8, out, p.multiProcessorCount, ghz);
    run<4>("LOP3", 4, out, p.multiProcessorCount, ghz);
    run<8>("IMAD lo", 4, out, p.multiProcessorCount, ghz);
    run<16>("FFMA", 4, out, p.multiProcessorCount, ghz);
    run<1 | 2>("DFMA + IMAD.WIDE(+xor)", 12, out, p.multiProcessorCount, ghz);
    run<1 | 4>("DFMA + LOP3", 8, out, p.multiProcessorCount, ghz);
    run<1 | 8>("DFMA + IMAD lo", 8, out, p.multiProcessorCount, ghz);
    run<1 | 16>("DFMA + FFMA", 8, out, p.multiProcessorCount, ghz);
    run<2 | 4>("IMAD.WIDE(+xor) + LOP3", 12, out, p.multiProcessorCount, ghz);
    run<1 | 2 | 4>("DFMA + IMAD.WIDE + LOP3", 16, out, p.multiProcessorCount, ghz);
    run<32>("F2F f64->f32->f64 + DADD", 12, out, p.multiProcessorCount, ghz);
    run<64>("MUFU.EX2", 4, out, p.multiProcessorCount, ghz);
    run<128>("F2I.U64.F64.CEIL (+lop)", 4, out, p.multiProcessorCount, ghz);
    run<256>("IMAD.HI.U32", 4, out, p.multiProcessorCount, ghz);
    run<1 | 64>("DFMA + MUFU.EX2", 8, out, p.multiProcessorCount, ghz);
    run<1 | 128>("DFMA + F2I.CEIL", 8, out, p.multiProcessorCount, ghz);
    run<1 | 256>("DFMA + IMAD.HI", 8, out, p.multiProcessorCount, ghz);
    run<2 | 64>("IMAD.WIDE(+xor) + MUFU.EX2", 12, out, p.multiProcessorCount, ghz);
    run<8 | 2>("IMAD lo + IMAD.WIDE(+xor)", 12, out, p.multiProcessorCount, ghz);
    run<512>("IMAD.WIDE (pure, 64-bit acc)", 4, out, p.multiProcessorCount, ghz);
    run<512 | 4>("IMAD.WIDE pure + LOP3", 8, out, p.multiProcessorCount, ghz);
    run<512 | 16>("IMAD.WIDE pure + FFMA", 8, out, p.multiProcessorCount, ghz);
    run<512 | 1>("IMAD.WIDE pure + DFMA", 8, out, p.multiProcessorCount, ghz);
    run<512 | 8>("IMAD.WIDE pure + IMAD lo", 8, out, p.multiProcessorCount, ghz);
    run<512 | 64>("IMAD.WIDE pure + MUFU.EX2", 8, out, p.multiProcessorCount, ghz);
    run<512 | 4 | 16>("IMAD.WIDE pure + LOP3 + FFMA", 12, out, p.multiProcessorCount, ghz);
    run<1024>("DADD", 4, out, p.multiProcessorCount, ghz);
    run<4 | 16>("LOP3 + FFMA", 8, out, p.multiProcessorCount, ghz);
    run<4 | 16 | 1>("LOP3 + FFMA + DFMA", 12, out, p.multiProcessorCount, ghz);
    return 0;
}
