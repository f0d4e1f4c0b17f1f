"""gen_mix.py -- writes mix.cu: a dependency-free replay of the instruction mix of ONE Box-Muller pair of the headline
kernel (sweep_philox_kernel<HARMONIC, FAST, series>), opcode counts taken from the ncu source page of the r02 capture
(profiles/r02_sweep_ncu_summary.md: 141 warp instructions per pair in the flat loop).

Every instruction class runs on its own set of independent accumulators (no instruction waits for another class, 4-8
independent chains inside a class), the classes are interleaved evenly, and the kernel runs at the sweep's own occupancy
(256 threads x 4 CTAs per SM, <= 64 registers).  What it measures: how many SMSP cycles the B200 needs to ISSUE this mix
when nothing but the issue / dispatch ports limits it -- the floor the real kernel (218 cycles per pair) is compared with.
"""
import collections, random, sys

MIX = collections.OrderedDict([     # SASS opcode -> count per pair (r02 capture)
    ("LOP3", 26), ("DFMA", 23), ("IMAD.WIDE", 16), ("DMUL", 12), ("VIADD", 8 + 1 + 2),   # VIADD + IADD3 + LEA.HI
    ("MOV", 6), ("FMUL", 4), ("F2I", 4), ("FSEL", 4), ("DADD", 3), ("SHF", 3 + 2), ("LDS64", 3), ("LDS128", 1),
    ("LDC", 2), ("ISETP", 2 + 2 + 2), ("FSETP", 2), ("F2F", 2), ("FFMA", 2), ("EX2", 2), ("RSQ64H", 1), ("BRA", 4 - 1),
    ("BSSY", 2),
])
assert sum(MIX.values()) + 1 == 141          # + the loop's own backward branch

# The same loop at the END of round 2 (125 instructions per pair: one-FFMA accept filter, table-folded Horner step,
# single-compare loop control; opcode counts from cuobjdump of the shipped kernel's flat series loop, hot path only).
# FADD and FFMA.RM are modelled as FFMA, LEA.HI / BSSY / BSYNC / not-taken BRA as integer adds.   --final selects it.
MIX_FINAL = collections.OrderedDict([
    ("LOP3", 26), ("DFMA", 23), ("IMAD.WIDE", 16), ("DMUL", 13), ("VIADD", 6 + 2 + 2 + 2 + 2),
    ("MOV", 2), ("FFMA", 2 + 2), ("FSEL", 4), ("DADD", 3), ("SHF", 3 + 2), ("LDS64", 2), ("LDS128", 2),
    ("ISETP", 1), ("FSETP", 4), ("F2F", 2), ("EX2", 2), ("RSQ64H", 1),
])
assert sum(MIX_FINAL.values()) + 1 == 125
if "--final" in sys.argv:
    MIX = MIX_FINAL

# Every result must be CONSUMED (ptxas deletes dead instructions whatever `volatile` says): instructions whose result
# the sweep uses elsewhere ("sinks": conversions, loads, shifts, moves, MUFU) hand their value to an instruction of the
# mix that exists anyway -- 32-bit integers to a LOP3 source, doubles to a DFMA addend, floats to an FMUL operand,
# predicates to the next SETP (SETP.AND chains) and finally to an FSEL -- so no instruction is added to the 141.
class Gen:
    def __init__(self):
        self.r32, self.f64, self.f32 = [], [], []       # pending temporaries by type
        self.n = 0
        self.decl = []

    def tmp(self, kind):
        self.n += 1
        name = f"{kind}{self.n}"
        self.decl.append({"u": "uint32_t", "x": "double", "y": "float"}[kind] + f" {name} = 0;")
        return name

    def emit(self, k, i):
        if k == "LOP3":
            src = self.r32.pop(0) if self.r32 else f"l[{(i + 3) % 8}]"
            return f'asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[{i % 8}]) : "r"({src}), "r"(k1));'
        if k == "DFMA":
            add = self.f64.pop(0) if self.f64 else "b"
            return f'asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[{i % 8}]) : "d"(a), "d"({add}));'
        if k == "DMUL":
            return f'asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(m[{i % 4}]) : "d"(a));'
        if k == "DADD":
            return f'asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(m[{i % 4}]) : "d"(b));'
        if k == "IMAD.WIDE" and NO_WIDE:       # control: the same mix with the wide multiplies replaced by 32-bit logic ops
            return f'asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[{i % 8}]) : "r"(l[{(i + 5) % 8}]), "r"(k2));'
        if k == "IMAD.WIDE":
            return f'asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[{i % 4}]) : "r"(l[{i % 8}]), "r"(k2));'
        if k in ("VIADD", "BSSY", "BRA"):     # BSSY / not-taken BRA: one issue slot each, modelled as an integer add
            return f'asm volatile("add.u32 %0, %0, %1;" : "+r"(v[{i % 4}]) : "r"(k1));'
        if k == "FMUL":
            src = self.f32.pop(0) if self.f32 else "fa"
            return f'asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[{i % 4}]) : "f"({src}));'
        if k == "FFMA":
            return f'asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[{i % 4}]) : "f"(fa), "f"(fb));'
        if k == "FSEL":
            return f'asm volatile("selp.f32 %0, %0, %1, p{1 + i % 2};" : "+f"(f[{i % 4}]) : "f"(fb));'
        if k == "ISETP":
            return f'asm volatile("setp.ne.and.u32 p1, %0, %1, p1;" :: "r"(v[{i % 4}]), "r"(k3));'
        if k == "FSETP":
            return f'asm volatile("setp.ge.or.f32 p2, %0, %1, p2;" :: "f"(f[{i % 4}]), "f"(fb));'
        if k == "MOV":
            t = self.tmp("u"); self.r32.append(t)
            return f'asm volatile("mov.b32 %0, %1;" : "=r"({t}) : "r"(v[{i % 4}]));'
        if k == "F2I":
            t = self.tmp("u"); self.r32.append(t)
            return f'asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"({t}) : "f"(f[{i % 4}]));'
        if k == "SHF":
            t = self.tmp("u"); self.r32.append(t)
            return f'asm volatile("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"({t}) : "r"(l[{i % 8}]), "r"(l[{(i + 1) % 8}]), "r"(k3));'
        if k == "F2F":
            t = self.tmp("y"); self.f32.append(t)
            return f'asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"({t}) : "d"(m[{i % 4}]));'
        if k == "EX2":
            t = self.tmp("y"); self.f32.append(t)
            return f'asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"({t}) : "f"(f[{i % 4}]));'
        if k == "LDS64":
            t = self.tmp("x"); self.f64.append(t)
            return f'asm volatile("ld.shared.f64 %0, [%1+{8 * (i % 8)}];" : "=d"({t}) : "r"(sa));'
        if k == "LDS128":
            t, u = self.tmp("x"), self.tmp("x"); self.f64 += [t, u]
            return f'asm volatile("ld.shared.v2.f64 {{%0, %1}}, [%2+64];" : "=d"({t}), "=d"({u}) : "r"(sa));'
        if k == "LDC":
            t = self.tmp("x"); self.f64.append(t)
            return f'asm volatile("ld.const.f64 %0, [ctab+{8 * (i % 4)}];" : "=d"({t}));'
        if k == "RSQ64H":
            t = self.tmp("x"); self.f64.append(t)
            return f'asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"({t}) : "d"(m[0]));'
        raise KeyError(k)


def schedule():
    """Evenly interleaved order (largest-remainder): every class is spread over the whole pair."""
    seq, acc = [], {k: 0.0 for k in MIX}
    total = sum(MIX.values())
    done = collections.Counter()
    for _ in range(total):
        for k in MIX:
            acc[k] += MIX[k] / total
        k = max((k for k in MIX if done[k] < MIX[k]), key=lambda k: acc[k])
        acc[k] -= 1.0
        seq.append((k, done[k]))
        done[k] += 1
    return seq


def main(out):
    g = Gen()
    # two passes over the schedule so that a consumer scheduled BEFORE its producer picks the value up one iteration later
    lines = [g.emit(k, i) for k, i in schedule()]
    body = "\n".join("            " + ln for ln in lines)
    decl = "\n".join("    " + d for d in g.decl)
    assert not g.r32 or len(g.r32) < 6, g.r32
    open(out, "w").write(TEMPLATE.replace("@BODY@", body).replace("@DECL@", decl).replace("@N@", str(sum(MIX.values()) + 1))
                         .replace("@LEFT@", " + ".join(["0.0"] + [f"(double){t}" for t in g.r32 + g.f64 + g.f32])))


TEMPLATE = r'''// mix.cu -- GENERATED by gen_mix.py: dependency-free replay of the headline kernel's per-pair instruction mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mix mix.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ double ctab[4] = {1.0, 2.0, 3.0, 4.0};

__global__ void __launch_bounds__(256, 4) mix_kernel(double *out, int iters, double a, double b, uint32_t k1, uint32_t k2,
                                                     uint32_t k3, float fa, float fb, const double *cptr, long long *cycles)
{
    __shared__ double s_tab[64];
    if (threadIdx.x < 64) s_tab[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(s_tab);
    double d[8], m[4];
    uint64_t w[4];
    uint32_t l[8], v[4];
    float f[4];
@DECL@
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x + i; l[i] = threadIdx.x * 7u + i; }
    for (int i = 0; i < 4; ++i) { m[i] = 1.0 + i; w[i] = threadIdx.x + i; v[i] = threadIdx.x * 3u + i; f[i] = 0.5f + i + threadIdx.x * 1e-3f; }
    asm volatile("{\n .reg .pred p1, p2;\n setp.eq.u32 p1, %0, 7;\n setp.eq.u32 p2, %0, 7;" :: "r"(k3));
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
@BODY@
    }
    const long long t1 = clock64();
    asm volatile("}");
    double acc = @LEFT@;
    for (int i = 0; i < 8; ++i) acc += d[i] + l[i];
    for (int i = 0; i < 4; ++i) acc += m[i] + (double)w[i] + v[i] + f[i];
    out[(size_t)blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount * 4, iters = 20000;      // exactly one resident wave: 8 warps per SMSP
    double *out; long long *cyc; const double *cptr;
    cudaMalloc(&out, sizeof(double) * grid * 256);
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaGetSymbolAddress((void **)&cptr, ctab);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mix_kernel, 256, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, mix_kernel);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        mix_kernel<<<grid, 256>>>(out, iters, 0.999999, 1e-9, 0x9E3779B9u, 0xD2511F53u, 7u, 0.999f, 0.25f, cptr, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        long long *h = new long long[grid];
        cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < grid; ++i) mean += (double)h[i] / grid;
        delete[] h;
        // 8 warps share one SMSP: the SMSP retires one warp-iteration every cycles / (iters * 8)
        printf("%s: %d regs, %d CTAs/SM, %.3f ms, %.1f SM cycles per iteration of one warp, %.1f SMSP cycles per "
               "warp-iteration (@N@ warp instructions incl. 16 IMAD.WIDE)\n", cudaGetErrorString(cudaGetLastError()),
               fa.numRegs, per_sm, ms, mean / iters, mean / iters / 8.0);
    }
    return 0;
}
'''

NO_WIDE = "--no-wide" in sys.argv
if __name__ == "__main__":
    main([a for a in sys.argv[1:] if not a.startswith("--")][0] if len(sys.argv) > 1 else "mix.cu")
