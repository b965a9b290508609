// Integer-multiply throughput probe for sm_100a: which SASS multiply is fast on B200?
// Each variant runs ITER iterations of 8 independent ops per thread; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int V>
__global__ void k(uint64_t* out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 77u;
    uint64_t acc[8];
    uint32_t lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = a + i; lo[i] = a * (i + 3); }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (V == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo[i]), "r"(b));          // IMAD.WIDE.U32 + 64-bit acc
            if (V == 1) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(lo[i]), "r"(b)); acc[i] ^= t; }  // IMAD.WIDE no acc (+LOP)
            if (V == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(b), "r"(a));                 // IMAD 32
            if (V == 3) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(b));                            // IMAD.HI.U32
            if (V == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(lo[i]) : "r"(b));                               // IADD
            if (V == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(lo[i]) : "r"(b));                      // SHF
            if (V == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(lo[i]) : "r"(b), "r"(a));            // LOP3
            if (V == 7) { double d = __longlong_as_double(acc[i]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5)); acc[i] = __double_as_longlong(d); }  // DFMA
            if (V == 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[i]) : "r"(b & 0xffff), "r"(a));        // IMAD 32 (16-bit operand)
            if (V == 9) { uint32_t h, l2; asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(l2), "=r"(h) : "r"(lo[i]), "r"(b)); lo[i] = l2 ^ h; }  // lo+hi pair
            if (V == 10) { float f = __uint_as_float(lo[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); lo[i] = __float_as_uint(f); }  // FFMA
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i] + lo[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
void run(const char* name, int ops_per_iter = 8) {
    uint64_t* d;
    cudaMalloc(&d, 148 * 8 * 256 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<148 * 8, 256>>>(d, 1);
    cudaEventRecord(e0);
    k<V><<<148 * 8, 256>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double warp_instr = 148.0 * 8 * 8 /*warps*/ * ITER * (double)ops_per_iter;
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("%-34s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, warp_instr / cycles / 148.0, clk_khz / 1000);
    cudaFree(d);
}
int main() {
    run<0>("mad.wide.u32 (IMAD.WIDE acc)");
    run<1>("mul.wide.u32 + xor");
    run<2>("mad.lo.u32 (IMAD)");
    run<3>("mul.hi.u32 (IMAD.HI)");
    run<4>("add.u32");
    run<5>("shf");
    run<6>("lop3");
    run<7>("fma.f64 (DFMA)");
    run<8>("mad.lo.u32 16-bit operand");
    run<9>("mul.lo + mul.hi pair", 16);
    run<10>("fma.f32 (FFMA)");
    return 0;
}
