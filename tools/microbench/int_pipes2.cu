// Second probe: throughput of the exact instruction shapes the field reduction compiles to (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITER 4096
template <int V>
__global__ void k(uint64_t* out, uint32_t seed) {
    uint32_t a = threadIdx.x * 2654435761u + seed, b = blockIdx.x * 40503u + 77u;
    uint64_t x[8];
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = ((uint64_t)a << 20) + i; lo[i] = a * (i + 3); hi[i] = b + i; }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (V == 0) x[i] = (x[i] >> 30) + a;                         // 64-bit shift right + add
            if (V == 1) x[i] = (x[i] & 0x1FFFFFFFFFFFFFFFULL) + (x[i] >> 61);   // Mersenne fold
            if (V == 2) x[i] = x[i] >= 0x1FFFFFFFFFFFFFFFULL ? x[i] - 0x1FFFFFFFFFFFFFFFULL + a : x[i] + b;  // cond subtract
            if (V == 3) x[i] = x[i] + ((uint64_t)b << 32 | a);           // 64-bit add
            if (V == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, 30;" : "+r"(lo[i]) : "r"(hi[i]));   // funnel shift
            if (V == 5) x[i] = ((x[i] & 0x3FFFFFFFULL) << 31) ^ a;       // mask + 64-bit shift left
            if (V == 6) { uint32_t p; asm volatile("{.reg .pred q; setp.gt.u32 q, %1, %2; selp.u32 %0, %1, %2, q;}" : "=r"(p) : "r"(lo[i]), "r"(hi[i])); lo[i] = p + 1; }
            if (V == 7) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(x[i]) : "r"((uint32_t)x[i] + a), "r"(b));   // IMAD.WIDE no acc, dependent
            if (V == 8) { asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a), "r"(b)); }  // add with carry chain
            if (V == 9) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[i]) : "r"(a), "r"(b));                // IMAD.WIDE acc
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + lo[i] + hi[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V>
void run(const char* name) {
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
    double ops = 148.0 * 8 * 8 * ITER * 8.0;
    double cycles = ms * 1e-3 * 1965e6;
    printf("%-40s %8.3f ms  %6.2f source-ops/clk/SM (warp granularity)\n", name, ms, ops / cycles / 148.0);
    cudaFree(d);
}
int main() {
    run<0>("u64 >> 30, + u32");
    run<1>("Mersenne fold (x&p)+(x>>61)");
    run<2>("cond subtract (x>=p ? x-p : x)");
    run<3>("u64 add");
    run<4>("shf.r.wrap (funnel)");
    run<5>("(x & mask30) << 31");
    run<6>("setp + selp");
    run<7>("mul.wide.u32 dependent");
    run<8>("add.cc / addc pair");
    run<9>("mad.wide.u32 acc");
    return 0;
}
