// Issue-rate microbenchmark of the integer instructions the Goldilocks NTT is made of (sm_100a).
// Each kernel runs ILP independent dependency chains of one instruction (or a fixed mix) per thread;
// result = warp-instructions per clock per SM sub-partition at 8 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
typedef unsigned long long u64;
#define ITERS 16384
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) k(u32* out, u32 seed) {
    u32 a[ILP], b[ILP], c[ILP], d[ILP];
    u64 w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i; c[i] = seed ^ (i * 0x9E3779B9u); d[i] = i; w[i] = ((u64)a[i] << 32) | b[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const u32 wl = (u32)w[i], wh = (u32)(w[i] >> 32);
            if (OP == 1) asm volatile("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,%3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(d[i]));   // IADD3 + IADD3.X
            if (OP == 2) asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));                  // IMAD
            if (OP == 3) asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(w[i]) : "r"(wl), "r"(c[i]));                   // IMAD.WIDE.U32, chain through the product
            if (OP == 4) asm volatile("mad.hi.u32 %0,%0,%1,%2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));                  // IMAD.HI.U32
            if (OP == 5) asm volatile("shf.l.wrap.b32 %0,%0,%1,%2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));              // SHF
            if (OP == 6) asm volatile("lop3.b32 %0,%0,%1,%2,0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));               // LOP3
            if (OP == 7) { asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(w[i]) : "r"(wl), "r"(c[i])); asm volatile("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,%0;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i])); }   // IMAD.WIDE + IADD3 + IADD3.X
            if (OP == 9) { asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(w[i]) : "r"(wl), "r"(c[i])); asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); }   // IMAD.WIDE + IMAD
            if (OP == 10) asm volatile("prmt.b32 %0,%0,%1,%2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));                   // PRMT
            if (OP == 11) { asm volatile("mad.lo.u32 %0,%0,%1,%2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); asm volatile("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,%0;" : "+r"(d[i]), "+r"(b[i]) : "r"(c[i])); }   // IMAD + IADD3 + IADD3.X
            if (OP == 12) asm volatile("add.cc.u32 %0,%0,%2; madc.lo.u32 %1,%1,%3,%0;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(d[i]));   // IADD3 + IMAD.X
            if (OP == 13) { asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(w[i]) : "r"(wl), "r"(c[i])); asm volatile("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,%0;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i])); asm volatile("shf.l.wrap.b32 %0,%0,%1,%2;" : "+r"(d[i]) : "r"(b[i]), "r"(c[i])); }   // IMAD.WIDE + IADD3 + IADD3.X + SHF
            if (OP == 14) { asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(w[i]) : "r"(wl), "r"(c[i])); asm volatile("mad.wide.u32 %0,%1,%2,%0;" : "+l"(w[i]) : "r"(wh), "r"(d[i])); asm volatile("add.cc.u32 %0,%0,%2; addc.u32 %1,%1,%0;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i])); }   // 2 IMAD.WIDE + IADD3 + IADD3.X
        }
    }
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i] + b[i] + d[i] + (u32)w[i] + (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int per_iter, u32* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 4;   // 4 CTAs of 8 warps per SM = 8 warps per sub-partition
    k<OP><<<blocks, 256>>>(out, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_inst = (double)blocks * 8 * ITERS * ILP * per_iter;
    const double clocks = ms * 1e-3 * 1.965e9;     // assumes the SM runs at its maximum clock while the kernel runs
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SMSP (at 1965 MHz)\n", name, ms, warp_inst / clocks / (148 * 4));
}

int main() {
    u32* out;
    cudaMalloc(&out, 148 * 4 * 256 * 4);
    for (int rep = 0; rep < 2; rep++) {
    run<1>("IADD3+IADD3.X", 2, out);
    run<2>("IMAD", 1, out);
    run<3>("IMAD.WIDE.U32", 1, out);
    run<4>("IMAD.HI.U32", 1, out);
    run<5>("SHF", 1, out);
    run<6>("LOP3", 1, out);
    run<7>("IMAD.WIDE+IADD3+IADD3.X", 3, out);
    run<9>("IMAD.WIDE+IMAD", 2, out);
    run<10>("PRMT", 1, out);
    run<11>("IMAD+IADD3+IADD3.X", 3, out);
    run<12>("IADD3+IMAD.X", 2, out);
    run<13>("IMAD.WIDE+IADD3+IADD3.X+SHF", 4, out);
    run<14>("2IMAD.WIDE+IADD3+IADD3.X", 4, out);
    }
    // warm long run to report clocks
    return 0;
}
