// Issue-rate microbenchmark, third edition: is the FP64 pipe usable as an exact wide-integer multiplier
// beside the IMAD.WIDE port?  (The 32-bit-lane kernels are bound by IMAD.WIDE: 4 per butterfly at ~4 port-cycles.)
// One opcode (or one fixed mix) per kernel, 8 independent chains per thread, 8 resident CTAs of 256 threads per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_fp64 ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define ILP 8
template <int KIND>
__global__ void k(long long *out, int a0, int b0, double d0)
{
    int v[ILP]; long long w[ILP]; double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x + i + a0; w[i] = v[i] * 7ll; d[i] = (double)(v[i]) * d0; }
    int b = b0 | 1;
    const double m = d0 * 0.999, c = d0 * 1e-9;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
            if (KIND == 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(m));
            if (KIND == 2) asm volatile("add.rm.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(c));
            if (KIND == 3) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                             asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b)); }
            if (KIND == 4) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                             asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(v[i]) : "r"(b)); }
            if (KIND == 5) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                             asm volatile("shr.s32 %0, %0, 1;" : "+r"(v[i])); }
            if (KIND == 6) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                             asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b));
                             asm volatile("shr.s32 %0, %0, 1;" : "+r"(v[i])); }
            if (KIND == 7) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b));
            if (KIND == 8) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));   // 2 DFMA : 1 WIDE
                             asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));
                             asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b)); }
            if (KIND == 9) { double t; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(v[i]));            // I2F.F64.S32
                             asm volatile("mov.b64 {%0, _}, %1;" : "=r"(v[i]) : "d"(t)); }
            if (KIND == 10) { double t; int x; asm volatile("xor.b32 %0, %1, 0x80000000;" : "=r"(x) : "r"(v[i]));  // LOP3 + DADD
                              asm volatile("mov.b64 %0, {%1, %2};" : "=d"(t) : "r"(x), "r"(0x43300000));
                              asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(t) : "d"(c));
                              asm volatile("mov.b64 {%0, _}, %1;" : "=r"(v[i]) : "d"(t)); }
            if (KIND == 11) { double t; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(v[i]));           // I2F + DFMA + WIDE
                              asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(t), "d"(m));
                              asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b)); }
        }
    }
    long long s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i] + w[i] + (long long)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND> void run(const char *name, double ops)
{
    long long *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 8, 256>>>(out, 1, 3, 1.0000001); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<KIND><<<148 * 8, 256>>>(out, 1, 3, 1.0000001);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = 5.0 * 148 * 8 * 256 * (double)ITERS * ILP * ops;
    std::printf("%-40s %7.2f T thread-instr/s = %6.1f per clk per SM @1.965GHz  (%.3f ms)\n", name, n / (ms * 1e-3) / 1e12,
                n / (ms * 1e-3) / 148 / 1.965e9, ms / 5);
    cudaFree(out);
}
int main()
{
    run<0>("DFMA", 1); run<1>("DMUL", 1); run<2>("DADD.RM", 1); run<7>("IMAD.WIDE", 1);
    run<3>("DFMA + IMAD.WIDE alternating", 2); run<4>("DFMA + IMAD alternating", 2); run<5>("DFMA + SHF alternating", 2);
    run<6>("DFMA + IMAD.WIDE + SHF", 3); run<8>("2 DFMA + IMAD.WIDE", 3);
    run<9>("I2F.F64.S32 (+mov)", 1); run<10>("LOP3 + DADD int->double", 1); run<11>("I2F.F64 + DFMA + IMAD.WIDE", 3);
    return 0;
}
