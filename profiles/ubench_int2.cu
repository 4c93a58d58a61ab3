// Issue-rate microbenchmark, second edition: one SASS opcode per kernel (checked with cuobjdump),
// 8 independent chains per thread, 8 resident CTAs of 256 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define ILP 8
template <int KIND>
__global__ void k(long long *out, int a0, int b0)
{
    int v[ILP]; long long w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x + i + a0; w[i] = v[i] * 7ll; }
    int b = b0 | 1;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(v[i]) : "r"(b));
            if (KIND == 1) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b));
            if (KIND == 2) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (KIND == 3) asm volatile("shf.r.clamp.b32 %0, %0, %1, 15;" : "+r"(v[i]) : "r"(b));
            if (KIND == 4) asm volatile("bfe.s32 %0, %0, 0, 17;" : "+r"(v[i]));
            if (KIND == 5) asm volatile("shr.s32 %0, %0, 1;" : "+r"(v[i]));
            if (KIND == 6) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(v[i]) : "r"(b));
            if (KIND == 7) { asm volatile("mad.lo.s32 %0, %0, %1, %0;" : "+r"(v[i]) : "r"(b)); asm volatile("shr.s32 %0, %0, 1;" : "+r"(v[i])); }
            if (KIND == 8) { asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(v[i]), "r"(b)); asm volatile("shr.s32 %0, %0, 1;" : "+r"(v[i])); }
            if (KIND == 9) asm volatile("add.s64 %0, %0, %1;" : "+l"(w[i]) : "l"((long long)b));
        }
    }
    long long s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND> void run(const char *name, double ops)
{
    long long *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(long long));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 8, 256>>>(out, 1, 3); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<KIND><<<148 * 8, 256>>>(out, 1, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = 5.0 * 148 * 8 * 256 * (double)ITERS * ILP * ops;
    std::printf("%-34s %7.2f T thread-instr/s = %6.1f per clk per SM @1.965GHz\n", name, n / (ms * 1e-3) / 1e12, n / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main()
{
    run<0>("IMAD (mad.lo)", 1); run<1>("IMAD.WIDE (mad.wide)", 1); run<2>("IMAD.HI (mul.hi)", 1);
    run<3>("SHF funnel", 1); run<4>("bfe.s32 (SGXT?)", 1); run<5>("SHF.R.S32 (shr)", 1); run<6>("LOP3", 1);
    run<7>("IMAD + SHF alternating", 2); run<8>("IMAD.WIDE + SHF alternating", 2); run<9>("add.s64", 1);
    return 0;
}
