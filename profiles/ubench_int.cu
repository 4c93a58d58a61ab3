// Integer issue-rate microbenchmark for B200 (SURVEY.md §8d "second ceiling"): how many
// thread-instructions per second the SMs sustain for the instruction kinds the butterfly uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_int ubench_int.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int KIND>
__global__ void k(int *out, int a0, int b0)
{
    int v[ILP];
    long long w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { v[i] = threadIdx.x + i + a0; w[i] = v[i]; }
    int b = b0 | 1, c = a0 ^ 5;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) v[i] = v[i] * b + c;                                   // IMAD
            if (KIND == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));   // IADD3
            if (KIND == 2) asm volatile("shr.s32 %0, %0, 1; xor.b32 %0, %0, %1;" : "+r"(v[i]) : "r"(b)); // SHF + LOP3
            if (KIND == 3) { v[i] = v[i] * b + c; asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(c)); } // IMAD + IADD
            if (KIND == 4) w[i] = (long long)(int)w[i] * (long long)b + w[i];     // IMAD.WIDE
            if (KIND == 5) { v[i] = v[i] * b + c; asm volatile("add.s32 %0, %0, %1; shr.s32 %0, %0, 1;" : "+r"(v[i]) : "r"(c)); } // 1 IMAD : 2 ALU
            if (KIND == 6) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(v[i]) : "r"(b));  // PRMT
            if (KIND == 7) asm volatile("shf.r.clamp.b32 %0, %0, %1, 3;" : "+r"(v[i]) : "r"(b)); // SHF funnel
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i] + (int)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND> void run(const char *name, double ops_per_iter)
{
    int *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(int));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND><<<148 * 8, 256>>>(out, 1, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<KIND><<<148 * 8, 256>>>(out, 1, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = 5.0 * 148 * 8 * 256 * (double)ITERS * ILP * ops_per_iter;
    std::printf("%-28s %8.2f T thread-instr/s  (%.3f ms)\n", name, n / (ms * 1e-3) / 1e12, ms / 5);
    cudaFree(out);
}

int main()
{
    run<0>("IMAD", 1); run<1>("IADD3", 1); run<2>("SHF+LOP3", 2); run<3>("IMAD+IADD (1:1)", 2);
    run<4>("IMAD.WIDE", 1); run<5>("IMAD+IADD+SHF (1:2)", 3); run<6>("PRMT", 1); run<7>("SHF.funnel", 1);
    return 0;
}
