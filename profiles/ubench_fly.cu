// Register-only butterfly-rate microbenchmark: the c2 kernel's 4-stage register round (round_regs<> of
// intfft_fast16.cu, unchanged) in a loop with no shared / global traffic.  Gives the arithmetic ceiling
// of the packed-16 kernels at the same residency (3 x 256 threads per SM), to compare with what the
// full kernel achieves (profiles/r01c/ubench_fly.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I intfftk_b200/csrc -o profiles/ubench_fly profiles/ubench_fly.cu
#include <cstdio>
#include "../intfftk_b200/csrc/intfft_fast16.cu"

namespace intfft { void count_launch(int) {} }
using namespace intfft;

template <int KIND>
__global__ void __launch_bounds__(256, 3) fly_loop(const uint32_t *in, uint32_t *out, const int2 *tw, int iters)
{
    extern __shared__ unsigned char smem[];
    const unsigned tid = threadIdx.x;
    int re[16], im[16], uwr[15], uwi[15];
#pragma unroll
    for (int i = 0; i < 16; ++i) unpack<true>(in[tid + 256 * i], 16, re[i], im[i]);
#pragma unroll
    for (int i = 0; i < 15; ++i) { const int2 w = tw[tid * 15 + i]; uwr[i] = w.x; uwi[i] = w.y; }
    for (int it = 0; it < iters; ++it) {
        if (KIND == 0) round_regs<8, 4, false, true, MODE_TRUNC, false>(re, im, TwRegs{uwr, uwi}, false, 16, 17);
        if (KIND == 1) round_regs<8, 4, true, true, MODE_TRUNC, false>(re, im, TwRegs{uwr, uwi}, false, 16, 17);
        if (KIND == 2) {     // with the pack / unpack of a round boundary, still no memory traffic
            round_regs<8, 4, false, true, MODE_TRUNC, true>(re, im, TwRegs{uwr, uwi}, false, 16, 17);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const uint32_t x = (m & 1) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632) : pack(re[m], im[m]);
                unpack<true>(x ^ (unsigned)it, 16, re[m], im[m]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(blockIdx.x * 256 + tid) * 16 + i] = pack(re[i], im[i]);
    if (iters < 0) smem[tid] = 0;
}

template <int KIND> void run(const char *name, int smem)
{
    const int grid = 148 * 3 * 4, iters = 512;
    uint32_t *in, *out; int2 *tw;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, (size_t)grid * 4096 * 4); cudaMalloc(&tw, 256 * 15 * 8);
    cudaMemset(in, 0x5a, 4096 * 4); cudaMemset(tw, 0x33, 256 * 15 * 8);
    cudaFuncSetAttribute(fly_loop<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fly_loop<KIND><<<grid, 256, smem>>>(in, out, tw, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    fly_loop<KIND><<<grid, 256, smem>>>(in, out, tw, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flies = (double)grid * 256 * iters * 32;          // 4 stages x 8 butterflies per thread and round
    std::printf("%-46s %8.3f ms  %7.2f G butterflies/s  = %6.2f ms for the 1.61 G butterflies of c2 (%s)\n", name, ms,
                flies / ms / 1e6, 1.610612736e9 / (flies / ms), cudaGetErrorString(cudaGetLastError()));
    cudaFree(in); cudaFree(out); cudaFree(tw);
}

int main()
{
    run<0>("DIF round, 3 CTAs/SM (72 KB smem each)", 72 * 1024);
    run<0>("DIF round, 6+ CTAs/SM (no smem)", 0);
    run<1>("DIT round, 3 CTAs/SM", 72 * 1024);
    run<2>("DIF round + pack/unpack, 3 CTAs/SM", 72 * 1024);
    return 0;
}
