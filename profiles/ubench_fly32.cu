// Register-only butterfly-rate microbenchmark for the 32-bit-lane kernels (c5: 18-bit DIT TRUNCATE):
//   KIND 0: round32<> of intfft_fast32.cuh unchanged (IMAD.WIDE products + funnel-shift slices)
//   KIND 1: the same values from 32-bit IMADs only: B = 2 Bh + Bl, so
//           (P2 -+ P1) >> 1 = (Bh1 W1 -+ Bh2 W2) + ((Bl1 W1 -+ Bl2 W2) >> 1)  (mod 2^32),
//           and the kept slice P(32 downto 16) is the top 17 bits of that word
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I intfftk_b200/csrc -o profiles/ubench_fly32 profiles/ubench_fly32.cu intfftk_b200/csrc/intfft_twiddle.cpp
#include <cstdio>
#include "../intfftk_b200/csrc/intfft_fast32.cuh"

using namespace intfft;
using namespace intfft::f32;

__device__ __forceinline__ int sra_n(int x, int n) { return x >> n; }

// DIT TRUNCATE butterfly, single-DSP slice, DATA_WIDTH 18 / TWDL_WIDTH 16, 32-bit products only.
// Values are carried as their floor-halves where that is what the next consumer reads.
__device__ __forceinline__ void fly_split(int &ar, int &ai, int &br, int &bi, int wr, int wi)
{
    // multiplier inputs (swapped re / im, int_dit2_fly.vhd:304-322): d1 = bi, d2 = br
    const int h1 = bi >> 1, h2 = br >> 1, l1 = bi & 1, l2 = br & 1;
    const int c_re = (int)((unsigned)l1 * (unsigned)wr - (unsigned)l2 * (unsigned)wi) >> 1;
    const int c_im = (int)((unsigned)l1 * (unsigned)wi + (unsigned)l2 * (unsigned)wr) >> 1;
    const int u_re = (int)((unsigned)h1 * (unsigned)wr - (unsigned)h2 * (unsigned)wi + (unsigned)c_re);
    const int u_im = (int)((unsigned)h1 * (unsigned)wi + (unsigned)h2 * (unsigned)wr + (unsigned)c_im);
    const int hi = u_re >> 15, hr = u_im >> 15;             // (BW >> 1), outputs swapped back
    const int xr = sra1(ar) + hr, xi = sra1(ai) + hi;
    br = msub2(hr, xr);
    bi = msub2(hi, xi);
    ar = xr;
    ai = xi;
}

// Three-multiply complex product (Gauss) on pre-shifted twiddle triples a = wr << 15, b = -(wr + wi) << 15,
// c = (wi - wr) << 15 (|wr +- wi| <= sqrt(2) * 32767 < 2^16, so the triples fit an int32):
//   k1 = a (dr + di);  re = k1 + di b = (dr wr - di wi) << 15;  im = k1 + dr c = (dr wi + di wr) << 15   (exact, mod 2^64)
// The kept slice wrap17(S >> 16) sits at bits 31..47 of the 64-bit word: one funnel shift + SGXT per output.
// 3 IMAD.WIDE + 1 IADD + 2 SHF instead of 4 IMAD.WIDE / IMAD.HI.
__device__ __forceinline__ int slice31(long long t)
{
    unsigned lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
    return sgxt32((int)__funnelshift_l(lo, hi, 1), 17);
}
__device__ __forceinline__ void fly_k3(int &ar, int &ai, int &br, int &bi, int a, int b, int c)
{
    const int dr = bi, di = br;                              // swapped re / im into the multiplier
    const long long k1 = (long long)a * (dr + di);
    const long long tr = k1 + (long long)di * b;
    const long long ti = k1 + (long long)dr * c;
    const int hi = slice31(tr), hr = slice31(ti);            // (BW >> 1), outputs swapped back
    const int xr = sra1(ar) + hr, xi = sra1(ai) + hi;
    br = msub2(hr, xr);
    bi = msub2(hi, xi);
    ar = xr;
    ai = xi;
}

// KIND_SINGLE_PRE with the twiddles shifted one bit less: the kept slice then starts at bit 31, so the LOW word is needed
// (one funnel shift per output) and ptxas cannot narrow the accumulating multiply to IMAD.HI — which issues at 30.7 per
// clock per SM against IMAD.WIDE's 39.3 (ubench_int2).  4 IMAD.WIDE + 2 SHF + 2 SGXT instead of 2 IMAD.WIDE + 2 IMAD.HI + 2 SGXT.
__device__ __forceinline__ void fly_wide(int &ar, int &ai, int &br, int &bi, int wr, int wi)
{
    const long long tr = (long long)bi * wr - (long long)br * wi;
    const long long ti = (long long)bi * wi + (long long)br * wr;
    const int hi = slice31(tr), hr = slice31(ti);
    const int xr = sra1(ar) + hr, xi = sra1(ai) + hi;
    br = msub2(hr, xr);
    bi = msub2(hi, xi);
    ar = xr;
    ai = xi;
}

template <int KIND>
__global__ void __launch_bounds__(256, 3) fly_loop(const int2 *in, int2 *out, const int2 *tw, int iters, const __grid_constant__ Fast32Params p)
{
    extern __shared__ unsigned char smem[];
    const unsigned tid = threadIdx.x;
    V re[16], im[16];
    int uwr[15], uwi[15];
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int2 v = in[tid + 256 * i]; re[i] = mk(sx(v.x, 18)); im[i] = mk(sx(v.y, 18)); }
#pragma unroll
    for (int i = 0; i < 15; ++i) { const int2 w = tw[tid * 15 + i]; uwr[i] = w.x; uwi[i] = w.y; }
    for (int it = 0; it < iters; ++it) {
        if (KIND == 0) round32<4, true, MODE_TRUNC, KIND_SINGLE>(re, im, p, 8, TwRegs32{uwr, uwi}, false, false);
        if (KIND == 2) round32<4, true, MODE_TRUNC, KIND_SINGLE_PRE>(re, im, p, 8, TwRegs32{uwr, uwi}, false, false);
        if (KIND == 3) {          // production butterfly, twiddles from a shared table (round B of the 8192-point kernel)
            const int2 *t = reinterpret_cast<const int2 *>(smem) + (tid & 15u);
            round32<4, true, MODE_TRUNC, KIND_SINGLE_PRE>(re, im, p, 4, TwSmem32{t, 16}, false, false);
        }
        if (KIND == 4) {          // three-multiply butterfly, twiddle triples from a shared table
            const int4 *t = reinterpret_cast<const int4 *>(smem) + (tid & 15u);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    if (m & (1 << q)) continue;
                    const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
                    const int4 tw3 = t[w * 16];
                    fly_k3(re[m].f, im[m].f, re[m | (1 << q)].f, im[m | (1 << q)].f, tw3.x, tw3.y, tw3.z);
                }
        }
        if (KIND == 5) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    if (m & (1 << q)) continue;
                    const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
                    fly_wide(re[m].f, im[m].f, re[m | (1 << q)].f, im[m | (1 << q)].f, uwr[w], uwi[w]);
                }
        }
        if (KIND == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    if (m & (1 << q)) continue;
                    const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
                    fly_split(re[m].f, im[m].f, re[m | (1 << q)].f, im[m | (1 << q)].f, uwr[w], uwi[w]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) out[(blockIdx.x * 256 + tid) * 16 + i] = make_int2(re[i].f, im[i].f);
    if (iters < 0) smem[tid] = 0;          // (the tables are never initialised: timing only)
}

template <int KIND> void run(const char *name, int smem)
{
    const int grid = 148 * 3 * 4, iters = 256;
    int2 *in, *out, *tw;
    cudaMalloc(&in, 4096 * 8); cudaMalloc(&out, (size_t)grid * 4096 * 8); cudaMalloc(&tw, 256 * 15 * 8);
    cudaMemset(in, 0x5a, 4096 * 8); cudaMemset(tw, 0x33, 256 * 15 * 8);
    Fast32Params p{};
    p.n = 13; p.dw = 18; p.format = 0; p.cm = cmult_consts(16, 1);
    cudaFuncSetAttribute(fly_loop<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fly_loop<KIND><<<grid, 256, smem>>>(in, out, tw, iters, p); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    fly_loop<KIND><<<grid, 256, smem>>>(in, out, tw, iters, p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flies = (double)grid * 256 * iters * 32;
    std::printf("%-52s %8.3f ms  %7.2f G butterflies/s  = %6.2f ms for the 5.9 G multiplying butterflies of c5 (%s)\n", name, ms,
                flies / ms / 1e6, 5.9e9 / (flies / ms), cudaGetErrorString(cudaGetLastError()));
    cudaFree(in); cudaFree(out); cudaFree(tw);
}

int main()
{
    run<0>("fly32 DIT TRUNC (IMAD.WIDE), 3 CTAs/SM", 72 * 1024);
    run<0>("fly32 DIT TRUNC (IMAD.WIDE), 6 CTAs/SM", 0);
    run<1>("split-operand 32-bit IMAD form, 3 CTAs/SM", 72 * 1024);
    run<1>("split-operand 32-bit IMAD form, 6 CTAs/SM", 0);
    run<2>("production KIND_SINGLE_PRE, register twiddles, 3 CTAs/SM", 72 * 1024);
    run<3>("production KIND_SINGLE_PRE, shared-table twiddles, 3 CTAs/SM", 72 * 1024);
    run<4>("three-multiply (Gauss) form, shared-table triples, 3 CTAs/SM", 72 * 1024);
    run<5>("pre-shift - 1: 4 IMAD.WIDE + funnel + SGXT, register twiddles", 72 * 1024);
    return 0;
}
