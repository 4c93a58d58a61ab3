// One-pass 8192-point kernel on 32-bit lanes (BASELINE c5: 8192-pt 18-bit DIT IFFT; every NFFT = 13 plan
// whose values fit 32 bits).  Replaces the strided-4 + contiguous-9 two-pass schedule: each sample is
// read from HBM once and written once (16 B instead of 32 B per sample at c5), and one of the two
// load / store / address-generation sequences disappears.
//
// Shape: one CTA of 512 threads per SM owns an 8192-sample frame.  A thread keeps 16 samples in
// registers; the 13 stages run as four register rounds
//     A: STAGE 0..3  (16 contiguous samples)        twiddles: kernel parameters (constant bank)
//     B: STAGE 4..7  (stride 16)                    twiddles: 15 x 16 shared table (depend on tid & 15)
//     C: STAGE 8..11 (stride 256)                   twiddles: 15 per-thread registers (depend on tid & 255)
//     D: STAGE 12    (8 pairs i, i + 4096)          twiddles: 8 per-thread registers
// DIT walks A -> D, DIF walks D -> A.  The three ownership changes (the delay-line commutations of
// int_delay_line.vhd:52-104) go through a double-buffered, skewed shared-memory tile: one CTA barrier
// each.  The next frame is prefetched with cp.async into thread-private staging slots while the current
// one is being computed.  Arithmetic = intfft_fast32.cuh (fly32 / cmul32), i.e. int_dif2_fly.vhd:142-373,
// int_dit2_fly.vhd:140-325, int_cmult_dsp48.vhd:182-190 / 307-317 and the dbl18 / dbl35 arrangements.
#include "intfft_fast32.cuh"

namespace intfft {

namespace f32 {

constexpr unsigned kTile13 = 9216;                         // >= phys8(8191) + 1
constexpr unsigned kStage13 = 16 * 512 * 8;                // 16 slots x 512 threads x 8 bytes
constexpr unsigned kSmem13 = kHead32 + 2 * kTile13 * 8 + kStage13;

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// register m of round D: pair index j = m >> 1 (sample t + 512 j), upper / lower half = m & 1
__host__ __device__ constexpr unsigned offD(int m) { return ((unsigned)(m >> 1) << 9) | ((unsigned)(m & 1) << 12); }

template <bool DIT, int MODE, int KIND>
__global__ void __launch_bounds__(512, 1) fast32_n13_kernel(const __grid_constant__ Fast32Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);                       // [15][16]
    int2(*work)[kTile13] = reinterpret_cast<int2(*)[kTile13]>(smem_raw + kHead32);
    unsigned char *stage = smem_raw + kHead32 + 2 * kTile13 * 8;

    const unsigned tid = threadIdx.x;
    const int esz = 2 * p.in_sb;                                                  // bytes per complex sample read

    // ---- batch-invariant twiddles ----
    if (tid < 240) {                                   // round B: table[w][tid & 15], w = (1 << q) - 1 + j, STAGE 4 + q
        const int w = tid >> 4, lo4 = tid & 15;
        const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
        const int j = w - ((1 << q) - 1);
        midtw[w * 16 + lo4] = __ldg(p.tw + (1u << (4 + q)) + lo4 + ((unsigned)j << 4));
    }
    int uwr[15], uwi[15];                              // round C: index (tid & 255) + 256 j at STAGE 8 + q
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < (1 << q); ++j) {
            const int2 w = __ldg(p.tw + (1u << (8 + q)) + (tid & 255u) + ((unsigned)j << 8));
            uwr[(1 << q) - 1 + j] = w.x;
            uwi[(1 << q) - 1 + j] = w.y;
        }
    int dwr[8], dwi[8];                                // round D: index tid + 512 j at STAGE 12
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int2 w = __ldg(p.tw + (1u << 12) + tid + ((unsigned)j << 9));
        dwr[j] = w.x;
        dwi[j] = w.y;
    }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }

    // ---- ownership (8-byte slots of the skewed tile) ----
    const unsigned pA = phys8(16u * tid);
    const unsigned pB = phys8((tid & 15u) | ((tid >> 4) << 8));
    const unsigned pC = phys8((tid & 255u) | ((tid >> 8) << 12));
    const unsigned pD = phys8(tid);

    // ---- prefetch of a frame's first-round samples into this thread's staging slots ----
    auto prefetch = [&](long long t) {
        const char *src = reinterpret_cast<const char *>(p.in) + (t << 13) * esz;
        if (DIT) {                                     // 16 contiguous samples = esz pieces of 16 bytes
            src += (size_t)(16u * tid) * esz;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j < esz) cp_async_16(stage + (j * 512 + tid) * 16, src + 16 * j);
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m)
                cp_async_elem(stage + (m * 512 + tid) * esz, src + (size_t)(tid + offD(m)) * esz, esz);
        }
        cp_async_commit();
    };
    if ((long long)blockIdx.x < p.n_tiles) prefetch(blockIdx.x);
    __syncthreads();

    const Stg stD = stage_of<DIT, MODE, KIND>(p, 12);

    int ex = 0;                                        // exchanges done so far: selects the buffer
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const long long g0 = tile << 13;
        V re[16], im[16];

        // ---- first round's samples: drain the staging slots, refill them with the next frame ----
        cp_async_wait_all();
        if (DIT) {
            if (p.in_sb == 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int4 v = reinterpret_cast<const int4 *>(stage)[j * 512 + tid];
                    re[2 * j] = mk(sx(v.x, p.dw)); im[2 * j] = mk(sx(v.y, p.dw));
                    re[2 * j + 1] = mk(sx(v.z, p.dw)); im[2 * j + 1] = mk(sx(v.w, p.dw));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 v = reinterpret_cast<const uint4 *>(stage)[j * 512 + tid];
                    const unsigned x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        re[4 * j + e] = mk(sx((int)x[e], p.dw));
                        im[4 * j + e] = mk(sx((int)x[e] >> 16, p.dw));
                    }
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                int a, b;
                if (p.in_sb == 4) {
                    const int2 v = reinterpret_cast<const int2 *>(stage)[m * 512 + tid];
                    a = v.x; b = v.y;
                } else {
                    const unsigned x = reinterpret_cast<const unsigned *>(stage)[m * 512 + tid];
                    a = (int)x; b = (int)x >> 16;
                }
                re[m] = mk(sx(a, p.dw));
                im[m] = mk(sx(b, p.dw));
            }
        }
        {
            const long long nt = tile + gridDim.x;
            if (nt < p.n_tiles) prefetch(nt);
        }

#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = DIT ? rr : 3 - rr;               // 0 = A, 1 = B, 2 = C, 3 = D
            if (r == 0) round32<4, DIT, MODE, KIND>(re, im, p, 0, TwRegs32{lwr, lwi}, true, false);
            else if (r == 1) round32<4, DIT, MODE, KIND>(re, im, p, 4, TwSmem32{midtw + (tid & 15u), 16}, false, false);
            else if (r == 2) round32<4, DIT, MODE, KIND>(re, im, p, 8, TwRegs32{uwr, uwi}, false, false);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    fly32<DIT, MODE, KIND>(stD, false, p.cm, re[2 * j], im[2 * j], re[2 * j + 1], im[2 * j + 1], dwr[j], dwi[j]);
            }
            if (rr == 3) break;
            // ---- ownership change r -> next round ----
            const int rn = DIT ? r + 1 : r - 1;
            int2 *sm = work[ex & 1];
            ++ex;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned a = r == 0 ? pA + m : (r == 1 ? pB + 18u * m : (r == 2 ? pC + 288u * m : pD + phys8(offD(m))));
                sm[a] = make_int2(re[m].f, im[m].f);
            }
            // rounds A, B, C never leave a 4096-sample half, and thread half == sample half in all three
            // layouts: those two changes only need the 256 threads of the half (named barrier 1 / 2)
            if (r == 3 || rn == 3) __syncthreads();
            else asm volatile("bar.sync %0, 256;" ::"r"(1 + (int)(tid >> 8)) : "memory");
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned a = rn == 0 ? pA + m : (rn == 1 ? pB + 18u * m : (rn == 2 ? pC + 288u * m : pD + phys8(offD(m))));
                const int2 v = sm[a];
                re[m] = mk(v.x);
                im[m] = mk(v.y);
            }
        }

        // ---- results ----
        if (DIT) {                                         // round D ownership: coalesced element stores
#pragma unroll
            for (int m = 0; m < 16; ++m) st_sample(p.out, g0 + tid + offD(m), p.out_sb, re[m].f, im[m].f);
        } else if (p.out_sb == 4) {                        // round A ownership: 16 contiguous samples per thread
            int4 *dst = reinterpret_cast<int4 *>(p.out) + ((g0 + 16u * tid) >> 1);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_int4(re[2 * j].f, im[2 * j].f, re[2 * j + 1].f, im[2 * j + 1].f);
        } else {
            uint4 *dst = reinterpret_cast<uint4 *>(p.out) + ((g0 + 16u * tid) >> 2);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                dst[j] = make_uint4(__byte_perm((unsigned)re[4 * j].f, (unsigned)im[4 * j].f, 0x5410),
                                    __byte_perm((unsigned)re[4 * j + 1].f, (unsigned)im[4 * j + 1].f, 0x5410),
                                    __byte_perm((unsigned)re[4 * j + 2].f, (unsigned)im[4 * j + 2].f, 0x5410),
                                    __byte_perm((unsigned)re[4 * j + 3].f, (unsigned)im[4 * j + 3].f, 0x5410));
        }
    }
}

template <typename K> cudaError_t launch_n13(K k, const Fast32Params &p, int grid, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem13);
    if (e != cudaSuccess) return e;
    k<<<grid, 512, kSmem13, st>>>(p);
    return cudaGetLastError();
}

template <bool DIT> cudaError_t launch_n13_dir(const Fast32Params &p, int mode, int kind, int grid, cudaStream_t st)
{
    switch (mode * 2 + kind) {
    case MODE_TRUNC * 2 + 0: return launch_n13(fast32_n13_kernel<DIT, MODE_TRUNC, KIND_SINGLE>, p, grid, st);
    case MODE_TRUNC * 2 + 1: return launch_n13(fast32_n13_kernel<DIT, MODE_TRUNC, KIND_MIXED>, p, grid, st);
    case MODE_ROUND * 2 + 0: return launch_n13(fast32_n13_kernel<DIT, MODE_ROUND, KIND_SINGLE>, p, grid, st);
    case MODE_ROUND * 2 + 1: return launch_n13(fast32_n13_kernel<DIT, MODE_ROUND, KIND_MIXED>, p, grid, st);
    case MODE_UNSCALED * 2 + 0: return launch_n13(fast32_n13_kernel<DIT, MODE_UNSCALED, KIND_SINGLE>, p, grid, st);
    default: return launch_n13(fast32_n13_kernel<DIT, MODE_UNSCALED, KIND_MIXED>, p, grid, st);
    }
}

}  // namespace f32

int f32_launch_n13(const f32::Fast32Params &p, bool dit, int mode, int kind, int grid, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return (int)(dit ? f32::launch_n13_dir<true>(p, mode, kind, grid, st) : f32::launch_n13_dir<false>(p, mode, kind, grid, st));
}

}  // namespace intfft
