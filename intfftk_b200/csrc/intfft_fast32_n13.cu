// One-pass 8192-point kernel on 32-bit lanes (BASELINE c5: 8192-pt 18-bit DIT IFFT; every NFFT = 13 plan
// whose values fit 32 bits).  Replaces the strided-4 + contiguous-9 two-pass schedule: each sample is
// read from HBM once and written once (16 B instead of 32 B per sample at c5), and one of the two
// load / store / address-generation sequences disappears.
//
// Shape: a CTA of 256 threads (three per SM, so one CTA's exchanges overlap the others' arithmetic) owns an
// 8192-sample frame and walks its two 4096-sample halves one after the other.  A thread keeps 16 samples
// in registers; per half, STAGE 0..11 run as three register rounds
//     A: STAGE 0..3  (16 contiguous samples)        twiddles: kernel parameters (constant bank)
//     B: STAGE 4..7  (stride 16)                    twiddles: 15 x 16 shared table (depend on tid & 15)
//     C: STAGE 8..11 (stride 256)                   twiddles: 15 per-thread registers
// and STAGE 12 pairs sample i of the lower half with i + 4096 of the upper half.  In round C's ownership
// (tid + 256 m) both samples of every STAGE-12 pair belong to the SAME thread, so that stage needs no
// exchange: one half's 16 samples wait in thread-private shared-memory slots while the other half is
// computed.  DIT walks A, B, C per half and then STAGE 12; DIF starts with STAGE 12 and walks C, B, A.
// Ownership changes (the delay-line commutations of int_delay_line.vhd:52-104): A <-> B stays inside a
// half-warp (__syncwarp), B <-> C goes through a skewed shared tile under CTA barriers.
// Arithmetic = intfft_fast32.cuh (fly32 / cmul32): int_dif2_fly.vhd:142-373, int_dit2_fly.vhd:140-325,
// int_cmult_dsp48.vhd:182-190 / 307-317 and the dbl18 / dbl35 arrangements.
#include <type_traits>

#include "intfft_fast32.cuh"

namespace intfft {

namespace f32 {

__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

constexpr int kCtas13 = 3;
constexpr unsigned kSmem13 = kHead32 + 2 * kTile8 * 8;   // table + exchange tile + slot tile (74 KB: 3 CTAs / SM)

// KIND = multiplier-arrangement policy of STAGE 8..12, KLO = of STAGE 0..7 (UNSCALED plans grow past the
// single-DSP limit only in their late stages: c5u runs STAGE 2..9 on the cheaper single-arrangement code)
__device__ __forceinline__ unsigned pA_of(unsigned tid) { return phys8(16u * tid); }   // a thread's 16 contiguous tile slots

struct TwPacked32 {       // {re:16 | im:16} of a twiddle pre-shifted by 16: re << 16 and im << 16 are a mask and a shift away
    const int (&x)[15];
    __device__ __forceinline__ void operator()(int w, int &wr, int &wi) const
    {
        wr = (int)((unsigned)x[w] & 0xffff0000u);
        wi = (int)((unsigned)x[w] << 16);
    }
};
template <bool DIT, int MODE, int KIND, int KLO>
__global__ void __launch_bounds__(256, kCtas13) fast32_n13_kernel(const __grid_constant__ Fast32Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);                       // [15][16]
    // ONE exchange tile serves both ownership changes of a half: A <-> B stays inside a warp's 512-sample
    // region, and the B-side of B <-> C is in place (a thread writes / reads exactly the slots it read /
    // will write as the A <-> B partner), so the only cross-warp hazards are the two guarded by CTA barriers
    int2 *P = reinterpret_cast<int2 *>(smem_raw + kHead32);                       // A <-> B (warp-local)
    int2 *Q = P;                                                                  // B <-> C
    int2 *S = P + kTile8;                                                         // thread-private slots (see below)

    const unsigned tid = threadIdx.x;

    // ---- batch-invariant twiddles ----
    if (tid < 240) {                                   // round B: table[w][tid & 15], w = (1 << q) - 1 + j, STAGE 4 + q
        const int w = tid >> 4, lo4 = tid & 15;
        const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
        const int j = w - ((1 << q) - 1);
        midtw[w * 16 + lo4] = __ldg(p.tw + (1u << (4 + q)) + lo4 + ((unsigned)j << 4));
    }
    // round C: index tid + 256 j at STAGE 8 + q.  KIND_SINGLE_PRE (TWDL_WIDTH <= 16, twiddles pre-shifted by 16): both
    // halves of a twiddle share ONE register, {re:16 | im:16}, and are split where they are used (LOP3 + SHF on the
    // ALU port, which this multiply-bound kernel leaves half idle): 15 registers instead of 30, no spill to local memory
    constexpr bool PACKC = KIND == KIND_SINGLE_PRE;
    int uwr[15], uwi[PACKC ? 1 : 15];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < (1 << q); ++j) {
            const int2 w = __ldg(p.tw + (1u << (8 + q)) + tid + ((unsigned)j << 8));
            if (PACKC) {
                uwr[(1 << q) - 1 + j] = (int)(((unsigned)w.x & 0xffff0000u) | ((unsigned)w.y >> 16));
            } else {
                uwr[(1 << q) - 1 + j] = w.x;
                uwi[PACKC ? 0 : (1 << q) - 1 + j] = w.y;
            }
        }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }
    const int2 *twD = p.tw + (1u << 12) + tid;         // STAGE 12: index tid + 256 m (read through L1 / L2 per frame)
    // KIND_SINGLE_PRE: the same 4096 twiddles packed into 16 KB, which stays resident in the 28 KB of L1 that three
    // 74 KB CTAs leave (the 32 KB int2 table does not: every frame fetched it from L2 again, 8 % long-scoreboard stalls)
    const unsigned *twP = p.tw16 + tid;
    auto tw12 = [&](int m) {
        if (KIND == KIND_SINGLE_PRE) {
            const unsigned x = __ldg(twP + 256 * m);
            return make_int2((int)(x & 0xffff0000u), (int)(x << 16));       // re << 16, im << 16: already pre-shifted
        }
        return __ldg(twD + 256 * m);
    };
    __syncthreads();

    const unsigned pA = pA_of(tid);
    const unsigned pB = phys8((tid & 15u) | ((tid >> 4) << 8));
    const unsigned pC = phys8(tid);
    const Stg stD = stage_of<DIT, MODE, KIND>(p, 12);

    // DIT: a thread's first-round input is 16 contiguous samples.  Both tiles give every thread the eight
    // 16-byte slots [pA, pA + 16) (tile-skewed, so LDS.128 / STS.128 on them are conflict-free); the WARP
    // fetches its 512 samples as 512 contiguous bytes per cp.async instruction, piece k of the warp landing in
    // the slot of the lane that owns it (packed 16-bit input: slots 4..7, so that parking the round-C results
    // over slots 0..7 never overwrites a piece that is still unread).  Lower halves land in P, upper halves in S,
    // where they trade places with the lower half's round-C results (which wait there for STAGE 12).
    int4 *ownP = reinterpret_cast<int4 *>(P + pA_of(tid));
    int4 *ownS = reinterpret_cast<int4 *>(S + pA_of(tid));
    const unsigned lane = tid & 31u, w0 = (tid & ~31u) << 4;
    // piece k = lane + 32 j: phys8 is additive over the disjoint bit fields, so every address is base + immediate
    const unsigned land4 = phys8(w0 + 2u * lane);                                 // in_sb == 4: + 72 j  (int2 units)
    const unsigned land2 = phys8(w0 + 16u * (lane >> 2)) + 8u + 2u * (lane & 3u); // in_sb == 2: + 144 j, slots 4..7
    auto prefetch = [&](int2 *tile_base, long long first_sample) {
        const char *src = reinterpret_cast<const char *>(p.in) + (first_sample + w0) * (2 * p.in_sb) + 16u * lane;
        if (p.in_sb == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) cp_async_16(tile_base + land4 + 72 * j, src + 512 * j);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async_16(tile_base + land2 + 144 * j, src + 512 * j);
        }
        cp_async_commit();
    };
    // frame counters are 32-bit (at most 2^27 frames): the kernel runs at its register cap, 64-bit ones cost spills
    const unsigned n_frames = (unsigned)p.n_tiles;
    if (DIT && blockIdx.x < n_frames) prefetch(P, (long long)blockIdx.x << 13);

    for (unsigned tile = blockIdx.x; tile < n_frames; tile += gridDim.x) {
        const long long g0 = (long long)tile << 13;
        V re[16], im[16];

        if (!DIT) {
            // ---- STAGE 12 first: lower-half results stay in registers, upper-half results wait in S ----
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                int ar, ai, br, bi;
                ld_sample(p.in, g0 + tid + 256u * m, p.in_sb, ar, ai);
                ld_sample(p.in, g0 + 4096 + tid + 256u * m, p.in_sb, br, bi);
                const int2 w = tw12(m);
                V xr = mk(sx(ar, p.dw)), xi = mk(sx(ai, p.dw)), yr = mk(sx(br, p.dw)), yi = mk(sx(bi, p.dw));
                fly32<DIT, MODE, KIND>(stD, false, p.cm, xr, xi, yr, yi, w.x, w.y);
                re[m] = xr;
                im[m] = xi;
                S[m * 256 + tid] = make_int2(yr.f, yi.f);
            }
        }

        if (DIT) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {                 // a real loop: keeps the body inside the 32 KB L1.5 I-cache
                // ---- this half's 16 contiguous samples were prefetched (cp.async) into private 16-byte slots:
                // ---- lower half -> this warp's region of P, upper half -> S, where they now trade places with
                // ---- the lower half's round-C results (which wait there for STAGE 12)
                cp_async_wait_all();
                __syncwarp();                              // the other lanes' pieces of this warp's region have landed too
                if (p.in_sb == 4) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int4 v;
                        if (h == 0) {
                            v = ownP[j];
                        } else {
                            v = ownS[j];
                            ownS[j] = make_int4(re[2 * j].f, im[2 * j].f, re[2 * j + 1].f, im[2 * j + 1].f);
                        }
                        re[2 * j] = mk(sx(v.x, p.dw)); im[2 * j] = mk(sx(v.y, p.dw));
                        re[2 * j + 1] = mk(sx(v.z, p.dw)); im[2 * j + 1] = mk(sx(v.w, p.dw));
                    }
                } else {                                   // packed 16-bit input: 4 pieces, parked in slots 4..7
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        int4 v;
                        if (h == 0) {
                            v = ownP[4 + j];
                        } else {
                            v = ownS[4 + j];
                            ownS[2 * j] = make_int4(re[4 * j].f, im[4 * j].f, re[4 * j + 1].f, im[4 * j + 1].f);
                            ownS[2 * j + 1] = make_int4(re[4 * j + 2].f, im[4 * j + 2].f, re[4 * j + 3].f, im[4 * j + 3].f);
                        }
                        const int x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            re[4 * j + e] = mk(sx(x[e], p.dw));
                            im[4 * j + e] = mk(sx(x[e] >> 16, p.dw));
                        }
                    }
                }
                if (h == 0) prefetch(S, g0 + 4096);         // every lane left the previous frame's STAGE 12 before the __syncwarp above
                round32<4, DIT, MODE, KLO>(re, im, p, 0, TwRegs32{lwr, lwi}, true, false);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; ++m) P[pA + m] = make_int2(re[m].f, im[m].f);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; ++m) { const int2 v = P[pB + 18u * m]; re[m] = mk(v.x); im[m] = mk(v.y); }
                round32<4, DIT, MODE, KLO>(re, im, p, 4, TwSmem32{midtw + (tid & 15u), 16}, false, false);
#pragma unroll
                for (int m = 0; m < 16; ++m) Q[pB + 18u * m] = make_int2(re[m].f, im[m].f);
                __syncthreads();
#pragma unroll
                for (int m = 0; m < 16; ++m) { const int2 v = Q[pC + 288u * m]; re[m] = mk(v.x); im[m] = mk(v.y); }
                __syncthreads();                           // the tile may be rewritten once every thread has read it
                // the tile is idle until the next frame's A -> B change: land that frame's lower half in it
                if (h == 1 && tile + gridDim.x < n_frames) prefetch(P, (long long)(tile + gridDim.x) << 13);
                if constexpr (PACKC) round32<4, DIT, MODE, KIND>(re, im, p, 8, TwPacked32{uwr}, false, false);
                else round32<4, DIT, MODE, KIND>(re, im, p, 8, TwRegs32{uwr, uwi}, false, false);
            }
            // ---- STAGE 12 between the parked lower half and the registers; coalesced stores ----
            auto stage12 = [&](auto out_sb_tag) {       // the container size is tested once per frame, not per store
                constexpr int OSB = decltype(out_sb_tag)::value;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int4 a = ownS[j];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int m = 2 * j + e;
                        const int2 w = tw12(m);
                        V xr = mk(e ? a.z : a.x), xi = mk(e ? a.w : a.y);
                        fly32<DIT, MODE, KIND>(stD, false, p.cm, xr, xi, re[m], im[m], w.x, w.y);
                        st_sample(p.out, g0 + tid + 256u * m, OSB, xr.f, xi.f);
                        st_sample(p.out, g0 + 4096 + tid + 256u * m, OSB, re[m].f, im[m].f);
                    }
                }
            };
            if (p.out_sb == 4) stage12(std::integral_constant<int, 4>{});
            else stage12(std::integral_constant<int, 2>{});
        } else {
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const long long gh = g0 + 4096 * h;
            {
                if (h == 1) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) { const int2 v = S[m * 256 + tid]; re[m] = mk(v.x); im[m] = mk(v.y); }
                }
                if constexpr (PACKC) round32<4, DIT, MODE, KIND>(re, im, p, 8, TwPacked32{uwr}, false, false);
                else round32<4, DIT, MODE, KIND>(re, im, p, 8, TwRegs32{uwr, uwi}, false, false);
                __syncthreads();                           // every warp is done with the previous half's A-side reads
#pragma unroll
                for (int m = 0; m < 16; ++m) Q[pC + 288u * m] = make_int2(re[m].f, im[m].f);
                __syncthreads();
#pragma unroll
                for (int m = 0; m < 16; ++m) { const int2 v = Q[pB + 18u * m]; re[m] = mk(v.x); im[m] = mk(v.y); }
                round32<4, DIT, MODE, KLO>(re, im, p, 4, TwSmem32{midtw + (tid & 15u), 16}, false, false);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; ++m) P[pB + 18u * m] = make_int2(re[m].f, im[m].f);
                __syncwarp();
#pragma unroll
                for (int m = 0; m < 16; ++m) { const int2 v = P[pA + m]; re[m] = mk(v.x); im[m] = mk(v.y); }
                round32<4, DIT, MODE, KLO>(re, im, p, 0, TwRegs32{lwr, lwi}, true, false);
                if (p.out_sb == 4) {
                    // 16 contiguous samples (128 bytes) per thread: back into the thread's own tile slots, then
                    // the warp (which owns these 512 samples) writes 512 contiguous bytes per instruction
#pragma unroll
                    for (int m = 0; m < 16; ++m) P[pA + m] = make_int2(re[m].f, im[m].f);
                    __syncwarp();
                    const unsigned w0 = (tid & ~31u) << 4;
                    int4 *dst = reinterpret_cast<int4 *>(reinterpret_cast<int2 *>(p.out) + gh + w0);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[lane + 32 * j] = *reinterpret_cast<const int4 *>(P + phys8(w0 + 2u * lane + 64u * j));
                } else {
                    uint4 *dst = reinterpret_cast<uint4 *>(p.out) + ((gh + 16u * tid) >> 2);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        dst[j] = make_uint4(__byte_perm((unsigned)re[4 * j].f, (unsigned)im[4 * j].f, 0x5410),
                                            __byte_perm((unsigned)re[4 * j + 1].f, (unsigned)im[4 * j + 1].f, 0x5410),
                                            __byte_perm((unsigned)re[4 * j + 2].f, (unsigned)im[4 * j + 2].f, 0x5410),
                                            __byte_perm((unsigned)re[4 * j + 3].f, (unsigned)im[4 * j + 3].f, 0x5410));
                }
            }
        }
        }
    }
}

template <typename K> cudaError_t launch_n13(K k, const Fast32Params &p, int grid, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem13);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, kSmem13, st>>>(p);
    return cudaGetLastError();
}

template <bool DIT> cudaError_t launch_n13_dir(const Fast32Params &p, int mode, int kind, int klo, int grid, cudaStream_t st)
{
    constexpr int S = KIND_SINGLE, M = KIND_MIXED;
    // DIT, TRUNCATE, every stage single-DSP (BASELINE c5): the pre-shifted-twiddle instance; the caller has put the
    // pre-shifted table / lowest-round twiddles into p (intfft_fast32_strided.cu: launch_fast32)
    if (kind == KIND_SINGLE_PRE) return launch_n13(fast32_n13_kernel<DIT, MODE_TRUNC, KIND_SINGLE_PRE, KIND_SINGLE_PRE>, p, grid, st);
    switch (mode * 4 + kind * 2 + klo) {
    case MODE_TRUNC * 4 + 0: return launch_n13(fast32_n13_kernel<DIT, MODE_TRUNC, S, S>, p, grid, st);
    case MODE_TRUNC * 4 + 2: return launch_n13(fast32_n13_kernel<DIT, MODE_TRUNC, M, S>, p, grid, st);
    case MODE_TRUNC * 4 + 3: return launch_n13(fast32_n13_kernel<DIT, MODE_TRUNC, M, M>, p, grid, st);
    case MODE_ROUND * 4 + 0: return launch_n13(fast32_n13_kernel<DIT, MODE_ROUND, S, S>, p, grid, st);
    case MODE_ROUND * 4 + 2: return launch_n13(fast32_n13_kernel<DIT, MODE_ROUND, M, S>, p, grid, st);
    case MODE_ROUND * 4 + 3: return launch_n13(fast32_n13_kernel<DIT, MODE_ROUND, M, M>, p, grid, st);
    case MODE_UNSCALED * 4 + 0: return launch_n13(fast32_n13_kernel<DIT, MODE_UNSCALED, S, S>, p, grid, st);
    case MODE_UNSCALED * 4 + 2: return launch_n13(fast32_n13_kernel<DIT, MODE_UNSCALED, M, S>, p, grid, st);
    case MODE_UNSCALED * 4 + 3: return launch_n13(fast32_n13_kernel<DIT, MODE_UNSCALED, M, M>, p, grid, st);
    default: return cudaErrorInvalidValue;          // single late stages with double early ones: fall back below
    }
}

}  // namespace f32

// kind / klo: arrangement policy of STAGE 8..12 / STAGE 0..7 (KIND_SINGLE or KIND_MIXED)
int f32_launch_n13(const f32::Fast32Params &p, bool dit, int mode, int kind, int klo, int grid, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (klo == f32::KIND_MIXED && kind != f32::KIND_SINGLE_PRE) kind = f32::KIND_MIXED;     // (single, mixed) is not instantiated: mixed covers it
    return (int)(dit ? f32::launch_n13_dir<true>(p, mode, kind, klo, grid, st) : f32::launch_n13_dir<false>(p, mode, kind, klo, grid, st));
}

}  // namespace intfft
