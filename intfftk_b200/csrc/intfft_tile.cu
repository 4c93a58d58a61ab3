// Generic fused stage-chain kernel: one CTA keeps a tile of 2^L complex samples in shared memory
// and runs every butterfly stage of the pass on it in register rounds of up to 4 stages
// (16 samples per thread), exchanging between rounds through XOR-swizzled shared memory.
//
// Stands in for the xCALC / xDELAYS generate loops of int_fftNk (src/vhdl/fft/int_fftNk.vhd:184-331)
// and int_ifftNk (src/vhdl/fft/int_ifftNk.vhd:183-330): the butterfly + twiddle ROM per stage, with
// the delay-line cross-commutation (src/vhdl/delay/int_delay_line.vhd:52-104) turned into index
// arithmetic — stage STAGE = s pairs samples whose in-place index differs in bit s, and the twiddle
// index is the in-place index modulo 2^s.
//
// Works for every generic combination that elaborates; the hot 16-bit scaled configuration has a
// specialised kernel in intfft_fast16.cu.
#include <cuda_runtime.h>

#include "intfft_arith.cuh"

namespace intfft {

namespace {

template <typename T> struct Elem;
template <> struct Elem<int32_t> { using type = int2; };
template <> struct Elem<int64_t> { using type = longlong2; };

// element index -> swizzled element index; linear over XOR, so swz(a | b) = swz(a) ^ swz(b) for
// disjoint a, b.  Keeps the three access patterns of a 4-bit round (lanes on the low bits, on the
// low bits with a hole, on bits 4..8) free of bank conflicts for 4/8/16-byte elements.
__device__ __forceinline__ unsigned swz(unsigned i) { return i ^ ((i >> 4) & 31u); }

template <typename T>
__device__ __forceinline__ void load_scalar_pair(const void *base, long long idx, int sb, T &re, T &im)
{
    if (sb == 2) {
        const short2 v = reinterpret_cast<const short2 *>(base)[idx];
        re = v.x; im = v.y;
    } else if (sb == 4) {
        const int2 v = reinterpret_cast<const int2 *>(base)[idx];
        re = v.x; im = v.y;
    } else {
        const longlong2 v = reinterpret_cast<const longlong2 *>(base)[idx];
        re = (T)v.x; im = (T)v.y;
    }
}

template <typename T>
__device__ __forceinline__ void store_scalar_pair(void *base, long long idx, int sb, T re, T im)
{
    if (sb == 2) {
        reinterpret_cast<short2 *>(base)[idx] = make_short2((short)re, (short)im);
    } else if (sb == 4) {
        reinterpret_cast<int2 *>(base)[idx] = make_int2((int)re, (int)im);
    } else {
        reinterpret_cast<longlong2 *>(base)[idx] = make_longlong2((long long)re, (long long)im);
    }
}

// Up to 4 stage bits of one round on the 16 register-resident samples of a thread.
// Register index bit q (< R) is local tile bit lo + q; bits >= R enumerate independent groups.
// One code body serves R = 1..4: steps for q >= R are skipped by a grid-uniform branch.
// The twiddles of a step are fetched (read-only path) before its butterflies so the loads overlap.
template <typename T, typename P, int MODE, bool DIT>
__device__ __forceinline__ void run_round(T (&re)[16], T (&im)[16], const PassParams &p, int lo, int R,
                                          unsigned gbase, const unsigned (&goff)[16])
{
#pragma unroll
    for (int step = 0; step < 4; ++step) {
        const int q = DIT ? step : 3 - step;
        if (q >= R) continue;
        const int s = p.pb + (lo + q - p.c);          // global bit == butterfly STAGE
        const StageInfo si = stage_info<DIT>(p, s);
        const unsigned kmask = (1u << s) - 1u;
        const unsigned kb = gbase & kmask;
        const int2 *tw = p.tw + (1u << s);
        int2 w[8];
        unsigned kk[8];
#pragma unroll
        for (int m = 0, j = 0; m < 16; ++m) {
            if (m & (1 << q)) continue;
            kk[j] = kb | (goff[m] & kmask);
            w[j] = (s >= 2) ? __ldg(tw + kk[j]) : make_int2(0, 0);
            ++j;
        }
#pragma unroll
        for (int m = 0, j = 0; m < 16; ++m) {
            if (m & (1 << q)) continue;
            const int mb = m | (1 << q);
            butterfly<T, P, MODE, DIT>(re[m], im[m], re[mb], im[mb], si, p.cm, kk[j], w[j]);
            ++j;
        }
    }
}

template <typename T, typename P, int MODE, bool DIT>
__global__ void __launch_bounds__(512) tile_kernel(const __grid_constant__ PassParams p)
{
    using E = typename Elem<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    E *tile = reinterpret_cast<E *>(smem_raw);

    const unsigned tid = threadIdx.x;
    const int tbits = p.L - 4;                        // log2(threads)
    const unsigned cmask = (1u << p.c) - 1u;
    const bool strided = p.c > 0;

    for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        // ---- where this tile lives in the flat stream ----
        long long origin;
        if (!strided) {
            origin = t << p.L;
        } else {
            const int mid_bits = p.pb - p.c, hi_bits = p.n - p.pb - p.g;
            const long long mid = t & ((1ll << mid_bits) - 1);
            const long long rest = t >> mid_bits;
            const long long hi = rest & ((1ll << hi_bits) - 1);
            const long long frame = rest >> hi_bits;
            origin = (frame << p.n) + (hi << (p.pb + p.g)) + (mid << p.c);
        }

        // ---- global -> shared (coalesced) ----
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const unsigned l = tid | ((unsigned)e << tbits);
            const long long gi = origin + ((long long)(l >> p.c) << p.pb) + (l & cmask);
            T re = 0, im = 0;
            if (gi < p.total) {
                load_scalar_pair<T>(p.in, gi, p.in_sb, re, im);
                if (p.in_wrap) { re = wrapw<T>(re, p.dw); im = wrapw<T>(im, p.dw); }
            }
            E v; v.x = re; v.y = im;
            tile[swz(l)] = v;
        }
        __syncthreads();

        // ---- register rounds ----
        for (int r = 0; r < p.nrounds; ++r) {
            const int lo = p.r_lo[r], R = p.r_n[r];
            const unsigned base = (tid & ((1u << lo) - 1u)) | ((tid >> lo) << (lo + R));
            const unsigned pbase = swz(base);
            const unsigned gbase = (unsigned)(((base >> p.c) << p.pb) + (base & cmask)) +
                                   (unsigned)(origin & ((1ll << p.n) - 1));
            unsigned poff[16], goff[16];
            T re[16], im[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (tbits + R));
                poff[m] = swz(off);
                goff[m] = ((off >> p.c) << p.pb) + (off & cmask);
                const E v = tile[pbase ^ poff[m]];
                re[m] = v.x; im[m] = v.y;
            }
            run_round<T, P, MODE, DIT>(re, im, p, lo, R, gbase, goff);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                E v; v.x = re[m]; v.y = im[m];
                tile[pbase ^ poff[m]] = v;
            }
            __syncthreads();
        }

        // ---- shared -> global (coalesced) ----
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const unsigned l = tid | ((unsigned)e << tbits);
            const long long gi = origin + ((long long)(l >> p.c) << p.pb) + (l & cmask);
            if (gi < p.total) {
                const E v = tile[swz(l)];
                store_scalar_pair<T>(p.out, gi, p.out_sb, (T)v.x, (T)v.y);
            }
        }
        __syncthreads();
    }
}

template <typename T, typename P>
cudaError_t launch_t(const PassDesc &pd, int mode, bool dit, int grid, cudaStream_t st)
{
    using K = void (*)(const PassParams);
    K k = nullptr;
    switch (mode * 2 + (dit ? 1 : 0)) {
    case 0: k = tile_kernel<T, P, MODE_TRUNC, false>; break;
    case 1: k = tile_kernel<T, P, MODE_TRUNC, true>; break;
    case 2: k = tile_kernel<T, P, MODE_ROUND, false>; break;
    case 3: k = tile_kernel<T, P, MODE_ROUND, true>; break;
    case 4: k = tile_kernel<T, P, MODE_UNSCALED, false>; break;
    default: k = tile_kernel<T, P, MODE_UNSCALED, true>; break;
    }
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pd.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, pd.threads, pd.smem_bytes, st>>>(pd.kp);
    return cudaGetLastError();
}

}  // namespace

int launch_tile_pass(const PassDesc &pd, int mode, bool dit, int num_sms, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // persistent-style grid: a few CTAs per SM, each striding over tiles
    const int per_sm = pd.smem_bytes > 100 * 1024 ? 1 : (pd.smem_bytes > 48 * 1024 ? 2 : 4);
    long long grid = (long long)num_sms * per_sm;
    if (grid > pd.kp.n_tiles) grid = pd.kp.n_tiles;
    if (grid < 1) grid = 1;
    cudaError_t e;
    switch (pd.lane) {
    case LANE_I32_P64: e = launch_t<int32_t, int64_t>(pd, mode, dit, (int)grid, st); break;
    case LANE_I64_P64: e = launch_t<int64_t, int64_t>(pd, mode, dit, (int)grid, st); break;
    default: e = launch_t<int64_t, __int128>(pd, mode, dit, (int)grid, st); break;
    }
    count_launch();
    return (int)e;
}

}  // namespace intfft
