// Specialised stage-chain kernel for the headline configuration class: SCALED / TRUNCATE mode,
// DATA_WIDTH <= 16, TWDL_WIDTH <= 16 (BASELINE c2: 4096-pt 16-bit scaled DIF), N = 2^8 .. 2^12.
//
// Same arithmetic as the generic kernel (intfft_tile.cu / intfft_arith.cuh) and the same reference
// rules (int_dif2_fly.vhd:144-164, 245-318, 348-366; int_dit2_fly.vhd:142-162, 221-322;
// int_cmult_dsp48.vhd:184-190 single-DSP slice), restructured for the integer issue ports:
//   * samples travel as packed {re:16, im:16} words (4 B / sample in HBM and in shared memory);
//   * all products are 32-bit: twiddles are pre-shifted by e = 33 - TWDL_WIDTH - DATA_WIDTH so that
//     wrap_DW((P2 +- P1) >> (TWDL_WIDTH-1)) is ONE arithmetic shift of the low 32 bits of the sum
//     of products (the slice P(DTW+TWD-2 downto TWD-1) then ends exactly at bit 31);
//   * TRUNCATE only ever consumes A>>1 and B>>1, so (A>>1)+(B>>1) is a shift-add (LEA.HI.SX32) and
//     (A>>1)-(B>>1) = sum - 2*(B>>1) goes to the multiply-add port (IMAD), balancing the two ports;
//   * each thread keeps 16 samples in registers for 4 consecutive stages; the twiddles of the two
//     upper rounds are per-thread constants hoisted out of the persistent frame loop, those of the
//     lowest round are kernel parameters (constant bank);
//   * DIF frames are staged HBM -> shared memory by the TMA engine (cp.async.bulk + mbarrier), one
//     frame ahead of the arithmetic;
//   * global memory is touched per WARP, never per thread: the lowest round's results (DIF) go back
//     into the thread's own tile slots and the warp stores its runs as 512 contiguous bytes per
//     instruction; the lowest round's input (DIT) is fetched by the warp with cp.async into a landing
//     tile one frame ahead (per-thread 64-byte accesses cost 7 % of the whole kernel at c2);
//   * rounds exchange through a double-buffered padded tile whose skew is additive, so every
//     shared-memory address is "per-round base register + compile-time immediate" and every access
//     pattern (32-bit at stride 256 / 16, 128-bit contiguous) is bank-conflict free for N = 4096.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "intfft_internal.h"

#ifndef FAST16_PART
#error "compile with -DFAST16_PART=1 and -DFAST16_PART=2 (see the Makefile)"
#endif
#include "intfft_taylor.cuh"
#include "intfft_tma.cuh"

namespace intfft {

namespace {

struct Fast16Params {
    const uint32_t *in;
    uint32_t *out;
    const int2 *twp;        // pre-shifted twiddles, entry (1 << s) + k
    long long n_tiles;      // tiles of 4096 samples
    long long total;        // batch * N samples
    int dw, sh_full, sh_half;
    int lw_r[16], lw_i[16]; // lowest-round twiddles: index (1 << s) - 1 + k, s = 2, 3
};

// sample index inside a 4096-sample tile -> word offset in the padded exchange tile.
// Additive skew: 4 words per 32 samples + 16 words per 256 samples.  For disjoint bit sets
// phys(a | b) = phys(a) + phys(b), so per-register offsets fold into instruction immediates.
__host__ __device__ constexpr unsigned phys(unsigned i) { return i + 4u * (i >> 5) + 16u * (i >> 8); }
constexpr unsigned kTileWords = 4864;   // >= phys(4095) + 1, multiple of 4
constexpr unsigned kSmemHead = 128 + 15 * 16 * 8;   // barriers + middle-round twiddle table

// ---- instruction-selection helpers (inline PTX keeps the front end from re-deriving 16-bit
// ---- value ranges and narrowing the arithmetic, and pins which issue port an op lands on) ----
__device__ __forceinline__ int sext_lo16(uint32_t x) { int r; asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(x)); return r; }
template <int S> __device__ __forceinline__ int sra(int x) { int r; asm("shr.s32 %0, %1, %2;" : "=r"(r) : "r"(x), "n"(S)); return r; }
__device__ __forceinline__ int msub2(int t, int x) { int r; asm("mad.lo.s32 %0, %1, -2, %2;" : "=r"(r) : "r"(t), "r"(x)); return r; }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// 16-byte piece with only the first `bytes` (0 or 16) read from global memory, the rest zero-filled
__device__ __forceinline__ void cp_async_16z(void *smem_dst, const void *gsrc, unsigned bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_group0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Only a pass's first round can meet values outside DATA_WIDTH (conv_std_logic_vector wrap of the stimulus):
// everything a later round reads from the exchange tile was produced here, already wrapped to DATA_WIDTH
// bits and sign-extended to 16, so those rounds take the cheap 16-bit unpack whatever DATA_WIDTH is.
template <bool DW16>
__device__ __forceinline__ void unpack(uint32_t x, int dw, int &re, int &im)
{
    if (DW16) {
        re = sext_lo16(x);
        im = sra<16>((int)x);
    } else {                                   // wrap to DATA_WIDTH bits (conv_std_logic_vector): one SGXT each
        asm("szext.clamp.s32 %0, %1, %2;" : "=r"(re) : "r"((int)x), "r"(dw));
        asm("szext.clamp.s32 %0, %1, %2;" : "=r"(im) : "r"((int)(x >> 16)), "r"(dw));
    }
}

__device__ __forceinline__ uint32_t pack(int re, int im) { return __byte_perm((unsigned)re, (unsigned)im, 0x5410); }

// -v for v >= 0, ~v for v < 0  (int_dif2_fly.vhd:299-303)
__device__ __forceinline__ int negq(int v) { return (v >> 31) - v; }

// RAWY (DIF, DW16 only): leave Y as the raw 32-bit sum of products; its value is raw >> 16, which the
// packer takes straight from the upper half-words (PRMT 0x7632) instead of two shifts + PRMT.
// ROUNDING helpers: (v + 1) >> 1 == (v >> 1) + v(0); the rounded DIFFERENCE can reach 2^(DW-1) and is kept
// in DW bits by the reference (int_dif2_fly.vhd:201-216), hence the sign-extension from bit DW-1
template <bool DW16> __device__ __forceinline__ int rnd_sum(int a, int b) { return sra<1>(a + b + 1); }
// rounded difference from the rounded sum: (a - b + 1) >> 1 == ((a + b + 1) >> 1) - b exactly (b is an integer
// inside the floor), which is one multiply-add-port instruction instead of an add and a shift on the ALU port
template <bool DW16> __device__ __forceinline__ int rnd_dif(int sum, int b, int sh_full)
{
    int d;
    asm("mad.lo.s32 %0, %1, -1, %2;" : "=r"(d) : "r"(b), "r"(sum));
    if (DW16) return sext_lo16((uint32_t)d);
    int r;
    asm("szext.clamp.s32 %0, %1, %2;" : "=r"(r) : "r"(d), "r"(32 - sh_full));      // sh_full = 32 - DATA_WIDTH
    return r;
}

template <bool DIT, bool DW16, int MODE, bool RAWY = false>
__device__ __forceinline__ void fly(int s, bool odd, int &ar, int &ai, int &br, int &bi, int wr, int wi,
                                    int sh_full, int sh_half)
{
    if (!DIT) {
        int xr, xi, sr, si;
        if (MODE == MODE_TRUNC) {
            const int tr = sra<1>(br), ti = sra<1>(bi);
            xr = sra<1>(ar) + tr;
            xi = sra<1>(ai) + ti;
            sr = msub2(tr, xr);                                  // (A>>1) - (B>>1)
            si = msub2(ti, xi);
        } else {
            xr = rnd_sum<DW16>(ar, br);
            xi = rnd_sum<DW16>(ai, bi);
            sr = rnd_dif<DW16>(xr, br, sh_full);
            si = rnd_dif<DW16>(xi, bi, sh_full);
        }
        ar = xr;
        ai = xi;
        if (s == 0) {
            br = sr;
            bi = si;
        } else if (s == 1) {
            br = odd ? si : sr;
            bi = odd ? negq(sr) : si;
        } else {
            const int pr = (int)((unsigned)sr * (unsigned)wr - (unsigned)si * (unsigned)wi);
            const int pi = (int)((unsigned)sr * (unsigned)wi + (unsigned)si * (unsigned)wr);
            br = RAWY ? pr : (DW16 ? sra<16>(pr) : (pr >> sh_full));
            bi = RAWY ? pi : (DW16 ? sra<16>(pi) : (pi >> sh_full));
        }
    } else if (MODE == MODE_TRUNC) {
        int hr, hi;                                             // BW >> 1
        if (s == 0) {
            hr = sra<1>(br);
            hi = sra<1>(bi);
        } else if (s == 1) {
            hr = sra<1>(odd ? negq(bi) : br);
            hi = sra<1>(odd ? br : bi);
        } else {                                                // multiplier fed with swapped re / im
            const int o_re = (int)((unsigned)bi * (unsigned)wr - (unsigned)br * (unsigned)wi);
            const int o_im = (int)((unsigned)bi * (unsigned)wi + (unsigned)br * (unsigned)wr);
            hi = DW16 ? sra<17>(o_re) : (o_re >> sh_half);
            hr = DW16 ? sra<17>(o_im) : (o_im >> sh_half);
        }
        const int xr = sra<1>(ar) + hr, xi = sra<1>(ai) + hi;
        br = msub2(hr, xr);
        bi = msub2(hi, xi);
        ar = xr;
        ai = xi;
    } else {                                                    // ROUNDING, DIT
        int wr_, wi_;                                           // BW
        if (s == 0) {
            wr_ = br;
            wi_ = bi;
        } else if (s == 1) {
            wr_ = odd ? negq(bi) : br;
            wi_ = odd ? br : bi;
        } else {
            const int o_re = (int)((unsigned)bi * (unsigned)wr - (unsigned)br * (unsigned)wi);
            const int o_im = (int)((unsigned)bi * (unsigned)wi + (unsigned)br * (unsigned)wr);
            wi_ = DW16 ? sra<16>(o_re) : (o_re >> sh_full);
            wr_ = DW16 ? sra<16>(o_im) : (o_im >> sh_full);
        }
        const int xr = rnd_sum<DW16>(ar, wr_), xi = rnd_sum<DW16>(ai, wi_);
        br = rnd_dif<DW16>(xr, wr_, sh_full);
        bi = rnd_dif<DW16>(xi, wi_, sh_full);
        ar = xr;
        ai = xi;
    }
}

// R stages (global bits LO .. LO+R-1) on the 16 register-resident samples
template <int LO, int R, bool DIT, bool DW16, int MODE, bool RAWLAST, typename TW>
__device__ __forceinline__ void round_regs(int (&re)[16], int (&im)[16], const TW &tw, bool tid_odd, int sh_full,
                                           int sh_half)
{
#pragma unroll
    for (int step = 0; step < R; ++step) {
        const int q = DIT ? step : R - 1 - step;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (m & (1 << q)) continue;
            const int j = m & ((1 << q) - 1);
            const int w = (1 << q) - 1 + j;
            const bool odd = (LO == 0) ? ((m & 1) != 0) : tid_odd;    // twiddle index bit 0 (STAGE = 1 only)
            int wr = 0, wi = 0;
            if (LO + q >= 2) tw(w, wr, wi);
            if (RAWLAST && step == R - 1)
                fly<DIT, DW16, MODE, true>(LO + q, odd, re[m], im[m], re[m | (1 << q)], im[m | (1 << q)], wr, wi, sh_full, sh_half);
            else
                fly<DIT, DW16, MODE>(LO + q, odd, re[m], im[m], re[m | (1 << q)], im[m | (1 << q)], wr, wi, sh_full, sh_half);
        }
    }
}

// twiddle sources for round_regs
struct TwRegs {
    const int (&r)[15];
    const int (&i)[15];
    __device__ __forceinline__ void operator()(int w, int &wr, int &wi) const { wr = r[w]; wi = i[w]; }
};
template <int PITCH> struct TwSmemP {   // table[w][low bits of tid] of pre-shifted (re, im); one LDS.64 per butterfly
    const int2 *t;
    __device__ __forceinline__ void operator()(int w, int &wr, int &wi) const
    {
        const int2 v = t[w * PITCH];
        wr = v.x;
        wi = v.y;
    }
};
using TwSmem = TwSmemP<16>;

// MIDSM: keep the middle round's 15 twiddles in a 1920-byte shared table instead of 30 registers, which
// brings the kernel under 85 registers so that three CTAs (24 warps) fit on one SM.
// NAT (4096-point DIF only): int_bitrev_order (buffers/int_bitrev_order.vhd:82-104) fused into the output.
// After the last round a thread holds in-place positions q = 16 tid + m; their natural positions
// rev12(q) = rev4(m) << 8 | rev8(tid) are scattered into the tile's (by then dead) TMA landing buffer,
// whose 16-byte chunks are XOR-swizzled by address bits 5..7 so that the scatter is 4-way instead of
// 8-way bank-conflicted, and after one CTA barrier the frame leaves as 16-byte stores in natural order.
// The landing buffer of the NEXT tile doubles as the previous tile's scatter buffer, so in this variant
// the TMA prefetch is issued after the tile's first CTA barrier (every thread has then left the
// previous tile) instead of at the top of the tile.
// PAIR (f2, int_fft_ifft_pair, main/int_fft_ifft_pair.vhd:209-283): int_fftNk and int_ifftNk in ONE kernel.  The
// FFT's last round leaves a thread with 16 contiguous samples of the bit-reversed spectrum — exactly what the IFFT's
// first round owns — so the spectrum never leaves the registers: the DIF chain runs without its output side, the
// DIT chain without its input side (same twiddle tables: W_s[k] does not depend on the direction).
// DWC: DATA_WIDTH as a compile-time constant for the common converter widths below 16 bits (12, 14; 0 = run-time).  With the
// shift amounts immediates, the product slice and the TRUNCATE halving fuse into one shift / one LEA.HI.SX32 exactly as in
// the 16-bit instance (with run-time amounts they stay two shifts and an add per operand: 0.40 of the roofline against 0.50).
template <int NLOG2, bool DIT, bool DW16, bool MIDSM, int MODE, bool NAT = false, bool PAIR = false, int DWC = 0>
__global__ void __launch_bounds__(256, MIDSM ? 3 : 2) fast16_kernel(const __grid_constant__ Fast16Params p)
{
    static_assert(DWC == 0 || !DW16, "DWC is for DATA_WIDTH < 16");
    static_assert(!NAT || (NLOG2 == 12 && !DIT), "fused natural-order output: 4096-point DIF only");
    static_assert(!PAIR || (!DIT && !NAT && NLOG2 >= 8), "pair: instantiated as the DIF kernel, 2^8 .. 2^12 points");
    constexpr int R0 = ((NLOG2 - 1) % 4) + 1;      // stages in the lowest round
    constexpr int NR = 1 + (NLOG2 - R0) / 4;       // rounds; round r > 0 covers bits R0+4(r-1) .. +3
    static_assert(NR >= 1 && NR <= 3, "supported: 2^3 .. 2^12 points");
    constexpr int NU = NR > 1 ? NR - 1 : 1;        // rounds above the lowest one (array extent, at least 1)
    constexpr bool TMA_IN = !DIT;                  // DIF: first round reads stride-256 words -> stage via TMA
    // DIF, 4-stage last round: results go back into the thread's own tile slots and leave warp-coalesced
    // (stored directly, a warp instruction would write 32 separate 16-byte pieces at a 64-byte pitch)
    // (COALESCE is decided per chain below: (!DIT || NR == 1) && !NAT — a one-round DIT ends in the lowest round as well)
    // DIT, 4-stage first round: a thread needs its own 16 contiguous samples (64 bytes).  Loaded directly, a
    // warp instruction would touch 32 separate 16-byte pieces at a 64-byte pitch, so the WARP fetches its
    // 2 KB as 512 contiguous bytes per cp.async instruction into a skewed landing tile, one tile ahead.
    constexpr bool CP_IN = DIT;
    static_assert(!MIDSM || NR == 3, "MIDSM is for the three-round schedules (R0 + 4 + 4 stages)");
    // dynamic shared memory: [bar 2 x u64 | pad to 128] [mid twiddles 15 x 16 x int2] [work 2 x kTileWords]
    //                        [stage 2 x 4096 (DIF only)]
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    uint32_t(*work)[kTileWords] = reinterpret_cast<uint32_t(*)[kTileWords]>(smem_raw + kSmemHead);
    uint32_t(*stage)[4096] = reinterpret_cast<uint32_t(*)[4096]>(smem_raw + kSmemHead + 2 * kTileWords * 4);
    uint32_t *land = reinterpret_cast<uint32_t *>(smem_raw + kSmemHead + 2 * kTileWords * 4);     // CP_IN

    const unsigned tid = threadIdx.x;
    const int dwv = DWC ? DWC : p.dw;              // DATA_WIDTH, sh_full = 32 - DATA_WIDTH, sh_half = 33 - DATA_WIDTH
    const int sh_full = DWC ? 32 - DWC : p.sh_full, sh_half = DWC ? 33 - DWC : p.sh_half;
    const bool tid_odd = tid & 1u;
    // In the lowest round a warp owns 16 >> R0 runs of 32 << R0 contiguous samples (a thread: 1 << R0 contiguous
    // samples per run).  16-byte piece q = lane + 32 c of the warp's 128 -> first sample inside the tile:
    auto warp_piece = [&](int c) {
        const unsigned q = (tid & 31u) + 32u * c;
        return ((tid & ~31u) << R0) + ((q >> (3 + R0)) << (8 + R0)) + 4u * (q & ((8u << R0) - 1u));
    };


    if (TMA_IN) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_init(&bar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && (long long)blockIdx.x < p.n_tiles) {
            const long long g0 = (long long)blockIdx.x << 12;
            const long long left = p.total - g0;
            const uint32_t bytes = (uint32_t)(left < 4096 ? left : 4096) * 4u;
            mbar_expect_tx(&bar[0], bytes);
            tma_load_1d(stage[0], p.in + g0, bytes, &bar[0]);
        }
    }

    auto prefetch_warp = [&](long long tile) {
        const long long g = tile << 12;
        if (g + 4096 <= p.total) {                                  // whole tile (all but possibly the last one): no predicates
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned i = warp_piece(j);
                cp_async_16z(land + phys(i), p.in + g + i, 16u);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned i = warp_piece(j);
                const bool in_range = g + i < p.total;              // p.total is a multiple of 4 samples
                cp_async_16z(land + phys(i), in_range ? p.in + g + i : p.in, in_range ? 16u : 0u);
            }
        }
        cp_async_commit_group();
    };
    if (CP_IN && (long long)blockIdx.x < p.n_tiles) prefetch_warp(blockIdx.x);

    // ---- per-thread constant twiddles of the upper rounds ----
    if (MIDSM) {                                   // table[w][low R0 bits of tid], w = (1 << q) - 1 + j, stage = R0 + q
        if (tid < (15u << R0)) {
            const int w = tid >> R0, lo4 = tid & ((1u << R0) - 1u);
            const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
            const int j = w - ((1 << q) - 1);
            midtw[(w << R0) + lo4] = __ldg(p.twp + (1u << (R0 + q)) + lo4 + ((unsigned)j << R0));
        }
        __syncthreads();
    }
    int uwr[NU][15], uwi[NU][15];
#pragma unroll
    for (int r = 1; r < NR; ++r) {
        if (MIDSM && r == 1) continue;
        const int lo = R0 + 4 * (r - 1);
        const unsigned low = tid & ((1u << lo) - 1u);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int2 w = __ldg(p.twp + (1u << (lo + q)) + low + ((unsigned)j << lo));
                uwr[r - 1][(1 << q) - 1 + j] = w.x;
                uwi[r - 1][(1 << q) - 1 + j] = w.y;
            }
    }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }

    int it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        uint32_t *sm = work[it & 1];
        const long long g0 = tile << 12;
        const bool full = g0 + 4096 <= p.total;

        auto prefetch_next = [&]() {                          // the next frame of this CTA
            // (the empty asm keeps the address / size arithmetic INSIDE thread 0's branch: left uniform, the front end
            // hoists some twenty uniform-datapath instructions in front of it, and all eight warps issue them every frame)
            long long nt = tile + gridDim.x;
            if (tid != 0) return;
            asm volatile("" : "+l"(nt));
            if (nt < p.n_tiles) {
                const long long left = p.total - (nt << 12);
                const uint32_t bytes = (uint32_t)(left < 4096 ? left : 4096) * 4u;
                if (NAT) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> TMA writes
                mbar_expect_tx(&bar[(it + 1) & 1], bytes);
                tma_load_1d(stage[(it + 1) & 1], p.in + (nt << 12), bytes, &bar[(it + 1) & 1]);
            }
        };
        // 8- and 16-point frames run in ONE register round, so nothing else orders the tiles: every thread must
        // have left the previous tile before its TMA landing buffer is refilled
        if (NR == 1) __syncthreads();
        if (TMA_IN) {
            if (!NAT) prefetch_next();
            mbar_wait(&bar[it & 1], (it >> 1) & 1);
        }

        int re[16], im[16];
        // one stage chain over the tile.  PH 0: the plain kernel; PH 1: first chain of a pair (DIF, results stay in
        // the registers); PH 2: second chain of a pair (DIT, first round's samples are already in the registers)
        auto chain = [&](auto dir_tag, auto ph_tag) {
        constexpr bool CD = decltype(dir_tag)::value;         // this chain's direction
        constexpr int PH = decltype(ph_tag)::value;
        constexpr bool TMA_IN = !CD && PH != 2;
        constexpr bool CP_IN = CD && PH == 0;
        constexpr bool COALESCE = (!CD || NR == 1) && !NAT && PH == 0;
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
            const int r = CD ? rr : NR - 1 - rr;                 // DIF walks the bits downwards
            const int lo = r == 0 ? 0 : R0 + 4 * (r - 1);
            const int R = r == 0 ? R0 : 4;
            const bool first = rr == 0, last = rr == NR - 1;
            const unsigned base = (tid & ((1u << lo) - 1u)) | ((tid >> lo) << (lo + R));
            const unsigned pbase = phys(base);

            // ---- fetch 16 samples ----
            if (PH == 2 && first) {
                // the spectrum is already here: DIF's lowest round and CD's lowest round own the same 16 samples
            } else if (r == 0 && R0 == 4) {
                if (first && CP_IN) { cp_async_wait_group0(); __syncwarp(); }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint4 v;
                    if (first && TMA_IN) {                        // 16-point DIF: straight from the (linear) TMA landing buffer
                        v = *reinterpret_cast<const uint4 *>(stage[it & 1] + 16 * tid + 4 * c);
                    } else if (first) {                           // CD: this thread's 64 bytes of the landed tile
                        v = *reinterpret_cast<const uint4 *>(land + pbase + phys(4 * c));
                    } else {
                        v = *reinterpret_cast<const uint4 *>(sm + pbase + phys(4 * c));
                    }
                    if (first) {
                        unpack<DW16>(v.x, dwv, re[4 * c + 0], im[4 * c + 0]);
                        unpack<DW16>(v.y, dwv, re[4 * c + 1], im[4 * c + 1]);
                        unpack<DW16>(v.z, dwv, re[4 * c + 2], im[4 * c + 2]);
                        unpack<DW16>(v.w, dwv, re[4 * c + 3], im[4 * c + 3]);
                    } else {
                        unpack<true>(v.x, dwv, re[4 * c + 0], im[4 * c + 0]);
                        unpack<true>(v.y, dwv, re[4 * c + 1], im[4 * c + 1]);
                        unpack<true>(v.z, dwv, re[4 * c + 2], im[4 * c + 2]);
                        unpack<true>(v.w, dwv, re[4 * c + 3], im[4 * c + 3]);
                    }
                }
                if (first && CP_IN) {                             // every lane has drained the warp's region
                    __syncwarp();
                    if (tile + gridDim.x < p.n_tiles) prefetch_warp(tile + gridDim.x);
                }
            } else {
                if (first && CP_IN) { cp_async_wait_group0(); __syncwarp(); }
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                    uint32_t x;
                    if (first && TMA_IN) x = stage[it & 1][base + off];
                    else if (first) x = land[pbase + phys(off)];          // CD: landed by the warp (zero-filled past the end)
                    else x = sm[pbase + phys(off)];
                    if (first) unpack<DW16>(x, dwv, re[m], im[m]);
                    else unpack<true>(x, dwv, re[m], im[m]);
                }
                if (first && CP_IN) {
                    __syncwarp();
                    if (tile + gridDim.x < p.n_tiles) prefetch_warp(tile + gridDim.x);
                }
            }

            // ---- butterflies ----
            // DIF upper rounds end with a multiply stage on register bit 0 (global bit >= 2): odd registers
            // then hold raw products and are packed from their upper half-words
            constexpr bool RAW = !CD && DW16 && R0 >= 2;
            if (r == 0) round_regs<0, R0, CD, DW16, MODE, false>(re, im, TwRegs{lwr, lwi}, tid_odd, sh_full, sh_half);
            else if (r == 1 && MIDSM) round_regs<R0, 4, CD, DW16, MODE, RAW>(re, im, TwSmemP<(1 << R0)>{midtw + (tid & ((1u << R0) - 1u))}, tid_odd, sh_full, sh_half);
            else if (r == 1) round_regs<R0, 4, CD, DW16, MODE, RAW>(re, im, TwRegs{uwr[0], uwi[0]}, tid_odd, sh_full, sh_half);
            else round_regs<R0 + 4, 4, CD, DW16, MODE, RAW>(re, im, TwRegs{uwr[NU - 1], uwi[NU - 1]}, tid_odd, sh_full, sh_half);

            // ---- hand the samples on: to the exchange tile, or to HBM after the last round ----
            if (PH == 1 && last) {
                // first chain of a pair: nothing leaves the registers
            } else if (r == 0 && R0 == 4) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = make_uint4(pack(re[4 * c], im[4 * c]), pack(re[4 * c + 1], im[4 * c + 1]),
                                               pack(re[4 * c + 2], im[4 * c + 2]), pack(re[4 * c + 3], im[4 * c + 3]));
                    if (last && NAT) {
                        const unsigned r8 = __brev(tid) >> 24;
                        uint32_t *nb = stage[it & 1] + (r8 ^ (((r8 >> 5) & 7u) << 2));
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int m = 4 * c + e;
                            nb[(((m & 1) << 3) | ((m & 2) << 1) | ((m & 4) >> 1) | ((m & 8) >> 3)) << 8] = w[e];
                        }
                    } else if (last && COALESCE) {
                        *reinterpret_cast<uint4 *>(sm + pbase + phys(4 * c)) = v;       // in place: the slots this thread read
                    } else if (last) {
                        const long long gi = g0 + 16 * tid + 4 * c;
                        if (full || gi < p.total) *reinterpret_cast<uint4 *>(p.out + gi) = v;
                    } else {
                        *reinterpret_cast<uint4 *>(sm + pbase + phys(4 * c)) = v;
                    }
                }
            } else {
                // (the whole-tile test is made ONCE: sixteen stores at immediate offsets from one pointer, instead of a
                // 64-bit compare and a predicate per store — 15 % of the DIT kernel's instructions)
                if (last && !COALESCE && full) {
                    uint32_t *dst = p.out + g0 + base;
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                        dst[off] = pack(re[m], im[m]);
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                        const uint32_t x = (RAW && r > 0 && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632)
                                                                     : pack(re[m], im[m]);
                        if (last && !COALESCE) { if ((g0 + base + off) < p.total) p.out[g0 + base + off] = x; }
                        else sm[pbase + phys(off)] = x;                       // last round: in place, stored below
                    }
                }
            }
            // The hand-over between the two LOWEST rounds of the 4+4+4 schedule stays inside a warp
            // (round bits 7..4 <-> 3..0 both keep tid >> 5 fixed), so a warp barrier is enough there;
            // the double-buffered tile makes one CTA barrier per frame sufficient for reuse safety.
            if (PH == 1 && last) {
                // (pair: the inverse chain follows)
            } else if (!last) {
                const bool warp_local = (NR == 3 && R0 == 4) && ((CD && rr == 0) || (!CD && rr == 1));
                if (warp_local) __syncwarp();
                else __syncthreads();
                if (NAT && rr == 0) prefetch_next();
            } else if (COALESCE) {
                // the warp's own runs of the tile, 512 contiguous bytes per store (up to the run length)
                __syncwarp();
                if (full) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const unsigned i = warp_piece(c);
                        *reinterpret_cast<uint4 *>(p.out + g0 + i) = *reinterpret_cast<const uint4 *>(sm + phys(i));
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const unsigned i = warp_piece(c);
                        if (g0 + i < p.total) *reinterpret_cast<uint4 *>(p.out + g0 + i) = *reinterpret_cast<const uint4 *>(sm + phys(i));
                    }
                }
            } else if (NAT) {
                __syncthreads();
                const uint4 *nb = reinterpret_cast<const uint4 *>(stage[it & 1]) + (tid ^ ((tid >> 3) & 7u));
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(p.out + g0 + 4 * (tid + 256 * c)) = nb[256 * c];
            }
        }
        };
        if (!PAIR) {
            chain(std::integral_constant<bool, DIT>{}, std::integral_constant<int, 0>{});
        } else {
            chain(std::false_type{}, std::integral_constant<int, 1>{});
            // every lane of the warp has read its round-0 samples of the forward chain out of the warp's region of
            // the exchange tile before the inverse chain's round 0 writes its results there
            __syncwarp();
            chain(std::true_type{}, std::integral_constant<int, 2>{});
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One-pass 8192-point packed-16 kernel (NFFT = 13, 16-bit scaled): replaces the strided-4 + contiguous-9 two-pass
// schedule, so every sample is read from HBM once and written once and one of the two load / exchange / store
// sequences disappears.  A 256-thread CTA owns a frame and walks its two 4096-sample blocks with the three
// register rounds of fast16_kernel<12>; STAGE 12 pairs sample i of the lower block with sample i of the upper one,
// and in the top round's ownership (tid + 256 m) both belong to the SAME thread, so that stage needs no exchange:
//   DIF: the frame lands as ONE 32 KB TMA bulk copy; STAGE 12 runs straight out of the landing buffer — the sums
//        stay in registers as the lower block's top-round input, the products go back into the thread's own slots
//        of the upper half of the landing buffer, which the upper block then reads like any TMA-landed tile.  The
//        next frame's copy is issued once every thread has read those slots (single-buffered, hidden behind the
//        upper block's last two rounds and the two other CTAs of the SM).
//   DIT: the lower block's top-round results wait in thread-private slots of the second exchange tile while the
//        upper block is computed, then STAGE 12 runs between those slots and the registers and both halves of the
//        frame leave as warp-coalesced 128-byte stores.  Blocks land per warp with cp.async, one block ahead.
// Same arithmetic (fly<>, round_regs<>) and twiddle placement as fast16_kernel; STAGE-12 twiddles are read through
// L1 / L2 per frame (16 per thread).
template <bool DIT, bool DW16, int MODE, int DWC = 0>
__global__ void __launch_bounds__(256, 3) fast16_n13_kernel(const __grid_constant__ Fast16Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    uint32_t(*work)[kTileWords] = reinterpret_cast<uint32_t(*)[kTileWords]>(smem_raw + kSmemHead);
    uint32_t(*stage)[4096] = reinterpret_cast<uint32_t(*)[4096]>(smem_raw + kSmemHead + 2 * kTileWords * 4);   // DIF
    uint32_t *land = reinterpret_cast<uint32_t *>(smem_raw + kSmemHead + 2 * kTileWords * 4);                    // DIT

    const unsigned tid = threadIdx.x;
    const int dwv = DWC ? DWC : p.dw;              // compile-time DATA_WIDTH (12 / 14) or the run-time one
    const int sh_full = DWC ? 32 - DWC : p.sh_full, sh_half = DWC ? 33 - DWC : p.sh_half;
    const bool tid_odd = tid & 1u;
    const long long n_frames = p.n_tiles;            // frames of 8192 samples
    constexpr bool RAW = !DIT && DW16;
    auto warp_piece = [&](int c) {                   // 16-byte piece lane + 32 c of the warp's 512 contiguous samples
        const unsigned q = (tid & 31u) + 32u * c;
        return ((tid & ~31u) << 4) + 4u * q;
    };

    if (!DIT) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && (long long)blockIdx.x < n_frames) {
            mbar_expect_tx(&bar[0], 32768u);
            tma_load_1d(stage[0], p.in + ((long long)blockIdx.x << 13), 32768u, &bar[0]);
        }
    }
    auto prefetch_warp = [&](long long block) {      // DIT: one 4096-sample block -> the skewed landing tile
        const long long g = block << 12;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned i = warp_piece(j);
            cp_async_16z(land + phys(i), p.in + g + i, 16u);
        }
        cp_async_commit_group();
    };
    if (DIT && (long long)blockIdx.x < n_frames) prefetch_warp((long long)blockIdx.x << 1);

    // ---- batch-invariant twiddles: STAGE 4..7 table, STAGE 8..11 registers, STAGE 2..3 parameters ----
    if (tid < 240) {
        const int w = tid >> 4, lo4 = tid & 15;
        const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
        const int j = w - ((1 << q) - 1);
        midtw[w * 16 + lo4] = __ldg(p.twp + (1u << (4 + q)) + lo4 + ((unsigned)j << 4));
    }
    int uwr[15], uwi[15];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < (1 << q); ++j) {
            const int2 w = __ldg(p.twp + (1u << (8 + q)) + tid + ((unsigned)j << 8));
            uwr[(1 << q) - 1 + j] = w.x;
            uwi[(1 << q) - 1 + j] = w.y;
        }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }
    const int2 *tw12 = p.twp + 4096 + tid;           // STAGE 12: index tid + 256 m
    __syncthreads();

    const unsigned pA = phys(16u * tid);                                    // round 0: 16 contiguous samples
    const unsigned pB = phys((tid & 15u) | ((tid >> 4) << 8));              // round 1: stride 16
    const unsigned pC = phys(tid);                                          // round 2: stride 256

    int it = 0;
    for (long long frame = blockIdx.x; frame < n_frames; frame += gridDim.x, ++it) {
        const long long g0 = frame << 13;
        int re[16], im[16];

        if (!DIT) {
            // ---- STAGE 12 out of the landing buffer: sums -> registers, products -> the thread's own upper slots ----
            mbar_wait(&bar[0], it & 1);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                int br, bi;
                unpack<DW16>(stage[0][tid + 256 * m], dwv, re[m], im[m]);
                unpack<DW16>(stage[1][tid + 256 * m], dwv, br, bi);
                const int2 w = __ldg(tw12 + 256 * m);
                fly<false, DW16, MODE, RAW>(12, false, re[m], im[m], br, bi, w.x, w.y, sh_full, sh_half);
                stage[1][tid + 256 * m] = RAW ? __byte_perm((unsigned)br, (unsigned)bi, 0x7632) : pack(br, bi);
            }
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t *sm = work[blk];
                if (blk) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) unpack<true>(stage[1][tid + 256 * m], dwv, re[m], im[m]);
                }
                // round 2: STAGE 11..8 (stride 256)
                round_regs<8, 4, false, DW16, MODE, RAW>(re, im, TwRegs{uwr, uwi}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    sm[pC + phys((unsigned)m << 8)] = (RAW && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632) : pack(re[m], im[m]);
                __syncthreads();
                // every thread has read block blk of the landing buffer (its STAGE-12 operands and, in the upper block,
                // its own product slots): that half of the NEXT frame may land — two 16 KB copies on one barrier, the
                // first one issued while the upper block of this frame is still being computed
                if (tid == 0 && frame + gridDim.x < n_frames) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic stores -> TMA writes
                    if (blk == 0) mbar_expect_tx(&bar[0], 32768u);
                    tma_load_1d(stage[blk], p.in + ((frame + gridDim.x) << 13) + 4096 * blk, 16384u, &bar[0]);
                }
                // round 1: STAGE 7..4 (stride 16)
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<true>(sm[pB + phys((unsigned)m << 4)], dwv, re[m], im[m]);
                round_regs<4, 4, false, DW16, MODE, RAW>(re, im, TwSmem{midtw + (tid & 15u)}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    sm[pB + phys((unsigned)m << 4)] = (RAW && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632) : pack(re[m], im[m]);
                __syncwarp();                                   // this hand-over stays inside the warp
                // round 0: STAGE 3..0 (16 contiguous samples)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(sm + pA + phys(4 * c));
                    unpack<true>(v.x, dwv, re[4 * c + 0], im[4 * c + 0]);
                    unpack<true>(v.y, dwv, re[4 * c + 1], im[4 * c + 1]);
                    unpack<true>(v.z, dwv, re[4 * c + 2], im[4 * c + 2]);
                    unpack<true>(v.w, dwv, re[4 * c + 3], im[4 * c + 3]);
                }
                round_regs<0, 4, false, DW16, MODE, false>(re, im, TwRegs{lwr, lwi}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4 *>(sm + pA + phys(4 * c)) =
                        make_uint4(pack(re[4 * c], im[4 * c]), pack(re[4 * c + 1], im[4 * c + 1]),
                                   pack(re[4 * c + 2], im[4 * c + 2]), pack(re[4 * c + 3], im[4 * c + 3]));
                __syncwarp();
                uint32_t *dst = p.out + g0 + 4096 * blk;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const unsigned i = warp_piece(c);
                    *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<const uint4 *>(sm + phys(i));
                }
            }
        } else {
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                uint32_t *sm = work[0];
                // round 0: STAGE 0..3 on the block the warp landed one block ago
                cp_async_wait_group0();
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(land + pA + phys(4 * c));
                    unpack<DW16>(v.x, dwv, re[4 * c + 0], im[4 * c + 0]);
                    unpack<DW16>(v.y, dwv, re[4 * c + 1], im[4 * c + 1]);
                    unpack<DW16>(v.z, dwv, re[4 * c + 2], im[4 * c + 2]);
                    unpack<DW16>(v.w, dwv, re[4 * c + 3], im[4 * c + 3]);
                }
                __syncwarp();                                   // every lane has drained the warp's region
                {
                    const long long nb = blk == 0 ? 2 * frame + 1 : 2 * (frame + gridDim.x);
                    if (nb < 2 * n_frames) prefetch_warp(nb);
                }
                round_regs<0, 4, true, DW16, MODE, false>(re, im, TwRegs{lwr, lwi}, tid_odd, sh_full, sh_half);
                // the exchange tile is single (the second one parks the lower block): the previous block's cross-warp
                // reads of it must be complete before this block writes
                __syncthreads();
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4 *>(sm + pA + phys(4 * c)) =
                        make_uint4(pack(re[4 * c], im[4 * c]), pack(re[4 * c + 1], im[4 * c + 1]),
                                   pack(re[4 * c + 2], im[4 * c + 2]), pack(re[4 * c + 3], im[4 * c + 3]));
                __syncwarp();
                // round 1: STAGE 4..7
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<true>(sm[pB + phys((unsigned)m << 4)], dwv, re[m], im[m]);
                round_regs<4, 4, true, DW16, MODE, false>(re, im, TwSmem{midtw + (tid & 15u)}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int m = 0; m < 16; ++m) sm[pB + phys((unsigned)m << 4)] = pack(re[m], im[m]);
                __syncthreads();
                // round 2: STAGE 8..11
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<true>(sm[pC + phys((unsigned)m << 8)], dwv, re[m], im[m]);
                round_regs<8, 4, true, DW16, MODE, false>(re, im, TwRegs{uwr, uwi}, tid_odd, sh_full, sh_half);
                if (blk == 0) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) work[1][m * 256 + tid] = pack(re[m], im[m]);     // thread-private slots
                } else {
                    // ---- STAGE 12 between the parked lower block and the registers; both halves leave coalesced ----
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        int ar, ai;
                        unpack<true>(work[1][m * 256 + tid], dwv, ar, ai);
                        const int2 w = __ldg(tw12 + 256 * m);
                        fly<true, DW16, MODE>(12, false, ar, ai, re[m], im[m], w.x, w.y, sh_full, sh_half);
                        p.out[g0 + tid + 256 * m] = pack(ar, ai);
                        p.out[g0 + 4096 + tid + 256 * m] = pack(re[m], im[m]);
                    }
                }
            }
        }
    }
}

template <bool DIT, bool DW16, int DWC = 0>
cudaError_t launch_n13_k(const Fast16Params &p, int mode, int grid, cudaStream_t st)
{
    const int smem = kSmemHead + 2 * kTileWords * 4 + (DIT ? (int)kTileWords * 4 : 2 * 4096 * 4);
    auto k = mode == MODE_ROUND ? fast16_n13_kernel<DIT, DW16, MODE_ROUND, DWC> : fast16_n13_kernel<DIT, DW16, MODE_TRUNC, DWC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// One-pass 16384-point packed-16 kernel (NFFT = 14, 16-bit scaled): the fast16_n13 pattern with FOUR 4096-sample
// blocks per frame.  It replaces the strided-4 + contiguous-10 schedule, whose strided pass does four stages at the
// price of a whole HBM round trip (2.1 GB in 0.34 ms = 6.2 TB/s: that pass runs AT the memory roofline).
// STAGE 13 pairs block b with b + 2, STAGE 12 block 2c with 2c + 1; in the top round's ownership (tid + 256 m) all four
// samples of such a radix-4 group belong to the SAME thread, so both stages need no exchange:
//   DIF: the frame lands as ONE 64 KB TMA bulk copy; STAGE 13 / 12 run straight out of the landing buffer — block 0's
//        results stay in registers, the other three go back into the thread's own slots, from where blocks 1..3 are
//        read like TMA-landed tiles.  The next frame's copy is issued once block 3 has been read (single-buffered).
//   DIT: blocks 0..2 park their top-round results in thread-private slots; after block 3 STAGE 12 / 13 run between the
//        parked blocks and the registers and the whole frame leaves as warp-coalesced 128-byte stores.
// 105 / 88 KB of shared memory: two CTAs per SM.  Twiddles of STAGE 12 / 13: three per radix-4 group, through L1 / L2.
template <bool DIT, bool DW16, int MODE, int DWC = 0>
__global__ void __launch_bounds__(256, 2) fast16_n14_kernel(const __grid_constant__ Fast16Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    uint32_t(*work)[kTileWords] = reinterpret_cast<uint32_t(*)[kTileWords]>(smem_raw + kSmemHead);
    uint32_t(*stage)[4096] = reinterpret_cast<uint32_t(*)[4096]>(smem_raw + kSmemHead + 2 * kTileWords * 4);   // DIF: 4 blocks
    uint32_t *land = reinterpret_cast<uint32_t *>(smem_raw + kSmemHead + kTileWords * 4);                        // DIT
    uint32_t(*park)[4096] = reinterpret_cast<uint32_t(*)[4096]>(smem_raw + kSmemHead + 2 * kTileWords * 4);     // DIT: 3 blocks

    const unsigned tid = threadIdx.x;
    const int dwv = DWC ? DWC : p.dw;              // compile-time DATA_WIDTH (12 / 14) or the run-time one
    const int sh_full = DWC ? 32 - DWC : p.sh_full, sh_half = DWC ? 33 - DWC : p.sh_half;
    const bool tid_odd = tid & 1u;
    const unsigned n_frames = (unsigned)p.n_tiles;   // frames of 16384 samples
    constexpr bool RAW = !DIT && DW16;
    auto warp_piece = [&](int c) {                   // 16-byte piece lane + 32 c of the warp's 512 contiguous samples
        const unsigned q = (tid & 31u) + 32u * c;
        return ((tid & ~31u) << 4) + 4u * q;
    };

    if (!DIT) {
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0 && blockIdx.x < n_frames) {
            mbar_expect_tx(&bar[0], 65536u);
            tma_load_1d(stage[0], p.in + ((long long)blockIdx.x << 14), 65536u, &bar[0]);
        }
    }
    auto prefetch_warp = [&](long long block) {      // DIT: one 4096-sample block -> the skewed landing tile
        const long long g = block << 12;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned i = warp_piece(j);
            cp_async_16z(land + phys(i), p.in + g + i, 16u);
        }
        cp_async_commit_group();
    };
    if (DIT && blockIdx.x < n_frames) prefetch_warp((long long)blockIdx.x << 2);

    // ---- batch-invariant twiddles: STAGE 4..7 table, STAGE 8..11 registers, STAGE 2..3 parameters ----
    if (tid < 240) {
        const int w = tid >> 4, lo4 = tid & 15;
        const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
        const int j = w - ((1 << q) - 1);
        midtw[w * 16 + lo4] = __ldg(p.twp + (1u << (4 + q)) + lo4 + ((unsigned)j << 4));
    }
    int uwr[15], uwi[15];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < (1 << q); ++j) {
            const int2 w = __ldg(p.twp + (1u << (8 + q)) + tid + ((unsigned)j << 8));
            uwr[(1 << q) - 1 + j] = w.x;
            uwi[(1 << q) - 1 + j] = w.y;
        }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }
    const int2 *tw12 = p.twp + 4096 + tid;           // STAGE 12: index tid + 256 m
    const int2 *tw13 = p.twp + 8192 + tid;           // STAGE 13: index tid + 256 m (+ 4096 for the odd blocks)
    __syncthreads();

    const unsigned pA = phys(16u * tid);                                    // round 0: 16 contiguous samples
    const unsigned pB = phys((tid & 15u) | ((tid >> 4) << 8));              // round 1: stride 16
    const unsigned pC = phys(tid);                                          // round 2: stride 256

    unsigned it = 0;
    for (unsigned frame = blockIdx.x; frame < n_frames; frame += gridDim.x, ++it) {
        const long long g0 = (long long)frame << 14;
        int re[16], im[16];

        if (!DIT) {
            // ---- STAGE 13 and 12 out of the landing buffer: block 0 -> registers, blocks 1..3 -> the thread's own slots ----
            mbar_wait(&bar[0], it & 1);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned i = tid + 256 * m;
                int r1, i1, r2, i2, r3, i3;
                unpack<DW16>(stage[0][i], dwv, re[m], im[m]);
                unpack<DW16>(stage[1][i], dwv, r1, i1);
                unpack<DW16>(stage[2][i], dwv, r2, i2);
                unpack<DW16>(stage[3][i], dwv, r3, i3);
                const int2 wa = __ldg(tw13 + 256 * m), wb = __ldg(tw13 + 4096 + 256 * m), wc = __ldg(tw12 + 256 * m);
                fly<false, DW16, MODE, false>(13, false, re[m], im[m], r2, i2, wa.x, wa.y, sh_full, sh_half);
                fly<false, DW16, MODE, false>(13, false, r1, i1, r3, i3, wb.x, wb.y, sh_full, sh_half);
                fly<false, DW16, MODE, RAW>(12, false, re[m], im[m], r1, i1, wc.x, wc.y, sh_full, sh_half);
                fly<false, DW16, MODE, RAW>(12, false, r2, i2, r3, i3, wc.x, wc.y, sh_full, sh_half);
                stage[1][i] = RAW ? __byte_perm((unsigned)r1, (unsigned)i1, 0x7632) : pack(r1, i1);
                stage[2][i] = pack(r2, i2);
                stage[3][i] = RAW ? __byte_perm((unsigned)r3, (unsigned)i3, 0x7632) : pack(r3, i3);
            }
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                uint32_t *sm = work[blk & 1];
                if (blk) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) unpack<true>(stage[blk][tid + 256 * m], dwv, re[m], im[m]);
                }
                // round 2: STAGE 11..8 (stride 256)
                round_regs<8, 4, false, DW16, MODE, RAW>(re, im, TwRegs{uwr, uwi}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    sm[pC + phys((unsigned)m << 8)] = (RAW && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632) : pack(re[m], im[m]);
                __syncthreads();
                // every thread has read block blk of the landing buffer (its STAGE 13 / 12 operands and, for blk > 0, its
                // own slots): that quarter of the NEXT frame may land — four 16 KB copies on one barrier, the first
                // three issued while later blocks of this frame are still being computed
                if (tid == 0 && frame + gridDim.x < n_frames) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic stores -> TMA writes
                    if (blk == 0) mbar_expect_tx(&bar[0], 65536u);
                    tma_load_1d(stage[blk], p.in + ((long long)(frame + gridDim.x) << 14) + 4096 * blk, 16384u, &bar[0]);
                }
                // round 1: STAGE 7..4 (stride 16)
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<true>(sm[pB + phys((unsigned)m << 4)], dwv, re[m], im[m]);
                round_regs<4, 4, false, DW16, MODE, RAW>(re, im, TwSmem{midtw + (tid & 15u)}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    sm[pB + phys((unsigned)m << 4)] = (RAW && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632) : pack(re[m], im[m]);
                __syncwarp();                                   // this hand-over stays inside the warp
                // round 0: STAGE 3..0 (16 contiguous samples)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(sm + pA + phys(4 * c));
                    unpack<true>(v.x, dwv, re[4 * c + 0], im[4 * c + 0]);
                    unpack<true>(v.y, dwv, re[4 * c + 1], im[4 * c + 1]);
                    unpack<true>(v.z, dwv, re[4 * c + 2], im[4 * c + 2]);
                    unpack<true>(v.w, dwv, re[4 * c + 3], im[4 * c + 3]);
                }
                round_regs<0, 4, false, DW16, MODE, false>(re, im, TwRegs{lwr, lwi}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4 *>(sm + pA + phys(4 * c)) =
                        make_uint4(pack(re[4 * c], im[4 * c]), pack(re[4 * c + 1], im[4 * c + 1]),
                                   pack(re[4 * c + 2], im[4 * c + 2]), pack(re[4 * c + 3], im[4 * c + 3]));
                __syncwarp();
                uint32_t *dst = p.out + g0 + 4096 * blk;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const unsigned i = warp_piece(c);
                    *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<const uint4 *>(sm + phys(i));
                }
            }
        } else {
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                uint32_t *sm = work[0];
                // round 0: STAGE 0..3 on the block the warp landed one block ago
                cp_async_wait_group0();
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(land + pA + phys(4 * c));
                    unpack<DW16>(v.x, dwv, re[4 * c + 0], im[4 * c + 0]);
                    unpack<DW16>(v.y, dwv, re[4 * c + 1], im[4 * c + 1]);
                    unpack<DW16>(v.z, dwv, re[4 * c + 2], im[4 * c + 2]);
                    unpack<DW16>(v.w, dwv, re[4 * c + 3], im[4 * c + 3]);
                }
                __syncwarp();                                   // every lane has drained the warp's region
                {
                    const long long nb = blk < 3 ? 4ll * frame + blk + 1 : 4ll * (frame + gridDim.x);
                    if (nb < 4ll * n_frames) prefetch_warp(nb);
                }
                round_regs<0, 4, true, DW16, MODE, false>(re, im, TwRegs{lwr, lwi}, tid_odd, sh_full, sh_half);
                // the exchange tile is single: the previous block's cross-warp reads of it must be complete before this
                // block writes
                __syncthreads();
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4 *>(sm + pA + phys(4 * c)) =
                        make_uint4(pack(re[4 * c], im[4 * c]), pack(re[4 * c + 1], im[4 * c + 1]),
                                   pack(re[4 * c + 2], im[4 * c + 2]), pack(re[4 * c + 3], im[4 * c + 3]));
                __syncwarp();
                // round 1: STAGE 4..7
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<true>(sm[pB + phys((unsigned)m << 4)], dwv, re[m], im[m]);
                round_regs<4, 4, true, DW16, MODE, false>(re, im, TwSmem{midtw + (tid & 15u)}, tid_odd, sh_full, sh_half);
#pragma unroll
                for (int m = 0; m < 16; ++m) sm[pB + phys((unsigned)m << 4)] = pack(re[m], im[m]);
                __syncthreads();
                // round 2: STAGE 8..11
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<true>(sm[pC + phys((unsigned)m << 8)], dwv, re[m], im[m]);
                round_regs<8, 4, true, DW16, MODE, false>(re, im, TwRegs{uwr, uwi}, tid_odd, sh_full, sh_half);
                if (blk < 3) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) park[blk][m * 256 + tid] = pack(re[m], im[m]);   // thread-private slots
                } else {
                    // ---- STAGE 12 then 13 between the three parked blocks and the registers; coalesced stores ----
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        int r0, i0, r1, i1, r2, i2;
                        unpack<true>(park[0][m * 256 + tid], dwv, r0, i0);
                        unpack<true>(park[1][m * 256 + tid], dwv, r1, i1);
                        unpack<true>(park[2][m * 256 + tid], dwv, r2, i2);
                        const int2 wc = __ldg(tw12 + 256 * m), wa = __ldg(tw13 + 256 * m), wb = __ldg(tw13 + 4096 + 256 * m);
                        fly<true, DW16, MODE>(12, false, r0, i0, r1, i1, wc.x, wc.y, sh_full, sh_half);
                        fly<true, DW16, MODE>(12, false, r2, i2, re[m], im[m], wc.x, wc.y, sh_full, sh_half);
                        fly<true, DW16, MODE>(13, false, r0, i0, r2, i2, wa.x, wa.y, sh_full, sh_half);
                        fly<true, DW16, MODE>(13, false, r1, i1, re[m], im[m], wb.x, wb.y, sh_full, sh_half);
                        uint32_t *dst = p.out + g0 + tid + 256 * m;
                        dst[0] = pack(r0, i0);
                        dst[4096] = pack(r1, i1);
                        dst[8192] = pack(r2, i2);
                        dst[12288] = pack(re[m], im[m]);
                    }
                }
            }
        }
    }
}

template <bool DIT, bool DW16, int DWC = 0>
cudaError_t launch_n14_k(const Fast16Params &p, int mode, int grid, cudaStream_t st)
{
    const int smem = DIT ? (int)(kSmemHead + 2 * kTileWords * 4 + 3 * 4096 * 4) : (int)(kSmemHead + 2 * kTileWords * 4 + 4 * 4096 * 4);
    auto k = mode == MODE_ROUND ? fast16_n14_kernel<DIT, DW16, MODE_ROUND, DWC> : fast16_n14_kernel<DIT, DW16, MODE_TRUNC, DWC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Strided pass for NFFT = 13..20 (16-bit scaled TRUNCATE): the top G = 4 or 8 stage bits of a frame.
// A tile is 2^G rows (stride 2^(NFFT-G) samples) by 2^(12-G) contiguous columns; a CTA keeps one
// column block ("mid") and walks over frames, so the between-pass twiddles — which depend on the
// column, not on the frame — are hoisted once per work unit exactly like in the single-tile kernel.
// The remaining NFFT-G bits are done by fast16_kernel<NFFT-G> on contiguous blocks (its twiddles
// W_s[k], s < 12, are the same table entries for every NFFT).
struct Strided16Params {
    const uint32_t *in;
    uint32_t *out;
    const int2 *twp;
    long long batch;
    int n;                  // NFFT
    int frames_per_unit;
    long long n_units;      // mid_count * ceil(batch / frames_per_unit)
    int dw, sh_full, sh_half;
    TaylorDev tay;          // .on: STAGE >= 11 twiddles are recomputed here (rom9 + Taylor MACs), twp ends at STAGE 11
};

// Shared memory of the strided pass: [head | exchange tile | 2 x landing tile]; all three tiles use the
// phys() skew, so a landing tile is read exactly like the exchange tile.
constexpr unsigned kStridedSmem = kSmemHead + 3 * kTileWords * 4;

__device__ __forceinline__ void cp_async_16(uint32_t *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// the 16 stores of a strided pass's last round: register m goes 2^SHIFT m rows below register 0; with the
// row pitch a compile-time constant (PB = NFFT - G = 9..12) every offset is an instruction immediate
template <int PB, int SHIFT>
__device__ __forceinline__ void store_rows(char *ptr, const int (&re)[16], const int (&im)[16])
{
#pragma unroll
    for (int m = 0; m < 16; ++m)
        *reinterpret_cast<uint32_t *>(ptr + ((((size_t)m << SHIFT) << PB) << 2)) = pack(re[m], im[m]);
}

template <int G, bool DIT, bool DW16, int MODE>
__global__ void __launch_bounds__(256, 3) fast16_strided_kernel(const __grid_constant__ Strided16Params p)
{
    constexpr int C = 12 - G;                       // log2 contiguous columns per tile
    constexpr int NR = G / 4;                       // rounds: local bits [8,12) and, for G = 8, [4,8)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    uint32_t *work = reinterpret_cast<uint32_t *>(smem_raw + kSmemHead);
    uint32_t(*land)[kTileWords] = reinterpret_cast<uint32_t(*)[kTileWords]>(smem_raw + kSmemHead + kTileWords * 4);

    const unsigned tid = threadIdx.x;
    const int sh_full = p.sh_full, sh_half = p.sh_half;
    const int pb = p.n - G;                         // lowest global bit of this pass
    const unsigned cmask = (1u << C) - 1u;

    int it = 0;
    // Units = (frame chunk, column block), column block fastest, dealt round-robin: the CTAs running at any moment
    // cover ALL column blocks of the same few frames, so the 64-byte row pieces of neighbouring blocks — two to a
    // 128-byte line — are fetched together and the frames stream through DRAM page by page.  (Giving every CTA one
    // contiguous range of (block, frame) items instead balances the load perfectly but loses that locality:
    // c4 1.27 -> 1.59 ms, measured r02.)  The launcher picks the chunk length that balances the round-robin deal.
    const int mid_bits = pb - C;
    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const unsigned mid = (unsigned)(u & ((1ll << mid_bits) - 1));
        const long long f0 = (u >> mid_bits) * p.frames_per_unit;
        const long long f1 = (f0 + p.frames_per_unit < p.batch) ? f0 + p.frames_per_unit : p.batch;
        // twiddle index of local position l (only its bits below the stage bit matter)
        auto kidx = [&](unsigned l) { return ((l >> C) << pb) | (mid << C) | (l & cmask); };

        // One frame's 2^G x 2^C column block -> a landing tile, as 16-byte cp.async pieces (4 per thread;
        // a piece never straddles a row because C >= 2).  The copy of frame f + 1 is in flight while
        // frame f is computed; the CTA barrier at the top of a frame publishes the landed pieces and
        // also orders the single exchange tile's reuse.
        // Addresses inside a frame are 32-bit byte offsets from a CTA-uniform 64-bit base (a frame is at
        // most 4 MB), so each access costs one integer add instead of a 64-bit multiply-add.
        auto prefetch = [&](uint32_t *dst, long long f) {
            const char *src = reinterpret_cast<const char *>(p.in + (f << p.n) + ((long long)mid << C));
            const unsigned l0 = 4u * tid;
            const unsigned b0 = ((l0 >> C) << (pb + 2)) + ((l0 & cmask) << 2);
#pragma unroll
            for (int j = 0; j < 4; ++j)       // piece j: local index l0 + 1024 j, i.e. 2^(10-C) j rows further down
                cp_async_16(dst + phys(4u * tid) + phys(1024u * j), src + (b0 + (((unsigned)j << (10 - C)) << (pb + 2))));
            cp_async_commit();
        };
        if (f0 < f1) prefetch(land[it & 1], f0);

        // ---- twiddles of this column block ----
        int uwr[15], uwi[15];                       // top round: local bits 8..11
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int sgl = pb + (8 + q - C);
                const int2 w = hoist_twiddle(p.twp, p.tay, sgl, kidx(tid | ((unsigned)j << 8)) & ((1u << sgl) - 1u));
                uwr[(1 << q) - 1 + j] = w.x;
                uwi[(1 << q) - 1 + j] = w.y;
            }
        if (NR == 2) {                              // lower round: local bits 4..7, table[w][tid & 15]
            __syncthreads();                        // previous unit's readers are done
            if (tid < 240) {
                const int w = tid >> 4, lo4 = tid & 15;
                const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
                const int j = w - ((1 << q) - 1);
                const int sgl = pb + (4 + q - C);
                midtw[w * 16 + lo4] = hoist_twiddle(p.twp, p.tay, sgl, kidx((unsigned)lo4 | ((unsigned)j << 4)) & ((1u << sgl) - 1u));
            }
        }

        for (long long f = f0; f < f1; ++f, ++it) {
            cp_async_wait_all();
            __syncthreads();
            const uint32_t *st = land[it & 1];
            if (f + 1 < f1) prefetch(land[(it + 1) & 1], f + 1);
            char *obase = reinterpret_cast<char *>(p.out + (f << p.n) + ((long long)mid << C));
            int re[16], im[16];
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) {
                const int r = DIT ? rr : NR - 1 - rr;              // r = 0: local bits 12-4*NR.., r = NR-1: 8..11
                const int lo = 12 - 4 * (NR - r);
                const bool first = rr == 0, last = rr == NR - 1;
                const unsigned base = (tid & ((1u << lo) - 1u)) | ((tid >> lo) << (lo + 4));
                const unsigned pbase = phys(base);
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const uint32_t x = (first ? st : work)[pbase + phys((unsigned)m << lo)];
                    if (first) unpack<DW16>(x, p.dw, re[m], im[m]);
                    else unpack<true>(x, p.dw, re[m], im[m]);
                }
                constexpr bool RAW = !DIT && DW16;
                if (lo == 8) round_regs<8, 4, DIT, DW16, MODE, RAW && (NR == 2)>(re, im, TwRegs{uwr, uwi}, false, sh_full, sh_half);
                else round_regs<4, 4, DIT, DW16, MODE, false>(re, im, TwSmem{midtw + (tid & 15u)}, false, sh_full, sh_half);
                if (last) {
                    // register m sits 2^(lo-C) m rows below register 0 (lo >= C in every geometry)
                    constexpr int SHIFT = (G == 8 && DIT) ? 4 : 0;
                    char *ptr = obase + (((base >> C) << (pb + 2)) + ((base & cmask) << 2));
                    switch (pb) {
                    case 9: store_rows<9, SHIFT>(ptr, re, im); break;
                    case 10: store_rows<10, SHIFT>(ptr, re, im); break;
                    case 11: store_rows<11, SHIFT>(ptr, re, im); break;
                    default: store_rows<12, SHIFT>(ptr, re, im); break;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const uint32_t x = (RAW && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632)
                                                            : pack(re[m], im[m]);
                        work[pbase + phys((unsigned)m << lo)] = x;
                    }
                    __syncthreads();
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Strided pass, G = 8 (NFFT 16..20: 256 rows x 16 columns), column block moved by the TMA engine in BOTH
// directions: one thread issues cp.async.bulk.tensor.2d for the frame's 256 x 64-byte block (the "stride-2^s
// shuffle" of the delay lines, int_delay_line.vhd:52-104, done by the copy engine instead of 4 cp.async + address
// arithmetic per thread), the last round's results go to a dense tile and leave as ONE tensor store instead of
// 16 four-byte STG per thread.  The whole batch is one 2-D tensor: 2^(NFFT-8) columns x (batch * 256) rows.
// G = 4 (one round per frame) runs at the HBM roofline: 2.1 GB in 0.337 ms = 6.2 TB/s, its warps wait on the landing
// barrier and a deeper landing ring changes nothing (tried: three tiles, two frames of lead — 0.980 vs 0.986 ms for NFFT 14).
// Dense landing tiles (TMA cannot skew): word l = 16 row + column.  Round "bits 8..11" touches l = tid + 256 m
// (conflict-free), round "bits 4..7" l = (tid & 15) + 256 (tid >> 4) + 16 m (two-way conflicts between the
// half-warps on the ONE access set of four that meets a dense tile; the exchange tile keeps the phys() skew).
constexpr unsigned kStridedTmaSmem = kSmemHead + kTileWords * 4 + 3 * 16384;

template <int G, bool DIT, bool DW16, int MODE, int DWC = 0>
__global__ void __launch_bounds__(256, 3) fast16_strided_tma_kernel(const __grid_constant__ Strided16Params p,
                                                                    const __grid_constant__ CUtensorMap map_in,
                                                                    const __grid_constant__ CUtensorMap map_out)
{
    constexpr int C = 12 - G;                       // G = 8: 256 rows x 16 columns; G = 4: 16 rows x 256 columns
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    uint32_t *work = reinterpret_cast<uint32_t *>(smem_raw + kSmemHead);
    uint32_t(*land)[4096] = reinterpret_cast<uint32_t(*)[4096]>(smem_raw + kSmemHead + kTileWords * 4);
    uint32_t *otile = land[2];

    const unsigned tid = threadIdx.x;
    const int dwv = DWC ? DWC : p.dw;              // compile-time DATA_WIDTH (12 / 14) or the run-time one
    const int sh_full = DWC ? 32 - DWC : p.sh_full, sh_half = DWC ? 33 - DWC : p.sh_half;
    const int pb = p.n - G;
    const unsigned cmask = (1u << C) - 1u;
    const int mid_bits = pb - C;

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_in)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_out)) : "memory");
    }
    __syncthreads();

    unsigned it = 0, phase = 0;                      // phase bit b = parity the next wait on bar[b] expects
    auto load = [&](unsigned buf, unsigned mid, long long f) {          // thread 0 only
        mbar_expect_tx(&bar[buf], 16384u);
        tma::load_2d(land[buf], &map_in, (int)(mid << C), (int)(f << G), &bar[buf]);
    };
    // ownerships inside the 256 x 16 block: round on local bits 8..11 / on local bits 4..7
    const unsigned base8 = tid, base4 = (tid & 15u) | ((tid >> 4) << 8);
    const unsigned pbase8 = phys(base8), pbase4 = phys(base4);

    for (long long u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        const unsigned mid = (unsigned)(u & ((1ll << mid_bits) - 1));
        const long long f0 = (u >> mid_bits) * p.frames_per_unit;
        const long long f1 = (f0 + p.frames_per_unit < p.batch) ? f0 + p.frames_per_unit : p.batch;
        auto kidx = [&](unsigned l) { return ((l >> C) << pb) | (mid << C) | (l & cmask); };
        if (tid == 0 && f0 < f1) load(it & 1u, mid, f0);

        int uwr[15], uwi[15];                       // round on local bits 8..11
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int sgl = pb + (8 + q - C);
                const int2 w = hoist_twiddle(p.twp, p.tay, sgl, kidx(tid | ((unsigned)j << 8)) & ((1u << sgl) - 1u));
                uwr[(1 << q) - 1 + j] = w.x;
                uwi[(1 << q) - 1 + j] = w.y;
            }
        if (G == 8) __syncthreads();                // previous unit's readers of the table are done
        if (G == 8 && tid < 240) {                  // round on local bits 4..7: table[w][tid & 15]
            const int w = tid >> 4, lo4 = tid & 15;
            const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
            const int j = w - ((1 << q) - 1);
            const int sgl = pb + (4 + q - C);
            midtw[w * 16 + lo4] = hoist_twiddle(p.twp, p.tay, sgl, kidx((unsigned)lo4 | ((unsigned)j << 4)) & ((1u << sgl) - 1u));
        }
        if (G == 8) __syncthreads();

        for (long long f = f0; f < f1; ++f, ++it) {
            const unsigned buf = it & 1u;
            // the other landing tile was drained in the previous frame's first round, which every thread left
            // before that frame's last barrier: refill it while this frame is computed
            if (tid == 0 && f + 1 < f1) load(buf ^ 1u, mid, f + 1);
            mbar_wait(&bar[buf], (phase >> buf) & 1u);
            phase ^= 1u << buf;
            const uint32_t *st = land[buf];
            int re[16], im[16];
            constexpr bool RAW = !DIT && DW16;
            if (G == 4) {                           // single round on local bits 8..11: landing tile -> output tile
#pragma unroll
                for (int m = 0; m < 16; ++m) unpack<DW16>(st[base8 + 256u * m], dwv, re[m], im[m]);
                round_regs<8, 4, DIT, DW16, MODE, false>(re, im, TwRegs{uwr, uwi}, false, sh_full, sh_half);
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();                    // the previous frame's tensor store has finished reading the output tile
#pragma unroll
                for (int m = 0; m < 16; ++m) otile[base8 + 256u * m] = pack(re[m], im[m]);
            } else {
                if (!DIT) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) unpack<DW16>(st[base8 + 256u * m], dwv, re[m], im[m]);
                    round_regs<8, 4, false, DW16, MODE, RAW>(re, im, TwRegs{uwr, uwi}, false, sh_full, sh_half);
#pragma unroll
                    for (int m = 0; m < 16; ++m)
                        work[pbase8 + phys((unsigned)m << 8)] = (RAW && (m & 1)) ? __byte_perm((unsigned)re[m], (unsigned)im[m], 0x7632) : pack(re[m], im[m]);
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) unpack<DW16>(st[base4 + 16u * m], dwv, re[m], im[m]);
                    round_regs<4, 4, true, DW16, MODE, false>(re, im, TwSmem{midtw + (tid & 15u)}, false, sh_full, sh_half);
#pragma unroll
                    for (int m = 0; m < 16; ++m) work[pbase4 + phys((unsigned)m << 4)] = pack(re[m], im[m]);
                }
                // the previous frame's tensor store has finished READING the output tile before anyone rewrites it
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
                if (!DIT) {
#pragma unroll
                    for (int m = 0; m < 16; ++m) unpack<true>(work[pbase4 + phys((unsigned)m << 4)], dwv, re[m], im[m]);
                    round_regs<4, 4, false, DW16, MODE, false>(re, im, TwSmem{midtw + (tid & 15u)}, false, sh_full, sh_half);
#pragma unroll
                    for (int m = 0; m < 16; ++m) otile[base4 + 16u * m] = pack(re[m], im[m]);
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) unpack<true>(work[pbase8 + phys((unsigned)m << 8)], dwv, re[m], im[m]);
                    round_regs<8, 4, true, DW16, MODE, false>(re, im, TwRegs{uwr, uwi}, false, sh_full, sh_half);
#pragma unroll
                    for (int m = 0; m < 16; ++m) otile[base8 + 256u * m] = pack(re[m], im[m]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the TMA engine
            __syncthreads();                        // also: every thread has left the exchange tile
            if (tid == 0) tma::store_2d(&map_out, (int)(mid << C), (int)(f << G), otile);
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int G, bool DIT, bool DW16, int DWC = 0>
cudaError_t launch_strided_tma_k(const Strided16Params &p, int mode, int grid, cudaStream_t st)
{
    CUtensorMap mi, mo;
    // the whole batch as a 2-D tensor of packed samples: 2^(NFFT-G) columns x (batch * 2^G) rows; box = 2^(12-G) x 2^G
    const uint64_t cols = (uint64_t)1 << (p.n - G), rows = (uint64_t)p.batch << G;
    if (!tma::make_map(&mi, p.in, 1, cols, rows, 1u << (12 - G), 1u << G) ||
        !tma::make_map(&mo, p.out, 1, cols, rows, 1u << (12 - G), 1u << G))
        return cudaErrorNotSupported;
    auto k = mode == MODE_ROUND ? fast16_strided_tma_kernel<G, DIT, DW16, MODE_ROUND, DWC> : fast16_strided_tma_kernel<G, DIT, DW16, MODE_TRUNC, DWC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStridedTmaSmem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, kStridedTmaSmem, st>>>(p, mi, mo);
    return cudaGetLastError();
}

template <int G, bool DIT, bool DW16>
cudaError_t launch_strided_k(const Strided16Params &p, int mode, int grid, cudaStream_t st)
{
    const int smem = (int)kStridedSmem;
    auto k = mode == MODE_ROUND ? fast16_strided_kernel<G, DIT, DW16, MODE_ROUND> : fast16_strided_kernel<G, DIT, DW16, MODE_TRUNC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NLOG2, bool DIT, bool DW16, bool NAT = false, bool PAIR = false, int DWC = 0>
cudaError_t launch_k(const Fast16Params &p, int mode, int grid, cudaStream_t st)
{
    const int smem = kSmemHead + 2 * kTileWords * 4 + (DIT ? (int)kTileWords * 4 : 2 * 4096 * 4);
    constexpr bool MIDSM = (NLOG2 >= 9 && NLOG2 <= 12);      // every three-round schedule: 80 registers, three CTAs per SM
    auto k = mode == MODE_ROUND ? fast16_kernel<NLOG2, DIT, DW16, MIDSM, MODE_ROUND, NAT, PAIR, DWC> : fast16_kernel<NLOG2, DIT, DW16, MIDSM, MODE_TRUNC, NAT, PAIR, DWC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

template <int NLOG2>
cudaError_t launch_n(const Fast16Params &p, int mode, bool dit, bool dw16, int grid, cudaStream_t st)
{
    if constexpr (NLOG2 >= 8) {                  // the converter widths 12 and 14 as compile-time constants (2^8 .. 2^12 points)
        if (p.dw == 12) return dit ? launch_k<NLOG2, true, false, false, false, 12>(p, mode, grid, st) : launch_k<NLOG2, false, false, false, false, 12>(p, mode, grid, st);
        if (p.dw == 14) return dit ? launch_k<NLOG2, true, false, false, false, 14>(p, mode, grid, st) : launch_k<NLOG2, false, false, false, false, 14>(p, mode, grid, st);
    }
    if (!dit) return dw16 ? launch_k<NLOG2, false, true>(p, mode, grid, st) : launch_k<NLOG2, false, false>(p, mode, grid, st);
    return dw16 ? launch_k<NLOG2, true, true>(p, mode, grid, st) : launch_k<NLOG2, true, false>(p, mode, grid, st);
}

}  // namespace

#if FAST16_PART == 1
bool fast16_supported(const intfft_generics &g)
{
    return g.format == 0 && g.use_fly == 1 && g.data_width <= 16 && g.twdl_width <= 16 &&
           g.nfft_log2 >= 3 && g.nfft_log2 <= 20;
}
#endif  // FAST16_PART == 1

// This file is compiled twice (Makefile: -DFAST16_PART=1 / 2) so that its two halves build in parallel: part 1 = the
// contiguous kernels and the fused pair, part 2 = the one-pass 8192 / 16384-point kernels and the strided passes.
#if FAST16_PART == 2
// Chunk length of the strided pass's round-robin deal.  Unit u = (chunk u / mids, column block u % mids) goes to
// CTA u % grid; a unit costs its frames plus `hoist` frame-times for fetching / recomputing the block's twiddles.
// Picks the length whose most loaded CTA finishes first (c4: 12 chunks of 22 frames = 6.9 waves instead of the old
// fixed "8 x grid units" rule's 14 chunks of 19 = 8.07 waves, whose ninth wave was almost empty).
int strided_frames_per_unit(long long mids, long long batch, long long grid, int hoist)
{
    struct Key { long long mids, batch, grid; int hoist, fpu; };
    static thread_local Key last{0, 0, 0, 0, 0};
    if (last.mids == mids && last.batch == batch && last.grid == grid && last.hoist == hoist) return last.fpu;
    long long best_cost = -1;
    int best_fpu = 1;
    std::vector<long long> load((size_t)grid);
    // candidates: up to ~12 waves of units (more only adds hoists), at most 64 of them
    long long max_chunks = (12 * grid + mids - 1) / mids;
    if (max_chunks < 48) max_chunks = 48;
    if (max_chunks > batch) max_chunks = batch;
    const long long step = (max_chunks + 63) / 64;
    for (long long want = 1; want <= max_chunks; want += (want < 16 ? 1 : step)) {
        const long long fpu = (batch + want - 1) / want;
        const long long chunks = (batch + fpu - 1) / fpu;
        const long long units = mids * chunks;
        std::fill(load.begin(), load.end(), 0);
        for (long long u = 0; u < units; ++u) {
            const long long c = u / mids;
            const long long frames = (c + 1) * fpu <= batch ? fpu : batch - c * fpu;
            load[(size_t)(u % grid)] += frames + hoist;
        }
        long long worst = 0;
        for (long long v : load) worst = v > worst ? v : worst;
        if (best_cost < 0 || worst < best_cost) { best_cost = worst; best_fpu = (int)fpu; }
    }
    last = Key{mids, batch, grid, hoist, best_fpu};
    return best_fpu;
}

// top-bits pass of an NFFT >= 13 plan: kp.g in {4, 8}, kp.pb = NFFT - kp.g
int launch_fast16_strided(const PassDesc &pd, int mode, bool dit, const int2 *twp, int num_sms, void *stream,
                          const TaylorDev *tay)
{
    Strided16Params p{};
    if (tay) p.tay = *tay;
    p.in = reinterpret_cast<const uint32_t *>(pd.kp.in);
    p.out = reinterpret_cast<uint32_t *>(pd.kp.out);
    p.twp = twp;
    p.n = pd.kp.n;
    p.batch = pd.kp.total >> pd.kp.n;
    p.dw = pd.kp.dw;
    p.sh_full = 32 - p.dw;
    p.sh_half = 33 - p.dw;
    const int G = pd.kp.g, C = 12 - G, mid_bits = p.n - G - C;
    const long long mids = 1ll << mid_bits;
    long long grid = 3ll * num_sms;
    p.frames_per_unit = strided_frames_per_unit(mids, p.batch, grid, p.tay.on ? 2 : 1);
    p.n_units = mids * ((p.batch + p.frames_per_unit - 1) / p.frames_per_unit);
    if (grid > p.n_units) grid = p.n_units;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool dw16 = p.dw == 16;
    cudaError_t e;
    // G = 8: the TMA-staged variant (INTFFT_STRIDED_TMA=0 keeps the cp.async / STG one)
    const char *tma_env = std::getenv("INTFFT_STRIDED_TMA");
    const bool tma_now = !(tma_env && tma_env[0] == '0');
    e = cudaErrorNotSupported;
    if (tma_now && G == 8) {
        if (p.dw == 12) e = dit ? launch_strided_tma_k<8, true, false, 12>(p, mode, (int)grid, st) : launch_strided_tma_k<8, false, false, 12>(p, mode, (int)grid, st);
        else if (p.dw == 14) e = dit ? launch_strided_tma_k<8, true, false, 14>(p, mode, (int)grid, st) : launch_strided_tma_k<8, false, false, 14>(p, mode, (int)grid, st);
        else if (!dit) e = dw16 ? launch_strided_tma_k<8, false, true>(p, mode, (int)grid, st) : launch_strided_tma_k<8, false, false>(p, mode, (int)grid, st);
        else e = dw16 ? launch_strided_tma_k<8, true, true>(p, mode, (int)grid, st) : launch_strided_tma_k<8, true, false>(p, mode, (int)grid, st);
    } else if (tma_now && G == 4) {
        if (p.dw == 12) e = dit ? launch_strided_tma_k<4, true, false, 12>(p, mode, (int)grid, st) : launch_strided_tma_k<4, false, false, 12>(p, mode, (int)grid, st);
        else if (p.dw == 14) e = dit ? launch_strided_tma_k<4, true, false, 14>(p, mode, (int)grid, st) : launch_strided_tma_k<4, false, false, 14>(p, mode, (int)grid, st);
        else if (!dit) e = dw16 ? launch_strided_tma_k<4, false, true>(p, mode, (int)grid, st) : launch_strided_tma_k<4, false, false>(p, mode, (int)grid, st);
        else e = dw16 ? launch_strided_tma_k<4, true, true>(p, mode, (int)grid, st) : launch_strided_tma_k<4, true, false>(p, mode, (int)grid, st);
    }
    if (e != cudaErrorNotSupported) {
        // launched (or failed for a real reason) on the TMA-staged kernel; cudaErrorNotSupported = no tensor map could be
        // encoded (driver without cuTensorMapEncodeTiled): the cp.async / STG kernels below do the same work
    } else if (G == 4) {
        if (!dit) e = dw16 ? launch_strided_k<4, false, true>(p, mode, (int)grid, st) : launch_strided_k<4, false, false>(p, mode, (int)grid, st);
        else e = dw16 ? launch_strided_k<4, true, true>(p, mode, (int)grid, st) : launch_strided_k<4, true, false>(p, mode, (int)grid, st);
    } else if (G == 8) {
        if (!dit) e = dw16 ? launch_strided_k<8, false, true>(p, mode, (int)grid, st) : launch_strided_k<8, false, false>(p, mode, (int)grid, st);
        else e = dw16 ? launch_strided_k<8, true, true>(p, mode, (int)grid, st) : launch_strided_k<8, true, false>(p, mode, (int)grid, st);
    } else {
        e = cudaErrorInvalidValue;
    }
    count_launch();
    return (int)e;
}

#endif  // FAST16_PART == 2
#if FAST16_PART == 1
int launch_fast16(const PassDesc &pd, int mode, bool dit, const int2 *twp, const int *lw_r, const int *lw_i,
                  int num_sms, void *stream)
{
    Fast16Params p{};
    p.in = reinterpret_cast<const uint32_t *>(pd.kp.in);
    p.out = reinterpret_cast<uint32_t *>(pd.kp.out);
    p.twp = twp;
    p.total = pd.kp.total;
    p.n_tiles = (pd.kp.total + 4095) >> 12;
    p.dw = pd.kp.dw;
    p.sh_full = 32 - p.dw;
    p.sh_half = 33 - p.dw;
    for (int i = 0; i < 16; ++i) { p.lw_r[i] = lw_r[i]; p.lw_i[i] = lw_i[i]; }
    long long grid = ((pd.kp.g >= 9 && pd.kp.g <= 12) ? 3ll : 2ll) * num_sms;   // three-round schedules: three CTAs per SM
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (grid < 1) grid = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool dw16 = p.dw == 16;
    cudaError_t e;
    switch (pd.kp.g) {          // stage bits of this (contiguous) pass; == NFFT for single-pass plans
    case 3: e = launch_n<3>(p, mode, dit, dw16, (int)grid, st); break;
    case 4: e = launch_n<4>(p, mode, dit, dw16, (int)grid, st); break;
    case 5: e = launch_n<5>(p, mode, dit, dw16, (int)grid, st); break;
    case 6: e = launch_n<6>(p, mode, dit, dw16, (int)grid, st); break;
    case 7: e = launch_n<7>(p, mode, dit, dw16, (int)grid, st); break;
    case 8: e = launch_n<8>(p, mode, dit, dw16, (int)grid, st); break;
    case 9: e = launch_n<9>(p, mode, dit, dw16, (int)grid, st); break;
    case 10: e = launch_n<10>(p, mode, dit, dw16, (int)grid, st); break;
    case 11: e = launch_n<11>(p, mode, dit, dw16, (int)grid, st); break;
    case 12:
        if (pd.natural && !dit) e = dw16 ? launch_k<12, false, true, true>(p, mode, (int)grid, st) : launch_k<12, false, false, true>(p, mode, (int)grid, st);
        else e = launch_n<12>(p, mode, dit, dw16, (int)grid, st);
        break;
    default: e = cudaErrorInvalidValue; break;
    }
    count_launch();
    return (int)e;
}

#endif  // FAST16_PART == 1
#if FAST16_PART == 2
// one-pass 16384-point packed-16 plan (kp.g == 14)
int launch_fast16_n14(const PassDesc &pd, int mode, bool dit, const int2 *twp, const int *lw_r, const int *lw_i,
                      int num_sms, void *stream)
{
    Fast16Params p{};
    p.in = reinterpret_cast<const uint32_t *>(pd.kp.in);
    p.out = reinterpret_cast<uint32_t *>(pd.kp.out);
    p.twp = twp;
    p.total = pd.kp.total;
    p.n_tiles = pd.kp.total >> 14;                  // frames
    p.dw = pd.kp.dw;
    p.sh_full = 32 - p.dw;
    p.sh_half = 33 - p.dw;
    for (int i = 0; i < 16; ++i) { p.lw_r[i] = lw_r[i]; p.lw_i[i] = lw_i[i]; }
    long long grid = 2ll * num_sms;
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (grid < 1) grid = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool dw16 = p.dw == 16;
    cudaError_t e;
    if (p.dw == 12) e = dit ? launch_n14_k<true, false, 12>(p, mode, (int)grid, st) : launch_n14_k<false, false, 12>(p, mode, (int)grid, st);
    else if (p.dw == 14) e = dit ? launch_n14_k<true, false, 14>(p, mode, (int)grid, st) : launch_n14_k<false, false, 14>(p, mode, (int)grid, st);
    else if (!dit) e = dw16 ? launch_n14_k<false, true>(p, mode, (int)grid, st) : launch_n14_k<false, false>(p, mode, (int)grid, st);
    else e = dw16 ? launch_n14_k<true, true>(p, mode, (int)grid, st) : launch_n14_k<true, false>(p, mode, (int)grid, st);
    count_launch();
    return (int)e;
}

// one-pass 8192-point packed-16 plan (kp.g == 13)
int launch_fast16_n13(const PassDesc &pd, int mode, bool dit, const int2 *twp, const int *lw_r, const int *lw_i,
                      int num_sms, void *stream)
{
    Fast16Params p{};
    p.in = reinterpret_cast<const uint32_t *>(pd.kp.in);
    p.out = reinterpret_cast<uint32_t *>(pd.kp.out);
    p.twp = twp;
    p.total = pd.kp.total;
    p.n_tiles = pd.kp.total >> 13;                  // frames
    p.dw = pd.kp.dw;
    p.sh_full = 32 - p.dw;
    p.sh_half = 33 - p.dw;
    for (int i = 0; i < 16; ++i) { p.lw_r[i] = lw_r[i]; p.lw_i[i] = lw_i[i]; }
    long long grid = 3ll * num_sms;
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (grid < 1) grid = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool dw16 = p.dw == 16;
    cudaError_t e;
    if (p.dw == 12) e = dit ? launch_n13_k<true, false, 12>(p, mode, (int)grid, st) : launch_n13_k<false, false, 12>(p, mode, (int)grid, st);
    else if (p.dw == 14) e = dit ? launch_n13_k<true, false, 14>(p, mode, (int)grid, st) : launch_n13_k<false, false, 14>(p, mode, (int)grid, st);
    else if (!dit) e = dw16 ? launch_n13_k<false, true>(p, mode, (int)grid, st) : launch_n13_k<false, false>(p, mode, (int)grid, st);
    else e = dw16 ? launch_n13_k<true, true>(p, mode, (int)grid, st) : launch_n13_k<true, false>(p, mode, (int)grid, st);
    count_launch();
    return (int)e;
}

#endif  // FAST16_PART == 2
#if FAST16_PART == 1
// f2: int_fftNk -> int_ifftNk of a packed-16 plan (2^8 .. 2^12 points, both cores on) as ONE kernel
bool fast16_pair_supported(const intfft_generics &g)
{
    return fast16_supported(g) && g.nfft_log2 >= 8 && g.nfft_log2 <= 12;
}

template <int NLOG2>
static cudaError_t launch_pair_n(const Fast16Params &p, int mode, bool dw16, int grid, cudaStream_t st)
{
    return dw16 ? launch_k<NLOG2, false, true, false, true>(p, mode, grid, st) : launch_k<NLOG2, false, false, false, true>(p, mode, grid, st);
}

int launch_fast16_pair(const PassDesc &pd, int mode, const int2 *twp, const int *lw_r, const int *lw_i, int num_sms,
                       void *stream)
{
    Fast16Params p{};
    p.in = reinterpret_cast<const uint32_t *>(pd.kp.in);
    p.out = reinterpret_cast<uint32_t *>(pd.kp.out);
    p.twp = twp;
    p.total = pd.kp.total;
    p.n_tiles = (pd.kp.total + 4095) >> 12;
    p.dw = pd.kp.dw;
    p.sh_full = 32 - p.dw;
    p.sh_half = 33 - p.dw;
    for (int i = 0; i < 16; ++i) { p.lw_r[i] = lw_r[i]; p.lw_i[i] = lw_i[i]; }
    long long grid = ((pd.kp.g >= 9 && pd.kp.g <= 12) ? 3ll : 2ll) * num_sms;   // three-round schedules: three CTAs per SM
    if (grid > p.n_tiles) grid = p.n_tiles;
    if (grid < 1) grid = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool dw16 = p.dw == 16;
    cudaError_t e;
    switch (pd.kp.g) {
    case 8: e = launch_pair_n<8>(p, mode, dw16, (int)grid, st); break;
    case 9: e = launch_pair_n<9>(p, mode, dw16, (int)grid, st); break;
    case 10: e = launch_pair_n<10>(p, mode, dw16, (int)grid, st); break;
    case 11: e = launch_pair_n<11>(p, mode, dw16, (int)grid, st); break;
    case 12: e = launch_pair_n<12>(p, mode, dw16, (int)grid, st); break;
    default: e = cudaErrorInvalidValue; break;
    }
    count_launch();
    return (int)e;
}

#endif  // FAST16_PART == 1
}  // namespace intfft
