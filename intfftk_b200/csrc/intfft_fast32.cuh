// Specialised stage-chain kernels for every plan whose values fit 32-bit lanes (all three output
// modes, both directions, TWDL_WIDTH 8..27, single- and double-DSP multiplier arrangements):
// BASELINE c5 (8192-pt 18-bit DIT), ROUNDING / UNSCALED 16-bit plans, 17..27-bit scaled plans.
//
// Same structure as the packed-16 kernels (intfft_fast16.cu): 4096-sample tiles, 256 threads,
// 16 samples per thread for 4 consecutive stages, twiddles hoisted (registers for the top round, a
// small shared table for the middle round, kernel parameters for the lowest round), one CTA barrier
// per tile.  Differences: samples are {re:int32, im:int32} in shared memory (no pack / unpack),
// products are 64-bit (IMAD.WIDE + funnel shift + SGXT), the multiplier arrangement of a stage is a
// grid-uniform run-time choice, and NFFT >= 13 is a strided top pass + a contiguous pass.
//
// Reference rules implemented: int_dif2_fly.vhd:142-373, int_dit2_fly.vhd:140-325,
// int_cmult_dsp48.vhd:182-190 / 307-317 (single), int_cmult_dbl18_dsp48.vhd:163-181 and
// int_cmult_dbl35_dsp48.vhd:155-168 (double; its 48-bit wrap cannot reach the kept bits when the
// result is <= 32 bits wide, so it is not materialised here).
#pragma once
#include <cuda_runtime.h>

#include "intfft_arith.cuh"
#include "intfft_taylor.cuh"
#include "intfft_tma.cuh"

namespace intfft {

namespace f32 {

struct Fast32Params {
    const void *in;
    void *out;
    const int2 *tw;          // raw twiddles, entry (1 << s) + k
    const unsigned *tw16;    // KIND_SINGLE_PRE: STAGE-12 twiddles packed {re:16 | im:16} (TWDL_WIDTH <= 16), entry k
    long long n_tiles;       // contiguous: tiles of 4096 samples
    long long total;         // frames * N
    long long batch;
    int n;                   // NFFT of the whole transform
    int dw, format;
    int in_sb, out_sb;       // scalar bytes of the containers read / written: 2 or 4
    int in_wrap;
    long long n_units;       // strided pass only: work items = column blocks per frame * batch
    CmultConsts cm;
    int lw_r[16], lw_i[16];  // lowest-round twiddles, index (1 << s) - 1 + k, s = 2, 3
    TaylorDev tay;           // strided pass: .on = STAGE >= 11 twiddles recomputed on the device, tw ends at STAGE 11
};

// element (8-byte) index inside a 4096-sample tile -> slot in the padded exchange tile; additive for
// disjoint bit sets, conflict-free for LDS.64 at stride 256 / 16 and for LDS.128 on 16 contiguous samples
__host__ __device__ constexpr unsigned phys8(unsigned i) { return i + 2u * (i >> 4); }
constexpr unsigned kTile8 = 4608;
constexpr unsigned kHead32 = 128 + 15 * 16 * 8;
constexpr unsigned kStage32 = 256 * 9 * 16;   // prefetch staging: 16 slots x 256 threads x 8 bytes, or 256 x (8 + 1 pad) x 16 bytes

// one scalar of a sample: the value and (TRUNCATE only; dead code elsewhere) its floor-half, which is
// all a TRUNCATE butterfly ever reads (inputs sliced (DTW-1 downto 1), int_dif2_fly.vhd:150-153)
struct V {
    int f, h;
};
__device__ __forceinline__ V mk(int f) { return V{f, f >> 1}; }

// multiplier arrangement policy of a kernel instance: every stage single-DSP, or chosen per stage
// KIND_SINGLE_PRE (TRUNCATE only): every stage single-DSP AND the twiddles arrive pre-shifted by
// e = 31 - sh_single (32 - TWDL_WIDTH below 19 bits), so that the multiplier's output slice starts at bit 32 of the
// 64-bit sum of products: the floor-half a TRUNCATE butterfly consumes is the HIGH WORD of the IMAD.WIDE chain plus
// one SGXT, instead of a funnel shift + an arithmetic shift (two ALU-port instructions fewer per butterfly, and a
// shorter dependency chain).  W << e always fits a 32-bit operand: |W| < 2^(TWDL_WIDTH-1).
enum { KIND_SINGLE = 0, KIND_MIXED = 1, KIND_SINGLE_PRE = 2 };

struct Stg {
    int s, ow, dtwc, kind;
    int k, sp;               // KIND_MIXED: pre-shift of each product and post-shift of their sum (grid-uniform)
};
template <bool DIT, int MODE, int KIND>
__device__ __forceinline__ Stg stage_of(const Fast32Params &p, int s)
{
    constexpr int FORMAT = MODE == MODE_UNSCALED ? 1 : 0;
    Stg st;
    st.s = s;
    const int ii = DIT ? s : p.n - 1 - s;
    const int dtw = p.dw + ii * FORMAT;
    st.ow = dtw + FORMAT;
    st.dtwc = DIT ? dtw : st.ow;
    st.kind = (KIND != KIND_MIXED) ? 0 : (st.dtwc < p.cm.lim_single ? 0 : 1);
    // the single arrangement is the double one with no pre-shift: (P2 +- P1) >> sh == ((P2 >> 0) +- (P1 >> 0)) >> sh,
    // so a pass that mixes both runs ONE branch-free code path with per-stage shift amounts
    st.k = st.kind ? p.cm.k_pre : 0;
    st.sp = st.kind ? p.cm.sh_post : p.cm.sh_single;
    return st;
}

__device__ __forceinline__ int negq32(int v) { return (v >> 31) - v; }

// bits [sh+w-1 : sh] of a 64-bit value, sign-extended: funnel-left by 64-sh-w (high word), then an
// arithmetic right shift by 32-w  (bfe.s32 with a register length costs three instructions instead)
__device__ __forceinline__ int field(long long t, int sh, int w)
{
    const int hi = (int)(((unsigned long long)t << (64 - sh - w)) >> 32);
    return hi >> (32 - w);
}
// low w bits of a 32-bit value, sign-extended
__device__ __forceinline__ int sx(int v, int w) { return sgxt32(v, w); }      // one SGXT (intfft_arith.cuh)

// exact signed 32 x 32 -> 64 (asm: next to word-wise shifts the front end otherwise emits IMAD.WIDE.U32 + fix-ups)
__device__ __forceinline__ long long mulw(int a, int b)
{
    long long r;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
// t >> k for 0 <= k < 32 (every pre-shift of the double arrangements is below 32)
// (register pairs are split / joined with mov.b64: built from shifts and ORs, the front end no longer sees
// a plain pair and turns every 64-bit add that follows into five or six instructions)
__device__ __forceinline__ long long sra64(long long t, int k)
{
    unsigned lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
    const unsigned rlo = __funnelshift_r(lo, hi, k);
    const int rhi = (int)hi >> k;
    long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(rlo), "r"(rhi));
    return r;
}

template <int MODE, int KIND>
__device__ __forceinline__ void cmul32(int dr, int di, int wr, int wi, const CmultConsts &cm, const Stg &st,
                                       V &o_re, V &o_im)
{
    if (KIND == KIND_MIXED) {                   // double (or single as its k = 0 case), int_cmult_dbl18/dbl35
        const long long tr = sra64(mulw(dr, wr), st.k) - sra64(mulw(di, wi), st.k);
        const long long ti = sra64(mulw(dr, wi), st.k) + sra64(mulw(di, wr), st.k);
        o_re.f = field(tr, st.sp, st.dtwc);
        o_im.f = field(ti, st.sp, st.dtwc);
        o_re.h = MODE == MODE_TRUNC ? field(tr, st.sp + 1, st.dtwc - 1) : 0;
        o_im.h = MODE == MODE_TRUNC ? field(ti, st.sp + 1, st.dtwc - 1) : 0;
        return;
    }
    const long long tr = (long long)dr * wr - (long long)di * wi;      // 2 x IMAD.WIDE
    const long long ti = (long long)dr * wi + (long long)di * wr;
    if (KIND == KIND_SINGLE_PRE) {               // MODE_TRUNC; wr / wi carry the factor 2^(31 - sh_single)
        unsigned rl, rh, il, ih;
        asm("mov.b64 {%0, %1}, %2;" : "=r"(rl), "=r"(rh) : "l"(tr));
        asm("mov.b64 {%0, %1}, %2;" : "=r"(il), "=r"(ih) : "l"(ti));
        o_re.h = sx((int)rh, st.dtwc - 1);       // wrap_dtwc(P >> sh) >> 1 == wrap_(dtwc-1)(P >> (sh + 1))
        o_im.h = sx((int)ih, st.dtwc - 1);
        o_re.f = sx((int)__funnelshift_l(rl, rh, 1), st.dtwc);     // only where a consumer needs the full value
        o_im.f = sx((int)__funnelshift_l(il, ih, 1), st.dtwc);
        return;
    }
    const int sh = cm.sh_single;
    o_re.f = field(tr, sh, st.dtwc);
    o_im.f = field(ti, sh, st.dtwc);
    if (MODE == MODE_TRUNC) {
        o_re.h = field(tr, sh + 1, st.dtwc - 1);
        o_im.h = field(ti, sh + 1, st.dtwc - 1);
    } else {
        o_re.h = o_im.h = 0;
    }
}

// inline PTX pins the instruction selection: ptxas fuses shr + add into LEA.HI.SX32 and keeps the
// multiply-subtract on the IMAD port (the front end would otherwise re-derive two shifts and two adds)
__device__ __forceinline__ int sra1(int x) { int r; asm("shr.s32 %0, %1, 1;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ int msub2(int t, int x) { int r; asm("mad.lo.s32 %0, %1, -2, %2;" : "=r"(r) : "r"(t), "r"(x)); return r; }

// a_half: the A operand is the product of a multiplying stage of this round (DIF only), whose floor-half is what the
// pre-shifted multiplier delivers directly — reading a.h then spares the product's full-value slice (a funnel shift and
// a sign extension) that sra1(a.f) would drag in.  Resolved at compile time at every call site.
template <int MODE> __device__ __forceinline__ void addsub32(const V &a, const V &b, int ow, V &x, V &y, bool a_half = false)
{
    int xf, yf;
    if (MODE == MODE_TRUNC) {                   // (A>>1) + (B>>1) as one shift-add; (A>>1) - (B>>1) = sum - 2 (B>>1)
        xf = (a_half ? a.h : sra1(a.f)) + b.h;  // so only the B operand's half is ever materialised
        yf = msub2(b.h, xf);                    // (as two subtractions ptxas re-balances the ports itself: IMAD.IADD for the
                                                // adds, separate shifts instead of LEA.HI — 13 % more instructions; measured r02)
    } else if (MODE == MODE_ROUND) {            // (v >> 1) + v(0) == (v + 1) >> 1
        xf = (int)((unsigned)a.f + (unsigned)b.f + 1u) >> 1;
        // (a - b + 1) >> 1 == ((a + b + 1) >> 1) - b exactly (ROUNDING plans keep a spare bit, so the sum cannot
        // overflow the lane); the rounded difference can reach 2^(ow-1) and is kept in ow bits by the reference
        yf = sx((int)((unsigned)xf - (unsigned)b.f), ow);
    } else {
        xf = (int)((unsigned)a.f + (unsigned)b.f);
        yf = (int)((unsigned)a.f - (unsigned)b.f);
    }
    x = mk(xf);
    y = mk(yf);
}

// MUL: the caller knows st.s >= 2 (every stage of a strided top pass), so the multiplier-free STAGE 0 / 1
// forms are not even compiled in and the butterflies of a round form one basic block
template <bool DIT, int MODE, int KIND, bool MUL = false>
__device__ __forceinline__ void fly32(const Stg &st, bool odd, const CmultConsts &cm, V &ar, V &ai, V &br, V &bi,
                                      int wr, int wi, bool a_half = false)
{
    if (!DIT) {
        V xr, xi, sr, si;
        addsub32<MODE>(ar, br, st.ow, xr, sr, a_half);
        addsub32<MODE>(ai, bi, st.ow, xi, si, a_half);
        ar = xr;
        ai = xi;
        if (!MUL && st.s == 0) {
            br = sr;
            bi = si;
        } else if (!MUL && st.s == 1) {
            br = odd ? si : sr;
            bi = odd ? mk(negq32(sr.f)) : si;
        } else {
            cmul32<MODE, KIND>(sr.f, si.f, wr, wi, cm, st, br, bi);
        }
    } else {
        V wr_, wi_;                                   // BW
        if (!MUL && st.s == 0) {
            wr_ = br;
            wi_ = bi;
        } else if (!MUL && st.s == 1) {
            wr_ = odd ? mk(negq32(bi.f)) : br;
            wi_ = odd ? br : bi;
        } else {                                      // DI_RE <= IB_IM, DI_IM <= IB_RE; DO_RE => bw_im, DO_IM => bw_re
            V o_re, o_im;
            cmul32<MODE, KIND>(bi.f, br.f, wr, wi, cm, st, o_re, o_im);
            wi_ = o_re;
            wr_ = o_im;
        }
        V xr, xi, yr, yi;
        addsub32<MODE>(ar, wr_, st.ow, xr, yr);
        addsub32<MODE>(ai, wi_, st.ow, xi, yi);
        ar = xr; ai = xi;
        br = yr; bi = yi;
    }
}

// ---- per-thread prefetch of the NEXT tile's first-round samples (cp.async into thread-private slots of
// ---- a 32 KB staging area: slot [m][tid]); no barrier is involved because a thread only ever reads the
// ---- slots it filled itself, and one staging buffer is enough because the slots are drained into
// ---- registers at the top of a tile before the copies for the following tile are issued
__device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gsrc, int bytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void stage_read(const unsigned char *stage, unsigned tid, int m, int sb, int &re, int &im)
{
    if (sb == 2) {
        const unsigned x = reinterpret_cast<const unsigned *>(stage)[m * 256 + tid];
        re = (int)(short)(x & 0xffffu);
        im = (int)x >> 16;
    } else {
        const int2 v = reinterpret_cast<const int2 *>(stage)[m * 256 + tid];
        re = v.x;
        im = v.y;
    }
}

struct TwRegs32 {
    const int (&r)[15];
    const int (&i)[15];
    __device__ __forceinline__ void operator()(int w, int &wr, int &wi) const { wr = r[w]; wi = i[w]; }
};
struct TwSmem32 {
    const int2 *t;
    int pitch;
    __device__ __forceinline__ void operator()(int w, int &wr, int &wi) const
    {
        const int2 v = t[w * pitch];
        wr = v.x;
        wi = v.y;
    }
};

// R stages on register bits 0..R-1 (global stage numbers S0 .. S0+R-1) of the 16 resident samples
template <int R, bool DIT, int MODE, int KIND, typename TW, bool MUL = false>
__device__ __forceinline__ void round32(V (&re)[16], V (&im)[16], const Fast32Params &p, int s0, const TW &tw,
                                        bool lo_is_zero, bool tid_odd)
{
#pragma unroll
    for (int step = 0; step < R; ++step) {
        const int q = DIT ? step : R - 1 - step;
        const Stg st = stage_of<DIT, MODE, KIND>(p, s0 + q);
        // A mixed pass still has stages on the single arrangement (c3's first pass: 3 of 8): those take the
        // single code — products accumulated in one IMAD.WIDE chain, no per-product pre-shift, 12 instructions
        // fewer per butterfly — behind a grid-uniform branch around the stage's eight butterflies.
        if (KIND == KIND_MIXED && st.kind == 0) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                if (m & (1 << q)) continue;
                const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
                int wr = 0, wi = 0;
                if (MUL || st.s >= 2) tw(w, wr, wi);
                const bool odd = lo_is_zero ? ((m & 1) != 0) : tid_odd;
                fly32<DIT, MODE, KIND_SINGLE, MUL>(st, odd, p.cm, re[m], im[m], re[m | (1 << q)], im[m | (1 << q)], wr, wi);
            }
            continue;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (m & (1 << q)) continue;
            const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
            int wr = 0, wi = 0;
            if (MUL || st.s >= 2) tw(w, wr, wi);
            const bool odd = lo_is_zero ? ((m & 1) != 0) : tid_odd;
            // DIF: the pair (m, m | 1 << q) holds the PRODUCTS of the round's previous step iff bit q + 1 of m is set
            // (and that step multiplied: STAGE >= 2)
            const bool a_half = !DIT && KIND == KIND_SINGLE_PRE && step > 0 && (m & (2 << q)) != 0 && (MUL || s0 + q + 1 >= 2);
            fly32<DIT, MODE, KIND, MUL>(st, odd, p.cm, re[m], im[m], re[m | (1 << q)], im[m | (1 << q)], wr, wi, a_half);
        }
    }
}

__device__ __forceinline__ void ld_sample(const void *base, long long idx, int sb, int &re, int &im)
{
    if (sb == 2) {
        const unsigned x = __ldg(reinterpret_cast<const unsigned *>(base) + idx);
        re = (int)(short)(x & 0xffffu);
        im = (int)x >> 16;
    } else {
        const int2 v = __ldg(reinterpret_cast<const int2 *>(base) + idx);
        re = v.x;
        im = v.y;
    }
}
__device__ __forceinline__ void st_sample(void *base, long long idx, int sb, int re, int im)
{
    if (sb == 2) reinterpret_cast<unsigned *>(base)[idx] = __byte_perm((unsigned)re, (unsigned)im, 0x5410);
    else reinterpret_cast<int2 *>(base)[idx] = make_int2(re, im);
}

// ------------------------------------------------------------------------------------------------
// contiguous pass: stage bits 0 .. NLOG2-1 of an NFFT = p.n transform (NLOG2 == p.n for one-pass plans)
template <int NLOG2, bool DIT, int MODE, int KIND>
__global__ void __launch_bounds__(256, 2) fast32_kernel(const __grid_constant__ Fast32Params p)
{
    constexpr int R0 = ((NLOG2 - 1) % 4) + 1;
    constexpr int NR = 1 + (NLOG2 - R0) / 4;
    static_assert(NR >= 1 && NR <= 3, "supported: 2^3 .. 2^12 points per tile");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);                  // [15][1 << R0]
    int2(*work)[kTile8] = reinterpret_cast<int2(*)[kTile8]>(smem_raw + kHead32);
    unsigned char *stage = smem_raw + kHead32 + 2 * kTile8 * 8;

    const unsigned tid = threadIdx.x;
    const bool tid_odd = tid & 1u;
    const int esz = 2 * p.in_sb;                                             // bytes per complex sample read

    // DIF, two or three rounds: the tile lands as ONE bulk TMA copy (cp.async.bulk + mbarrier, dense, read at stride 256) instead
    // of 16 element-sized cp.async per thread; the next tile's copy is issued behind the tile's first CTA barrier, when
    // every thread has drained the landing area (single-buffered, like the packed-16 NAT variant)
    constexpr bool TMAIN = !DIT && NR >= 2;      // (a one-round DIF has no CTA barrier to issue the next copy behind)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    unsigned tma_phase = 0;
    if (TMAIN) {
        if (tid == 0) { tma::mbar_init(bar, 1); tma::mbar_fence_init(); }
        __syncthreads();
    }

    // first-round ownership (the same for every tile): local index of register m
    constexpr int RF = DIT ? 0 : NR - 1;                                      // first round processed
    constexpr int LOF = RF == 0 ? 0 : R0 + 4 * (RF - 1);
    constexpr int RRF = RF == 0 ? R0 : 4;
    const unsigned basef = (tid & ((1u << LOF) - 1u)) | ((tid >> LOF) << (LOF + RRF));
    // In the lowest round a warp owns 16 >> R0 runs of 32 << R0 contiguous samples (a thread: 1 << R0 contiguous
    // samples of each run).  Read or written per thread, a warp instruction would touch up to 32 lines, so both the
    // DIT input and the DIF output move per WARP: 16-byte piece q = lane + 32 j of the warp's 256 pieces (two
    // int32 samples each) is at tile-local sample warp_piece(j), 512 contiguous bytes per instruction.
    const unsigned lane = tid & 31u, wbase = tid & ~31u;
    auto warp_piece = [&](int j) {
        const unsigned q = lane + 32u * j;
        return (wbase << R0) + ((q >> (4 + R0)) << (8 + R0)) + 2u * (q & ((16u << R0) - 1u));
    };
    // DIT input lands (cp.async, one tile ahead) in a tile-skewed copy of the block (32-bit containers, every R0),
    // or, for packed 16-bit input with a 4-stage first round, in a thread-major table with a 9-slot pitch
    const bool pieces = (DIT || NR == 1) && (R0 == 4 || p.in_sb == 4);   // (a one-round DIF starts in the lowest round too)
    int2 *land = reinterpret_cast<int2 *>(stage);
    auto prefetch = [&](long long t) {
        if (TMAIN) {
            if (tid == 0) {
                const unsigned bytes = 4096u * esz;
                tma::mbar_expect_tx(bar, bytes);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(tma::smem_u32(stage)), "l"(reinterpret_cast<const char *>(p.in) + (t << 12) * esz), "r"(bytes),
                               "r"(tma::smem_u32(bar)) : "memory");
            }
            return;
        }
        const char *src = reinterpret_cast<const char *>(p.in) + ((t << 12) + basef) * esz;
        if (pieces) {
            const char *wsrc = reinterpret_cast<const char *>(p.in) + ((t << 12) + (wbase << 4)) * esz + 16u * lane;
            int4 *st16 = reinterpret_cast<int4 *>(stage);
            if (p.in_sb == 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned i = warp_piece(j);
                    const unsigned a = (unsigned)__cvta_generic_to_shared(land + phys8(i));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(reinterpret_cast<const int2 *>(p.in) + (t << 12) + i) : "memory");
                }
            } else {                                       // four packed samples per piece: owner k >> 2, slot k & 3
                int4 *d = st16 + (wbase + (lane >> 2)) * 9 + (lane & 3u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned a = (unsigned)__cvta_generic_to_shared(d + 72 * j);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(wsrc + 512 * j) : "memory");
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned off = ((unsigned)(m & ((1 << RRF) - 1)) << LOF) | ((unsigned)(m >> RRF) << (8 + RRF));
                cp_async_elem(stage + (m * 256 + tid) * esz, src + (long long)off * esz, esz);
            }
        }
        cp_async_commit();
    };
    bool staged = false;
    if ((long long)blockIdx.x < p.n_tiles && (((long long)blockIdx.x + 1) << 12) <= p.total) {
        prefetch(blockIdx.x);
        staged = true;
    }

    // ---- batch-invariant twiddles: round 1 -> shared table, round 2 -> registers ----
    for (unsigned e = tid; e < 15u << R0; e += 256) {
        const int w = e >> R0, low = e & ((1u << R0) - 1u);
        const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
        const int j = w - ((1 << q) - 1);
        midtw[e] = __ldg(p.tw + (1u << (R0 + q)) + low + ((unsigned)j << R0));
    }
    int uwr[15], uwi[15];
    if (NR == 3) {
        const int lo = R0 + 4;
        const unsigned low = tid & ((1u << lo) - 1u);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int2 w = __ldg(p.tw + (1u << (lo + q)) + low + ((unsigned)j << lo));
                uwr[(1 << q) - 1 + j] = w.x;
                uwi[(1 << q) - 1 + j] = w.y;
            }
    }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }
    __syncthreads();

    int it = 0;
    for (long long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        int2 *sm = work[it & 1];
        const long long g0 = tile << 12;
        const bool full = g0 + 4096 <= p.total;
        V re[16], im[16];
#pragma unroll
        for (int rr = 0; rr < NR; ++rr) {
            const int r = DIT ? rr : NR - 1 - rr;
            const int lo = r == 0 ? 0 : R0 + 4 * (r - 1);
            const int R = r == 0 ? R0 : 4;
            const bool first = rr == 0, last = rr == NR - 1;
            const unsigned base = (tid & ((1u << lo) - 1u)) | ((tid >> lo) << (lo + R));
            const unsigned pbase = phys8(base);

            if (TMAIN && first && staged) {           // this tile was landed by the TMA engine: dense, element base + off
                tma::mbar_wait(bar, tma_phase);
                tma_phase ^= 1u;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned e = base + (((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R)));
                    int a, b;
                    if (p.in_sb == 2) {
                        const unsigned x = reinterpret_cast<const unsigned *>(stage)[e];
                        a = (int)(short)(x & 0xffffu);
                        b = (int)x >> 16;
                    } else {
                        const int2 v = reinterpret_cast<const int2 *>(stage)[e];
                        a = v.x;
                        b = v.y;
                    }
                    if (p.in_wrap) { a = sx(a, p.dw); b = sx(b, p.dw); }
                    re[m] = mk(a);
                    im[m] = mk(b);
                }
                staged = false;                       // (re-armed behind the first barrier below)
            } else if (first && staged) {             // this tile was prefetched into the thread's slots
                cp_async_wait_all();
                if (pieces) {
                    __syncwarp();                          // the other lanes' pieces of this warp's block have landed too
                    const int4 *st16 = reinterpret_cast<const int4 *>(stage) + tid * 9;
                    int a[16], b[16];
                    if (p.in_sb == 4 && R0 == 4) {
                        const int4 *own = reinterpret_cast<const int4 *>(land + pbase);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int4 v = own[j];
                            a[2 * j] = v.x; b[2 * j] = v.y; a[2 * j + 1] = v.z; b[2 * j + 1] = v.w;
                        }
                    } else if (p.in_sb == 4) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                            const int2 v = land[pbase + phys8(off)];
                            a[m] = v.x; b[m] = v.y;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int4 v = st16[j];
                            const int x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) { a[4 * j + e] = (int)(short)(x[e] & 0xffff); b[4 * j + e] = x[e] >> 16; }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        if (p.in_wrap) { a[m] = sx(a[m], p.dw); b[m] = sx(b[m], p.dw); }
                        re[m] = mk(a[m]);
                        im[m] = mk(b[m]);
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        int a, b;
                        stage_read(stage, tid, m, p.in_sb, a, b);
                        if (p.in_wrap) { a = sx(a, p.dw); b = sx(b, p.dw); }
                        re[m] = mk(a);
                        im[m] = mk(b);
                    }
                }
                const long long nt = tile + gridDim.x;    // refill the slots with this CTA's next tile
                staged = nt < p.n_tiles && ((nt + 1) << 12) <= p.total;
                if (pieces) __syncwarp();                 // every lane has drained its slots of the warp's block
                if (staged) prefetch(nt);
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                    int a = 0, b = 0;
                    if (first) {
                        if ((g0 + base + off) < p.total) {
                            ld_sample(p.in, g0 + base + off, p.in_sb, a, b);
                            if (p.in_wrap) { a = sx(a, p.dw); b = sx(b, p.dw); }
                        }
                    } else {
                        const int2 v = sm[pbase + phys8(off)];
                        a = v.x;
                        b = v.y;
                    }
                    re[m] = mk(a);
                    im[m] = mk(b);
                }
            }

            if (r == 0) round32<R0, DIT, MODE, KIND>(re, im, p, 0, TwRegs32{lwr, lwi}, true, tid_odd);
            else if (r == 1) round32<4, DIT, MODE, KIND>(re, im, p, R0, TwSmem32{midtw + (tid & ((1u << R0) - 1u)), 1 << R0}, false, tid_odd);
            else round32<4, DIT, MODE, KIND>(re, im, p, R0 + 4, TwRegs32{uwr, uwi}, false, tid_odd);

            if (last && full && (!DIT || NR == 1) && p.out_sb == 4) {
                // DIF results go back into the thread's own tile slots (the ones it read for this round) and
                // leave per warp, 512 contiguous bytes per instruction (up to the run length)
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                    sm[pbase + phys8(off)] = make_int2(re[m].f, im[m].f);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const unsigned i = warp_piece(j);
                    *reinterpret_cast<int4 *>(reinterpret_cast<int2 *>(p.out) + g0 + i) = *reinterpret_cast<const int4 *>(sm + phys8(i));
                }
            } else if (last && full && (!DIT || NR == 1)) {
                // packed 16-bit output (16-bit data with TWDL_WIDTH > 16): the packed word takes the first half of
                // the thread's own slot; the warp then gathers four of them per 16-byte store
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                    sm[pbase + phys8(off)].x = (int)__byte_perm((unsigned)re[m].f, (unsigned)im[m].f, 0x5410);
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned q = lane + 32u * j;     // piece of four samples among the warp's 128
                    const unsigned i = (wbase << R0) + ((q >> (3 + R0)) << (8 + R0)) + 4u * (q & ((8u << R0) - 1u));
                    const int4 a = *reinterpret_cast<const int4 *>(sm + phys8(i)), b = *reinterpret_cast<const int4 *>(sm + phys8(i) + 2);
                    *reinterpret_cast<int4 *>(reinterpret_cast<unsigned *>(p.out) + g0 + i) = make_int4(a.x, a.z, b.x, b.z);
                }
            } else if (last && full) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                    st_sample(p.out, g0 + base + off, p.out_sb, re[m].f, im[m].f);
                }
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const unsigned off = ((unsigned)(m & ((1 << R) - 1)) << lo) | ((unsigned)(m >> R) << (8 + R));
                    if (last) {
                        if ((g0 + base + off) < p.total) st_sample(p.out, g0 + base + off, p.out_sb, re[m].f, im[m].f);
                    } else {
                        sm[pbase + phys8(off)] = make_int2(re[m].f, im[m].f);
                    }
                }
            }
            if (!last) {
                const bool warp_local = (NR == 3 && R0 == 4) && ((DIT && rr == 0) || (!DIT && rr == 1));
                if (warp_local) __syncwarp();
                else __syncthreads();
                if (TMAIN && rr == 0) {               // every thread has drained the landing area: next tile, if whole
                    const long long nt = tile + gridDim.x;
                    staged = nt < p.n_tiles && ((nt + 1) << 12) <= p.total;
                    if (staged) prefetch(nt);
                }
            }
        }
    }
}

// the 16 stores of a strided pass's last round: register m goes 2^SHIFT m rows below register 0; with the
// row pitch a compile-time constant (PB = NFFT - G) every offset is an instruction immediate, and the
// container size is tested once, not per sample
template <int PB, int SHIFT>
__device__ __forceinline__ void store_rows32(char *ptr, int sb, const V (&re)[16], const V (&im)[16])
{
    if (sb == 2) {
#pragma unroll
        for (int m = 0; m < 16; ++m)
            *reinterpret_cast<unsigned *>(ptr + ((((size_t)m << SHIFT) << PB) << 2)) = __byte_perm((unsigned)re[m].f, (unsigned)im[m].f, 0x5410);
    } else {
#pragma unroll
        for (int m = 0; m < 16; ++m)
            *reinterpret_cast<int2 *>(ptr + ((((size_t)m << SHIFT) << PB) << 3)) = make_int2(re[m].f, im[m].f);
    }
}

// ------------------------------------------------------------------------------------------------
// strided pass: the top G = 4 or 8 stage bits of an NFFT = 13..20 transform (see intfft_fast16.cu)
template <int G, bool DIT, int MODE, int KIND>
__global__ void __launch_bounds__(256, 2) fast32_strided_kernel(const __grid_constant__ Fast32Params p)
{
    constexpr int C = 12 - G;
    constexpr int NR = G / 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    int2(*work)[kTile8] = reinterpret_cast<int2(*)[kTile8]>(smem_raw + kHead32);
    unsigned char *stage = smem_raw + kHead32 + 2 * kTile8 * 8;

    const unsigned tid = threadIdx.x;
    const int esz = 2 * p.in_sb;
    const int pb = p.n - G;
    const unsigned cmask = (1u << C) - 1u;
    const long long row_stride = 1ll << pb;

    int it = 0;
    // work items w = mid * batch + frame; CTA b owns the contiguous range [b T / G, (b + 1) T / G) (see intfft_fast16.cu)
    long long w = p.n_units * blockIdx.x / gridDim.x;
    const long long w_end = p.n_units * (blockIdx.x + 1) / gridDim.x;
    while (w < w_end) {
        const unsigned mid = (unsigned)(w / p.batch);
        const long long f0 = w - (long long)mid * p.batch;
        const long long f1 = (f0 + (w_end - w) < p.batch) ? f0 + (w_end - w) : p.batch;
        w += f1 - f0;
        auto kidx = [&](unsigned l) { return ((l >> C) << pb) | (mid << C) | (l & cmask); };

        int uwr[15], uwi[15];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int sgl = pb + (8 + q - C);
                const int2 w = hoist_twiddle(p.tw, p.tay, sgl, kidx(tid | ((unsigned)j << 8)) & ((1u << sgl) - 1u));
                uwr[(1 << q) - 1 + j] = w.x;
                uwi[(1 << q) - 1 + j] = w.y;
            }
        if (NR == 2) {
            __syncthreads();
            if (tid < 240) {
                const int w = tid >> 4, lo4 = tid & 15;
                const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
                const int j = w - ((1 << q) - 1);
                const int sgl = pb + (4 + q - C);
                midtw[w * 16 + lo4] = hoist_twiddle(p.tw, p.tay, sgl, kidx((unsigned)lo4 | ((unsigned)j << 4)) & ((1u << sgl) - 1u));
            }
            __syncthreads();
        }

        // first-round ownership inside the 2^G x 2^C tile, and the prefetch of one frame's column block
        constexpr int LOF = DIT ? 12 - 4 * NR : 8;
        const unsigned basef = (tid & ((1u << LOF) - 1u)) | ((tid >> LOF) << (LOF + 4));
        auto prefetch = [&](long long f) {
            const char *src = reinterpret_cast<const char *>(p.in) + ((f << p.n) + ((long long)mid << C)) * esz;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const unsigned l = basef | ((unsigned)m << LOF);
                cp_async_elem(stage + (m * 256 + tid) * esz, src + ((long long)(l >> C) * row_stride + (l & cmask)) * esz, esz);
            }
            cp_async_commit();
        };
        if (f0 < f1) prefetch(f0);

        for (long long f = f0; f < f1; ++f, ++it) {
            int2 *sm = work[it & 1];
            const long long gbase = (f << p.n) + ((long long)mid << C);
            V re[16], im[16];
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) {
                const int r = DIT ? rr : NR - 1 - rr;
                const int lo = 12 - 4 * (NR - r);
                const bool first = rr == 0, last = rr == NR - 1;
                const unsigned base = (tid & ((1u << lo) - 1u)) | ((tid >> lo) << (lo + 4));
                const unsigned pbase = phys8(base);
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    int a, b;
                    if (first) {
                        if (m == 0) cp_async_wait_all();
                        stage_read(stage, tid, m, p.in_sb, a, b);
                        if (p.in_wrap) { a = sx(a, p.dw); b = sx(b, p.dw); }
                    } else {
                        const int2 v = sm[pbase + phys8((unsigned)m << lo)];
                        a = v.x;
                        b = v.y;
                    }
                    re[m] = mk(a);
                    im[m] = mk(b);
                }
                if (first && f + 1 < f1) prefetch(f + 1);
                // every stage of this pass is a multiplying one: STAGE >= NFFT - 8 >= 5
                if (lo == 8) round32<4, DIT, MODE, KIND, TwRegs32, true>(re, im, p, pb + 8 - C, TwRegs32{uwr, uwi}, false, false);
                else round32<4, DIT, MODE, KIND, TwSmem32, true>(re, im, p, pb + 4 - C, TwSmem32{midtw + (tid & 15u), 16}, false, false);
                if (last) {
                    constexpr int SHIFT = (G == 8 && DIT) ? 4 : 0;          // lo - C of the last round
                    char *ptr = reinterpret_cast<char *>(p.out) +
                                (gbase + (long long)(base >> C) * row_stride + (base & cmask)) * (2 * p.out_sb);
                    switch (pb) {
                    case 8: store_rows32<8, SHIFT>(ptr, p.out_sb, re, im); break;
                    case 9: store_rows32<9, SHIFT>(ptr, p.out_sb, re, im); break;
                    case 10: store_rows32<10, SHIFT>(ptr, p.out_sb, re, im); break;
                    case 11: store_rows32<11, SHIFT>(ptr, p.out_sb, re, im); break;
                    default: store_rows32<12, SHIFT>(ptr, p.out_sb, re, im); break;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) sm[pbase + phys8((unsigned)m << lo)] = make_int2(re[m].f, im[m].f);
                    __syncthreads();
                }
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Strided pass, G = 8, column block moved by 2-D TMA in both directions (see intfft_tma.cuh; packed-16 twin:
// fast16_strided_tma_kernel).  Shared memory: [head | ONE exchange tile | two dense landing tiles]; a frame's
// landing tile, drained by the first round, takes the last round's results and is the source of the frame's
// tensor store, so a third tile is not needed and two CTAs still fit on an SM (102 KB each).
// A sample is 1 or 2 words (in_sb / out_sb = 2 or 4); dense tile: sample l = 16 row + column at l * (words per sample).
constexpr unsigned kLand32 = 4096 * 8;
constexpr unsigned kStridedTmaSmem32 = kHead32 + kTile8 * 8 + 2 * kLand32;

template <int G, bool DIT, int MODE, int KIND>
__global__ void __launch_bounds__(256, 2) fast32_strided_tma_kernel(const __grid_constant__ Fast32Params p,
                                                                    const __grid_constant__ CUtensorMap map_in,
                                                                    const __grid_constant__ CUtensorMap map_out)
{
    constexpr int C = 12 - G;                       // G = 8: 256 rows x 16 columns; G = 4: 16 rows x 256 columns
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw + 128);
    int2 *work = reinterpret_cast<int2 *>(smem_raw + kHead32);
    unsigned char *land0 = smem_raw + kHead32 + kTile8 * 8;

    const unsigned tid = threadIdx.x;
    const int pb = p.n - G;
    const unsigned cmask = (1u << C) - 1u;
    const int iw = p.in_sb >> 1;                              // words per sample in the input container

    if (tid == 0) {
        tma::mbar_init(&bar[0], 1);
        tma::mbar_init(&bar[1], 1);
        tma::mbar_fence_init();
        tma::prefetch_map(&map_in);
        tma::prefetch_map(&map_out);
    }
    __syncthreads();

    unsigned it = 0, phase = 0;
    auto load = [&](unsigned buf, unsigned mid, long long f) {          // thread 0 only
        tma::mbar_expect_tx(&bar[buf], 4096u * 4u * iw);
        tma::load_2d(land0 + buf * kLand32, &map_in, (int)(mid << C), (int)(f << G), &bar[buf]);
    };
    const unsigned base8 = tid, base4 = (tid & 15u) | ((tid >> 4) << 8);
    const unsigned pbase8 = phys8(base8), pbase4 = phys8(base4);
    constexpr unsigned baseF_is8 = (G == 4 || !DIT) ? 1u : 0u;           // first round: bits 8..11 (DIF, G = 4) or 4..7 (DIT)
    const unsigned baseF = baseF_is8 ? base8 : base4, stepF = baseF_is8 ? 256u : 16u;
    const unsigned baseL = G == 4 ? base8 : (baseF_is8 ? base4 : base8), stepL = G == 4 ? 256u : (baseF_is8 ? 16u : 256u);

    // work items w = mid * batch + frame; CTA b owns the contiguous range [b T / G, (b + 1) T / G) (see intfft_fast16.cu)
    long long w = p.n_units * blockIdx.x / gridDim.x;
    const long long w_end = p.n_units * (blockIdx.x + 1) / gridDim.x;
    while (w < w_end) {
        const unsigned mid = (unsigned)(w / p.batch);
        const long long f0 = w - (long long)mid * p.batch;
        const long long f1 = (f0 + (w_end - w) < p.batch) ? f0 + (w_end - w) : p.batch;
        w += f1 - f0;
        auto kidx = [&](unsigned l) { return ((l >> C) << pb) | (mid << C) | (l & cmask); };
        if (tid == 0) {
            tma::store_wait_read<1>();               // the store two frames back read from the tile about to be refilled
            load(it & 1u, mid, f0);
        }

        int uwr[15], uwi[15];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int sgl = pb + (8 + q - C);
                const int2 tw = hoist_twiddle(p.tw, p.tay, sgl, kidx(tid | ((unsigned)j << 8)) & ((1u << sgl) - 1u));
                uwr[(1 << q) - 1 + j] = tw.x;
                uwi[(1 << q) - 1 + j] = tw.y;
            }
        if (G == 8) __syncthreads();
        if (G == 8 && tid < 240) {
            const int ww = tid >> 4, lo4 = tid & 15;
            const int q = ww >= 7 ? 3 : (ww >= 3 ? 2 : (ww >= 1 ? 1 : 0));
            const int j = ww - ((1 << q) - 1);
            const int sgl = pb + (4 + q - C);
            midtw[ww * 16 + lo4] = hoist_twiddle(p.tw, p.tay, sgl, kidx((unsigned)lo4 | ((unsigned)j << 4)) & ((1u << sgl) - 1u));
        }
        if (G == 8) __syncthreads();

        for (long long f = f0; f < f1; ++f, ++it) {
            const unsigned buf = it & 1u;
            if (tid == 0 && f + 1 < f1) {
                tma::store_wait_read<0>();           // the previous frame's store has read the other tile: refill it
                load(buf ^ 1u, mid, f + 1);
            }
            tma::mbar_wait(&bar[buf], (phase >> buf) & 1u);
            phase ^= 1u << buf;
            unsigned char *tile = land0 + buf * kLand32;
            V re[16], im[16];
            // ---- first round, from the dense landing tile ----
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                int a, b;
                const unsigned l = baseF + stepF * m;
                if (p.in_sb == 2) {
                    const unsigned x = reinterpret_cast<const unsigned *>(tile)[l];
                    a = (int)(short)(x & 0xffffu);
                    b = (int)x >> 16;
                } else {
                    const int2 v = reinterpret_cast<const int2 *>(tile)[l];
                    a = v.x;
                    b = v.y;
                }
                if (p.in_wrap) { a = sx(a, p.dw); b = sx(b, p.dw); }
                re[m] = mk(a);
                im[m] = mk(b);
            }
            if (G == 4) {                            // the pass's only round; results go straight back into the tile
                round32<4, DIT, MODE, KIND, TwRegs32, true>(re, im, p, pb + 8 - C, TwRegs32{uwr, uwi}, false, false);
                __syncthreads();                     // every thread has drained the landing tile
            } else {
                if (!DIT) round32<4, DIT, MODE, KIND, TwRegs32, true>(re, im, p, pb + 8 - C, TwRegs32{uwr, uwi}, false, false);
                else round32<4, DIT, MODE, KIND, TwSmem32, true>(re, im, p, pb + 4 - C, TwSmem32{midtw + (tid & 15u), 16}, false, false);
                {
                    const unsigned pbase = baseF_is8 ? pbase8 : pbase4;
#pragma unroll
                    for (int m = 0; m < 16; ++m) work[pbase + phys8(stepF * m)] = make_int2(re[m].f, im[m].f);
                }
                __syncthreads();                         // exchange tile complete; every thread has drained the landing tile
                {
                    const unsigned pbase = baseF_is8 ? pbase4 : pbase8;
#pragma unroll
                    for (int m = 0; m < 16; ++m) { const int2 v = work[pbase + phys8(stepL * m)]; re[m] = mk(v.x); im[m] = mk(v.y); }
                }
                if (!DIT) round32<4, DIT, MODE, KIND, TwSmem32, true>(re, im, p, pb + 4 - C, TwSmem32{midtw + (tid & 15u), 16}, false, false);
                else round32<4, DIT, MODE, KIND, TwRegs32, true>(re, im, p, pb + 8 - C, TwRegs32{uwr, uwi}, false, false);
            }
            // ---- results into the (drained) landing tile, dense, in the output container ----
            if (p.out_sb == 2) {
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    reinterpret_cast<unsigned *>(tile)[baseL + stepL * m] = __byte_perm((unsigned)re[m].f, (unsigned)im[m].f, 0x5410);
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) reinterpret_cast<int2 *>(tile)[baseL + stepL * m] = make_int2(re[m].f, im[m].f);
            }
            tma::fence_async();
            __syncthreads();                         // also: every thread has left the exchange tile
            if (tid == 0) tma::store_2d(&map_out, (int)(mid << C), (int)(f << G), tile);
        }
    }
    if (tid == 0) tma::store_wait_all();
}

template <int G, typename K> cudaError_t launch_any_tma(K k, const Fast32Params &p, int grid, cudaStream_t st)
{
    CUtensorMap mi, mo;
    const uint64_t rows = (uint64_t)p.batch << G;
    const uint32_t cols = 1u << (12 - G);
    if (!tma::make_map(&mi, p.in, p.in_sb >> 1, (uint64_t)1 << (p.n - G), rows, cols, 1u << G) ||
        !tma::make_map(&mo, p.out, p.out_sb >> 1, (uint64_t)1 << (p.n - G), rows, cols, 1u << G))
        return cudaErrorNotSupported;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStridedTmaSmem32);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, kStridedTmaSmem32, st>>>(p, mi, mo);
    return cudaGetLastError();
}
template <int G, bool DIT> cudaError_t launch_strided_tma(const Fast32Params &p, int mode, int kind, int grid, cudaStream_t st)
{
    if (kind == KIND_SINGLE_PRE) return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_TRUNC, KIND_SINGLE_PRE>, p, grid, st);
    switch (mode * 2 + kind) {
    case MODE_TRUNC * 2 + 0: return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_TRUNC, KIND_SINGLE>, p, grid, st);
    case MODE_TRUNC * 2 + 1: return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_TRUNC, KIND_MIXED>, p, grid, st);
    case MODE_ROUND * 2 + 0: return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_ROUND, KIND_SINGLE>, p, grid, st);
    case MODE_ROUND * 2 + 1: return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_ROUND, KIND_MIXED>, p, grid, st);
    case MODE_UNSCALED * 2 + 0: return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_UNSCALED, KIND_SINGLE>, p, grid, st);
    default: return launch_any_tma<G>(fast32_strided_tma_kernel<G, DIT, MODE_UNSCALED, KIND_MIXED>, p, grid, st);
    }
}

template <typename K> cudaError_t launch_any(K k, const Fast32Params &p, int grid, cudaStream_t st)
{
    const int smem = kHead32 + 2 * kTile8 * 8 + kStage32;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

// one translation unit per direction keeps the build parallel
template <int NLOG2, bool DIT> cudaError_t launch_contig(const Fast32Params &p, int mode, int kind, int grid, cudaStream_t st)
{
    // TRUNCATE, every stage single-DSP, pre-shifted twiddles in p (launch_fast32)
    if (kind == KIND_SINGLE_PRE) return launch_any(fast32_kernel<NLOG2, DIT, MODE_TRUNC, KIND_SINGLE_PRE>, p, grid, st);
    switch (mode * 2 + kind) {
    case MODE_TRUNC * 2 + 0: return launch_any(fast32_kernel<NLOG2, DIT, MODE_TRUNC, KIND_SINGLE>, p, grid, st);
    case MODE_TRUNC * 2 + 1: return launch_any(fast32_kernel<NLOG2, DIT, MODE_TRUNC, KIND_MIXED>, p, grid, st);
    case MODE_ROUND * 2 + 0: return launch_any(fast32_kernel<NLOG2, DIT, MODE_ROUND, KIND_SINGLE>, p, grid, st);
    case MODE_ROUND * 2 + 1: return launch_any(fast32_kernel<NLOG2, DIT, MODE_ROUND, KIND_MIXED>, p, grid, st);
    case MODE_UNSCALED * 2 + 0: return launch_any(fast32_kernel<NLOG2, DIT, MODE_UNSCALED, KIND_SINGLE>, p, grid, st);
    default: return launch_any(fast32_kernel<NLOG2, DIT, MODE_UNSCALED, KIND_MIXED>, p, grid, st);
    }
}
template <bool DIT> cudaError_t launch_contig_n(const Fast32Params &p, int bits, int mode, int kind, int grid, cudaStream_t st)
{
    switch (bits) {
    case 3: return launch_contig<3, DIT>(p, mode, kind, grid, st);
    case 4: return launch_contig<4, DIT>(p, mode, kind, grid, st);
    case 5: return launch_contig<5, DIT>(p, mode, kind, grid, st);
    case 6: return launch_contig<6, DIT>(p, mode, kind, grid, st);
    case 7: return launch_contig<7, DIT>(p, mode, kind, grid, st);
    case 8: return launch_contig<8, DIT>(p, mode, kind, grid, st);
    case 9: return launch_contig<9, DIT>(p, mode, kind, grid, st);
    case 10: return launch_contig<10, DIT>(p, mode, kind, grid, st);
    case 11: return launch_contig<11, DIT>(p, mode, kind, grid, st);
    case 12: return launch_contig<12, DIT>(p, mode, kind, grid, st);
    default: return cudaErrorInvalidValue;
    }
}
template <int G, bool DIT> cudaError_t launch_strided(const Fast32Params &p, int mode, int kind, int grid, cudaStream_t st)
{
    if (kind == KIND_SINGLE_PRE) return launch_any(fast32_strided_kernel<G, DIT, MODE_TRUNC, KIND_SINGLE_PRE>, p, grid, st);
    switch (mode * 2 + kind) {
    case MODE_TRUNC * 2 + 0: return launch_any(fast32_strided_kernel<G, DIT, MODE_TRUNC, KIND_SINGLE>, p, grid, st);
    case MODE_TRUNC * 2 + 1: return launch_any(fast32_strided_kernel<G, DIT, MODE_TRUNC, KIND_MIXED>, p, grid, st);
    case MODE_ROUND * 2 + 0: return launch_any(fast32_strided_kernel<G, DIT, MODE_ROUND, KIND_SINGLE>, p, grid, st);
    case MODE_ROUND * 2 + 1: return launch_any(fast32_strided_kernel<G, DIT, MODE_ROUND, KIND_MIXED>, p, grid, st);
    case MODE_UNSCALED * 2 + 0: return launch_any(fast32_strided_kernel<G, DIT, MODE_UNSCALED, KIND_SINGLE>, p, grid, st);
    default: return launch_any(fast32_strided_kernel<G, DIT, MODE_UNSCALED, KIND_MIXED>, p, grid, st);
    }
}

}  // namespace f32

// implemented in intfft_fast32_dif.cu / _dit.cu / _strided.cu
int f32_launch_contig_dif(const f32::Fast32Params &p, int bits, int mode, int kind, int grid, void *stream);
int f32_launch_contig_dit(const f32::Fast32Params &p, int bits, int mode, int kind, int grid, void *stream);
int f32_launch_strided(const f32::Fast32Params &p, int g, bool dit, int mode, int kind, int grid, void *stream);
int f32_launch_n13(const f32::Fast32Params &p, bool dit, int mode, int kind, int klo, int grid, void *stream);   // intfft_fast32_n13.cu

}  // namespace intfft
