// 64-bit-lane stage-chain kernel for the lowest eight stage bits (STAGE 7 .. 0) of plans whose values
// outgrow 32 bits: BASELINE c3 (65536-pt, 24-bit UNSCALED: widths 24 -> 40) runs its first eight
// stages on the 32-bit-lane strided kernel (intfft_fast32.cuh) and its last eight here.
//
// Reference rules implemented (paths relative to the reference root):
//   src/vhdl/fft/int_dif2_fly.vhd:142-373, src/vhdl/fft/int_dit2_fly.vhd:140-325      butterflies
//   src/vhdl/math/cmult/int_cmult_dsp48.vhd:182-190 / 307-317                          single DSP48 pair
//   src/vhdl/math/cmult/int_cmult_dbl18_dsp48.vhd:163-181, int_cmult_dbl35_dsp48.vhd:155-168   double
//   src/vhdl/math/cmult/int_cmult_trpl18_dsp48.vhd:151-155, int_cmult_trpl52_dsp48.vhd:150-170 triple
// The double arrangement's 48-bit wrap is not materialised: the kept slice ends at bit
// sh_post + dtwc - 1 <= 46 (dtwc < 45 / 43 with sh_post = 3 / 5; dtwc < 36 with sh_post = 12).
//
// Shape: 256 contiguous samples are an independent sub-transform of these eight stages, so the kernel
// is warp-centric and has NO CTA barrier: a warp owns 512 samples (two sub-blocks, one per half-warp),
// a thread keeps 16 samples in registers for four stages, and the single ownership change
// (stride-16 <-> 16 contiguous) goes through a warp-private, skewed shared-memory tile under
// __syncwarp.  Twiddles of STAGE 4..7 depend only on lane & 15 -> hoisted into registers for the whole
// persistent loop; those of STAGE 2, 3 are kernel parameters (constant bank).
// Products are exact 64-bit: D = hi' * 2^32 + (signed) lo, D * W = mul.wide.s32(lo, W) + ((hi' * W) << 32).
#include <cuda_runtime.h>

#include <cstdlib>

#include "intfft_arith.cuh"

namespace intfft {

namespace {

struct Fast64Params {
    const void *in;
    void *out;
    const int2 *tw;          // raw twiddles, entry (1 << s) + k
    int64_t total;         // frames * N samples (a multiple of 256)
    int64_t n_chunks;      // chunks of 512 samples (the last one may hold a single sub-block)
    int n;                   // NFFT of the whole transform
    int dw, format;
    int in_sb, out_sb;       // scalar bytes of the containers read / written
    int in_wrap;
    int prefetch;            // 1: input (32-bit containers) is staged through the per-warp landing area
    CmultConsts cm;
    int lw_r[16], lw_i[16];  // STAGE 2, 3 twiddles, index (1 << s) - 1 + k
};

constexpr int kWarpSlots = 544;      // 512 samples + one 16-byte slot of skew per 16 samples

// sample index inside a warp's 512-sample chunk -> 16-byte slot; additive for disjoint bit sets
__host__ __device__ constexpr unsigned phys64(unsigned i) { return i + (i >> 4); }

// Input prefetch (32-bit containers, 8 bytes per sample: c3's second pass): the NEXT chunk's 4 KB land in a
// warp-private area while the current chunk is computed — the warp fetches 512 contiguous bytes per cp.async
// instruction, and nobody waits on HBM at the top of a chunk (without it this kernel spent more warp-cycles on
// the long scoreboard than on anything else: two CTAs of 128-register threads cannot hide a DRAM round trip).
// 8-byte element i sits at land[i + 2 * (i >> 4)]: one 16-byte slot of skew per 16 samples, so that both the
// stride-16 (DIF, LDS.64) and the 16-contiguous (DIT, LDS.128) ownerships read it without bank conflicts.
__host__ __device__ constexpr unsigned physL(unsigned i) { return i + 2u * (i >> 4); }
constexpr int kLandElems = 576;      // physL(511) + 1 rounded up to whole 16-byte slots: 4608 bytes per warp
__device__ __forceinline__ void cp_async_16z(void *smem_dst, const void *gsrc, unsigned bytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}

struct Stg64 {
    int s, ow, dtwc;
    int k, sp;               // KIND 3: pre-shift of each product / post-shift of their sum for THIS stage
};
template <bool DIT, int MODE> __device__ __forceinline__ Stg64 stage64(const Fast64Params &p, int s)
{
    constexpr int FORMAT = MODE == MODE_UNSCALED ? 1 : 0;
    Stg64 st;
    st.s = s;
    const int ii = DIT ? s : p.n - 1 - s;
    const int dtw = p.dw + ii * FORMAT;
    st.ow = dtw + FORMAT;
    st.dtwc = DIT ? dtw : st.ow;
    // One form covers all three arrangements: wrap(((P2 >> k) -+ (P1 >> k)) >> sp).  single: k = 0, sp = sh_single;
    // double: k = k_pre, sp = sh_post; triple: k = sh_single, sp = 0 — its per-product wrap to dtwc bits
    // (int_cmult_trpl18_dsp48.vhd:151-155) commutes with the subtraction, wrapping being arithmetic mod 2^dtwc
    const bool dbl = st.dtwc >= p.cm.lim_single, trpl = st.dtwc >= p.cm.lim_dbl;
    st.k = trpl ? p.cm.sh_single : (dbl ? p.cm.k_pre : 0);
    st.sp = trpl ? 0 : (dbl ? p.cm.sh_post : p.cm.sh_single);
    return st;
}

// register pairs are split / joined with mov.b64: built from shifts and ORs, the front end no longer sees a
// plain pair and spends five or six instructions on every 64-bit add that follows
__device__ __forceinline__ unsigned lo32(int64_t v) { return (unsigned)v; }
__device__ __forceinline__ int hi32(int64_t v) { return (int)(v >> 32); }
__device__ __forceinline__ int64_t mk64(unsigned lo, int hi) { int64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }

// d = hi * 2^32 + (int)lo with hi corrected for the sign of the low word
struct Split {
    int lo, hi;
};
__device__ __forceinline__ Split split64(int64_t d)
{
    Split r;
    r.lo = (int)lo32(d);
    r.hi = hi32(d) + (int)((unsigned)r.lo >> 31);
    return r;
}
// exact signed 32 x 32 -> 64 (asm: the front end otherwise widens both operands and emits a 64 x 64 multiply)
__device__ __forceinline__ int64_t mulw(int a, int b)
{
    int64_t r;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ int64_t madw(int a, int b, int64_t c)
{
    int64_t r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
// acc + d * w (low 64 bits, exact when the true value fits): IMAD.WIDE + IMAD
__device__ __forceinline__ int64_t mad64x32(const Split &d, int w, int64_t acc)
{
    const int64_t t = madw(d.lo, w, acc);
    return mk64(lo32(t), hi32(t) + d.hi * w);
}
__device__ __forceinline__ int64_t mul64x32(const Split &d, int w)
{
    const int64_t t = mulw(d.lo, w);
    return mk64(lo32(t), hi32(t) + d.hi * w);
}

// t >> sh, 0 <= sh < 32 (every pre / post shift of the multiplier arrangements is below 32)
__device__ __forceinline__ int64_t sra64(int64_t t, int sh)
{
    return mk64(__funnelshift_r(lo32(t), (unsigned)hi32(t), sh), hi32(t) >> sh);
}
// keep the low w bits of v, sign-extended, 32 < w <= 64
__device__ __forceinline__ int64_t wrap_hi(int64_t v, int w)
{
    return mk64(lo32(v), sgxt32(hi32(v), w - 32));
}
// bits [sh + w - 1 : sh] of t, sign-extended; 0 <= sh < 32 < w, sh + w <= 64 (all grid-uniform)
__device__ __forceinline__ int64_t field64(int64_t t, int sh, int w)
{
    return mk64(__funnelshift_r(lo32(t), (unsigned)hi32(t), sh), (int)((unsigned)hi32(t) << (64 - sh - w)) >> (64 - w));
}

// keep the low w bits of v, sign-extended, for ANY 2 <= w <= 64 (grid-uniform w, branch-free): the low word is
// sign-extended from min(w, 32) bits, the high word from w - 32 bits or taken from the low word's sign
__device__ __forceinline__ int64_t wrap_any(int64_t v, int w)
{
    const int lo = sgxt32((int)lo32(v), w);                      // w >= 32: unchanged
    const int hi_w = sgxt32(hi32(v), w > 32 ? w - 32 : 32);
    return mk64((unsigned)lo, w > 32 ? hi_w : (lo >> 31));
}

// The double / per-stage arrangements work on explicit half-word pairs with carry chains: as int64_t the
// intermediate values had to be re-joined into aligned register pairs after every word-wise shift, which cost
// two to three moves per butterfly.
struct W2 {
    unsigned lo;
    int hi;                  // value = hi * 2^32 + lo
};
__device__ __forceinline__ W2 mulx(const Split &d, int w)
{
    const int64_t t = mulw(d.lo, w);
    W2 r;
    r.lo = lo32(t);
    r.hi = hi32(t) + d.hi * w;
    return r;
}
__device__ __forceinline__ W2 srax(W2 a, int k)        // 0 <= k < 32
{
    W2 r;
    r.lo = __funnelshift_r(a.lo, (unsigned)a.hi, k);
    r.hi = a.hi >> k;
    return r;
}
__device__ __forceinline__ W2 subx(W2 a, W2 b)
{
    W2 r;
    asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
    return r;
}
__device__ __forceinline__ W2 addx(W2 a, W2 b)
{
    W2 r;
    asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
    return r;
}
// bits [sh + w - 1 : sh] of t, sign-extended; 0 <= sh < 32 < w
__device__ __forceinline__ int64_t fieldx(W2 t, int sh, int w)
{
    return mk64(__funnelshift_r(t.lo, (unsigned)t.hi, sh), sgxt32(t.hi >> sh, w - 32));
}
// the same for any 2 <= w <= 64 (branch-free)
__device__ __forceinline__ int64_t fieldx_any(W2 t, int sh, int w)
{
    const W2 v = srax(t, sh);
    const int lo = sgxt32((int)v.lo, w);                          // w >= 32: unchanged
    const int hi_w = sgxt32(v.hi, w > 32 ? w - 32 : 32);
    return mk64((unsigned)lo, w > 32 ? hi_w : (lo >> 31));
}

// KIND: 0 single, 1 double, 2 triple (the same for every multiplying stage of the pass, every width beyond 32
// bits); 3 = arrangement chosen per stage, any width (plans whose STAGE 7..0 cross the 32-bit line or a limit)
template <int KIND>
__device__ __forceinline__ void cmul64(int64_t dr, int64_t di, int wr, int wi, const CmultConsts &cm, const Stg64 &st,
                                       int64_t &o_re, int64_t &o_im)
{
    const int dtwc = st.dtwc;
    const Split r = split64(dr), i = split64(di);
    if (KIND == 3) {
        const W2 tr = subx(srax(mulx(r, wr), st.k), srax(mulx(i, wi), st.k));
        const W2 ti = addx(srax(mulx(r, wi), st.k), srax(mulx(i, wr), st.k));
        o_re = fieldx_any(tr, st.sp, dtwc);
        o_im = fieldx_any(ti, st.sp, dtwc);
    } else if (KIND == 0) {
        const int64_t tr = mad64x32(i, -wi, mul64x32(r, wr));
        const int64_t ti = mad64x32(i, wr, mul64x32(r, wi));
        o_re = field64(tr, cm.sh_single, dtwc);
        o_im = field64(ti, cm.sh_single, dtwc);
    } else if (KIND == 1) {
        const W2 tr = subx(srax(mulx(r, wr), cm.k_pre), srax(mulx(i, wi), cm.k_pre));
        const W2 ti = addx(srax(mulx(r, wi), cm.k_pre), srax(mulx(i, wr), cm.k_pre));
        o_re = fieldx(tr, cm.sh_post, dtwc);
        o_im = fieldx(ti, cm.sh_post, dtwc);
    } else {
        const int64_t a = field64(mul64x32(r, wr), cm.sh_single, dtwc), b = field64(mul64x32(i, wi), cm.sh_single, dtwc);
        const int64_t c = field64(mul64x32(r, wi), cm.sh_single, dtwc), d = field64(mul64x32(i, wr), cm.sh_single, dtwc);
        o_re = wrap_hi((int64_t)((uint64_t)a - (uint64_t)b), dtwc);
        o_im = wrap_hi((int64_t)((uint64_t)c + (uint64_t)d), dtwc);
    }
}

// -v for v >= 0, ~v for v < 0; the result always fits the operand's width
__device__ __forceinline__ int64_t negq64(int64_t v) { return (v >> 63) - v; }

// `ow` > 32: only the ROUNDING difference can leave it (see addsub<> in intfft_arith.cuh)
template <int MODE, bool ANYW = false> __device__ __forceinline__ void addsub64(int64_t a, int64_t b, int ow, int64_t &ad, int64_t &su)
{
    if (MODE == MODE_TRUNC) {
        const int64_t ha = sra64(a, 1), hb = sra64(b, 1);
        ad = ha + hb;
        su = ha - hb;
    } else if (MODE == MODE_ROUND) {
        const int64_t s = (int64_t)((uint64_t)a + (uint64_t)b + 1u), d = (int64_t)((uint64_t)a - (uint64_t)b + 1u);
        ad = sra64(s, 1);
        su = ANYW ? wrap_any(sra64(d, 1), ow) : wrap_hi(sra64(d, 1), ow);
    } else {
        ad = (int64_t)((uint64_t)a + (uint64_t)b);
        su = (int64_t)((uint64_t)a - (uint64_t)b);
    }
}

// MUL: the caller knows st.s >= 2 (every stage of a strided top pass): no per-butterfly stage branches
template <bool DIT, int MODE, int KIND, bool MUL = false>
__device__ __forceinline__ void fly64(const Stg64 &st, bool odd, const CmultConsts &cm, int64_t &ar, int64_t &ai,
                                      int64_t &br, int64_t &bi, int wr, int wi)
{
    using T = int64_t;
    if (!DIT) {
        T xr, xi, sr, si;
        addsub64<MODE, KIND == 3>(ar, br, st.ow, xr, sr);
        addsub64<MODE, KIND == 3>(ai, bi, st.ow, xi, si);
        ar = xr;
        ai = xi;
        if (!MUL && st.s == 0) {
            br = sr;
            bi = si;
        } else if (!MUL && st.s == 1) {
            br = odd ? si : sr;
            bi = odd ? negq64(sr) : si;
        } else {
            cmul64<KIND>(sr, si, wr, wi, cm, st, br, bi);
        }
    } else {
        T wr_, wi_;
        if (!MUL && st.s == 0) {
            wr_ = br;
            wi_ = bi;
        } else if (!MUL && st.s == 1) {
            wr_ = odd ? negq64(bi) : br;
            wi_ = odd ? br : bi;
        } else {                                      // multiplier fed with swapped re / im, outputs swapped back
            T o_re, o_im;
            cmul64<KIND>(bi, br, wr, wi, cm, st, o_re, o_im);
            wi_ = o_re;
            wr_ = o_im;
        }
        T xr, xi, yr, yi;
        addsub64<MODE, KIND == 3>(ar, wr_, st.ow, xr, yr);
        addsub64<MODE, KIND == 3>(ai, wi_, st.ow, xi, yi);
        ar = xr; ai = xi;
        br = yr; bi = yi;
    }
}

// four stages (global STAGE S0 .. S0+3) on the register index bits 0..3
template <int S0, bool DIT, int MODE, int KIND>
__device__ __forceinline__ void round64(int64_t (&re)[16], int64_t (&im)[16], const Fast64Params &p,
                                        const int (&twr)[15], const int (&twi)[15])
{
#pragma unroll
    for (int step = 0; step < 4; ++step) {
        const int q = DIT ? step : 3 - step;
        const Stg64 st = stage64<DIT, MODE>(p, S0 + q);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (m & (1 << q)) continue;
            const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
            int wr = 0, wi = 0;
            if (S0 + q >= 2) { wr = twr[w]; wi = twi[w]; }
            fly64<DIT, MODE, KIND>(st, (m & 1) != 0, p.cm, re[m], im[m], re[m | (1 << q)], im[m | (1 << q)], wr, wi);
        }
    }
}

template <bool DIT, int MODE, int KIND, bool PF>
__global__ void __launch_bounds__(256, 2) fast64_kernel(const __grid_constant__ Fast64Params p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    longlong2 *sm = reinterpret_cast<longlong2 *>(smem_raw) + warp * kWarpSlots;
    const unsigned sub = lane >> 4, l4 = lane & 15u;

    // STAGE 4..7 twiddles: index = (position mod 2^s) = l4 + 16 * j
    int uwr[15], uwi[15];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < (1 << q); ++j) {
            const int2 w = __ldg(p.tw + (1u << (4 + q)) + l4 + ((unsigned)j << 4));
            uwr[(1 << q) - 1 + j] = w.x;
            uwi[(1 << q) - 1 + j] = w.y;
        }
    int lwr[15], lwi[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) { lwr[i] = p.lw_r[i]; lwi[i] = p.lw_i[i]; }

    // layouts inside the chunk: A = stride-16 ownership (register bits = sample bits 4..7),
    //                           B = 16 contiguous samples  (register bits = sample bits 0..3)
    const unsigned baseA = sub * 256u + l4, baseB = sub * 256u + 16u * l4;
    const unsigned pA = phys64(baseA), pB = phys64(baseB);

    constexpr bool pf = PF;                            // in_sb == 4 (the launcher sized the landing area)
    int2 *land = reinterpret_cast<int2 *>(smem_raw + 8 * kWarpSlots * 16) + warp * kLandElems;
    // piece q = lane + 32 j holds samples 2q, 2q + 1: physL(2 lane + 64 j) = physL(2 lane) + 72 j
    auto prefetch = [&](unsigned chunk) {
        const int64_t g = (int64_t)chunk << 9;
        const char *src = reinterpret_cast<const char *>(p.in) + g * 8 + 16u * lane;
        int2 *dst = land + physL(2u * lane);
        if (g + 512 <= p.total) {                      // whole chunk (all but possibly the last one): no predicates
#pragma unroll
            for (int j = 0; j < 8; ++j) cp_async_16z(dst + 72 * j, src + 512 * j, 16u);
        } else {                                       // total is a multiple of 256 samples: only the lower sub-block exists
#pragma unroll
            for (int j = 0; j < 4; ++j) cp_async_16z(dst + 72 * j, src + 512 * j, 16u);
#pragma unroll
            for (int j = 4; j < 8; ++j) cp_async_16z(dst + 72 * j, p.in, 0u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // chunk counters are 32-bit (total <= 2^40 samples = 2^31 chunks): the 64-bit ones were spilled to local memory and
    // the loop test waited on their reload every chunk
    const unsigned n_chunks = (unsigned)p.n_chunks, chunk_step = gridDim.x * 8u;
    if (pf && blockIdx.x * 8u + warp < n_chunks) prefetch(blockIdx.x * 8u + warp);

    for (unsigned chunk = blockIdx.x * 8u + warp; chunk < n_chunks; chunk += chunk_step) {
        const int64_t g0 = (int64_t)chunk << 9;
        const bool active = g0 + sub * 256 < p.total;
        int64_t re[16], im[16];
        // ---- first round: from the landing area (prefetched one chunk ago) or straight from HBM ----
        if (pf) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();                              // the other lanes' pieces have landed too
            if (DIT) {
                const int4 *own = reinterpret_cast<const int4 *>(land + physL(baseB));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int4 v = own[j];
                    re[2 * j] = v.x; im[2 * j] = v.y;
                    re[2 * j + 1] = v.z; im[2 * j + 1] = v.w;
                }
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int2 v = land[physL(baseA) + 18u * m];          // physL(baseA + 16 m)
                    re[m] = v.x;
                    im[m] = v.y;
                }
            }
            __syncwarp();                              // every lane has drained the area
            const unsigned next = chunk + chunk_step;
            if (next < n_chunks && next > chunk) prefetch(next);
        } else if (p.in_sb == 4) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                int2 v = make_int2(0, 0);
                if (active) v = __ldg(reinterpret_cast<const int2 *>(p.in) + g0 + (DIT ? baseB + m : baseA + 16u * m));
                re[m] = v.x;
                im[m] = v.y;
            }
        } else if (p.in_sb == 8) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                longlong2 v = make_longlong2(0, 0);
                if (active) v = __ldg(reinterpret_cast<const longlong2 *>(p.in) + g0 + (DIT ? baseB + m : baseA + 16u * m));
                re[m] = v.x;
                im[m] = v.y;
            }
        } else {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                short2 v = make_short2(0, 0);
                if (active) v = __ldg(reinterpret_cast<const short2 *>(p.in) + g0 + (DIT ? baseB + m : baseA + 16u * m));
                re[m] = v.x;
                im[m] = v.y;
            }
        }
        if (p.in_wrap) {
#pragma unroll
            for (int m = 0; m < 16; ++m) { re[m] = wrapw<int64_t>(re[m], p.dw); im[m] = wrapw<int64_t>(im[m], p.dw); }
        }
        if (DIT) round64<0, DIT, MODE, KIND>(re, im, p, lwr, lwi);
        else round64<4, DIT, MODE, KIND>(re, im, p, uwr, uwi);
        // ---- ownership change inside the warp ----
        __syncwarp();                                  // previous chunk's reads of the tile are complete
#pragma unroll
        for (int m = 0; m < 16; ++m) sm[DIT ? pB + m : pA + phys64(16u * m)] = make_longlong2(re[m], im[m]);
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const longlong2 v = sm[DIT ? pA + phys64(16u * m) : pB + m];
            re[m] = v.x;
            im[m] = v.y;
        }
        if (DIT) round64<4, DIT, MODE, KIND>(re, im, p, uwr, uwi);
        else round64<0, DIT, MODE, KIND>(re, im, p, lwr, lwi);
        if (!DIT && p.out_sb == 8) {
            // DIF leaves 16 contiguous samples (256 bytes) in each lane: stored directly, one instruction
            // would touch 32 different lines.  The lane's own tile slots (the ones it read for this
            // round) take the results in place, and the warp then writes 512 contiguous bytes per
            // instruction.
#pragma unroll
            for (int m = 0; m < 16; ++m) sm[pB + m] = make_longlong2(re[m], im[m]);
            __syncwarp();
            // total is a multiple of 256 samples: pieces 0..7 (the lower sub-block) always exist, 8..15 iff the chunk is whole
            longlong2 *dst = reinterpret_cast<longlong2 *>(p.out) + g0 + lane;
            const longlong2 *src = sm + phys64(lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[32 * j] = src[34 * j];       // phys64(lane + 32 j)
            if (g0 + 512 <= p.total) {
#pragma unroll
                for (int j = 8; j < 16; ++j) dst[32 * j] = src[34 * j];
            }
        } else if (active) {
            if (p.out_sb == 8) {
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    reinterpret_cast<longlong2 *>(p.out)[g0 + (DIT ? baseA + 16u * m : baseB + m)] = make_longlong2(re[m], im[m]);
            } else {
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    reinterpret_cast<int2 *>(p.out)[g0 + (DIT ? baseA + 16u * m : baseB + m)] = make_int2((int)re[m], (int)im[m]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Strided pass on 64-bit lanes: the top G = 4 or 8 stage bits of a 2^12 / 2^16-point plan whose values have
// outgrown 32 bits there (wide DIT plans end with it, DIF plans wider than 24 bits start with it).  Same
// geometry as the packed-16 / 32-bit strided kernels: a tile is 2^G rows (pitch 2^(NFFT-G) samples) by
// 2^(12-G) contiguous columns, a CTA keeps one column block and walks over frames so that the twiddles of
// the block are fetched once.  16 samples per thread, arrangement and wrap width per stage (instance 3).
// The single 64 KB exchange tile needs no skew: in both ownerships the lanes of a quarter-warp hold eight
// consecutive 16-byte elements.
struct Strided64Params {
    const void *in;
    void *out;
    const int2 *tw;
    int64_t batch;
    int n, dw, format, in_sb, out_sb, in_wrap;
    int64_t n_units;         // work items = column blocks per frame * batch
    CmultConsts cm;
};

template <int G, bool DIT, int MODE>
__global__ void __launch_bounds__(256, 2) fast64_strided_kernel(const __grid_constant__ Strided64Params p)
{
    constexpr int C = 12 - G, NR = G / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int2 *midtw = reinterpret_cast<int2 *>(smem_raw);                          // [15][16]
    longlong2 *tile = reinterpret_cast<longlong2 *>(smem_raw + 2048);

    const unsigned tid = threadIdx.x;
    const int pb = p.n - G;
    const unsigned cmask = (1u << C) - 1u;
    Fast64Params sp{};                              // what stage64<> reads
    sp.n = p.n; sp.dw = p.dw; sp.format = p.format; sp.cm = p.cm;

    // work items w = mid * batch + frame; CTA b owns the contiguous range [b T / G, (b + 1) T / G) (see intfft_fast16.cu)
    int64_t w = p.n_units * blockIdx.x / gridDim.x;
    const int64_t w_end = p.n_units * (blockIdx.x + 1) / gridDim.x;
    while (w < w_end) {
        const unsigned mid = (unsigned)(w / p.batch);
        const int64_t f0 = w - (int64_t)mid * p.batch;
        const int64_t f1 = (f0 + (w_end - w) < p.batch) ? f0 + (w_end - w) : p.batch;
        w += f1 - f0;
        auto kidx = [&](unsigned l) { return ((l >> C) << pb) | (mid << C) | (l & cmask); };

        int uwr[15], uwi[15];                       // round on local bits 8..11: STAGE pb + (8 + q - C)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < (1 << q); ++j) {
                const int sgl = pb + (8 + q - C);
                const int2 w = __ldg(p.tw + (1u << sgl) + (kidx(tid | ((unsigned)j << 8)) & ((1u << sgl) - 1u)));
                uwr[(1 << q) - 1 + j] = w.x;
                uwi[(1 << q) - 1 + j] = w.y;
            }
        if (NR == 2) {                              // round on local bits 4..7: table[w][tid & 15]
            __syncthreads();
            if (tid < 240) {
                const int w = tid >> 4, lo4 = tid & 15;
                const int q = w >= 7 ? 3 : (w >= 3 ? 2 : (w >= 1 ? 1 : 0));
                const int j = w - ((1 << q) - 1);
                const int sgl = pb + (4 + q - C);
                midtw[w * 16 + lo4] = __ldg(p.tw + (1u << sgl) + (kidx((unsigned)lo4 | ((unsigned)j << 4)) & ((1u << sgl) - 1u)));
            }
            __syncthreads();
        }

        for (int64_t f = f0; f < f1; ++f) {
            const int64_t gbase = (f << p.n) + ((int64_t)mid << C);
            int64_t re[16], im[16];
#pragma unroll
            for (int rr = 0; rr < NR; ++rr) {
                const int r = DIT ? rr : NR - 1 - rr;
                const int lo = 12 - 4 * (NR - r);
                const bool first = rr == 0, last = rr == NR - 1;
                const unsigned base = (tid & ((1u << lo) - 1u)) | ((tid >> lo) << (lo + 4));
                // register m sits 2^(lo-C) m rows below register 0 (lo >= C in every geometry)
                const int64_t g0 = gbase + ((int64_t)(base >> C) << pb) + (base & cmask);
                const int64_t gstep = (int64_t)1 << (pb + lo - C);
                if (first) {
                    if (p.in_sb == 4) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) { const int2 v = __ldg(reinterpret_cast<const int2 *>(p.in) + g0 + m * gstep); re[m] = v.x; im[m] = v.y; }
                    } else if (p.in_sb == 8) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) { const longlong2 v = __ldg(reinterpret_cast<const longlong2 *>(p.in) + g0 + m * gstep); re[m] = v.x; im[m] = v.y; }
                    } else {
#pragma unroll
                        for (int m = 0; m < 16; ++m) { const short2 v = __ldg(reinterpret_cast<const short2 *>(p.in) + g0 + m * gstep); re[m] = v.x; im[m] = v.y; }
                    }
                    if (p.in_wrap) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) { re[m] = wrapw<int64_t>(re[m], p.dw); im[m] = wrapw<int64_t>(im[m], p.dw); }
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 16; ++m) { const longlong2 v = tile[base + ((unsigned)m << lo)]; re[m] = v.x; im[m] = v.y; }
                }
                // ---- four multiplying stages on the register bits ----
                const int s0 = pb + (lo - C);
#pragma unroll
                for (int step = 0; step < 4; ++step) {
                    const int q = DIT ? step : 3 - step;
                    const Stg64 st = stage64<DIT, MODE>(sp, s0 + q);
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        if (m & (1 << q)) continue;
                        const int w = (1 << q) - 1 + (m & ((1 << q) - 1));
                        int wr, wi;
                        if (lo == 8) { wr = uwr[w]; wi = uwi[w]; }
                        else { const int2 t = midtw[w * 16 + (tid & 15u)]; wr = t.x; wi = t.y; }
                        fly64<DIT, MODE, 3, true>(st, false, p.cm, re[m], im[m], re[m | (1 << q)], im[m | (1 << q)], wr, wi);
                    }
                }
                if (last) {
                    if (p.out_sb == 8) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) reinterpret_cast<longlong2 *>(p.out)[g0 + m * gstep] = make_longlong2(re[m], im[m]);
                    } else {
#pragma unroll
                        for (int m = 0; m < 16; ++m) reinterpret_cast<int2 *>(p.out)[g0 + m * gstep] = make_int2((int)re[m], (int)im[m]);
                    }
                } else {
                    __syncthreads();                                   // the previous frame's readers have left the tile
#pragma unroll
                    for (int m = 0; m < 16; ++m) tile[base + ((unsigned)m << lo)] = make_longlong2(re[m], im[m]);
                    __syncthreads();
                }
            }
        }
    }
}

template <int G, bool DIT> cudaError_t launch_strided64_k(const Strided64Params &p, int mode, int grid, cudaStream_t st)
{
    using K = void (*)(const Strided64Params);
    K k = mode == MODE_TRUNC ? (K)fast64_strided_kernel<G, DIT, MODE_TRUNC>
        : (mode == MODE_ROUND ? (K)fast64_strided_kernel<G, DIT, MODE_ROUND> : (K)fast64_strided_kernel<G, DIT, MODE_UNSCALED>);
    const int smem = 2048 + 4096 * 16;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

template <bool DIT, int MODE> cudaError_t launch_k(const Fast64Params &p, int kind, int grid, cudaStream_t st)
{
    using K = void (*)(const Fast64Params);
    // KIND 0 (single DSP48 pair) needs dtwc < 28 and so never meets the "every width beyond 32 bits" rule of the
    // specialised instances; plans with such stages run the per-stage instance (3)
    if (kind < 1 || kind > 3) return cudaErrorInvalidValue;
    K k;
    if (p.prefetch) k = kind == 1 ? (K)fast64_kernel<DIT, MODE, 1, true> : (kind == 2 ? (K)fast64_kernel<DIT, MODE, 2, true> : (K)fast64_kernel<DIT, MODE, 3, true>);
    else k = kind == 1 ? (K)fast64_kernel<DIT, MODE, 1, false> : (kind == 2 ? (K)fast64_kernel<DIT, MODE, 2, false> : (K)fast64_kernel<DIT, MODE, 3, false>);
    const int smem = 8 * kWarpSlots * 16 + (p.prefetch ? 8 * kLandElems * 8 : 0);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace

// Which instance runs STAGE 7..0 of this plan: 1 / 2 when every stage of the pass is wider than 32 bits and all
// multiplying stages (STAGE 2..7) share the double / triple arrangement; 3 (arrangement per stage, any width)
// when some stage still wraps at <= 32 bits or the arrangements differ.
int fast64_uniform_kind(const PassParams &kp, bool dit)
{
    int kind = -1;
    bool all_wide = true, uniform = true;
    for (int s = 0; s < 8; ++s) {
        const int ii = dit ? s : kp.n - 1 - s;
        const int dtw = kp.dw + ii * kp.format;
        const int dtwc = dit ? dtw : dtw + kp.format;
        if (dtwc <= 32 || dtw + kp.format <= 32) all_wide = false;
        if (s < 2) continue;
        const int k = dtwc < kp.cm.lim_single ? 0 : (dtwc < kp.cm.lim_dbl ? 1 : 2);
        if (kind >= 0 && k != kind) uniform = false;
        kind = k;
    }
    if (all_wide && uniform && kind >= 1) return kind;
    return 3;
}

int launch_fast64(const PassDesc &pd, int mode, bool dit, const int2 *tw, const int *lw_r, const int *lw_i,
                  int num_sms, void *stream)
{
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Fast64Params p{};
    p.in = pd.kp.in;
    p.out = pd.kp.out;
    p.tw = tw;
    p.total = pd.kp.total;
    p.n_chunks = (pd.kp.total + 511) >> 9;
    p.n = pd.kp.n;
    p.dw = pd.kp.dw;
    p.format = pd.kp.format;
    p.in_sb = pd.kp.in_sb;
    p.out_sb = pd.kp.out_sb;
    p.in_wrap = pd.kp.in_wrap;
    p.prefetch = (p.in_sb == 4 && getenv("INTFFT_F64_NO_PREFETCH") == nullptr) ? 1 : 0;
    p.cm = pd.kp.cm;
    for (int i = 0; i < 16; ++i) { p.lw_r[i] = lw_r[i]; p.lw_i[i] = lw_i[i]; }
    const int kind = fast64_uniform_kind(pd.kp, dit);
    if (kind < 0) return (int)cudaErrorInvalidValue;
    int64_t grid = 2ll * num_sms;
    const int64_t need = (p.n_chunks + 7) / 8;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaError_t e;
    switch (mode * 2 + (dit ? 1 : 0)) {
    case MODE_TRUNC * 2 + 0: e = launch_k<false, MODE_TRUNC>(p, kind, (int)grid, st); break;
    case MODE_TRUNC * 2 + 1: e = launch_k<true, MODE_TRUNC>(p, kind, (int)grid, st); break;
    case MODE_ROUND * 2 + 0: e = launch_k<false, MODE_ROUND>(p, kind, (int)grid, st); break;
    case MODE_ROUND * 2 + 1: e = launch_k<true, MODE_ROUND>(p, kind, (int)grid, st); break;
    case MODE_UNSCALED * 2 + 0: e = launch_k<false, MODE_UNSCALED>(p, kind, (int)grid, st); break;
    default: e = launch_k<true, MODE_UNSCALED>(p, kind, (int)grid, st); break;
    }
    count_launch();
    return (int)e;
}

// top-bits pass of a wide plan on 64-bit lanes: kp.g in {4, 8}, kp.pb = NFFT - kp.g, every stage multiplies
int launch_fast64_strided(const PassDesc &pd, int mode, bool dit, const int2 *tw, int num_sms, void *stream)
{
    Strided64Params p{};
    p.in = pd.kp.in;
    p.out = pd.kp.out;
    p.tw = tw;
    p.n = pd.kp.n;
    p.batch = pd.kp.total >> pd.kp.n;
    p.dw = pd.kp.dw;
    p.format = pd.kp.format;
    p.in_sb = pd.kp.in_sb;
    p.out_sb = pd.kp.out_sb;
    p.in_wrap = pd.kp.in_wrap;
    p.cm = pd.kp.cm;
    const int G = pd.kp.g, C = 12 - G, mid_bits = p.n - G - C;
    const int64_t mids = (int64_t)1 << mid_bits;
    int64_t grid = 2ll * num_sms;
    p.n_units = mids * p.batch;
    if (grid > p.n_units) grid = p.n_units;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (G == 4) e = dit ? launch_strided64_k<4, true>(p, mode, (int)grid, st) : launch_strided64_k<4, false>(p, mode, (int)grid, st);
    else if (G == 8) e = dit ? launch_strided64_k<8, true>(p, mode, (int)grid, st) : launch_strided64_k<8, false>(p, mode, (int)grid, st);
    else e = cudaErrorInvalidValue;
    count_launch();
    return (int)e;
}

}  // namespace intfft
