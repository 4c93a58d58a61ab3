// Device arithmetic of one radix-2 butterfly, generic over lane / product width.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   int_dif2_fly.vhd:142-373   DIF: X = A+B, Y = (A-B)*W with per-mode pre-shift / round / growth
//   int_dit2_fly.vhd:140-325   DIT: BW = B*conj(W) (re/im swapped in and out of the multiplier),
//                              X = A+BW, Y = A-BW
//   int_cmult_dsp48.vhd:182-434 and the five int_cmult_*_dsp48 variants: where the floor is taken
//   int_dif2_fly.vhd:281-304 / int_dit2_fly.vhd:252-281: the multiplier-free -j / +j with the
//                              "not(x)" (off by one) negation of negative operands
#pragma once
#include <cstdint>

#include "intfft_internal.h"

namespace intfft {

template <typename T> struct LaneBits;
template <> struct LaneBits<int32_t> { static constexpr int bits = 32; using U = uint32_t; };
template <> struct LaneBits<int64_t> { static constexpr int bits = 64; using U = uint64_t; };

// keep the low w bits, sign-extended (a VHDL slice): szext is one SGXT (bfe.s32 with a register length is
// PRMT + SHF + SGXT).
__device__ __forceinline__ int32_t sgxt32(int32_t v, int w)
{
    int32_t r;
    asm("szext.clamp.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(w));      // one SGXT; w >= 32 copies v
    return r;
}
template <typename T> __device__ __forceinline__ T wrapw(T v, int w);
template <> __device__ __forceinline__ int32_t wrapw<int32_t>(int32_t v, int w) { return sgxt32(v, w); }
template <> __device__ __forceinline__ int64_t wrapw<int64_t>(int64_t v, int w)
{
    if (w <= 32) return (int64_t)sgxt32((int32_t)v, w);          // w is uniform: no divergence
    const int32_t hi = sgxt32((int32_t)(v >> 32), w - 32);
    return (int64_t)(((uint64_t)(uint32_t)hi << 32) | (uint32_t)v);
}

__device__ __forceinline__ int64_t wrap48_64(int64_t v) { return (int64_t)((uint64_t)v << 16) >> 16; }

// negation used by the STAGE = 1 butterflies: -v for v >= 0, ~v for v < 0, kept to w bits
template <typename T> __device__ __forceinline__ T negq(T v, int w)
{
    return wrapw<T>(v >= 0 ? (T)(0 - v) : (T)~v, w);
}

// `ow` = output width.  Only the ROUNDING difference can leave it: A - B = 2^DTW - 1 rounds up to
// 2^(DTW-1), which the reference keeps in DTW bits (rnd(DTW downto 1) + '1', int_dif2_fly.vhd:201-216),
// i.e. it wraps to -2^(DTW-1).
template <int MODE, typename T>
__device__ __forceinline__ void addsub(T a, T b, int ow, T &ad, T &su)
{
    using U = typename LaneBits<T>::U;
    if (MODE == MODE_TRUNC) {          // inputs sliced (DTW-1 downto 1)
        const T ha = a >> 1, hb = b >> 1;
        ad = ha + hb;
        su = ha - hb;
    } else if (MODE == MODE_ROUND) {   // (v >> 1) + v(0) on the exact sum / difference
        const T s = (T)((U)a + (U)b), d = (T)((U)a - (U)b);
        ad = (s >> 1) + (s & 1);
        su = wrapw<T>((d >> 1) + (d & 1), ow);
    } else {                           // exact, one bit of growth
        ad = (T)((U)a + (U)b);
        su = (T)((U)a - (U)b);
    }
}

// DO = DI * W, truncated the way the selected DSP48 arrangement does it, kept to dtwc bits.
// `kind` is uniform across the grid for a given stage: 0 single, 1 double, 2 triple.
template <typename T, typename P>
__device__ __forceinline__ void cmult(T d_re, T d_im, int w_re, int w_im, const CmultConsts &cm,
                                      int kind, int dtwc, T &o_re, T &o_im)
{
    if (kind == 2 && dtwc > cm.trpl_awd) {      // trpl18 above 61 / 59 bits: the multiplier's data port cuts the operand
        d_re = wrapw<T>(d_re, cm.trpl_awd);
        d_im = wrapw<T>(d_im, cm.trpl_awd);
    }
    const P p_rr = (P)d_re * (P)w_re, p_ii = (P)d_im * (P)w_im;
    const P p_ri = (P)d_re * (P)w_im, p_ir = (P)d_im * (P)w_re;
    if (kind == 0) {
        o_re = wrapw<T>((T)((p_rr - p_ii) >> cm.sh_single), dtwc);
        o_im = wrapw<T>((T)((p_ri + p_ir) >> cm.sh_single), dtwc);
    } else if (kind == 1) {
        const int64_t a = wrap48_64((int64_t)(p_rr >> cm.k_pre)), b = wrap48_64((int64_t)(p_ii >> cm.k_pre));
        const int64_t c = wrap48_64((int64_t)(p_ri >> cm.k_pre)), d = wrap48_64((int64_t)(p_ir >> cm.k_pre));
        o_re = wrapw<T>((T)(wrap48_64(a - b) >> cm.sh_post), dtwc);
        o_im = wrapw<T>((T)(wrap48_64(c + d) >> cm.sh_post), dtwc);
    } else {
        using U = typename LaneBits<T>::U;
        const T a = wrapw<T>((T)(p_rr >> cm.sh_single), dtwc), b = wrapw<T>((T)(p_ii >> cm.sh_single), dtwc);
        const T c = wrapw<T>((T)(p_ri >> cm.sh_single), dtwc), d = wrapw<T>((T)(p_ir >> cm.sh_single), dtwc);
        o_re = wrapw<T>((T)((U)a - (U)b), dtwc);
        o_im = wrapw<T>((T)((U)c + (U)d), dtwc);
    }
}

struct StageInfo {
    int s;       // butterfly STAGE number == global bit being paired
    int dtw;     // input width of this stage
    int ow;      // output width
    int kind;    // multiplier arrangement for this stage
};

template <bool DIT>
__device__ __forceinline__ StageInfo stage_info(const PassParams &p, int s)
{
    StageInfo si;
    si.s = s;
    const int ii = DIT ? s : p.n - 1 - s;
    si.dtw = p.dw + ii * p.format;
    si.ow = si.dtw + p.format;
    const int dtwc = DIT ? si.dtw : si.ow;
    si.kind = dtwc < p.cm.lim_single ? 0 : (dtwc < p.cm.lim_dbl ? 1 : 2);
    return si;
}

// One butterfly in place on (a, b).  k = twiddle index (position mod 2^s), w = W_s[k] (s >= 2).
template <typename T, typename P, int MODE, bool DIT>
__device__ __forceinline__ void butterfly(T &a_re, T &a_im, T &b_re, T &b_im, const StageInfo &si,
                                          const CmultConsts &cm, unsigned k, int2 w)
{
    if (!DIT) {
        T ad_re, ad_im, su_re, su_im;
        addsub<MODE, T>(a_re, b_re, si.ow, ad_re, su_re);
        addsub<MODE, T>(a_im, b_im, si.ow, ad_im, su_im);
        a_re = ad_re;
        a_im = ad_im;
        if (si.s == 0) {
            b_re = su_re;
            b_im = su_im;
        } else if (si.s == 1) {
            const bool odd = k & 1;
            b_re = odd ? su_im : su_re;
            b_im = odd ? negq<T>(su_re, si.ow) : su_im;
        } else {
            cmult<T, P>(su_re, su_im, w.x, w.y, cm, si.kind, si.ow, b_re, b_im);
        }
    } else {
        T bw_re, bw_im;
        if (si.s == 0) {
            bw_re = b_re;
            bw_im = b_im;
        } else if (si.s == 1) {
            const bool odd = k & 1;
            bw_re = odd ? negq<T>(b_im, si.dtw) : b_re;
            bw_im = odd ? b_re : b_im;
        } else {
            T o_re, o_im;
            cmult<T, P>(b_im, b_re, w.x, w.y, cm, si.kind, si.dtw, o_re, o_im);
            bw_im = o_re;
            bw_re = o_im;
        }
        T x_re, x_im, y_re, y_im;
        addsub<MODE, T>(a_re, bw_re, si.ow, x_re, y_re);
        addsub<MODE, T>(a_im, bw_im, si.ow, x_im, y_im);
        a_re = x_re; a_im = x_im;
        b_re = y_re; b_im = y_im;
    }
}

}  // namespace intfft
