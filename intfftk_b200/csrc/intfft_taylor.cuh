// On-device twiddle source for STAGE >= 11: rom_twiddle_int's 512-entry coarse ROM plus the first-order
// Taylor step of row_twiddle_tay on its two DSP48 MACs, recomputed where a kernel hoists its twiddles
// instead of being read from a 2^NFFT-entry table (N > 64K: the tables of a 2^20-point plan are 2 x 8 MB,
// the ROM is 4 KB and stays in L1).
//
// Reference behaviour reproduced (same rules as the host generator, intfft_twiddle.cpp, which stays the
// source of every table for STAGE < 12 and of the intfft_twiddles test hook):
//   src/vhdl/twiddle/rom_twiddle_int.vhd:215-246  addrx = top 9 bits of the quadrant index, cnt = the rest
//   src/vhdl/twiddle/rom_twiddle_int.vhd:174-184  second quadrant: (re, im) <- (im, -re), BEFORE the refinement
//   src/vhdl/twiddle/row_twiddle_tay.vhd:123-148  XSHIFT 21 / 23, MATHPI = round(pi * 2^(13 - ii - del))
//   src/vhdl/twiddle/row_twiddle_tay.vhd:213-247  mpx = ((MATHPI * cnt) mod 2^16) >> 1
//   src/vhdl/twiddle/row_twiddle_tay.vhd:260-268, 304-312, 374-382  P = C -+ A * B on 48-bit accumulators
//   src/vhdl/twiddle/row_twiddle_tay.vhd:178-196  round half up at bit XSHIFT, keep TWDL_WIDTH bits
#pragma once
#include <cuda_runtime.h>

#include "intfft_arith.cuh"
#include "intfft_internal.h"

namespace intfft {

// bits [sh, 48) of the 48-bit accumulator held in the low 48 bits of acc, sign-extended (48 - sh <= 32)
__device__ __forceinline__ int tay_field48(long long acc, int sh)
{
    unsigned lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(acc));
    return sgxt32((int)__funnelshift_r(lo, hi, sh), 48 - sh);
}

// W_s[k] for 11 <= s <= 19, 0 <= k < 2^s (raw, TWDL_WIDTH bits, not pre-shifted).  Every operand fits 32 bits
// (|ROM| < 2^26, mpx < 2^15), so each MAC is one shift + one IMAD.WIDE; the DSP48's 48-bit wrap and the rounding
// slice are one funnel shift + one sign extension.
__device__ __forceinline__ int2 taylor_twiddle(const TaylorDev &t, int s, unsigned k)
{
    const unsigned quad = k >> (s - 1);
    const unsigned a = k & ((1u << (s - 1)) - 1u);
    const int fb = s - 10;                                   // counter bits handed to the Taylor block
    const unsigned coarse = a >> fb, cnt = a & ((1u << fb) - 1u);
    const int2 q = __ldg(t.rom9 + coarse);                   // (c_i, s_i) = mg * (cos, -sin)(i pi / 1024)
    const int lo = quad ? q.y : q.x;                                         // -> WW_RE half
    const int hi = quad ? sgxt32(-q.x, t.tw) : q.y;                          // -> WW_IM half (wrap_TW(-c))
    const int mpx = (int)(((unsigned)t.mathpi[s - 11] * cnt) & 0xFFFFu) >> 1;
    const long long acc_im = (long long)(-lo) * mpx + ((long long)hi << t.xs);      // C - A * B
    const long long acc_re = (long long)hi * mpx + ((long long)lo << t.xs);         // C + A * B
    const int t_im = tay_field48(acc_im, t.xs - 1), t_re = tay_field48(acc_re, t.xs - 1);
    return make_int2(sgxt32((t_re + 1) >> 1, t.tw), sgxt32((t_im + 1) >> 1, t.tw));   // (t >> 1) + (t & 1)
}

// what a strided-pass hoist reads: table entry (1 << s) + k, or the Taylor recomputation shifted like the
// table the kernel would otherwise have been given (pre-shifted tables hold W << e)
__device__ __forceinline__ int2 hoist_twiddle(const int2 *table, const TaylorDev &t, int s, unsigned k)
{
    if (t.on && s >= 11) {
        const int2 w = taylor_twiddle(t, s, k);
        return make_int2((int)((unsigned)w.x << t.e), (int)((unsigned)w.y << t.e));
    }
    return __ldg(table + (1u << s) + k);
}

}  // namespace intfft
