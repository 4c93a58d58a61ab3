// Host-side fixed-point twiddle generator of the product: the tables rom_twiddle_int streams.
//
// Reference behaviour reproduced (paths relative to the reference root):
//   src/vhdl/twiddle/rom_twiddle_int.vhd:118-131  ROM depth: STAGE-1 for STAGE <= 10, else 9
//   src/vhdl/twiddle/rom_twiddle_int.vhd:135-159  quarter-wave ROM, amplitude 2^(AWD-1)-1 (AWD < 18)
//                                                  or 2^(AWD-2)-1, INTEGER() rounding
//   src/vhdl/twiddle/rom_twiddle_int.vhd:174-184  second quadrant: (re, im) <- (im, -re)
//   src/vhdl/twiddle/rom_twiddle_int.vhd:215-246  STAGE >= 11: 512-entry coarse ROM + counter to
//   src/vhdl/twiddle/row_twiddle_tay.vhd:123-268  first-order Taylor step on two DSP48 MACs
//
// Built as whole tables (coarse quadrant first, then the refinement sweep) — elaboration-time work
// in the reference (math_real at elaboration), plan-creation work here.
#include <cmath>
#include <cstdint>
#include <vector>

#include "intfft_internal.h"

namespace intfft {

namespace {

inline int64_t sext(int64_t v, int bits)
{
    const int sh = 64 - bits;
    return (int64_t)((uint64_t)v << sh) >> sh;
}

// quarter-wave table with 2^depth entries: angle = i * pi / 2^(depth+1)
void quarter_wave(int depth, int awd, std::vector<int64_t> &c, std::vector<int64_t> &s)
{
    const double amp = std::ldexp(1.0, awd < 18 ? awd - 1 : awd - 2) - 1.0;
    const double den = std::ldexp(1.0, depth + 1);
    const size_t cnt = (size_t)1 << depth;
    c.resize(cnt);
    s.resize(cnt);
    for (size_t i = 0; i < cnt; ++i) {
        const double ang = ((double)i * M_PI) / den;
        c[i] = std::llround(amp * std::cos(ang));
        s[i] = std::llround(amp * std::sin(-ang));
    }
}

}  // namespace

void twiddle_stage_table(int stage, int awd, int xser, int32_t *re, int32_t *im)
{
    const int depth = stage <= 10 ? stage - 1 : 9;
    std::vector<int64_t> qc, qs;
    quarter_wave(depth, awd, qc, qs);
    const int64_t quarter = (int64_t)1 << (stage - 1);    // entries per quadrant of this stage
    const int fine_bits = stage <= 10 ? 0 : stage - 10;   // counter bits handed to the Taylor block
    const int64_t fine = (int64_t)1 << fine_bits;

    // Taylor constants (row_twiddle_tay.vhd:123-148)
    const int xshift = xser ? 21 : 23;
    const int64_t mathpi = fine_bits ? std::llround(M_PI * std::ldexp(1.0, 13 - (stage - 11) - (xser ? 2 : 0))) : 0;

    for (int quad = 0; quad < 2; ++quad) {
        for (int64_t coarse = 0; coarse < (quarter >> fine_bits); ++coarse) {
            // coarse point after the quadrant mux
            int64_t lo = quad ? qs[coarse] : qc[coarse];               // -> WW_RE half
            int64_t hi = quad ? sext(-qc[coarse], awd) : qs[coarse];   // -> WW_IM half
            const int64_t k0 = quad * quarter + coarse * fine;
            if (!fine_bits) {
                re[k0] = (int32_t)lo;
                im[k0] = (int32_t)hi;
                continue;
            }
            for (int64_t cnt = 0; cnt < fine; ++cnt) {
                const int64_t mpx = ((mathpi * cnt) & 0xFFFF) >> 1;
                // 48-bit accumulators of the two DSP48s
                const int64_t acc_im = sext(hi * ((int64_t)1 << xshift) - lo * mpx, 48);
                const int64_t acc_re = sext(lo * ((int64_t)1 << xshift) + hi * mpx, 48);
                const int64_t t_im = acc_im >> (xshift - 1), t_re = acc_re >> (xshift - 1);
                re[k0 + cnt] = (int32_t)sext((t_re >> 1) + (t_re & 1), awd);
                im[k0 + cnt] = (int32_t)sext((t_im >> 1) + (t_im & 1), awd);
            }
        }
    }
}

void taylor_consts(int awd, int xser, int32_t *rom_c, int32_t *rom_s, int *mathpi, int *xshift)
{
    std::vector<int64_t> qc, qs;
    quarter_wave(9, awd, qc, qs);
    for (int i = 0; i < 512; ++i) { rom_c[i] = (int32_t)qc[i]; rom_s[i] = (int32_t)qs[i]; }
    for (int stage = 11; stage <= 19; ++stage)
        mathpi[stage - 11] = (int)std::llround(M_PI * std::ldexp(1.0, 13 - (stage - 11) - (xser ? 2 : 0)));
    *xshift = xser ? 21 : 23;
}

CmultConsts cmult_consts(int tw, int xser)
{
    CmultConsts c{};
    if (tw < 19) {                       // int_cmult_dsp48.vhd:182, dbl18 / trpl18 family
        const int awd = xser ? 44 : 42;  // int_cmult_dbl18_dsp48.vhd:129
        c.lim_single = xser ? 28 : 26;
        c.lim_dbl = xser ? 45 : 43;
        c.lim_none = xser ? 79 : 77;
        c.sh_single = tw - 1;
        c.k_pre = awd + tw - 48;         // :174-175
        c.sh_post = 47 - awd;            // :163
        c.trpl_awd = xser ? 61 : 59;     // int_cmult_trpl18_dsp48.vhd:117-129 (find_widthA)
        c.trpl_pwd = xser ? 79 : 77;     // :131-143 (find_widthP)
    } else {                             // int_cmult_dsp48.vhd:307, dbl35 / trpl52 family
        c.lim_single = 19;
        c.lim_dbl = 36;
        c.lim_none = 53;
        c.sh_single = tw - 2;
        c.k_pre = tw - 14;               // int_cmult_dbl35_dsp48.vhd:155-156
        c.sh_post = 12;                  // :160
        c.trpl_awd = 64;                 // trpl52: SXT(M1_AA, 52) with DTW < 53 never cuts
        c.trpl_pwd = 128;
    }
    return c;
}

}  // namespace intfft
