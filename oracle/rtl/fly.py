"""int_dif2_fly / int_dit2_fly (src/vhdl/fft/) transcribed at signal level on top of netlist.py —
TEST INFRASTRUCTURE ONLY.  Every slice, generic and port association is the reference's; the registers
are dropped.  `dt_sw` is the STAGE = 1 toggle (0 on even beats of a frame, 1 on odd ones).

All signals are bit vectors (unsigned python ints); widths as declared in the entities.
"""
from __future__ import annotations

from .dsp48 import bits, mask
from .netlist import int_addsub_dsp48, int_cmult_dsp48


def _rnd(v: int, DTW: int) -> int:
    """pr_rnd: v(DTW downto 1) + v(0), kept in DTW bits."""
    return (bits(v, DTW, 1) + (v & 1)) & mask(DTW)


def _negq(v: int, w: int) -> int:
    """"not(v) + '1'" when the sign bit is '0', "not(v)" otherwise (int_dif2_fly.vhd:299-303)."""
    inv = v ^ mask(w)
    return (inv + 1) & mask(w) if not (v >> (w - 1)) & 1 else inv


def int_dif2_fly(IA_RE, IA_IM, IB_RE, IB_IM, WW_RE, WW_IM, dt_sw, STAGE, DTW, TFW, SCALE, RNDMODE, XSER):
    """src/vhdl/fft/int_dif2_fly.vhd:142-373.  Inputs DTW bits; returns (OA_RE, OA_IM, OB_RE, OB_IM), DTW+1-SCALE bits."""
    OW = DTW + 1 - SCALE
    if RNDMODE == 0 and SCALE == 1:                 # xTRUNC :144-164: DSPW => DTW-1, inputs (DTW-1 downto 1)
        ad_re, ad_im, su_re, su_im = int_addsub_dsp48(bits(IA_RE, DTW - 1, 1), bits(IA_IM, DTW - 1, 1),
                                                      bits(IB_RE, DTW - 1, 1), bits(IB_IM, DTW - 1, 1), DTW - 1, XSER)
    elif RNDMODE == 1 and SCALE == 1:               # xROUND :167-219
        r = int_addsub_dsp48(IA_RE, IA_IM, IB_RE, IB_IM, DTW, XSER)
        ad_re, ad_im, su_re, su_im = (_rnd(v, DTW) for v in r)
    else:                                           # xUNSCALED :221-241
        assert RNDMODE == 0, "RNDMODE = 1 with SCALE = 0 double-drives wz_re (:331-338): does not elaborate"
        ad_re, ad_im, su_re, su_im = int_addsub_dsp48(IA_RE, IA_IM, IB_RE, IB_IM, DTW, XSER)
    if STAGE == 0:                                  # xST0 :245-255
        return ad_re, ad_im, su_re, su_im
    if STAGE == 1:                                  # xST1 :259-318
        if dt_sw == 0:
            return ad_re, ad_im, su_re, su_im
        return ad_re, ad_im, su_im, _negq(su_re, OW)
    db = int_cmult_dsp48(su_re, su_im, WW_RE, WW_IM, OW, TFW, XSER)     # xSTn :322-373, DTW => DTW+1-SCALE
    assert db is not None, "no multiplier is generated for these widths"
    return ad_re, ad_im, db[0], db[1]


def int_dit2_fly(IA_RE, IA_IM, IB_RE, IB_IM, WW_RE, WW_IM, dt_sw, STAGE, DTW, TFW, SCALE, RNDMODE, XSER):
    """src/vhdl/fft/int_dit2_fly.vhd:140-325.  Inputs DTW bits; returns (OA_RE, OA_IM, OB_RE, OB_IM), DTW+1-SCALE bits."""
    if STAGE == 0:                                  # xST0 :221-230
        bw_re, bw_im = IB_RE, IB_IM
    elif STAGE == 1:                                # xST1 :234-286
        if dt_sw == 0:
            bw_re, bw_im = IB_RE, IB_IM
        else:
            bw_im, bw_re = IB_RE, _negq(IB_IM, DTW)
    else:                                           # xSTn :289-325: DI_RE => IB_IM, DI_IM => IB_RE, DO_RE => bw_im, DO_IM => bw_re
        do = int_cmult_dsp48(IB_IM, IB_RE, WW_RE, WW_IM, DTW, TFW, XSER)
        assert do is not None, "no multiplier is generated for these widths"
        bw_im, bw_re = do
    az_re, az_im = IA_RE, IA_IM
    if SCALE == 0 or RNDMODE == 0:                  # xUNSCALED :142-162: DSPW => DTW-SCALE, inputs (DTW-1 downto SCALE)
        return int_addsub_dsp48(bits(az_re, DTW - 1, SCALE), bits(az_im, DTW - 1, SCALE),
                                bits(bw_re, DTW - 1, SCALE), bits(bw_im, DTW - 1, SCALE), DTW - SCALE, XSER)
    r = int_addsub_dsp48(az_re, az_im, bw_re, bw_im, DTW, XSER)          # xROUND :164-217
    return tuple(_rnd(v, DTW) for v in r)
