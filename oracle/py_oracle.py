"""Second, independent CPU restatement of the intfftk integer FFT/IFFT (pure Python ints).

TEST INFRASTRUCTURE ONLY — never imported by the product package (intfftk_b200/).
PARITY UNPINNED by the reference (no golden vectors / asserting tests upstream); see the header
of oracle/intfft_oracle.c.  This module exists so that two restatements written in *different
formulations* must agree bit for bit:

  * intfft_oracle.c  — in-place radix-2 indexing (ia, ia+half), twiddles from a per-stage cache;
  * this file        — the reference's own streaming picture: two lanes of N/2 beats, a butterfly
                       per beat, a per-stage beat counter for the twiddle ROM, and the
                       cross-commutation between stages done as explicit block swaps, the way
                       int_delay_line does it (src/vhdl/delay/int_delay_line.vhd:52-104) and
                       math/fn_radix2.m models it (fn_rev2rdx / fn_rdx2rev, :51-89).

Arbitrary-precision Python ints + explicit wrap() make every bit-slice of the VHDL literal.
Small sizes only (pure-Python loops).
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass(frozen=True)
class Generics:
    """Entity generics of int_fftNk / int_ifftNk (src/vhdl/fft/int_fftNk.vhd:73-84)."""
    nfft_log2: int
    data_width: int = 16
    twdl_width: int = 16
    format: int = 0          # 1 unscaled, 0 scaled
    rndmode: int = 0         # 0 truncate, 1 round
    xser: int = 1            # 1 "NEW", 0 "OLD"
    use_fly: int = 1
    direction: int = 0       # 0 int_fftNk (DIF), 1 int_ifftNk (DIT)


def wrap(v: int, w: int) -> int:
    """VHDL slice (w-1 downto 0) read back as a signed number."""
    v &= (1 << w) - 1
    return v - (1 << w) if v >> (w - 1) else v


def vhdl_integer(x: float) -> int:
    """VHDL INTEGER(real): round to nearest, halves away from zero."""
    return int(math.copysign(math.floor(abs(x) + 0.5), x))


# ---------------------------------------------------------------- twiddles
def _rom(depth: int, i: int, awd: int) -> tuple[int, int]:
    # rom_twiddle_int.vhd:135-159
    mg = (2.0 ** (awd - 1)) - 1.0 if awd < 18 else (2.0 ** (awd - 2)) - 1.0
    ang = (float(i) * math.pi) / (2.0 ** (depth + 1))
    return vhdl_integer(mg * math.cos(ang)), vhdl_integer(mg * math.sin(-ang))


def twiddle(stage: int, cnt_value: int, awd: int, xser: int) -> tuple[int, int]:
    """(WW_RE, WW_IM) for counter value cnt_value of rom_twiddle_int(STAGE=stage), stage >= 2."""
    div = (cnt_value >> (stage - 1)) & 1                       # rom_twiddle_int.vhd:189
    addr = cnt_value & ((1 << (stage - 1)) - 1)                # :188
    if stage < 11:                                             # xSTD :205-212
        re, im = _rom(stage - 1, addr, awd)
        count = 0
    else:                                                      # xLNG :215-246
        re, im = _rom(9, addr >> (stage - 10), awd)
        count = addr & ((1 << (stage - 10)) - 1)
    if div:                                                    # pr_ww :174-184
        re, im = im, wrap(-re, awd)
    if stage < 11:
        return re, im
    # row_twiddle_tay.vhd — rom_ww = {im, re}: sin_aa <- low half (re), cos_aa <- high half (im)
    ii = stage - 11
    xshift = 21 if xser else 23                                # :123-132
    mathpi = vhdl_integer(math.pi * 2.0 ** (13 - ii - (2 if xser else 0)))  # :134-148
    mpi = (mathpi * count) & 0xFFFF                            # rom_pi entry, :213
    mpx = mpi >> 1                                             # :247
    sin_aa, cos_aa = re, im                                    # :250-251
    cos_prod = wrap((cos_aa << xshift) - sin_aa * mpx, 48)     # MULT_ADD: C - A*B (:304-312)
    sin_prod = wrap((sin_aa << xshift) + cos_aa * mpx, 48)     # MULT_SUB: C + A*B (:374-382)
    cos_pdt = cos_prod >> (xshift - 1)                         # :201-202
    sin_pdt = sin_prod >> (xshift - 1)
    cos_rnd = (cos_pdt >> 1) + (cos_pdt & 1)                   # pr_rnd :181-196
    sin_rnd = (sin_pdt >> 1) + (sin_pdt & 1)
    return wrap(sin_rnd, awd), wrap(cos_rnd, awd)              # rom_re <= sin_rnd, rom_im <= cos_rnd (:174-175)


# ---------------------------------------------------------------- complex multiplier
def _half(p2: int, p1: int, sub: bool, dtw: int, twd: int, xser: int) -> int:
    """One int_cmult*_dsp48 instance: MP_12 = M2_AA*M2_BB -/+ M1_AA*M1_BB, sliced per variant."""
    new = bool(xser)
    sgn = -1 if sub else 1
    if twd < 19:                                               # int_cmult_dsp48.vhd:182
        if dtw < (28 if new else 26):                          # single: P(DTW+TWD-2 downto TWD-1)
            return wrap((p2 + sgn * p1) >> (twd - 1), dtw)
        if dtw < (45 if new else 43):                          # int_cmult_dbl18_dsp48.vhd
            awd, pwd = (44, 62) if new else (42, 60)
            lo = pwd - 48 - (18 - twd)                         # :174-175
            t = wrap(wrap(p2 >> lo, 48) + sgn * wrap(p1 >> lo, 48), 48)
            return wrap(t >> (47 - awd), dtw)                  # :163
        if dtw < (79 if new else 77):                          # int_cmult_trpl18_dsp48.vhd:151-152
            return wrap(wrap(p2 >> (twd - 1), dtw) + sgn * wrap(p1 >> (twd - 1), dtw), dtw)
        raise ValueError("no multiplier generated")
    if twd < (28 if new else 26):                              # int_cmult_dsp48.vhd:307
        if dtw < 19:                                           # :316-317
            return wrap((p2 + sgn * p1) >> (twd - 2), dtw)
        if dtw < 36:                                           # int_cmult_dbl35_dsp48.vhd:155-160
            pwd, bwd = (62, 27) if new else (60, 25)
            lo = pwd - 48 - (bwd - twd) - 1
            t = wrap(wrap(p2 >> lo, 48) + sgn * wrap(p1 >> lo, 48), 48)
            return wrap(t >> 12, dtw)
        if dtw < 53:                                           # int_cmult_trpl52_dsp48.vhd:167-168
            return wrap(wrap(p2 >> (twd - 2), dtw) + sgn * wrap(p1 >> (twd - 2), dtw), dtw)
    raise ValueError("no multiplier generated")


def cmult(di_re: int, di_im: int, ww_re: int, ww_im: int, dtw: int, twd: int, xser: int):
    if twd < 19 and (45 if xser else 43) <= dtw < (79 if xser else 77):
        # trpl18: the 61x18 / 59x18 multiplier takes SXT(M1_AA, AWD) (int_cmult_trpl18_dsp48.vhd:161-162), so data
        # wider than AWD is cut there, and the product slice (:151-152) must lie inside its 79 / 77 bits
        awd, pwd = (61, 79) if xser else (59, 77)
        if dtw + twd - 2 > pwd - 1:
            raise ValueError("no multiplier generated (trpl18 product slice out of range)")
        di_re, di_im = wrap(di_re, awd), wrap(di_im, awd)
    do_re = _half(di_re * ww_re, di_im * ww_im, True, dtw, twd, xser)    # xMDSP_RE, XALU "SUB"
    do_im = _half(di_re * ww_im, di_im * ww_re, False, dtw, twd, xser)   # xMDSP_IM, XALU "ADD"
    return do_re, do_im


# ---------------------------------------------------------------- butterflies
def _negq(v: int, w: int) -> int:
    # int_dif2_fly.vhd:299-303: sign bit 0 -> not(x)+1, else not(x)
    return wrap(-v if v >= 0 else ~v, w)


def _addsub(g: Generics, dtw: int, a: int, b: int) -> tuple[int, int]:
    ow = dtw + g.format
    if g.format == 0 and g.rndmode == 0:
        return (a >> 1) + (b >> 1), (a >> 1) - (b >> 1)
    if g.format == 0:
        s, d = a + b, a - b
        return wrap((s >> 1) + (s & 1), ow), wrap((d >> 1) + (d & 1), ow)
    return wrap(a + b, ow), wrap(a - b, ow)


def dif_fly(g: Generics, stage: int, dtw: int, beat: int, a, b):
    """int_dif2_fly.vhd: returns (OA, OB) for inputs IA=a, IB=b (complex as (re, im) tuples)."""
    ow = dtw + g.format
    ad_re, su_re = _addsub(g, dtw, a[0], b[0])
    ad_im, su_im = _addsub(g, dtw, a[1], b[1])
    if stage == 0:
        ob = (su_re, su_im)
    elif stage == 1:
        ob = (su_re, su_im) if beat % 2 == 0 else (su_im, _negq(su_re, ow))
    else:
        w = twiddle(stage, beat % (1 << stage), g.twdl_width, g.xser)
        ob = cmult(su_re, su_im, w[0], w[1], ow, g.twdl_width, g.xser)
    return (ad_re, ad_im), ob


def dit_fly(g: Generics, stage: int, dtw: int, beat: int, a, b):
    """int_dit2_fly.vhd: returns (OA, OB)."""
    if stage == 0:
        bw = b
    elif stage == 1:
        bw = b if beat % 2 == 0 else (_negq(b[1], dtw), b[0])
    else:
        w = twiddle(stage, beat % (1 << stage), g.twdl_width, g.xser)
        do_re, do_im = cmult(b[1], b[0], w[0], w[1], dtw, g.twdl_width, g.xser)  # :304-322
        bw = (do_im, do_re)
    x_re, y_re = _addsub(g, dtw, a[0], bw[0])
    x_im, y_im = _addsub(g, dtw, a[1], bw[1])
    return (x_re, x_im), (y_re, y_im)


# ---------------------------------------------------------------- cross-commutation
def commute(la: list, lb: list, blocks: int) -> tuple[list, list]:
    """Block swap between lanes with `blocks` output blocks per lane.
    Output block j takes its A part from the first half and its B part from the second half of a
    2*size window of lane A (j even) or lane B (j odd) — int_delay_line.vhd:60-104 diagrams,
    fn_radix2.m:51-69."""
    size = len(la) // blocks
    oa, ob = [], []
    for j in range(blocks):
        src = la if j % 2 == 0 else lb
        start = 2 * (j // 2) * size
        oa.extend(src[start:start + size])
        ob.extend(src[start + size:start + 2 * size])
    return oa, ob


# ---------------------------------------------------------------- cores
def transform(g: Generics, frame: list) -> list:
    """One frame through int_fftNk (direction 0) or int_ifftNk (direction 1).
    `frame` is a list of N (re, im) int tuples in the flat stream order of include/intfft.h."""
    n = g.nfft_log2
    N = 1 << n
    assert len(frame) == N
    data = [(wrap(r, g.data_width), wrap(i, g.data_width)) for r, i in frame]
    if not g.use_fly:
        if g.format:
            m = (1 << g.data_width) - 1
            return [(r & m, i & m) for r, i in data]
        return data
    if g.direction == 0:
        la, lb = data[:N // 2], data[N // 2:]                   # int_fftNk.vhd:15-17
    else:
        la, lb = data[0::2], data[1::2]                         # int_ifftNk.vhd:15-17
    for ii in range(n):
        stage = ii if g.direction else n - 1 - ii
        dtw = g.data_width + ii * g.format
        oa, ob = [], []
        for beat in range(N // 2):
            fly = dit_fly if g.direction else dif_fly
            x, y = fly(g, stage, dtw, beat, la[beat], lb[beat])
            oa.append(x)
            ob.append(y)
        if ii < n - 1:
            # FFT delay line STAGE=ii: blocks of 2^(NFFT-ii-2) (int_fftNk.vhd:281-324);
            # IFFT delay line STAGE=NFFT-ii-2: blocks of 2^ii (int_ifftNk.vhd:289-312)
            blocks = (1 << (ii + 1)) if g.direction == 0 else (1 << (n - 1 - ii))
            la, lb = commute(oa, ob, blocks)
        else:
            la, lb = oa, ob
    if g.direction == 0:
        out = [None] * N
        out[0::2], out[1::2] = la, lb                           # int_fftNk.vhd:19-21
        return out
    return la + lb                                              # int_ifftNk.vhd:19-21
