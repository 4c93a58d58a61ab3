"""Port-map level transcription of the reference's arithmetic wrappers onto the DSP48 model (dsp48.py)
— TEST INFRASTRUCTURE ONLY.

Each function below is ONE entity of the reference, wired slice by slice as its architecture wires it:
which bits of which operand go to the A / B / C ports, which OPMODE / ALUMODE each slice runs, how PCOUT /
CARRYCASCOUT chain, and which bits of which P make up the result.  Nothing is simplified to `a * b`:
the point of this file is to CHECK that oracle/intfft_oracle.c's "exact product, then this slice"
reading of these entities is what the primitive-level netlist computes (tests/test_oracle_rtl.py).
Pipeline registers are dropped (value function only, see dsp48.py).

All signals are python ints holding std_logic_vector bit patterns (unsigned); paths are relative to
/root/reference/src/vhdl/.
"""
from __future__ import annotations

from .dsp48 import DSP48, bits, mask, sxt, to_bits, to_signed

Z = 0  # (others => '0')


def _series(xser: str) -> str:
    assert xser in ("OLD", "NEW")
    return "E1" if xser == "OLD" else "E2"


def _op(series: str, opmode7: str) -> str:
    """The reference writes "1010101" for DSP48E1 and "001010101" for DSP48E2 (W mux = 00)."""
    return opmode7 if series == "E1" else "00" + opmode7


# ---------------------------------------------------------------------------------------------------
# math/mults: wide multipliers as 17-bit-limb cascades
# ---------------------------------------------------------------------------------------------------
def mlt42x18_dsp48e1(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt42x18_dsp48e1.vhd:82-123,191.  MLT_A 42 bits, MLT_B 18 bits -> MLT_P 60 bits."""
    return _mlt_a2(MLT_A, MLT_B, "E1", 42)


def mlt44x18_dsp48e2(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt44x18_dsp48e2.vhd:82-123,191.  MLT_A 44 bits, MLT_B 18 bits -> MLT_P 62 bits."""
    return _mlt_a2(MLT_A, MLT_B, "E2", 44)


def _mlt_a2(MLT_A, MLT_B, series, aw):
    ma = aw - 17                                  # 25 (E1) / 27 (E2): the multiplier's A width
    dsp = DSP48(series)
    dspA_M2 = bits(MLT_A, 16, 0)                  # dspA_M2(16 downto 0) <= MLT_A(16 downto 0); (29 downto 17) <= '0'
    dspA_M1 = sxt(bits(MLT_A, aw - 1, 17), ma, 30)   # dspA_M1(ma-1 downto 0) <= MLT_A(aw-1 downto 17); rest = sign
    dspB_12 = MLT_B
    # xDSP_M2: OPMODE "0000101": P = A*B; PCOUT => dspP_12
    dspP_M2, _ = dsp(A=dspA_M2, B=dspB_12, OPMODE=_op(series, "0000101"))
    # xDSP_M1: OPMODE "1010101": P = (PCIN >> 17) + A*B; PCIN => dspP_12
    dspP_M1, _ = dsp(A=dspA_M1, B=dspB_12, PCIN=dspP_M2, OPMODE=_op(series, "1010101"))
    # MLT_P(16 downto 0) <= dspP_M2(16 downto 0); MLT_P(aw+17 downto 17) <= dspP_M1(aw downto 0)
    return bits(dspP_M2, 16, 0) | (bits(dspP_M1, aw, 0) << 17)


def mlt59x18_dsp48e1(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt59x18_dsp48e1.vhd:87-135,203,271.  59 x 18 -> 77 bits, three slices."""
    return _mlt_a3(MLT_A, MLT_B, "E1", 59)


def mlt61x18_dsp48e2(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt61x18_dsp48e2.vhd:87-135,203,271.  61 x 18 -> 79 bits, three slices."""
    return _mlt_a3(MLT_A, MLT_B, "E2", 61)


def _mlt_a3(MLT_A, MLT_B, series, aw):
    ma = aw - 34
    dsp = DSP48(series)
    dspA_M3 = bits(MLT_A, 16, 0)
    dspA_M2 = bits(MLT_A, 33, 17)
    dspA_M1 = sxt(bits(MLT_A, aw - 1, 34), ma, 30)
    dspB_12 = MLT_B                               # dspB_ZZ is dspB_12 one clock later (same beat at xDSP_M1)
    dspP_M3, _ = dsp(A=dspA_M3, B=dspB_12, OPMODE=_op(series, "0000101"))                    # PCOUT => dspP_23
    dspP_M2, _ = dsp(A=dspA_M2, B=dspB_12, PCIN=dspP_M3, OPMODE=_op(series, "1010101"))      # PCOUT => dspP_12
    dspP_M1, _ = dsp(A=dspA_M1, B=dspB_12, PCIN=dspP_M2, OPMODE=_op(series, "1010101"))
    # MLT_P(16:0) <= dspP_MZ(16:0) [= dspP_M3]; MLT_P(33:17) <= dspP_M2(16:0); MLT_P(aw+17:34) <= dspP_M1(aw-34+17 .. 0)
    return bits(dspP_M3, 16, 0) | (bits(dspP_M2, 16, 0) << 17) | (bits(dspP_M1, aw - 17, 0) << 34)


def mlt35x25_dsp48e1(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt35x25_dsp48e1.vhd:82-124,192.  MLT_A 35 bits (on the B ports, limbs), MLT_B 25 bits (A port)."""
    return _mlt_b2(MLT_A, MLT_B, "E1", 25)


def mlt35x27_dsp48e2(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt35x27_dsp48e2.vhd:83-125,193.  35 x 27 -> 62 bits."""
    return _mlt_b2(MLT_A, MLT_B, "E2", 27)


def _mlt_b2(MLT_A, MLT_B, series, bw):
    dsp = DSP48(series)
    dspA_12 = sxt(MLT_B, bw, 30)                  # dspA_12(bw-1 downto 0) <= MLT_B; upper bits = MLT_B(bw-1)
    dspB_M2 = bits(MLT_A, 16, 0)                  # dspB_M2(16 downto 0) <= MLT_A(16 downto 0); dspB_M2(17) <= '0'
    dspB_M1 = bits(MLT_A, 34, 17)                 # 18 bits, signed
    dspP_M2, _ = dsp(A=dspA_12, B=dspB_M2, OPMODE=_op(series, "0000101"))
    dspP_M1, _ = dsp(A=dspA_12, B=dspB_M1, PCIN=dspP_M2, OPMODE=_op(series, "1010101"))
    # MLT_P(16:0) <= dspP_M2(16:0); MLT_P(bw+34:17) <= dspP_M1(bw+17:0)   (42:0 for E1, 44:0 for E2)
    return bits(dspP_M2, 16, 0) | (bits(dspP_M1, bw + 17, 0) << 17)


def mlt52x25_dsp48e1(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt52x25_dsp48e1.vhd:87-136,204,273.  52 x 25 -> 77 bits."""
    return _mlt_b3(MLT_A, MLT_B, "E1", 25)


def mlt52x27_dsp48e2(MLT_A: int, MLT_B: int) -> int:
    """math/mults/mlt52x27_dsp48e2.vhd:87-136,204,273.  52 x 27 -> 79 bits."""
    return _mlt_b3(MLT_A, MLT_B, "E2", 27)


def _mlt_b3(MLT_A, MLT_B, series, bw):
    dsp = DSP48(series)
    dspA_12 = sxt(MLT_B, bw, 30)
    dspB_M3 = bits(MLT_A, 16, 0)
    dspB_M2 = bits(MLT_A, 33, 17)
    dspB_M1 = bits(MLT_A, 51, 34)                 # 18 bits, signed
    dspP_M3, _ = dsp(A=dspA_12, B=dspB_M3, OPMODE=_op(series, "0000101"))
    dspP_M2, _ = dsp(A=dspA_12, B=dspB_M2, PCIN=dspP_M3, OPMODE=_op(series, "1010101"))
    dspP_M1, _ = dsp(A=dspA_12, B=dspB_M1, PCIN=dspP_M2, OPMODE=_op(series, "1010101"))
    return bits(dspP_M3, 16, 0) | (bits(dspP_M2, 16, 0) << 17) | (bits(dspP_M1, bw + 17, 0) << 34)


# ---------------------------------------------------------------------------------------------------
# math/cmult: "half" complex multipliers  MP_12 = M2_AA*M2_BB +- M1_AA*M1_BB
# ---------------------------------------------------------------------------------------------------
def _alumode(XALU: str) -> str:
    return {"ADD": "0000", "SUB": "0011"}[XALU]    # Z + (X + Y) / Z - (X + Y)


def int_cmult18x25_dsp48(M1_AA, M1_BB, M2_AA, M2_BB, MAW, MBW, XALU, XSER) -> int:
    """math/cmult/int_cmult18x25_dsp48.vhd:112-126,161,229 (OLD) / 301,369 (NEW).  MP_12: 48 bits."""
    series = _series(XSER)
    dsp = DSP48(series)
    dspB_M1, dspB_M2 = sxt(M1_BB, MBW, 18), sxt(M2_BB, MBW, 18)
    dspA_M1, dspA_M2 = sxt(M1_AA, MAW, 30), sxt(M2_AA, MAW, 30)
    dspP_M2, _ = dsp(A=dspA_M2, B=dspB_M2, OPMODE=_op(series, "0000101"))                       # PCOUT => dspP_M2
    dspP_M1, _ = dsp(A=dspA_M1, B=dspB_M1, PCIN=dspP_M2, OPMODE=_op(series, "0010101"), ALUMODE=_alumode(XALU))
    return dspP_M1


def _final_add_48(series, dsp1_48, dsp2_48, XALU):
    """xDSP_ADD: A:B = dsp1_48, C = dsp2_48, OPMODE "0110011": P = C +- A:B."""
    dsp = DSP48(series)
    p, _ = dsp(A=bits(dsp1_48, 47, 18), B=bits(dsp1_48, 17, 0), C=dsp2_48, OPMODE=_op(series, "0110011"),
               ALUMODE=_alumode(XALU))
    return p


def int_cmult_dbl18_dsp48(M1_AA, M1_BB, M2_AA, M2_BB, MAW, MBW, XALU, XSER) -> int:
    """math/cmult/int_cmult_dbl18_dsp48.vhd:129-181,234,323.  MP_12: MAW bits."""
    series = _series(XSER)
    AWD, PWD = (44, 62) if XSER == "NEW" else (42, 60)
    mlt = mlt44x18_dsp48e2 if XSER == "NEW" else mlt42x18_dsp48e1
    dspP_M1 = mlt(sxt(M1_AA, MAW, AWD), sxt(M1_BB, MBW, 18))
    dspP_M2 = mlt(sxt(M2_AA, MAW, AWD), sxt(M2_BB, MBW, 18))
    hi, lo = PWD - 1 - (18 - MBW), PWD - 48 - (18 - MBW)        # :174-175
    dsp1_48, dsp2_48 = bits(dspP_M1, hi, lo), bits(dspP_M2, hi, lo)
    dspP_12 = _final_add_48(series, dsp1_48, dsp2_48, XALU)
    return bits(dspP_12, 47 - 1 - (AWD - MAW), 47 - AWD)        # :163


def int_cmult_dbl35_dsp48(M1_AA, M1_BB, M2_AA, M2_BB, MAW, MBW, XALU, XSER) -> int:
    """math/cmult/int_cmult_dbl35_dsp48.vhd:155-182,235,324.  MP_12: MAW bits."""
    series = _series(XSER)
    PWD, BWD = (62, 27) if XSER == "NEW" else (60, 25)
    mlt = mlt35x27_dsp48e2 if XSER == "NEW" else mlt35x25_dsp48e1
    dspP_M1 = mlt(sxt(M1_AA, MAW, 35), sxt(M1_BB, MBW, BWD))
    dspP_M2 = mlt(sxt(M2_AA, MAW, 35), sxt(M2_BB, MBW, BWD))
    hi, lo = PWD - 1 - (BWD - MBW) - 1, PWD - 48 - (BWD - MBW) - 1   # :163-164
    dsp1_48, dsp2_48 = bits(dspP_M1, hi, lo), bits(dspP_M2, hi, lo)
    dspP_12 = _final_add_48(series, dsp1_48, dsp2_48, XALU)
    return bits(dspP_12, 47 - 1 - (35 - MAW), 47 - 35)          # :168


def _final_add_trpl(series, dsp1_48, dsp2_48, MAW, XALU):
    """xDT48 (MAW < 49): one slice on sign-extended operands; xDT96 (MAW > 48): two slices, the low one's
    CARRYCASCOUT into the high one's CARRYCASCIN (CARRYINSEL "010") — int_cmult_trpl18_dsp48.vhd:170-292 / :368-445."""
    dsp = DSP48(series)
    if MAW < 49:
        dsp1_DT, dsp2_DT = sxt(dsp1_48, MAW, 48), sxt(dsp2_48, MAW, 48)
        p, _ = dsp(A=bits(dsp1_DT, 47, 18), B=bits(dsp1_DT, 17, 0), C=dsp2_DT, OPMODE=_op(series, "0110011"),
                   ALUMODE=_alumode(XALU))
        return bits(p, MAW - 1, 0)
    dsp1_LO, dsp2_LO = bits(dsp1_48, 47, 0), bits(dsp2_48, 47, 0)
    dsp1_HI = sxt(bits(dsp1_48, MAW - 1, 48), MAW - 48, 48)
    dsp2_HI = sxt(bits(dsp2_48, MAW - 1, 48), MAW - 48, 48)
    p_lo, cy = dsp(A=bits(dsp1_LO, 47, 18), B=bits(dsp1_LO, 17, 0), C=dsp2_LO, OPMODE=_op(series, "0110011"),
                   ALUMODE=_alumode(XALU))
    p_hi, _ = dsp(A=bits(dsp1_HI, 47, 18), B=bits(dsp1_HI, 17, 0), C=dsp2_HI, OPMODE=_op(series, "0110011"),
                  ALUMODE=_alumode(XALU), CARRYINSEL="010", CARRYCASCIN=cy)
    return p_lo | (bits(p_hi, MAW - 1 - 48, 0) << 48)


def int_cmult_trpl18_dsp48(M1_AA, M1_BB, M2_AA, M2_BB, MAW, MBW, XALU, XSER) -> int:
    """math/cmult/int_cmult_trpl18_dsp48.vhd:151-167 + final adder.  MP_12: MAW bits."""
    series = _series(XSER)
    AWD, PWD = (61, 79) if XSER == "NEW" else (59, 77)
    if MAW + MBW - 2 > PWD - 1:
        return None                                             # slice :151 out of range: elaboration error
    mlt = mlt61x18_dsp48e2 if XSER == "NEW" else mlt59x18_dsp48e1
    dspP_M1 = mlt(sxt(M1_AA, MAW, AWD), sxt(M1_BB, MBW, 18))
    dspP_M2 = mlt(sxt(M2_AA, MAW, AWD), sxt(M2_BB, MBW, 18))
    dsp1_48 = bits(dspP_M1, MAW + MBW - 2, MBW - 1)             # :151
    dsp2_48 = bits(dspP_M2, MAW + MBW - 2, MBW - 1)             # :152
    return _final_add_trpl(series, dsp1_48, dsp2_48, MAW, XALU)


def int_cmult_trpl52_dsp48(M1_AA, M1_BB, M2_AA, M2_BB, MAW, MBW, XALU, XSER) -> int:
    """math/cmult/int_cmult_trpl52_dsp48.vhd:159-170 + final adder.  MP_12: MAW bits."""
    series = _series(XSER)
    BWD = 27 if XSER == "NEW" else 25
    mlt = mlt52x27_dsp48e2 if XSER == "NEW" else mlt52x25_dsp48e1
    dspP_M1 = mlt(sxt(M1_AA, MAW, 52), sxt(M1_BB, MBW, BWD))
    dspP_M2 = mlt(sxt(M2_AA, MAW, 52), sxt(M2_BB, MBW, BWD))
    dsp1_48 = bits(dspP_M1, MAW + MBW - 2 - 1, MBW - 1 - 1)     # :166
    dsp2_48 = bits(dspP_M2, MAW + MBW - 2 - 1, MBW - 1 - 1)     # :167
    return _final_add_trpl(series, dsp1_48, dsp2_48, MAW, XALU)


def int_cmult_dsp48(DI_RE, DI_IM, WW_RE, WW_IM, DTW, TWD, XSER):
    """math/cmult/int_cmult_dsp48.vhd:182-434: the dispatcher.  Returns (DO_RE, DO_IM) as DTW-bit vectors, or
    None when no generate branch matches (the entity then drives nothing: "does not elaborate")."""
    SNGL, DBL, TRPL = (28, 45, 79) if XSER == "NEW" else (26, 43, 77)
    TWD_DSP = 28 if XSER == "NEW" else 26
    if TWD < 19:                                                                    # xGEN_TWD18
        if DTW < SNGL:                                                              # xGEN_SNGL :184-224
            P_RE = int_cmult18x25_dsp48(DI_IM, WW_IM, DI_RE, WW_RE, DTW, TWD, "SUB", XSER)
            P_IM = int_cmult18x25_dsp48(DI_IM, WW_RE, DI_RE, WW_IM, DTW, TWD, "ADD", XSER)
            return bits(P_RE, DTW + TWD - 2, TWD - 1), bits(P_IM, DTW + TWD - 2, TWD - 1)
        f = int_cmult_dbl18_dsp48 if DTW < DBL else (int_cmult_trpl18_dsp48 if DTW < TRPL else None)
        if f is None:
            return None
        r = (f(DI_IM, WW_IM, DI_RE, WW_RE, DTW, TWD, "SUB", XSER),                  # xMDSP_RE
             f(DI_IM, WW_RE, DI_RE, WW_IM, DTW, TWD, "ADD", XSER))                  # xMDSP_IM
        return None if r[0] is None else r
    if TWD < TWD_DSP:                                                               # xGEN_TWD25
        if DTW < 19:                                                                # xGEN_SNGL :309-352 (A = twiddle)
            P_RE = int_cmult18x25_dsp48(WW_IM, DI_IM, WW_RE, DI_RE, TWD, DTW, "SUB", XSER)
            P_IM = int_cmult18x25_dsp48(WW_RE, DI_IM, WW_IM, DI_RE, TWD, DTW, "ADD", XSER)
            return bits(P_RE, DTW + TWD - 3, TWD - 2), bits(P_IM, DTW + TWD - 3, TWD - 2)
        f = int_cmult_dbl35_dsp48 if DTW < 36 else (int_cmult_trpl52_dsp48 if DTW < 53 else None)
        if f is None:
            return None
        return (f(DI_IM, WW_IM, DI_RE, WW_RE, DTW, TWD, "SUB", XSER),
                f(DI_IM, WW_RE, DI_RE, WW_IM, DTW, TWD, "ADD", XSER))
    return None


# ---------------------------------------------------------------------------------------------------
# math/int_addsub_dsp48.vhd: OX = IA + IB, OY = IA - IB, DSPW -> DSPW+1 bits
# ---------------------------------------------------------------------------------------------------
def int_addsub_dsp48(IA_RE, IA_IM, IB_RE, IB_IM, DSPW, XSER):
    """Returns (OX_RE, OX_IM, OY_RE, OY_IM), each DSPW+1 bits.  xGEN_LOW (DSPW < 24, one TWO24 slice per
    output pair, :713-1018), xGEN_HIGH (24..47, ONE48, :112-710), xGEN_DBL (> 47, two cascaded slices, :1021-2190)."""
    series = _series(XSER)
    op = _op(series, "0110011")                                 # P = C +- A:B
    if DSPW < 24:
        dsp = DSP48(series, "TWO24")
        dspC_XY = sxt(IA_RE, DSPW, 24) | (sxt(IA_IM, DSPW, 24) << 24)
        dspAB = sxt(IB_RE, DSPW, 24) | (sxt(IB_IM, DSPW, 24) << 24)
        dspP_XX, _ = dsp(A=bits(dspAB, 47, 18), B=bits(dspAB, 17, 0), C=dspC_XY, OPMODE=op, ALUMODE="0000")
        dspP_YY, _ = dsp(A=bits(dspAB, 47, 18), B=bits(dspAB, 17, 0), C=dspC_XY, OPMODE=op, ALUMODE="0011")
        return (bits(dspP_XX, DSPW, 0), bits(dspP_XX, DSPW + 24, 24), bits(dspP_YY, DSPW, 0), bits(dspP_YY, DSPW + 24, 24))
    if DSPW < 48:
        dsp = DSP48(series, "ONE48")
        out = []
        for alumode in ("0000", "0011"):
            for ia, ib in ((IA_RE, IB_RE), (IA_IM, IB_IM)):
                ab, c = sxt(ib, DSPW, 48), sxt(ia, DSPW, 48)   # A:B = IB sign-extended, C = IA sign-extended
                p, _ = dsp(A=bits(ab, 47, 18), B=bits(ab, 17, 0), C=c, OPMODE=op, ALUMODE=alumode)
                out.append(bits(p, DSPW, 0))
        return tuple(out)
    dsp = DSP48(series, "ONE48")
    out = []
    for alumode in ("0000", "0011"):
        for ia, ib in ((IA_RE, IB_RE), (IA_IM, IB_IM)):
            a96, b96 = sxt(ia, DSPW, 96), sxt(ib, DSPW, 96)
            p1, cy = dsp(A=bits(b96, 47, 18), B=bits(b96, 17, 0), C=bits(a96, 47, 0), OPMODE=op, ALUMODE=alumode)
            p2, _ = dsp(A=bits(b96, 95, 66), B=bits(b96, 65, 48), C=bits(a96, 95, 48), OPMODE=op, ALUMODE=alumode,
                        CARRYINSEL="010", CARRYCASCIN=cy)
            out.append(p1 | (bits(p2, DSPW - 48, 0) << 48))
    return tuple(out)


# ---------------------------------------------------------------------------------------------------
# twiddle/row_twiddle_tay.vhd: first-order Taylor refinement of a coarse twiddle
# ---------------------------------------------------------------------------------------------------
def row_twiddle_tay(rom_ww: int, rom_cnt: int, AWD: int, XSER: str, ii: int, USE_MLT: bool = False):
    """rom_ww = im & re (2*AWD bits), rom_cnt = ii+1 bits.  Returns (rom_re, rom_im) as AWD-bit vectors.
    twiddle/row_twiddle_tay.vhd:123-148 (XSHIFT, MATHPI), :199-247 (mpi / mpx), :250-268 (A / C ports),
    :304-312 + :374-382 (OLD) / :454-462 + :524-530 (NEW) (the two MACs), :174-196 (rounding)."""
    import math
    series = _series(XSER)
    XSHIFT = 21 if XSER == "NEW" else 23
    del_val = 2 if XSER == "NEW" else 0
    MATHPI = round(math.pi * 2.0 ** (13 - ii - del_val))            # INTEGER(MATH_PI * 2.0**(13-ii-del_val))
    cnt_exp = rom_cnt & mask(ii + 1)                                # cnt_exp(ii downto 0) <= rom_cnt, upper bits '0'
    if not USE_MLT:
        mpi = (MATHPI * cnt_exp) & 0xFFFF                           # rom_pi(jj) = conv_std_logic_vector(MATHPI*jj, 16)
    else:
        mpi = ((MATHPI & 0xFFFF) * (cnt_exp & 0xFF)) & mask(24)     # unsigned(std_pi) * unsigned(cnt_exp)
    mpx = bits(mpi, 17, 1)                                          # mpx <= '0' & mpi(17 downto 1)
    sin_aa = sxt(bits(rom_ww, AWD - 1, 0), AWD, 30)                 # low half  (re)
    cos_aa = sxt(bits(rom_ww, 2 * AWD - 1, AWD), AWD, 30)           # high half (im)
    # C ports: the AWD-bit value at bits XSHIFT .. XSHIFT+AWD-1, sign above, zeros below
    cos_cc = to_bits(to_signed(cos_aa, 30) << XSHIFT, 48)
    sin_cc = to_bits(to_signed(sin_aa, 30) << XSHIFT, 48)
    dsp = DSP48(series)
    op = _op(series, "0110101")                                     # P = C +- A*B
    cos_prod, _ = dsp(A=sin_aa, B=mpx, C=cos_cc, OPMODE=op, ALUMODE="0011")    # MULT_ADD
    sin_prod, _ = dsp(A=cos_aa, B=mpx, C=sin_cc, OPMODE=op, ALUMODE="0000")    # MULT_SUB
    cos_pdt, sin_pdt = bits(cos_prod, 47, XSHIFT - 1), bits(sin_prod, 47, XSHIFT - 1)
    w = 48 - XSHIFT                                                 # cos_rnd is (47-XSHIFT downto 0)
    cos_rnd = (bits(cos_pdt, 48 - XSHIFT, 1) + (cos_pdt & 1)) & mask(w)
    sin_rnd = (bits(sin_pdt, 48 - XSHIFT, 1) + (sin_pdt & 1)) & mask(w)
    return bits(sin_rnd, AWD - 1, 0), bits(cos_rnd, AWD - 1, 0)     # rom_re <= sin_rnd, rom_im <= cos_rnd
