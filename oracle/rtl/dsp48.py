"""Bit-level behavioural model of the Xilinx DSP48E1 / DSP48E2 slice — TEST INFRASTRUCTURE ONLY.

The reference (hukenovs/intfftk) builds every multiplier and adder of the butterfly chain from these
vendor primitives (`library unisim; use unisim.vcomponents.DSP48E1 / DSP48E2`), which are NOT part of
the reference repository.  What is restated here is the primitive's published function (Xilinx UG479
"7 Series DSP48E1 Slice" and UG579 "UltraScale Architecture DSP Slice": the OPMODE / ALUMODE / INMODE /
CARRYINSEL tables and the SIMD carry-chain rules), restricted to the configurations the reference
instantiates:

  OPMODE (E1: 7 bits ZZZ YY XX; E2: 9 bits WW ZZZ YY XX, WW is always "00" in the reference)
      "0000101"  P = A*B                         mults/*.vhd (lowest limb), cmult18x25 M2, ...
      "0010101"  P = PCIN +- A*B                 math/cmult/int_cmult18x25_dsp48.vhd:xDSP_M1
      "1010101"  P = (PCIN >> 17) + A*B          math/mults/mlt42x18_dsp48e1.vhd:xDSP_M1 (limb cascade)
      "0110011"  P = C +- A:B                    math/int_addsub_dsp48.vhd, cmult_dbl*/trpl* final adders
      "0110101"  P = C +- A*B                    twiddle/row_twiddle_tay.vhd MULT_ADD / MULT_SUB
  ALUMODE  "0000" Z + X + Y + CIN,   "0011" Z - (X + Y + CIN)   (also 0001 / 0010 for completeness)
  INMODE   "00000" only (multiplier A input = A[24:0] (E1) / A[26:0] (E2), no pre-adder)
  CARRYINSEL "000" (CARRYIN pin) and "010" (CARRYCASCIN, 96-bit adders)
  USE_SIMD ONE48 / TWO24 (FOUR12 modelled as well)

The model is combinational: pipeline registers (AREG, MREG, PREG, ...) only delay values, and every
wrapper of the reference aligns its operands so that the same frame beat meets in the ALU, so the value
function is what matters.  The subtracter is modelled the way the silicon does it — not(not(Z) + X + Y + CIN)
— so that CARRYCASCOUT of a low slice carries the right borrow into the high slice of a 96-bit subtraction.

Every port is an UNSIGNED python int holding the bit vector of the stated width (A 30, B 18, C 48,
PCIN 48, P 48).
"""
from __future__ import annotations

M48 = (1 << 48) - 1


def mask(w: int) -> int:
    return (1 << w) - 1


def to_signed(v: int, w: int) -> int:
    """Bit vector (unsigned int of w bits) -> two's-complement value."""
    v &= mask(w)
    return v - (1 << w) if v >> (w - 1) else v


def to_bits(v: int, w: int) -> int:
    """Two's-complement value -> bit vector of w bits (wraps)."""
    return v & mask(w)


def bits(v: int, hi: int, lo: int) -> int:
    """VHDL slice v(hi downto lo) of a bit vector."""
    assert hi >= lo >= 0
    return (v >> lo) & mask(hi - lo + 1)


def sxt(v: int, w_from: int, w_to: int) -> int:
    """ieee.std_logic_arith SXT(v, w_to): sign-extend (or truncate) a w_from-bit vector to w_to bits."""
    return to_bits(to_signed(v, w_from), w_to)


class DSP48:
    """One slice.  series = "E1" (25 x 18 multiplier) or "E2" (27 x 18)."""

    def __init__(self, series: str, use_simd: str = "ONE48"):
        assert series in ("E1", "E2") and use_simd in ("ONE48", "TWO24", "FOUR12")
        self.series = series
        self.use_simd = use_simd
        self.mult_a_bits = 25 if series == "E1" else 27

    def __call__(self, A=0, B=0, C=0, D=0, PCIN=0, OPMODE="0000101", ALUMODE="0000", INMODE="00000",
                 CARRYIN=0, CARRYINSEL="000", CARRYCASCIN=0, P_prev=0):
        """Returns (P, CARRYCASCOUT).  PCOUT == P.  P_prev = the slice's own P register (X/Z = P modes)."""
        assert 0 <= A < (1 << 30) and 0 <= B < (1 << 18) and 0 <= C <= M48 and 0 <= PCIN <= M48
        assert INMODE.strip("0") == "", "only INMODE = 0 is used by the reference"
        if self.series == "E2":
            assert len(OPMODE) == 9 and OPMODE[:2] == "00", "W multiplexer is unused (00) in the reference"
            op = OPMODE[2:]
        else:
            assert len(OPMODE) == 7
            op = OPMODE
        zsel, ysel, xsel = op[0:3], op[3:5], op[5:7]
        # 25/27 x 18 two's-complement multiplier; its 43/45-bit product is sign-extended to 48 bits
        m = to_bits(to_signed(bits(A, self.mult_a_bits - 1, 0), self.mult_a_bits) * to_signed(B, 18), 48)
        if xsel == "01":
            assert ysel == "01", "X = M requires Y = M (the two partial products)"
            x, y = m, 0            # X + Y = M
        else:
            x = {"00": 0, "10": P_prev, "11": ((A << 18) | B) & M48}[xsel]
            y = {"00": 0, "10": M48, "11": C}[ysel]
        z = {"000": 0, "001": PCIN, "010": P_prev, "011": C, "100": P_prev,
             "101": to_bits(to_signed(PCIN, 48) >> 17, 48), "110": to_bits(to_signed(P_prev, 48) >> 17, 48)}[zsel]
        cin = {"000": CARRYIN, "010": CARRYCASCIN}[CARRYINSEL]
        if ALUMODE in ("0001", "0011"):
            z ^= M48
        lanes = {"ONE48": 1, "TWO24": 2, "FOUR12": 4}[self.use_simd]
        lw = 48 // lanes
        p, cout = 0, 0
        for i in range(lanes):
            s = bits(x, lw * i + lw - 1, lw * i) + bits(y, lw * i + lw - 1, lw * i) + bits(z, lw * i + lw - 1, lw * i)
            s += cin if i == 0 else 0          # SIMD lanes above the first get no carry in
            p |= (s & mask(lw)) << (lw * i)
            cout = (s >> lw) & 1               # CARRYOUT of the top lane (valid for two-operand adds)
        if ALUMODE in ("0010", "0011"):
            p ^= M48
        return p, cout
