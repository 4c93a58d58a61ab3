"""The C oracle against the reference's arithmetic ONE LEVEL BELOW where it restates it: a bit-level DSP48E1 /
DSP48E2 model (oracle/rtl/dsp48.py) with the reference's wrappers wired onto it port map by port map
(oracle/rtl/netlist.py, fly.py).  oracle/intfft_oracle.c reads those wrappers as "exact product, then this
slice" / "exact sum"; these tests check that reading operand by operand, for every multiplier arrangement,
both device families (XSER) and the width corners of each arrangement.

What this pins: rows a5 (complex multiplier, 6 arrangements), a6 (add/sub), a7 (wide multipliers), a9 (the
two Taylor MACs) and the per-butterfly slices of a3 / a4.  What it cannot pin: anything an
RTL simulator would add over the port maps — register alignment between data and twiddles (assumption A1).
"""
import random

import pytest

from oracle import c_oracle as co
from oracle.rtl import netlist as nl
from oracle.rtl import fly as rfly
from oracle.rtl.dsp48 import DSP48, to_bits, to_signed, mask


def corners(w):
    lo, hi = -(1 << (w - 1)), (1 << (w - 1)) - 1
    return [lo, lo + 1, -1, 0, 1, hi - 1, hi, 0x5555555555555555555555 & hi, -(0x2AAAAAAAAAAAAAAAAAAAAA & hi) - 1]


def operands(rng, w, n):
    out = corners(w)
    while len(out) < n:
        # mix full-range values with small ones (sign-extension paths) and single-bit patterns (limb boundaries)
        k = rng.random()
        if k < 0.6:
            out.append(rng.randrange(-(1 << (w - 1)), 1 << (w - 1)))
        elif k < 0.8:
            out.append(rng.randrange(-(1 << min(w - 1, 17)), 1 << min(w - 1, 17)))
        else:
            b = rng.randrange(0, w - 1)
            out.append(rng.choice((1, -1)) * (1 << b) + rng.randrange(-2, 3))
            out[-1] = max(-(1 << (w - 1)), min((1 << (w - 1)) - 1, out[-1]))
    return out


# --------------------------------------------------------------------------------------------------------------
def test_dsp48_primitive_modes():
    """The OPMODE / ALUMODE combinations the reference uses, against their documented meaning."""
    rng = random.Random(1)
    for series, ma in (("E1", 25), ("E2", 27)):
        d = DSP48(series)
        pre = "" if series == "E1" else "00"
        for _ in range(300):
            a, b = rng.randrange(-(1 << (ma - 1)), 1 << (ma - 1)), rng.randrange(-(1 << 17), 1 << 17)
            c, pc = rng.randrange(-(1 << 47), 1 << 47), rng.randrange(-(1 << 47), 1 << 47)
            A, B, C, PC = to_bits(a, 30), to_bits(b, 18), to_bits(c, 48), to_bits(pc, 48)
            assert to_signed(d(A=A, B=B, OPMODE=pre + "0000101")[0], 48) == a * b
            assert d(A=A, B=B, PCIN=PC, OPMODE=pre + "0010101")[0] == to_bits(pc + a * b, 48)
            assert d(A=A, B=B, PCIN=PC, OPMODE=pre + "0010101", ALUMODE="0011")[0] == to_bits(pc - a * b, 48)
            assert d(A=A, B=B, PCIN=PC, OPMODE=pre + "1010101")[0] == to_bits((pc >> 17) + a * b, 48)
            assert d(A=A, B=B, C=C, OPMODE=pre + "0110101")[0] == to_bits(c + a * b, 48)
            assert d(A=A, B=B, C=C, OPMODE=pre + "0110101", ALUMODE="0011")[0] == to_bits(c - a * b, 48)
            ab = to_signed((A << 18 | B) & mask(48), 48)
            assert d(A=A, B=B, C=C, OPMODE=pre + "0110011")[0] == to_bits(c + ab, 48)
            assert d(A=A, B=B, C=C, OPMODE=pre + "0110011", ALUMODE="0011")[0] == to_bits(c - ab, 48)
    # SIMD: the carry chain is cut at bit 24 / 12
    d2 = DSP48("E2", "TWO24")
    p, _ = d2(A=to_bits(-1, 30), B=to_bits(-1, 18), C=1 | (1 << 24), OPMODE="000110011")     # (-1, -1) + (1, 1) per lane
    assert p == 0
    d1 = DSP48("E2", "ONE48")
    p, _ = d1(A=to_bits(-1, 30), B=to_bits(-1, 18), C=1, OPMODE="000110011")
    assert p == 0


@pytest.mark.parametrize("name,aw,bw", [("mlt42x18_dsp48e1", 42, 18), ("mlt44x18_dsp48e2", 44, 18),
                                        ("mlt59x18_dsp48e1", 59, 18), ("mlt61x18_dsp48e2", 61, 18),
                                        ("mlt35x25_dsp48e1", 35, 25), ("mlt35x27_dsp48e2", 35, 27),
                                        ("mlt52x25_dsp48e1", 52, 25), ("mlt52x27_dsp48e2", 52, 27)])
def test_wide_multipliers_are_exact(name, aw, bw):
    """a7: the 17-bit-limb cascades (P = (PCIN >> 17) + A*B, low limbs taken from each P) give the exact signed
    product — the assumption intfft_oracle.c makes when it writes `(i128)d * w`."""
    rng = random.Random(hash(name) & 0xFFFF)
    f = getattr(nl, name)
    for a in operands(rng, aw, 120):
        for b in operands(rng, bw, 12):
            got = to_signed(f(to_bits(a, aw), to_bits(b, bw)), aw + bw)
            assert got == a * b, (name, a, b)


def _variant_cases():
    cases = []
    for xser in ("OLD", "NEW"):
        sn, db, tr = (28, 45, 79) if xser == "NEW" else (26, 43, 77)
        for twd in (8, 12, 16, 17, 18):
            for dtw in sorted({8, 16, 18, sn - 1, sn, sn + 1, 32, 38, db - 1, db, db + 1, 48, 49, 52, 64}):
                if dtw <= 64:
                    cases.append((xser, dtw, twd))
        tmax = 27 if xser == "NEW" else 25
        for twd in sorted({19, 20, 24, tmax}):
            for dtw in (8, 16, 18, 19, 24, 32, 35, 36, 40, 48, 49, 52):
                cases.append((xser, dtw, twd))
    return cases


@pytest.mark.parametrize("xser,dtw,twd", _variant_cases())
def test_cmult_against_primitive_netlist(xser, dtw, twd):
    """a5: int_cmult_dsp48 and its five variants, netlist vs oracle, random + corner operands."""
    rng = random.Random(dtw * 1000 + twd * 10 + (xser == "NEW"))
    # twiddles as rom_twiddle_int produces them (|w| <= 2^(twd-1) - 1 below 18 bits, 2^(twd-2) - 1 from 18 on), plus raw corners
    wmax = (1 << (twd - 1)) - 1 if twd < 18 else (1 << (twd - 2)) - 1
    ws = [(wmax, 0), (0, -wmax), (-wmax, 0), (wmax, -wmax), (-(1 << (twd - 1)), (1 << (twd - 1)) - 1)]
    ws += [(rng.randrange(-wmax, wmax + 1), rng.randrange(-wmax, wmax + 1)) for _ in range(10)]
    ds = list(zip(operands(rng, dtw, 40), reversed(operands(rng, dtw, 40))))
    if nl.int_cmult_dsp48(0, 0, 0, 0, dtw, twd, xser) is None:
        # no generate branch matches, or a slice of the selected variant is out of range (trpl18: product slice beyond
        # the 77 / 79-bit product): the entity does not elaborate, and the oracle must refuse it as well
        with pytest.raises(ValueError):
            co.cmult(dtw, twd, 1 if xser == "NEW" else 0, 0, 0, 0, 0)
        return
    for d_re, d_im in ds:
        for w_re, w_im in ws:
            r = nl.int_cmult_dsp48(to_bits(d_re, dtw), to_bits(d_im, dtw), to_bits(w_re, twd), to_bits(w_im, twd),
                                   dtw, twd, xser)
            assert r is not None
            want = co.cmult(dtw, twd, 1 if xser == "NEW" else 0, d_re, d_im, w_re, w_im)
            assert (to_signed(r[0], dtw), to_signed(r[1], dtw)) == want, (xser, dtw, twd, d_re, d_im, w_re, w_im)


def test_cmult_elaboration_limits_match():
    """Where no generate branch of int_cmult_dsp48 matches, the oracle's orc_validate refuses the plan too."""
    for xser in ("OLD", "NEW"):
        x = 1 if xser == "NEW" else 0
        for twd in range(8, 30):
            for dtw in (8, 18, 19, 35, 36, 52, 53, 60, 64):
                r = nl.int_cmult_dsp48(0, 0, 0, 0, dtw, twd, xser)
                try:
                    co.cmult(dtw, twd, x, 0, 0, 0, 0)
                    ok = True
                except ValueError:
                    ok = False
                assert ok == (r is not None), (xser, dtw, twd)


@pytest.mark.parametrize("xser", ["OLD", "NEW"])
@pytest.mark.parametrize("dspw", [7, 15, 16, 22, 23, 24, 25, 31, 40, 47, 48, 49, 63, 64, 95])
def test_addsub_is_exact(xser, dspw):
    """a6: TWO24 SIMD (< 24 bits), ONE48 (24..47) and the two-slice carry-cascaded form (>= 48) all give the exact
    sum and difference in DSPW + 1 bits."""
    rng = random.Random(dspw)
    vals = operands(rng, dspw, 60)
    for i, a_re in enumerate(vals):
        a_im, b_re, b_im = vals[-1 - i], vals[(i * 7 + 3) % len(vals)], vals[(i * 11 + 5) % len(vals)]
        r = nl.int_addsub_dsp48(to_bits(a_re, dspw), to_bits(a_im, dspw), to_bits(b_re, dspw), to_bits(b_im, dspw), dspw, xser)
        got = tuple(to_signed(v, dspw + 1) for v in r)
        assert got == (a_re + b_re, a_im + b_im, a_re - b_re, a_im - b_im)


@pytest.mark.parametrize("xser", ["OLD", "NEW"])
@pytest.mark.parametrize("awd", [8, 12, 16, 17, 18, 24, 25])
def test_taylor_macs_against_primitive_netlist(xser, awd):
    """a9: row_twiddle_tay's two MACs + rounding on the DSP48 model == orc_twiddle for every STAGE >= 11 (NFFT up
    to 19 as the reference elaborates it; the ROM and the quadrant logic in front of it come from the oracle)."""
    import math
    g = co.generics(19, 16, awd, 0, 0, 1 if xser == "NEW" else 0)
    rng = random.Random(awd)
    mg = (1 << (awd - 1)) - 1 if awd < 18 else (1 << (awd - 2)) - 1
    for s in range(11, 19):
        ii = s - 11
        re_t, im_t = co.twiddle_table(g, s)
        ks = [0, 1, 2, (1 << (s - 10)) - 1, 1 << (s - 10), (1 << (s - 1)) - 1, 1 << (s - 1), (1 << (s - 1)) + 1, (1 << s) - 1]
        ks += [rng.randrange(1 << s) for _ in range(60)]
        for k in ks:
            q, a = k >> (s - 1), k & ((1 << (s - 1)) - 1)
            addrx, cnt = a >> (s - 10), a & ((1 << (s - 10)) - 1)
            ang = addrx * math.pi / 1024.0                              # ROM depth 9: rom_twiddle_int.vhd:149
            c, sn = round(mg * math.cos(ang)), round(mg * math.sin(-ang))
            lo, hi = (c, sn) if q == 0 else (sn, -c)                    # quadrant logic :174-184 (re, im) before the refinement
            rom_ww = to_bits(lo, awd) | (to_bits(hi, awd) << awd)       # rom_ww = im & re
            for use_mlt in (False, True):
                r = nl.row_twiddle_tay(rom_ww, cnt, awd, xser, ii, use_mlt)
                assert (to_signed(r[0], awd), to_signed(r[1], awd)) == (int(re_t[k]), int(im_t[k])), (xser, awd, s, k, use_mlt)


def _fly_cases():
    out = []
    for direction in (0, 1):
        for fmt, rnd in ((0, 0), (0, 1), (1, 0)):
            for xser in (0, 1):
                for dtw, twd in ((8, 8), (16, 16), (18, 16), (23, 16), (24, 16), (27, 16), (28, 18), (30, 16), (40, 16),
                                 (47, 16), (48, 16), (18, 24), (24, 24), (36, 24), (46, 12)):
                    out.append((direction, fmt, rnd, xser, dtw, twd))
    return out


@pytest.mark.parametrize("direction,fmt,rnd,xser,dtw,twd", _fly_cases())
def test_butterflies_against_primitive_netlist(direction, fmt, rnd, xser, dtw, twd):
    """a3 / a4: int_dif2_fly / int_dit2_fly wired from int_addsub_dsp48 + int_cmult_dsp48 netlists with the reference's
    input slices, vs the oracle's fly_dif / fly_dit, STAGE 0, 1 (both toggle states) and several multiplying stages."""
    if direction == 0 and fmt == 1 and rnd == 1:
        pytest.skip("does not elaborate")
    scale = 1 - fmt
    xs = "NEW" if xser else "OLD"
    g = co.generics(16, 16, twd, fmt, rnd, xser, 1, direction)
    rng = random.Random(dtw * 7 + twd)
    ow = dtw + 1 - scale
    dtwc = ow if direction == 0 else dtw
    if nl.int_cmult_dsp48(0, 0, 0, 0, dtwc, twd, xs) is None or dtw + 1 > 64:
        pytest.skip("no multiplier for these widths")
    vals = operands(rng, dtw, 24)
    for s in (0, 1, 2, 5, 10, 11, 14):
        re_t = im_t = None
        if s >= 2:
            re_t, im_t = co.twiddle_table(g, s)
        for i in range(len(vals)):
            k = rng.randrange(1 << s) if s else 0
            if s == 1:
                k = i & 1
            a_re, a_im, b_re, b_im = vals[i], vals[-1 - i], vals[(5 * i + 1) % len(vals)], vals[(3 * i + 2) % len(vals)]
            w_re, w_im = (int(re_t[k]), int(im_t[k])) if s >= 2 else (0, 0)
            f = rfly.int_dif2_fly if direction == 0 else rfly.int_dit2_fly
            r = f(to_bits(a_re, dtw), to_bits(a_im, dtw), to_bits(b_re, dtw), to_bits(b_im, dtw),
                  to_bits(w_re, twd), to_bits(w_im, twd), k & 1, s, dtw, twd, scale, rnd, xs)
            got = tuple(to_signed(v, ow) for v in r)
            want = co.fly(g, s, dtw, k, a_re, a_im, b_re, b_im)
            assert got == want, (direction, fmt, rnd, xs, dtw, twd, s, k, (a_re, a_im, b_re, b_im))
