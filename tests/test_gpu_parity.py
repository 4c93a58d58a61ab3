"""Parity of the CUDA path (through the C-ABI) against the CPU oracle: bit-exact, same seeded inputs.

Every test here needs a B200 (`-m gpu`).  Inputs are full-scale uniform (so two's-complement wrap in
the multipliers is exercised, as in the reference) unless a test says otherwise."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")


@pytest.fixture(scope="module")
def ib():
    _need_gpu()
    import intfftk_b200
    intfftk_b200.lib()          # raises if the CUDA library is missing: no fallback
    return intfftk_b200


def _run_both(ib, co, batch, seed=1, threads=0, via="host", **kw):
    g = ib.Generics(**{k: v for k, v in kw.items() if k != "direction"})
    direction = kw.get("direction", 0)
    og = co.generics(g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, 1 if g.XSER == "NEW" else 0,
                     g.USE_FLY, direction)
    n = 1 << g.NFFT
    x = co.fill_random(batch * n * 2, g.DATA_WIDTH, seed).reshape(batch, n, 2)
    want = co.batch(og, x, threads)
    core = ib.Core(g, batch, direction)
    if via == "host":
        got = core.exec_host(x)
    else:
        d_in = torch.from_numpy(x).cuda()
        got = core.exec(d_in).cpu().numpy()
    core.close()
    return got, want


SMALL = []
for nfft in (3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13):
    for direction in (0, 1):
        for fmt, rnd in ((0, 0), (0, 1), (1, 0)):
            SMALL.append((nfft, direction, fmt, rnd))


@pytest.mark.parametrize("nfft,direction,fmt,rnd", SMALL)
def test_parity_16bit_all_modes(ib, oracle, nfft, direction, fmt, rnd):
    batch = max(3, 5000 >> nfft)
    got, want = _run_both(ib, oracle, batch, seed=nfft * 8 + direction * 4 + fmt * 2 + rnd, NFFT=nfft, DATA_WIDTH=16,
                          TWDL_WIDTH=16, FORMAT=fmt, RNDMODE=rnd, direction=direction)
    assert got.dtype == want.dtype
    assert np.array_equal(got, want)


WIDTHS = [(8, 8, "NEW"), (12, 16, "OLD"), (18, 16, "NEW"), (18, 18, "NEW"), (24, 17, "OLD"), (25, 16, "OLD"),
          (27, 16, "NEW"), (28, 16, "NEW"), (31, 12, "NEW"), (32, 16, "OLD"),
          (16, 19, "NEW"), (18, 24, "NEW"), (19, 25, "OLD"), (30, 27, "NEW"),          # TW >= 19: single25 / dbl35
          (36, 16, "NEW"), (40, 18, "NEW"), (44, 16, "OLD"), (45, 16, "NEW"), (50, 16, "NEW"),   # dbl18 / trpl18
          (36, 22, "NEW"), (48, 19, "OLD"), (52, 27, "NEW"), (60, 16, "NEW"), (64, 10, "OLD"),
          (62, 16, "OLD"), (63, 16, "NEW"), (64, 8, "NEW")]   # trpl18 beyond its 59 / 61-bit data port (operand cut)


@pytest.mark.parametrize("dw,tw,xser", WIDTHS)
@pytest.mark.parametrize("direction", [0, 1])
def test_parity_widths_and_multiplier_variants(ib, oracle, dw, tw, xser, direction):
    for fmt, rnd in ((0, 0), (0, 1), (1, 0)):
        nfft = 7
        if tw >= 19 and dw + fmt * nfft + (1 - direction) * fmt > 52:
            continue                                    # no trpl52 beyond 52 bits
        if dw + fmt * nfft + (rnd if not fmt else 0) > 64:
            continue                                    # beyond the 64-bit lanes
        if ib.validate(ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd, XSER=xser), direction):
            continue                                    # trpl18 product slice out of range: does not elaborate
        got, want = _run_both(ib, oracle, 9, seed=dw * 100 + tw, NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw,
                              FORMAT=fmt, RNDMODE=rnd, XSER=xser, direction=direction)
        assert np.array_equal(got, want), (fmt, rnd)


@pytest.mark.parametrize("nfft", [8, 9, 10, 11, 12])
@pytest.mark.parametrize("direction", [0, 1])
def test_parity_packed16_kernel_widths(ib, oracle, nfft, direction):
    """The specialised packed-16 kernel (scaled TRUNCATE, DW <= 16, TW <= 16) at narrower widths,
    ragged batches (last tile only partly filled) and both XSER settings."""
    for dw, tw, xser in ((16, 16, "OLD"), (12, 16, "NEW"), (16, 12, "NEW"), (10, 9, "OLD"), (8, 8, "NEW"), (15, 16, "NEW")):
        batch = (3 << (12 - nfft)) + 1
        got, want = _run_both(ib, oracle, batch, seed=nfft + dw, via="device", NFFT=nfft, DATA_WIDTH=dw,
                              TWDL_WIDTH=tw, FORMAT=0, RNDMODE=0, XSER=xser, direction=direction)
        assert np.array_equal(got, want), (dw, tw, xser)


@pytest.mark.parametrize("nfft", [3, 4, 5, 6, 7])
@pytest.mark.parametrize("direction", [0, 1])
def test_parity_tiny_frames_all_width_classes(ib, oracle, nfft, direction):
    """8..128-point frames on the specialised kernels: packed-16 with DATA_WIDTH < 16 and a narrow twiddle,
    32-bit lanes (scaled, rounding, unscaled growth, packed 16-bit in / 32-bit out), ragged last tile."""
    cases = [dict(DATA_WIDTH=12, TWDL_WIDTH=14, FORMAT=0, RNDMODE=0), dict(DATA_WIDTH=9, TWDL_WIDTH=16, FORMAT=0, RNDMODE=1),
             dict(DATA_WIDTH=18, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0), dict(DATA_WIDTH=24, TWDL_WIDTH=18, FORMAT=0, RNDMODE=1),
             dict(DATA_WIDTH=20, TWDL_WIDTH=16, FORMAT=1, RNDMODE=0), dict(DATA_WIDTH=16, TWDL_WIDTH=16, FORMAT=1, RNDMODE=0),
             dict(DATA_WIDTH=16, TWDL_WIDTH=20, FORMAT=0, RNDMODE=0)]
    for kw in cases:
        batch = (9000 >> nfft) + 3
        got, want = _run_both(ib, oracle, batch, seed=nfft * 11 + direction, via="device", NFFT=nfft, direction=direction, **kw)
        assert got.dtype == want.dtype and np.array_equal(got, want), kw


def test_packed16_kernel_matches_generic_kernel(ib, oracle, monkeypatch):
    """Same plan through both device kernels (INTFFT_DISABLE_FAST16 selects the generic one)."""
    g = ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0)
    x = torch.from_numpy(oracle.fill_random(40 * 4096 * 2, 16, 123).reshape(40, 4096, 2)).cuda()
    for direction in (0, 1):
        fast = ib.Core(g, 40, direction)
        monkeypatch.setenv("INTFFT_DISABLE_FAST16", "1")
        slow = ib.Core(g, 40, direction)
        monkeypatch.delenv("INTFFT_DISABLE_FAST16")
        assert torch.equal(fast.exec(x), slow.exec(x))
        fast.close(); slow.close()


@pytest.mark.parametrize("nfft,dw,fmt,direction", [(14, 16, 0, 0), (14, 16, 0, 1), (15, 18, 1, 0), (16, 24, 1, 0),
                                                   (16, 16, 0, 1), (17, 16, 0, 0), (18, 12, 1, 1)])
def test_parity_multipass(ib, oracle, nfft, dw, fmt, direction):
    got, want = _run_both(ib, oracle, 3, seed=nfft, via="device", NFFT=nfft, DATA_WIDTH=dw, FORMAT=fmt,
                          direction=direction)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("nfft", [13, 14, 15, 16, 17, 18, 19])
@pytest.mark.parametrize("direction", [0, 1])
def test_parity_packed16_two_pass(ib, oracle, nfft, direction):
    """NFFT 13..19, 16-bit scaled: strided packed-16 pass (top 4 / 8 bits) + contiguous packed-16 pass."""
    for dw, tw, batch in ((16, 16, 3), (12, 14, 2)):
        got, want = _run_both(ib, oracle, batch, seed=nfft * 3 + direction, via="device", NFFT=nfft, DATA_WIDTH=dw,
                              TWDL_WIDTH=tw, FORMAT=0, RNDMODE=0, direction=direction)
        assert np.array_equal(got, want), (dw, tw)


LANE32 = [  # (dw, tw, xser): 32-bit-lane kernels; covers single / dbl18 / single25 / dbl35 arrangements
    (16, 16, "NEW"), (18, 16, "NEW"), (18, 18, "OLD"), (20, 17, "NEW"), (24, 16, "OLD"), (25, 16, "OLD"), (26, 16, "OLD"),
    (27, 16, "NEW"), (28, 16, "NEW"), (31, 12, "NEW"), (32, 16, "OLD"), (9, 8, "NEW"),
    (16, 19, "NEW"), (18, 24, "NEW"), (19, 25, "OLD"), (24, 27, "NEW"), (30, 20, "OLD")]


@pytest.mark.parametrize("dw,tw,xser", LANE32)
@pytest.mark.parametrize("direction", [0, 1])
def test_parity_lane32_kernels(ib, oracle, dw, tw, xser, direction):
    """Every mode through the 32-bit-lane kernels: one-pass sizes (8..12) and two-pass sizes (13, 14, 17)."""
    for nfft in (8, 9, 10, 11, 12, 13, 14, 17):
        for fmt, rnd in ((0, 0), (0, 1), (1, 0)):
            if dw + fmt * nfft + (rnd if not fmt else 0) > 32:
                continue                               # not a 32-bit-lane plan (covered elsewhere)
            if direction == 0 and fmt == 1 and rnd == 1:
                continue
            batch = 2 if nfft >= 13 else (2 << (12 - nfft)) + 1
            got, want = _run_both(ib, oracle, batch, seed=dw * 7 + tw + nfft, via="device", NFFT=nfft, DATA_WIDTH=dw,
                                  TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd, XSER=xser, direction=direction)
            assert np.array_equal(got, want), (nfft, fmt, rnd)


def test_lane32_kernels_match_generic_kernel(ib, oracle, monkeypatch):
    for kw in (dict(NFFT=13, DATA_WIDTH=18, FORMAT=0), dict(NFFT=12, DATA_WIDTH=16, FORMAT=1),
               dict(NFFT=10, DATA_WIDTH=20, FORMAT=0, RNDMODE=1)):
        g = ib.Generics(**kw)
        n = 1 << g.NFFT
        x = torch.from_numpy(oracle.fill_random(6 * n * 2, g.DATA_WIDTH, 5).reshape(6, n, 2)).cuda()
        for direction in (0, 1):
            fast = ib.Core(g, 6, direction)
            monkeypatch.setenv("INTFFT_DISABLE_FAST16", "1")
            slow = ib.Core(g, 6, direction)
            monkeypatch.delenv("INTFFT_DISABLE_FAST16")
            assert torch.equal(fast.exec(x), slow.exec(x))
            fast.close(); slow.close()


LANE64 = [  # (nfft, dw, tw, xser, fmt, rnd): plans whose STAGE 7..0 run on the 64-bit-lane warp-centric kernel
    (8, 33, 16, "NEW", 0, 0), (8, 36, 16, "OLD", 0, 1), (8, 40, 18, "NEW", 1, 0), (8, 44, 16, "NEW", 0, 0),   # dbl18
    (8, 46, 16, "NEW", 0, 0), (8, 47, 16, "OLD", 0, 1), (8, 45, 12, "NEW", 1, 0),                             # trpl18
    (8, 33, 24, "NEW", 0, 0), (8, 35, 27, "NEW", 0, 1), (8, 34, 19, "OLD", 0, 0),                             # dbl35
    (8, 36, 24, "NEW", 0, 1), (8, 38, 22, "OLD", 1, 0), (8, 40, 24, "NEW", 0, 0),                             # trpl52
    (12, 30, 16, "NEW", 1, 0), (12, 34, 16, "NEW", 1, 0), (12, 40, 16, "OLD", 0, 0), (13, 36, 16, "NEW", 0, 1),
    (14, 33, 17, "NEW", 0, 0), (15, 26, 16, "NEW", 1, 0), (16, 24, 16, "NEW", 1, 0), (16, 24, 16, "OLD", 1, 0),
    (16, 38, 20, "NEW", 0, 0),
    # STAGE 7..0 crossing the 32-bit line and / or mixing single and double arrangements: the per-stage instance
    (12, 24, 16, "NEW", 1, 0), (12, 27, 16, "NEW", 1, 0), (12, 31, 16, "OLD", 1, 0), (8, 28, 16, "NEW", 1, 0),
    (16, 18, 16, "NEW", 1, 0), (16, 20, 17, "OLD", 1, 0), (12, 32, 16, "NEW", 0, 1), (8, 30, 20, "NEW", 1, 0),
    (12, 22, 18, "NEW", 1, 0), (8, 32, 16, "OLD", 0, 1), (12, 26, 24, "NEW", 1, 0),
    # double and triple arrangements inside STAGE 7..0 (the IFFT of c3's spectrum: 40 -> 56 bits)
    (16, 40, 16, "NEW", 1, 0), (8, 41, 16, "NEW", 1, 0), (12, 38, 16, "OLD", 1, 0), (8, 34, 20, "NEW", 1, 0)]


@pytest.mark.parametrize("nfft,dw,tw,xser,fmt,rnd", LANE64)
@pytest.mark.parametrize("direction", [0, 1])
def test_parity_lane64_kernel(ib, oracle, nfft, dw, tw, xser, fmt, rnd, direction):
    """64-bit-lane kernel (intfft_fast64.cu): double / triple multiplier arrangements of both TWDL_WIDTH
    families, all three modes, both directions, ragged batches; c3's generics are the (16, 24) rows."""
    if direction == 0 and fmt == 1 and rnd == 1:
        pytest.skip("does not elaborate")
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd, XSER=xser)
    if ib.validate(g, direction) != 0:
        pytest.skip("generics do not elaborate / exceed 64-bit lanes")
    batch = 2 if nfft >= 12 else 7
    got, want = _run_both(ib, oracle, batch, seed=nfft * 64 + dw, via="device", NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw,
                          FORMAT=fmt, RNDMODE=rnd, XSER=xser, direction=direction)
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_lane64_kernel_matches_generic_kernel(ib, oracle, monkeypatch):
    for kw in (dict(NFFT=16, DATA_WIDTH=24, FORMAT=1), dict(NFFT=8, DATA_WIDTH=40, FORMAT=0), dict(NFFT=12, DATA_WIDTH=35, FORMAT=0, RNDMODE=1)):
        g = ib.Generics(**kw)
        n = 1 << g.NFFT
        x = torch.from_numpy(oracle.fill_random(3 * n * 2, g.DATA_WIDTH, 5).reshape(3, n, 2)).cuda()
        for direction in (0, 1):
            if ib.validate(g, direction) != 0:
                continue
            fast = ib.Core(g, 3, direction)
            monkeypatch.setenv("INTFFT_DISABLE_FAST16", "1")
            slow = ib.Core(g, 3, direction)
            monkeypatch.delenv("INTFFT_DISABLE_FAST16")
            assert torch.equal(fast.exec(x), slow.exec(x))
            fast.close(); slow.close()


@pytest.mark.parametrize("xser", ["NEW", "OLD"])
def test_parity_nfft20_taylor_extension(ib, oracle, xser):
    """BASELINE config c4 shape (one frame): 2^20 points, Taylor twiddles on STAGE 11..19."""
    got, want = _run_both(ib, oracle, 1, seed=20, via="device", NFFT=20, DATA_WIDTH=16, FORMAT=0, XSER=xser)
    assert np.array_equal(got, want)
    got, want = _run_both(ib, oracle, 2, seed=21, via="device", NFFT=20, DATA_WIDTH=16, FORMAT=0, XSER=xser, direction=1)
    assert np.array_equal(got, want)


def test_parity_c2_sample(ib, oracle):
    """BASELINE config c2 generics, 256 frames, device path."""
    got, want = _run_both(ib, oracle, 256, seed=0x696E7466, via="device", NFFT=12, DATA_WIDTH=16, FORMAT=0)
    assert np.array_equal(got, want)


def test_parity_c5_sample(ib, oracle):
    """BASELINE config c5 generics (8192-pt 18-bit DIT), scaled and unscaled."""
    for fmt in (0, 1):
        got, want = _run_both(ib, oracle, 64, seed=5 + fmt, via="device", NFFT=13, DATA_WIDTH=18, FORMAT=fmt,
                              direction=1)
        assert np.array_equal(got, want)


def test_rounding_difference_wrap_edge(ib, oracle):
    """ROUNDING mode: the rounded difference 2^(DTW-1) wraps to -2^(DTW-1) (both device kernel families)."""
    for nfft, dw, tw in ((7, 9, 8), (10, 9, 8), (12, 16, 16), (13, 12, 10)):
        n = 1 << nfft
        hi, lo = (1 << (dw - 1)) - 1, -(1 << (dw - 1))
        x = np.empty((3, n, 2), oracle.scalar_dtype(dw))
        x[:, : n // 2] = (hi, lo)
        x[:, n // 2:] = (lo, hi)
        x[1, ::2] = (lo, hi)
        x[2, 1::3] = (hi, hi)
        for direction in (0, 1):
            g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=0, RNDMODE=1)
            want = oracle.batch(oracle.generics(nfft, dw, tw, 0, 1, 1, 1, direction), x)
            core = ib.Core(g, 3, direction)
            got = core.exec_host(x)
            core.close()
            assert np.array_equal(got, want), (nfft, dw, direction)


@pytest.mark.parametrize("kw", [dict(NFFT=7, DATA_WIDTH=16, FORMAT=1), dict(NFFT=12, DATA_WIDTH=16, FORMAT=0),
                                dict(NFFT=10, DATA_WIDTH=12, FORMAT=1, XSER="OLD"), dict(NFFT=9, DATA_WIDTH=18, FORMAT=0, RNDMODE=1),
                                dict(NFFT=14, DATA_WIDTH=16, FORMAT=0)])
def test_fft_ifft_pair(ib, oracle, kw):
    """f2: int_fft_ifft_pair = int_fftNk -> int_ifftNk(DATA_WIDTH + FORMAT*NFFT), spectrum kept on the device."""
    g = ib.Generics(**kw)
    n, batch = 1 << g.NFFT, 5
    xs = 1 if g.XSER == "NEW" else 0
    x = oracle.fill_random(batch * n * 2, g.DATA_WIDTH - 1, 11).reshape(batch, n, 2)     # one bit of headroom
    x = x.astype(oracle.scalar_dtype(g.DATA_WIDTH))
    mid = oracle.batch(oracle.generics(g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, xs, 1, 0), x)
    want = oracle.batch(oracle.generics(g.NFFT, g.out_width, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, xs, 1, 1), mid)
    pair = ib.Pair(g, batch)
    got = pair.exec(torch.from_numpy(x).cuda()).cpu().numpy()
    pair.close()
    assert got.dtype == want.dtype and np.array_equal(got, want)


def test_cpp_host_replays_testbench_files(ib, oracle, tmp_path):
    """f3: host/intfft_host.cpp (stand-in for src/vhdl/tb/*.vhd) with the reference's file formats:
    di_single.dat ("re im" per line, math/fft_single.m:94-98) and the two-lane four-column format of
    di_double.dat / dout_pair.dat (tb/fft_double_test.vhd:154-161, 207-214)."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "intfftk_b200", "intfft_host")
    assert os.path.exists(exe), "intfft_host not built"
    nfft, n, frames = 7, 128, 3
    x = oracle.fill_random(frames * n * 2, 15, 42).reshape(frames, n, 2).astype(np.int16)
    # --- single-lane format, three testbench modes
    np.savetxt(tmp_path / "di_single.dat", x.reshape(-1, 2), fmt="%d")
    for mode, (fmt, rnd) in (("UNSCALED", (1, 0)), ("ROUNDING", (0, 1)), ("TRUNCATE", (0, 0))):
        subprocess.check_call([exe, "--nfft", str(nfft), "--mode", mode, str(tmp_path / "di_single.dat"), str(tmp_path / "do.dat")])
        got = np.loadtxt(tmp_path / "do.dat", dtype=np.int64).reshape(frames, n, 2)
        want = oracle.batch(oracle.generics(nfft, 16, 16, fmt, rnd, 1, 1, 0), x)
        assert np.array_equal(got, want.astype(np.int64)), mode
    # --- two-lane format: FFT (halves in, even/odd out)
    beats = np.stack([x[:, : n // 2, 0], x[:, n // 2:, 0], x[:, : n // 2, 1], x[:, n // 2:, 1]], -1).reshape(-1, 4)
    np.savetxt(tmp_path / "di_double.dat", beats, fmt="%d")
    subprocess.check_call([exe, "--nfft", str(nfft), "--mode", "UNSCALED", "--lanes", str(tmp_path / "di_double.dat"), str(tmp_path / "do2.dat")])
    got = np.loadtxt(tmp_path / "do2.dat", dtype=np.int64).reshape(frames, n // 2, 4)
    want = oracle.batch(oracle.generics(nfft, 16, 16, 1, 0, 1, 1, 0), x).astype(np.int64)
    assert np.array_equal(got[..., 0], want[:, 0::2, 0]) and np.array_equal(got[..., 1], want[:, 1::2, 0])
    assert np.array_equal(got[..., 2], want[:, 0::2, 1]) and np.array_equal(got[..., 3], want[:, 1::2, 1])
    # --- int_fft_ifft_pair with the top-17-bit dump of fft_double_test
    subprocess.check_call([exe, "--nfft", str(nfft), "--mode", "UNSCALED", "--pair", "--lanes", "--top17",
                           str(tmp_path / "di_double.dat"), str(tmp_path / "dout_pair.dat")])
    got = np.loadtxt(tmp_path / "dout_pair.dat", dtype=np.int64).reshape(frames, n // 2, 4)
    mid = oracle.batch(oracle.generics(nfft, 16, 16, 1, 0, 1, 1, 0), x)
    fin = oracle.batch(oracle.generics(nfft, 16 + nfft, 16, 1, 0, 1, 1, 1), mid).astype(np.int64) >> (16 + 2 * nfft - 17)
    assert np.array_equal(got[..., 0], fin[:, : n // 2, 0]) and np.array_equal(got[..., 1], fin[:, n // 2:, 0])
    assert np.array_equal(got[..., 2], fin[:, : n // 2, 1]) and np.array_equal(got[..., 3], fin[:, n // 2:, 1])


@pytest.mark.parametrize("kw", [dict(NFFT=7, DATA_WIDTH=16, FORMAT=1), dict(NFFT=12, DATA_WIDTH=16, FORMAT=0),
                                dict(NFFT=13, DATA_WIDTH=18, FORMAT=0), dict(NFFT=16, DATA_WIDTH=24, FORMAT=1)])
def test_single_path_natural_order(ib, oracle, kw):
    """f1: int_fft_single_path = core + int_bitrev_order: natural order in and out, both directions; the
    scaled forward result must also sit within a few LSB of numpy's fft / N at natural bin positions."""
    g = ib.Generics(**kw)
    n, batch = 1 << g.NFFT, 3
    x = oracle.fill_random(batch * n * 2, g.DATA_WIDTH - 1, 3).reshape(batch, n, 2).astype(oracle.scalar_dtype(g.DATA_WIDTH))
    og = oracle.generics(g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, 1, 1, 0)
    fwd = ib.Core(g, batch, 0)
    got = fwd.exec_natural(torch.from_numpy(x).cuda()).cpu().numpy()
    want = oracle.bitrev(g.NFFT, oracle.batch(og, x))
    assert np.array_equal(got, want)
    if g.FORMAT == 0:
        ref = np.fft.fft(x[..., 0].astype(np.float64) + 1j * x[..., 1], axis=1) / n
        assert np.abs((got[..., 0] + 1j * got[..., 1]) - ref).max() < 16.0
    # inverse core on a natural-order spectrum
    gi = ib.Generics(**dict(kw, DATA_WIDTH=g.out_width)) if g.FORMAT == 0 or g.out_width + g.NFFT <= 64 else None
    if gi is not None and ib.validate(gi, 1) == 0:
        inv = ib.Core(gi, batch, 1)
        spec = want                                   # natural-order spectrum
        got_i = inv.exec_natural(torch.from_numpy(spec).cuda()).cpu().numpy()
        ogi = oracle.generics(gi.NFFT, gi.DATA_WIDTH, gi.TWDL_WIDTH, gi.FORMAT, gi.RNDMODE, 1, 1, 1)
        want_i = oracle.batch(ogi, oracle.bitrev(g.NFFT, spec))
        assert np.array_equal(got_i, want_i)
        inv.close()
    fwd.close()


def test_use_fly_bypass(ib, oracle):
    for fmt in (0, 1):
        got, want = _run_both(ib, oracle, 5, seed=3, NFFT=9, DATA_WIDTH=12, FORMAT=fmt, USE_FLY=0)
        assert np.array_equal(got, want)


def test_ragged_batches_and_inplace(ib, oracle):
    """Batches that do not fill the last tile, and d_in == d_out."""
    for nfft, batch in ((3, 1), (3, 513), (5, 129), (9, 7), (11, 3)):
        got, want = _run_both(ib, oracle, batch, seed=batch, via="device", NFFT=nfft, DATA_WIDTH=16, FORMAT=0)
        assert np.array_equal(got, want), (nfft, batch)
    g = ib.Generics(NFFT=10, DATA_WIDTH=16, FORMAT=0)
    x = oracle.fill_random(11 * 1024 * 2, 16, 77).reshape(11, 1024, 2)
    core = ib.Core(g, 11, 0)
    d = torch.from_numpy(x).cuda()
    core.exec(d, d)
    assert np.array_equal(d.cpu().numpy(), oracle.batch(oracle.generics(10), x))


def test_invalid_generics_fail_like_elaboration(ib):
    for kw in (dict(NFFT=2), dict(NFFT=21), dict(TWDL_WIDTH=28), dict(TWDL_WIDTH=26, XSER="OLD"), dict(DATA_WIDTH=7),
               dict(FORMAT=1, RNDMODE=1), dict(TWDL_WIDTH=20, DATA_WIDTH=53, FORMAT=0)):
        base = dict(NFFT=8, DATA_WIDTH=16, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0)
        base.update(kw)
        with pytest.raises(ib.IntfftError) as ei:
            ib.Core(ib.Generics(**base), 4, 0)
        assert ei.value.status == -1, kw
    with pytest.raises(ib.IntfftError) as ei:
        ib.Core(ib.Generics(NFFT=10, DATA_WIDTH=60, FORMAT=1, TWDL_WIDTH=8), 4, 0)
    assert ei.value.status == -4
    with pytest.raises(ib.IntfftError) as ei:      # trpl18: product slice beyond the 79-bit product (does not elaborate)
        ib.Core(ib.Generics(NFFT=10, DATA_WIDTH=60, FORMAT=1), 4, 0)
    assert ei.value.status == -1


def test_device_stimulus_and_checksum_match_oracle(ib, oracle):
    for width, dt in ((16, torch.int16), (18, torch.int32), (40, torch.int64)):
        d = torch.empty(100003, dtype=dt, device="cuda")
        ib.fill_random(d, width, 0x696E7466)
        h = oracle.fill_random(100003, width, 0x696E7466)
        assert np.array_equal(d.cpu().numpy(), h)
        assert ib.checksum(d) == oracle.checksum(h)


@pytest.mark.parametrize("nfft", [3, 4, 7, 10, 12, 13, 16])
def test_bitrev_order(ib, oracle, nfft):
    for dt in (np.int16, np.int32, np.int64):
        batch = 3
        x = np.arange(batch * (1 << nfft) * 2, dtype=np.int64).astype(dt).reshape(batch, 1 << nfft, 2)
        got = ib.bitrev_order(torch.from_numpy(x).cuda(), nfft).cpu().numpy()
        assert np.array_equal(got, oracle.bitrev(nfft, x))


def test_exec_rejects_misaligned_device_buffers(ib):
    g = ib.Generics(NFFT=8, DATA_WIDTH=16, FORMAT=0)
    core = ib.Core(g, 2, 0)
    pad = torch.zeros(2 * 256 * 2 + 2, dtype=torch.int16, device="cuda")
    with pytest.raises(ib.IntfftError):
        core.exec(pad[2:], core.new_output())
    core.close()


def test_bitrev_order_unaligned_buffers_take_the_scalar_path(ib, oracle):
    """The 16-byte reorder kernel needs 16-byte-aligned buffers; anything else must still be reordered."""
    nfft, batch = 12, 3
    n = 1 << nfft
    x = np.arange(batch * n * 2, dtype=np.int64).astype(np.int16).reshape(batch, n, 2)
    pad = torch.zeros(batch * n * 2 + 2, dtype=torch.int16, device="cuda")
    src = pad[2:]                                   # 4 bytes past a 256-byte-aligned allocation
    src.copy_(torch.from_numpy(x).reshape(-1))
    assert src.data_ptr() % 16 == 4
    got = ib.bitrev_order(src, nfft).cpu().numpy().reshape(batch, n, 2)
    assert np.array_equal(got, oracle.bitrev(nfft, x))


def test_full_c2_batch_checksum_matches_oracle(ib, oracle):
    """The WHOLE BASELINE c2 batch (65536 frames x 4096 points): device stimulus -> device FFT -> device checksum
    against the same stimulus pushed through the multithreaded C oracle on the host."""
    nfft, batch, seed = 12, 65536, 0x696E7466
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=16, FORMAT=0)
    core = ib.Core(g, batch, 0)
    x = core.new_input()
    ib.fill_random(x, 16, seed)
    y = core.exec(x)
    got = ib.checksum(y)
    hx = oracle.fill_random(batch * (1 << nfft) * 2, 16, seed).reshape(batch, 1 << nfft, 2)
    assert oracle.checksum(hx) == ib.checksum(x)
    want = oracle.batch(oracle.generics(nfft), hx)
    assert oracle.checksum(want) == got
    # and the natural-order wrapper on the same batch: a permutation of every frame
    z = core.exec_natural(x)
    for f in (0, 777, 65535):
        assert np.array_equal(z[f].cpu().numpy()[None], oracle.bitrev(nfft, want[f][None]))
    core.close()


def test_fft_ifft_pair_roundtrip_full_c2_batch(ib, oracle):
    """Size-independent property at the full c2 size (65536 x 4096): FFT then IFFT returns x / N
    within the truncation bias; plus a checksum-of-frames comparison with the oracle on a sample."""
    nfft, batch = 12, 65536
    n = 1 << nfft
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=16, FORMAT=0)
    fwd, inv = ib.Core(g, batch, 0), ib.Core(g, batch, 1)
    x = fwd.new_input()
    ib.fill_random(x, 15, 99)                      # 15-bit amplitude: sqrt(2) headroom, no wrap
    y = fwd.exec(x)
    z = inv.exec(y)
    err = z.to(torch.float32) - x.to(torch.float32) / n
    assert float(err.abs().max()) < 24.0
    assert float(err.pow(2).mean().sqrt()) < 3.0
    # sampled frames against the oracle, bit-exact
    og = oracle.generics(nfft)
    for f in (0, 1, 4095, 32768, 65535):
        want = oracle.batch(og, x[f].cpu().numpy()[None])
        assert np.array_equal(y[f].cpu().numpy()[None], want)
    fwd.close(); inv.close()


# ---- round 2: ring-staged host pipeline, pair through host buffers, multi-device API, plan re-entrancy ---------------
@pytest.mark.parametrize("kw,direction,batch", [
    (dict(NFFT=12, DATA_WIDTH=16, FORMAT=0), 0, 40000),      # 5 chunks of 32 MiB through the 3-slot ring, ragged tail
    (dict(NFFT=13, DATA_WIDTH=18, FORMAT=0), 1, 2500),       # 64 KiB frames, 512 per chunk
    (dict(NFFT=16, DATA_WIDTH=24, FORMAT=1), 0, 150),        # two-pass plan, 64-bit output container
    (dict(NFFT=20, DATA_WIDTH=16, FORMAT=0), 0, 19),         # 4 MiB frames: 8 per chunk
])
def test_exec_host_ring_matches_device_path(ib, oracle, kw, direction, batch):
    """intfft_exec_host streams the batch through a ring of three chunk buffers: every chunk boundary / slot reuse
    must give what one intfft_exec over the whole batch gives; sampled frames are checked against the oracle."""
    g = ib.Generics(**kw)
    n = 1 << g.NFFT
    core = ib.Core(g, batch, direction)
    x = core.new_input()
    ib.fill_random(x, g.DATA_WIDTH, 77)
    want = core.exec(x).cpu()
    hx = x.cpu().numpy()
    got = core.exec_host(hx)
    assert torch.equal(torch.from_numpy(got), want)
    xs = 1 if g.XSER == "NEW" else 0
    og = oracle.generics(g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, xs, 1, direction)
    for f in (0, batch // 2, batch - 1):
        assert np.array_equal(got[f][None], oracle.batch(og, hx[f][None]))
    core.close()


@pytest.mark.parametrize("kw", [dict(NFFT=10, DATA_WIDTH=16, FORMAT=0), dict(NFFT=7, DATA_WIDTH=16, FORMAT=1),
                                dict(NFFT=13, DATA_WIDTH=14, FORMAT=0, RNDMODE=1)])
def test_pair_exec_host(ib, oracle, kw):
    """intfft_pair_exec_host: host buffers in and out, the spectrum stays on the device."""
    g = ib.Generics(**kw)
    n, batch = 1 << g.NFFT, 37
    xs = 1 if g.XSER == "NEW" else 0
    x = oracle.fill_random(batch * n * 2, g.DATA_WIDTH - 1, 5).reshape(batch, n, 2).astype(oracle.scalar_dtype(g.DATA_WIDTH))
    mid = oracle.batch(oracle.generics(g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, xs, 1, 0), x)
    want = oracle.batch(oracle.generics(g.NFFT, g.out_width, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, xs, 1, 1), mid)
    pair = ib.Pair(g, batch)
    got = pair.exec_host(x)
    assert got.dtype == want.dtype and np.array_equal(got, want)
    got2 = pair.exec(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got2, want)
    pair.close()


def test_multi_device_api_one_host_batch(ib, oracle):
    """intfft_multi_*: one host batch, sharded inside the library over the devices it is given (the same device
    twice when the box has one GPU: the sharding, the per-device pipelines and their host threads are what is tested)."""
    ndev = torch.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    g = ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0)
    batch = 1001
    m = ib.Multi(g, batch, 0, devices)
    sh = m.shards()
    assert [s[0] for s in sh] == devices
    assert sh[0][1] == 0 and sum(s[2] for s in sh) == batch
    assert all(sh[i][1] + sh[i][2] == sh[i + 1][1] for i in range(len(sh) - 1))
    assert [(s[1], s[1] + s[2]) for s in sh] == [ib.shard_range(batch, i, len(devices)) for i in range(len(devices))]
    hin = ib.HostBuffer((batch, 4096, 2), np.int16)
    hout = ib.HostBuffer((batch, 4096, 2), np.int16)
    hin.array[...] = oracle.fill_random(batch * 4096 * 2, 16, 9).reshape(batch, 4096, 2)
    m.exec_host_ptr(hin.ptr, hout.ptr)
    want = oracle.batch(oracle.generics(12), hin.array)
    assert np.array_equal(hout.array, want)
    m.close(); hin.close(); hout.close()
    with pytest.raises(ib.IntfftError):
        ib.Multi(g, 2, 0, [0, 0, 0])             # fewer frames than devices


def test_plan_is_reentrant_on_the_device_path(ib, oracle):
    """exec entry points no longer write into the plan — exec / exec_natural on ONE plan from several host
    threads and streams at once must each give their own result."""
    import threading
    g = ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0)
    batch = 4096
    core = ib.Core(g, batch, 0)
    x = core.new_input()
    ib.fill_random(x, 16, 3)
    want_rev = core.exec(x).clone()
    want_nat = core.exec_natural(x).clone()
    torch.cuda.synchronize()
    assert not torch.equal(want_rev, want_nat)
    errs = []

    def worker(natural):
        try:
            s = torch.cuda.Stream()
            y = core.new_output()
            for _ in range(30):
                (core.exec_natural if natural else core.exec)(x, y, stream=s.cuda_stream)
                s.synchronize()
                if not torch.equal(y, want_nat if natural else want_rev):
                    errs.append(natural)
                    return
        except Exception as e:       # pragma: no cover
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i % 2 == 1,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs
    core.close()


def test_python_mirror_rejects_wrong_tensors(ib):
    """exec_natural / Pair.exec check device, contiguity, element size and count like exec does."""
    g = ib.Generics(NFFT=8, DATA_WIDTH=16, FORMAT=0)
    core = ib.Core(g, 4, 0)
    x = core.new_input()
    for bad in (torch.empty((4, 256, 2), dtype=torch.int32, device="cuda"), torch.empty((3, 256, 2), dtype=torch.int16, device="cuda"),
                torch.empty((4, 256, 2), dtype=torch.int16)):
        with pytest.raises(ib.IntfftError):
            core.exec_natural(x, bad)
        with pytest.raises(ib.IntfftError):
            core.exec(bad if bad.is_cuda else x, bad)
    with pytest.raises(ib.IntfftError):
        core.exec_natural(x, x)
    pair = ib.Pair(g, 4)
    with pytest.raises(ib.IntfftError):
        pair.exec(x, torch.empty((4, 256, 2), dtype=torch.int32, device="cuda"))
    pair.close(); core.close()


# ---- on-device Taylor twiddles (intfft_taylor.cuh): STAGE >= 11 recomputed where the strided kernels hoist them ----
@pytest.mark.parametrize("tw,xser", [(16, "NEW"), (16, "OLD"), (8, "NEW"), (12, "OLD"), (17, "NEW"), (18, "NEW"),
                                      (19, "NEW"), (24, "OLD"), (25, "OLD"), (27, "NEW")])
def test_device_taylor_function_equals_table(ib, tw, xser):
    """The kernels' device function over whole stages == the host generator == row_twiddle_tay
    (rom_twiddle_int.vhd:215-246, row_twiddle_tay.vhd:123-382), STAGE 11..19, both XSER, narrow and wide twiddles."""
    g = ib.Generics(TWDL_WIDTH=tw, XSER=xser)
    for stage in (11, 12, 13, 15, 16, 18, 19):
        dre, dim = ib.twiddles_device(g, stage)
        hre, him = ib.twiddles(g, stage)
        assert np.array_equal(dre, hre) and np.array_equal(dim, him), (tw, xser, stage)


@pytest.mark.parametrize("kw,batch", [
    (dict(NFFT=17, DATA_WIDTH=16, FORMAT=0), 3),                    # packed-16 strided-8 + 9 bits (default threshold)
    (dict(NFFT=19, DATA_WIDTH=12, FORMAT=0, RNDMODE=1), 2),
    (dict(NFFT=18, DATA_WIDTH=18, FORMAT=0), 2),                    # 32-bit lanes, pre-shifted twiddles in the DIT pass
    (dict(NFFT=17, DATA_WIDTH=14, FORMAT=1), 2),                    # UNSCALED growth to 31 bits, mixed arrangements
    (dict(NFFT=17, DATA_WIDTH=18, TWDL_WIDTH=24, FORMAT=0, XSER="OLD"), 2),
])
def test_device_taylor_plans_match_oracle(ib, oracle, kw, batch):
    """NFFT >= 17 plans on the strided kernels run WITHOUT STAGE >= 12 tables: bit-exact vs the oracle, both directions."""
    for direction in (0, 1):
        if ib.validate(ib.Generics(**kw), direction) != 0:
            continue
        got, want = _run_both(ib, oracle, batch, seed=170 + direction, via="device", direction=direction, **kw)
        assert np.array_equal(got, want), (kw, direction)


@pytest.mark.parametrize("kw", [dict(NFFT=14, DATA_WIDTH=16, FORMAT=0), dict(NFFT=16, DATA_WIDTH=16, FORMAT=0),
                                dict(NFFT=15, DATA_WIDTH=18, FORMAT=0), dict(NFFT=16, DATA_WIDTH=16, FORMAT=1)])
def test_device_taylor_equals_table_path(ib, kw, monkeypatch):
    """Same plan with the threshold lowered (device Taylor from NFFT 13 on) and raised (tables only): identical output."""
    g = ib.Generics(**kw)
    n = 1 << g.NFFT
    for direction in (0, 1):
        x = ib.fill_random(torch.empty(4 * n * 2, dtype=torch.int16 if g.DATA_WIDTH <= 16 else torch.int32, device="cuda"),
                           g.DATA_WIDTH, 99 + direction).reshape(4, n, 2)
        monkeypatch.setenv("INTFFT_TAYLOR_MIN_NFFT", "13")
        a = ib.Core(g, 4, direction)
        monkeypatch.setenv("INTFFT_TAYLOR_MIN_NFFT", "99")
        b = ib.Core(g, 4, direction)
        monkeypatch.delenv("INTFFT_TAYLOR_MIN_NFFT")
        assert torch.equal(a.exec(x), b.exec(x)), (kw, direction)
        a.close(); b.close()


# ---- strided passes: 2-D TMA (default) and the cp.async / STG form (INTFFT_STRIDED_TMA=0) must agree ----
@pytest.mark.parametrize("kw", [
    dict(NFFT=20, DATA_WIDTH=16, FORMAT=0),                    # c4 geometry: packed-16, 256 rows x 16 columns
    dict(NFFT=17, DATA_WIDTH=12, FORMAT=0, RNDMODE=1),         # DATA_WIDTH < 16 unpack, ROUNDING
    dict(NFFT=15, DATA_WIDTH=16, FORMAT=0),                    # packed-16, 16 rows x 256 columns
    dict(NFFT=16, DATA_WIDTH=24, FORMAT=1),                    # c3: 32-bit lanes, 8-byte samples, mixed arrangements
    dict(NFFT=18, DATA_WIDTH=18, FORMAT=0),                    # 32-bit lanes, pre-shifted twiddles in the DIT pass
    dict(NFFT=14, DATA_WIDTH=18, FORMAT=0),                    # 32-bit lanes, 16 rows x 256 columns (UINT64 tensor map)
    dict(NFFT=16, DATA_WIDTH=16, FORMAT=1),                    # packed input container, 32-bit output container
])
def test_strided_tma_equals_cp_async_form(ib, oracle, kw, monkeypatch):
    g = ib.Generics(**kw)
    n = 1 << g.NFFT
    og_args = (g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, 1 if g.XSER == "NEW" else 0, g.USE_FLY)
    for direction in (0, 1):
        if ib.validate(g, direction) != 0:
            continue
        x = oracle.fill_random(3 * n * 2, g.DATA_WIDTH, 31 + direction).reshape(3, n, 2)
        d_in = torch.from_numpy(x).cuda()
        core = ib.Core(g, 3, direction)
        monkeypatch.setenv("INTFFT_STRIDED_TMA", "0")
        a = core.exec(d_in).clone()
        monkeypatch.setenv("INTFFT_STRIDED_TMA", "1")
        b = core.exec(d_in).clone()
        monkeypatch.delenv("INTFFT_STRIDED_TMA")
        core.close()
        assert torch.equal(a, b), (kw, direction)
        want = oracle.batch(oracle.generics(*og_args, direction), x, 0)
        assert np.array_equal(b.cpu().numpy(), want), (kw, direction)


# ---- BASELINE c3 / c4 / c5 at their FULL sizes: sampled frames bit-exact against the oracle, and the exact
# ---- size-independent property of this path — frames are independent, so the whole-batch result must equal the
# ---- results of the two half-batches run through separate plans (different work splits, same integers)
@pytest.mark.parametrize("name,kw,direction,batch", [
    ("c3", dict(NFFT=16, DATA_WIDTH=24, FORMAT=1), 0, 4096),
    ("c4", dict(NFFT=20, DATA_WIDTH=16, FORMAT=0), 0, 256),
    ("c5", dict(NFFT=13, DATA_WIDTH=18, FORMAT=0), 1, 131072),
    ("c5u", dict(NFFT=13, DATA_WIDTH=18, FORMAT=1), 1, 131072),
])
def test_full_size_baseline_configs(ib, oracle, name, kw, direction, batch):
    g = ib.Generics(**kw)
    n = 1 << g.NFFT
    core = ib.Core(g, batch, direction)
    x = core.new_input()
    ib.fill_random(x, g.DATA_WIDTH, 0x696E7466 + len(name))
    y = core.exec(x)
    # (1) sampled frames: first, last, and a few inside
    idx = sorted({0, 1, batch // 3, batch // 2, batch - 2, batch - 1})
    sel = torch.tensor(idx, device=x.device)
    hx = x.index_select(0, sel).cpu().numpy()
    og = oracle.generics(g.NFFT, g.DATA_WIDTH, g.TWDL_WIDTH, g.FORMAT, g.RNDMODE, 1, 1, direction)
    assert np.array_equal(y.index_select(0, sel).cpu().numpy(), oracle.batch(og, hx, 0)), name
    # (2) frame independence at full size: two half-batch plans reproduce the whole-batch output bit for bit
    half = batch // 2
    lo, hi = ib.Core(g, half, direction), ib.Core(g, batch - half, direction)
    y2 = torch.empty_like(y)
    lo.exec(x[:half], y2[:half])
    hi.exec(x[half:], y2[half:])
    assert ib.checksum(y2) == ib.checksum(y), name
    assert torch.equal(y2[half - 1:half + 1], y[half - 1:half + 1]), name
    for c in (core, lo, hi):
        c.close()


# ---- one-pass 8192- / 16384-point packed-16 kernels against the oracle AND against the two-pass schedule they replace ----
@pytest.mark.parametrize("nfft,env", [(13, "INTFFT_N13_TWO_PASS"), (14, "INTFFT_N14_TWO_PASS")])
@pytest.mark.parametrize("dw,rnd", [(16, 0), (12, 0), (16, 1), (9, 1)])
def test_one_pass_packed16_kernels(ib, oracle, nfft, env, dw, rnd, monkeypatch):
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, FORMAT=0, RNDMODE=rnd)
    n, batch = 1 << nfft, 7                                    # several frames per CTA walk, odd count
    for direction in (0, 1):
        x = oracle.fill_random(batch * n * 2, dw, 40 + nfft + direction).reshape(batch, n, 2)
        d_in = torch.from_numpy(x).cuda()
        one = ib.Core(g, batch, direction)
        assert f"fast16_n{nfft}" in ib.describe(g, batch, direction)
        monkeypatch.setenv(env, "1")
        two = ib.Core(g, batch, direction)
        monkeypatch.delenv(env)
        a, b = one.exec(d_in), two.exec(d_in)
        one.close(); two.close()
        assert torch.equal(a, b), (nfft, dw, rnd, direction)
        want = oracle.batch(oracle.generics(nfft, dw, 16, 0, rnd, 1, 1, direction), x, 0)
        assert np.array_equal(a.cpu().numpy(), want), (nfft, dw, rnd, direction)


# ---- seeded fuzz over the whole generic space: whatever elaborates must match the oracle bit for bit ----
def _fuzz_cases(count, seed):
    import random
    rng = random.Random(seed)
    cases = []
    while len(cases) < count:
        nfft = rng.choice([3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 12, 13, 13, 14, 15, 16, 17])
        xser = rng.choice(["NEW", "OLD"])
        tw = rng.randint(8, 27 if xser == "NEW" else 25)
        fmt, rnd = rng.choice([(0, 0), (0, 1), (1, 0)])
        dw = rng.randint(8, 40 if nfft <= 12 else 32)
        if dw + fmt * nfft > 64:
            continue
        cases.append((nfft, dw, tw, xser, fmt, rnd, rng.randint(0, 1), rng.randint(1, 5 if nfft <= 12 else 2), rng.randint(0, 1 << 30)))
    return cases


# (INTFFT_FUZZ_CASES / INTFFT_FUZZ_SEED scale the hunt; the default set found the TWDL_WIDTH < 16 bug of the packed STAGE-12
# twiddles in the one-pass 8192-point kernel)
@pytest.mark.parametrize("case", _fuzz_cases(int(os.environ.get("INTFFT_FUZZ_CASES", "400")), int(os.environ.get("INTFFT_FUZZ_SEED", str(0x5EED)), 0)),
                         ids=lambda c: "n%d-dw%d-tw%d-%s-f%d-r%d-d%d" % c[:7])
def test_fuzz_generics_against_oracle(ib, oracle, case):
    nfft, dw, tw, xser, fmt, rnd, direction, batch, seed = case
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, XSER=xser, FORMAT=fmt, RNDMODE=rnd)
    st = ib.validate(g, direction)
    if st != 0:
        # does not elaborate (or needs lanes beyond 64 bits): the library must refuse to build the plan as well
        with pytest.raises(ib.IntfftError):
            ib.Core(g, batch, direction)
        return
    got, want = _run_both(ib, oracle, batch, seed=seed, via="device", NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, XSER=xser,
                          FORMAT=fmt, RNDMODE=rnd, direction=direction)
    assert np.array_equal(got, want), case


@pytest.mark.parametrize("dw,tw,direction", [(18, 12, 0), (18, 12, 1), (20, 9, 0), (24, 15, 1), (16, 14, 1)])
def test_n13_32bit_lanes_narrow_twiddles(ib, oracle, dw, tw, direction):
    """Regression (found by the fuzz test): NFFT 13 on the 32-bit one-pass kernel with TWDL_WIDTH < 16 must not take the
    packed STAGE-12 twiddle path, which rebuilds W << 16 and is right for 16-bit twiddles only."""
    kw = dict(NFFT=13, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=0, direction=direction)
    if dw == 16:
        kw["RNDMODE"] = 0
        kw["FORMAT"] = 1                      # 16-bit data on 32-bit lanes: UNSCALED
    got, want = _run_both(ib, oracle, 3, seed=77, via="device", **kw)
    assert np.array_equal(got, want)


def _fuzz_batches(count, seed):
    import random
    rng = random.Random(seed)
    cases = []
    for _ in range(count):
        nfft = rng.randint(3, 14)
        dw = rng.choice([9, 12, 16, 16, 16, 18, 18, 24, 30])
        fmt, rnd = rng.choice([(0, 0), (0, 0), (0, 1), (1, 0)])
        if dw + fmt * nfft > 64:
            fmt = 0
        hi = max(1, (1 << 21) >> nfft)                       # up to 2^21 samples: several waves of tiles, ragged tails
        batch = rng.choice([1, 2, 3, rng.randint(1, hi), rng.randint(1, hi), hi - 1 if hi > 1 else 1])
        cases.append((nfft, dw, fmt, rnd, rng.randint(0, 1), batch, rng.randint(0, 1 << 30)))
    return cases


@pytest.mark.parametrize("case", _fuzz_batches(int(os.environ.get("INTFFT_FUZZ_BATCHES", "120")), int(os.environ.get("INTFFT_FUZZ_SEED", "7"), 0)),
                         ids=lambda c: "n%d-dw%d-f%d-r%d-d%d-b%d" % c[:6])
def test_fuzz_batch_sizes_against_oracle(ib, oracle, case):
    """Ragged batches of every size: partial tiles, partial chunks, grids smaller and larger than the persistent launch
    (the whole-tile fast paths of the store / prefetch loops must leave the tails to the guarded ones)."""
    nfft, dw, fmt, rnd, direction, batch, seed = case
    if ib.validate(ib.Generics(NFFT=nfft, DATA_WIDTH=dw, FORMAT=fmt, RNDMODE=rnd), direction) != 0:
        pytest.skip("does not elaborate")
    got, want = _run_both(ib, oracle, batch, seed=seed, via="device", NFFT=nfft, DATA_WIDTH=dw, FORMAT=fmt, RNDMODE=rnd,
                          direction=direction)
    assert np.array_equal(got, want), case


@pytest.mark.parametrize("kw,direction,batch", [
    (dict(NFFT=12, DATA_WIDTH=16, FORMAT=0), 0, 64),           # one launch, bulk-TMA input
    (dict(NFFT=17, DATA_WIDTH=16, FORMAT=0), 0, 2),            # two launches, tensor maps passed by value, on-device Taylor
    (dict(NFFT=16, DATA_WIDTH=24, FORMAT=1), 0, 2),            # c3 chain, plan-owned intermediate
    (dict(NFFT=13, DATA_WIDTH=18, FORMAT=0), 1, 8),            # c5 kernel
])
def test_exec_is_cuda_graph_capturable(ib, kw, direction, batch):
    """intfft_exec allocates nothing, synchronises nothing and keeps no per-call host state, so a launch-bound caller can
    capture it into a CUDA graph and replay it (include/intfft.h: threading / stream contract)."""
    g = ib.Generics(**kw)
    core = ib.Core(g, batch, direction)
    x, y = core.new_input(), core.new_output()
    ib.fill_random(x, g.DATA_WIDTH, 5)
    core.exec(x, y)
    torch.cuda.synchronize()
    ref = y.clone()
    y.zero_()
    s = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            core.exec(x, y, stream=s.cuda_stream)
    for _ in range(2):
        y.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, ref)
    core.close()
