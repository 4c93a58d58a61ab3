"""Size-independent properties of the oracle (what SURVEY.md §4/§8c lists as the pins the build has
to create itself): distance to a float FFT, FFT->IFFT round trip, impulse -> twiddle read-back,
bit-reversed stream order, and the structural identity between math/fn_radix2.m's lane model and
in-place indexing."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import fn_radix2 as fr


def _bitrev_idx(n):
    idx = np.arange(1 << n)
    rev = np.zeros_like(idx)
    for b in range(n):
        rev |= ((idx >> b) & 1) << (n - 1 - b)
    return rev


def _rand(rng, batch, n, width, headroom=1.0):
    hi = int(((1 << (width - 1)) - 1) / headroom)
    return rng.integers(-hi, hi + 1, size=(batch, 1 << n, 2)).astype(co.scalar_dtype(width))


@pytest.mark.parametrize("nfft", [7, 10, 12, 13])
def test_scaled_fft_close_to_float_fft(nfft):
    """Scaled 16-bit DIF vs numpy fft / N: ~2 LSB rms. Tolerance: rms < 3, max < 16 LSB."""
    rng = np.random.default_rng(nfft)
    x = _rand(rng, 4, nfft, 16, headroom=1.5)
    y = co.batch(co.generics(nfft), x).astype(np.float64)
    xc = x[..., 0].astype(np.float64) + 1j * x[..., 1]
    ref = np.fft.fft(xc, axis=1)[:, _bitrev_idx(nfft)] / (1 << nfft)
    err = (y[..., 0] + 1j * y[..., 1]) - ref
    assert np.sqrt(np.mean(np.abs(err) ** 2)) < 3.0
    assert np.abs(err).max() < 16.0


@pytest.mark.parametrize("nfft,dw", [(8, 16), (12, 24), (16, 24)])
def test_unscaled_fft_close_to_float_fft(nfft, dw):
    """Unscaled DIF vs numpy fft: relative error ~1.6e-4 (twiddle quantisation, amplitude 2^15-1)."""
    rng = np.random.default_rng(nfft)
    # headroom: a full-scale complex sample rotated by 45 degrees exceeds the 1-bit-per-stage growth
    # and wraps in the reference too (int_cmult_dsp48.vhd:189-190 slices, no saturation)
    x = _rand(rng, 2, nfft, dw, headroom=1.5)
    y = co.batch(co.generics(nfft, data_width=dw, format=1), x).astype(np.float64)
    xc = x[..., 0].astype(np.float64) + 1j * x[..., 1]
    ref = np.fft.fft(xc, axis=1)[:, _bitrev_idx(nfft)]
    got = y[..., 0] + 1j * y[..., 1]
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel < 5e-4


@pytest.mark.parametrize("nfft", [6, 10, 13])
def test_fft_ifft_roundtrip(nfft):
    """int_fft_ifft_pair style loop-back: IFFT(FFT(x)) ~= x / N (scaled), within a few LSB."""
    rng = np.random.default_rng(100 + nfft)
    x = _rand(rng, 3, nfft, 16, headroom=1.5)     # sqrt(2) headroom: no wrap in the first cmult
    y = co.batch(co.generics(nfft), x)
    z = co.batch(co.generics(nfft, direction=1), y).astype(np.float64)
    err = z - x.astype(np.float64) / (1 << nfft)
    # truncation is biased (floor): the bias piles up on the first few output samples (~ -10 LSB)
    assert np.sqrt(np.mean(err ** 2)) < 3.0
    assert np.abs(err).max() < 20.0


def test_unscaled_roundtrip_gain():
    """Unscaled pair: IFFT(FFT(x)) ~= x * N * (mg/2^15)^2-ish; check relative error only."""
    nfft = 9
    rng = np.random.default_rng(9)
    x = _rand(rng, 2, nfft, 12, headroom=1.5)
    y = co.batch(co.generics(nfft, data_width=12, format=1), x)
    z = co.batch(co.generics(nfft, data_width=12 + nfft, format=1, direction=1), y).astype(np.float64)
    ref = x.astype(np.float64) * (1 << nfft)
    assert np.linalg.norm(z - ref) / np.linalg.norm(ref) < 1e-3


@pytest.mark.parametrize("stage", [3, 5, 10, 11, 12])
def test_impulse_reads_back_twiddles(stage):
    """Unscaled DIF, NFFT = stage+1, x[k] = 2^15: the first butterfly stage (STAGE = stage) writes
    A - B = 2^15 times W_k >> 15 = W_k exactly into d[N/2 + k]; the remaining stages transform the
    upper half on its own, so out[N/2:] must equal the (NFFT-1)-stage transform of an impulse W_k."""
    nfft = stage + 1
    n = 1 << nfft
    g = co.generics(nfft, data_width=18, format=1)
    re, im = co.twiddle_table(g, stage)
    g2 = co.generics(nfft - 1, data_width=19, format=1)
    for k in (0, 1, (1 << stage) // 3, (1 << stage) - 1):
        x = np.zeros((1, n, 2), np.int32)
        x[0, k, 0] = 1 << 15
        y = co.batch(g, x)
        x2 = np.zeros((1, n // 2, 2), np.int32)
        x2[0, k] = (re[k], im[k])
        y2 = co.batch(g2, x2)
        assert np.array_equal(y[0, n // 2:].astype(np.int64), y2[0].astype(np.int64))


def test_stream_order_is_bitreversed_and_natural():
    """A pure tone lands on out[bitrev(bin)] for the FFT; the IFFT of a single bit-reversed bin is a tone."""
    nfft, n = 8, 256
    t = np.arange(n)
    tone = np.round(8000 * np.exp(2j * np.pi * 37 * t / n))
    x = np.stack([tone.real, tone.imag], -1).astype(np.int16)[None]
    y = co.batch(co.generics(nfft), x).astype(np.float64)
    mag = np.hypot(y[0, :, 0], y[0, :, 1])
    assert int(np.argmax(mag)) == int(_bitrev_idx(nfft)[37])
    spec = np.zeros((1, n, 2), np.int16)
    spec[0, _bitrev_idx(nfft)[5], 0] = 16000
    z = co.batch(co.generics(nfft, direction=1), spec).astype(np.float64)
    zc = z[0, :, 0] + 1j * z[0, :, 1]
    ref = 16000.0 / n * np.exp(2j * np.pi * 5 * t / n)
    assert np.abs(zc - ref).max() < 4.0


@pytest.mark.parametrize("n", [8, 64, 1024])
def test_fn_radix2_restatement_matches_numpy_fft(n):
    """math/fn_radix2.m (float model) == fft / N*ifft; and its pre-bitrevorder stream is the in-place
    DIF order — the structural identity the in-place C oracle relies on."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.allclose(fr.fn_radix2(x, n, "FWD"), np.fft.fft(x))
    assert np.allclose(fr.fn_radix2(x, n, "INV"), np.fft.ifft(x) * n)
    nl = int(np.log2(n))
    # textbook in-place DIF, natural in -> bit-reversed out
    d = x.copy()
    for ii in range(nl):
        half = n >> (ii + 1)
        for p in range(n // 2):
            j = p % half
            ia = (p // half) * 2 * half + j
            ib = ia + half
            w = np.exp(-2j * np.pi * j / (2 * half))
            d[ia], d[ib] = d[ia] + d[ib], (d[ia] - d[ib]) * w
    assert np.allclose(fr.fft_dif(x, n, bitrev_out=False), d)
    # in-place DIT, bit-reversed in -> natural out
    e = d.copy()
    for ii in range(nl):
        half = 1 << ii
        for p in range(n // 2):
            j = p % half
            ia = (p // half) * 2 * half + j
            ib = ia + half
            w = np.exp(+2j * np.pi * j / (2 * half))
            e[ia], e[ib] = e[ia] + e[ib] * w, e[ia] - e[ib] * w
    assert np.allclose(fr.fft_dit(d, n, bitrev_in=False), e)
    assert np.allclose(e, x * n)


def test_fill_random_and_checksum_are_deterministic():
    a = co.fill_random(4096, 16, 0x696E7466)
    b = co.fill_random(4096, 16, 0x696E7466)
    assert np.array_equal(a, b) and a.dtype == np.int16
    assert a.min() < -30000 and a.max() > 30000
    c = co.fill_random(4096, 18, 1)
    assert c.dtype == np.int32 and c.min() >= -(1 << 17) and c.max() < (1 << 17)
    assert co.checksum(a) == co.checksum(b) != co.checksum(a[::-1].copy())


def test_batch_threads_agree():
    rng = np.random.default_rng(0)
    x = _rand(rng, 37, 8, 16)
    g = co.generics(8)
    assert np.array_equal(co.batch(g, x, threads=1), co.batch(g, x, threads=5))
