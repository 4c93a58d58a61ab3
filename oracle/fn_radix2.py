"""NumPy restatement of the reference's floating-point structural model math/fn_radix2.m.

TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Octave is not available in the build image, so this is
the stand-in for timing BASELINE.json configs[0] ("1024-pt ... via math/fn_radix2.m in Octave") and
for the structure test (lane commutation + twiddle indexing == in-place indexing).  It keeps the
reference's control flow — two lanes, per-stage twiddle expansion, butterfly, cross-commutation,
interleave, bit-reverse — in double precision; it does NO quantisation, exactly like the original
(fn_radix2.m:93-107,136-148).
"""
from __future__ import annotations

import numpy as np


def bitrevorder(x: np.ndarray) -> np.ndarray:
    n = int(np.log2(len(x)))
    idx = np.arange(len(x))
    rev = np.zeros_like(idx)
    for b in range(n):
        rev |= ((idx >> b) & 1) << (n - 1 - b)
    return x[rev]


def _commute(ia, ib, blocks):
    # fn_rev2rdx / fn_rdx2rev (fn_radix2.m:51-89): same body, different block count
    half = len(ia)
    size = half // blocks
    oa = np.empty_like(ia)
    ob = np.empty_like(ib)
    for i in range(size):                 # the reference loops element-wise (:54-68); kept that way
        for j in range(blocks):
            stp = 2 * (j // 2) * size
            src = ia if j % 2 == 0 else ib
            oa[i + size * j] = src[i + stp]
            ob[i + size * j] = src[i + stp + size]
    return oa, ob


def _twiddle(n_pts, sign):
    k = np.arange(n_pts // 2)             # fn_twiddle_dif / fn_twiddle_dit (:93-107)
    return np.cos(k * 2 * np.pi / n_pts) + sign * 1j * np.sin(k * 2 * np.pi / n_pts)


def _twiddle_stage(ww, cnt, n_pts):
    stp = (n_pts // 2) // cnt             # fn_twiddleN_dif / _dit (:109-128)
    wo = np.empty(n_pts // 2, complex)
    for n in range(stp):
        for k in range(cnt):
            wo[n + stp * k] = ww[n * cnt]
    return wo


def fft_dif(din: np.ndarray, n_pts: int, bitrev_out: bool = True) -> np.ndarray:
    """fn_fft_dif (:152-190). bitrev_out=False returns the core's own (bit-reversed) stream."""
    nl = int(np.log2(n_pts))
    ta, tb = din[:n_pts // 2].astype(complex), din[n_pts // 2:].astype(complex)
    ww = _twiddle(n_pts, -1)
    for i in range(1, nl + 1):
        wx = _twiddle_stage(ww, 2 ** (i - 1), n_pts)
        oa, ob = ta + tb, (ta - tb) * wx  # fn_fly_dif (:136-139)
        if i < nl:
            ta, tb = _commute(oa, ob, 2 ** i)
    oo = np.empty(n_pts, complex)
    oo[0::2], oo[1::2] = oa, ob
    return bitrevorder(oo) if bitrev_out else oo


def fft_dit(din: np.ndarray, n_pts: int, bitrev_in: bool = True) -> np.ndarray:
    """fn_fft_dit (:193-232). bitrev_in=False takes the core's own (bit-reversed) input stream."""
    nl = int(np.log2(n_pts))
    dx = bitrevorder(np.asarray(din)) if bitrev_in else np.asarray(din)
    ta, tb = dx[0::2].astype(complex), dx[1::2].astype(complex)
    ww = _twiddle(n_pts, +1)
    for i in range(1, nl + 1):
        wx = _twiddle_stage(ww, 2 ** (nl - i), n_pts)
        oa, ob = ta + tb * wx, ta - tb * wx  # fn_fly_dit (:145-148)
        if i < nl:
            ta, tb = _commute(oa, ob, 2 ** (nl - i))
    return np.concatenate([oa, ob])


def fn_radix2(din, n_pts, mode):
    if mode == "FWD":
        return fft_dif(np.asarray(din), n_pts)
    if mode == "INV":
        return fft_dit(np.asarray(din), n_pts)
    raise ValueError("MODE must be FWD or INV")
