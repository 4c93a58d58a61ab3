"""ctypes binding of oracle/libintfft_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libintfft_oracle.so")


class OrcGenerics(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "nfft_log2", "data_width", "twdl_width", "format", "rndmode", "xser", "use_fly", "direction")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "intfft_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libintfft_oracle.so"])
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        P = ctypes.POINTER
        L.orc_validate.argtypes = [P(OrcGenerics)]
        L.orc_twiddle.argtypes = [P(OrcGenerics), ctypes.c_int, ctypes.c_int64,
                                  P(ctypes.c_int64), P(ctypes.c_int64)]
        L.orc_twiddle_table.argtypes = [P(OrcGenerics), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_cmult.argtypes = [ctypes.c_int] * 3 + [ctypes.c_int64] * 4 + [P(ctypes.c_int64)] * 2
        L.orc_fly.argtypes = [P(OrcGenerics), ctypes.c_int, ctypes.c_int, ctypes.c_int64, P(ctypes.c_int64)]
        L.orc_transform.argtypes = [P(OrcGenerics)] + [ctypes.c_void_p] * 4
        L.orc_batch.argtypes = [P(OrcGenerics), ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_fill_random.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        L.orc_fill_random.restype = None
        L.orc_checksum.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        L.orc_checksum.restype = ctypes.c_uint64
        L.orc_bitrev.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_bitrev.restype = None
        _lib = L
    return _lib


def generics(nfft_log2, data_width=16, twdl_width=16, format=0, rndmode=0, xser=1, use_fly=1,
             direction=0) -> OrcGenerics:
    return OrcGenerics(nfft_log2, data_width, twdl_width, format, rndmode, xser, use_fly, direction)


def scalar_dtype(width: int):
    return np.int16 if width <= 16 else (np.int32 if width <= 32 else np.int64)


def validate(g: OrcGenerics) -> int:
    return lib().orc_validate(ctypes.byref(g))


def twiddle_table(g: OrcGenerics, stage: int):
    n = 1 << stage
    re = np.empty(n, np.int32)
    im = np.empty(n, np.int32)
    st = lib().orc_twiddle_table(ctypes.byref(g), stage, re.ctypes.data, im.ctypes.data)
    if st:
        raise ValueError(f"orc_twiddle_table status {st}")
    return re, im


def cmult(dtw: int, twd: int, xser_new: int, d_re: int, d_im: int, w_re: int, w_im: int):
    """int_cmult_dsp48(DTW, TWD, XSER) on one operand pair (two's-complement values in, values out)."""
    o_re, o_im = ctypes.c_int64(), ctypes.c_int64()
    st = lib().orc_cmult(dtw, twd, xser_new, d_re, d_im, w_re, w_im, ctypes.byref(o_re), ctypes.byref(o_im))
    if st:
        raise ValueError(f"orc_cmult status {st}")
    return o_re.value, o_im.value


def fly(g: OrcGenerics, stage: int, dtw: int, k: int, a_re: int, a_im: int, b_re: int, b_im: int):
    """One butterfly of STAGE `stage` at input width dtw, beat k (twiddle generated inside)."""
    ab = (ctypes.c_int64 * 4)(a_re, a_im, b_re, b_im)
    lib().orc_fly(ctypes.byref(g), stage, dtw, k, ab)
    return tuple(int(v) for v in ab)


def transform(g: OrcGenerics, re, im):
    """One frame on int64 arrays (most literal path: twiddles recomputed per butterfly)."""
    re = np.ascontiguousarray(re, np.int64)
    im = np.ascontiguousarray(im, np.int64)
    ore = np.empty_like(re)
    oim = np.empty_like(im)
    st = lib().orc_transform(ctypes.byref(g), re.ctypes.data, im.ctypes.data, ore.ctypes.data, oim.ctypes.data)
    if st:
        raise ValueError(f"orc_transform status {st}")
    return ore, oim


def batch(g: OrcGenerics, x: np.ndarray, threads: int = 0) -> np.ndarray:
    """x: [batch, N, 2] in the input container dtype -> [batch, N, 2] in the output container."""
    n = 1 << g.nfft_log2
    x = np.ascontiguousarray(x, scalar_dtype(g.data_width)).reshape(-1, n, 2)
    out = np.empty(x.shape, scalar_dtype(g.data_width + g.format * g.nfft_log2))
    st = lib().orc_batch(ctypes.byref(g), x.shape[0], x.ctypes.data, out.ctypes.data, threads)
    if st < 0:
        raise ValueError(f"orc_batch status {st}")
    return out


def fill_random(n_scalars: int, width: int, seed: int, dtype=None) -> np.ndarray:
    dt = np.dtype(dtype or scalar_dtype(width))
    buf = np.empty(n_scalars, dt)
    lib().orc_fill_random(buf.ctypes.data, n_scalars, dt.itemsize, width, seed & (2**64 - 1))
    return buf


def checksum(buf: np.ndarray) -> int:
    buf = np.ascontiguousarray(buf)
    return int(lib().orc_checksum(buf.ctypes.data, buf.size, buf.dtype.itemsize))


def bitrev(nfft_log2: int, x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x).reshape(-1, 1 << nfft_log2, 2)
    out = np.empty_like(x)
    lib().orc_bitrev(nfft_log2, x.dtype.itemsize, x.shape[0], x.ctypes.data, out.ctypes.data)
    return out
