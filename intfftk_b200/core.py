"""ctypes binding of libintfft_b200.so + the host-side mirror of the reference's entity interface.

Names follow the reference: generics NFFT / DATA_WIDTH / TWDL_WIDTH / FORMAT / RNDMODE / XSER /
USE_FLY (int_fftNk.vhd:73-84), the testbench's legacy MODE strings (tb/fft_signle_test.vhd:81-88,
109-115), frames, lanes.  All compute goes through the C-ABI; nothing here computes a transform.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, asdict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libintfft_b200.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4


class IntfftError(RuntimeError):
    def __init__(self, status: int, what: str):
        self.status = status
        super().__init__(f"{what}: {_strerror(status)} (status {status})")


class _CGenerics(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in (
        "nfft_log2", "data_width", "twdl_width", "format", "rndmode", "xser", "use_fly", "direction")]


class _CLayout(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int64), ("batch", ctypes.c_int64),
                ("in_width", ctypes.c_int32), ("out_width", ctypes.c_int32),
                ("in_scalar_bytes", ctypes.c_int32), ("out_scalar_bytes", ctypes.c_int32),
                ("in_bytes", ctypes.c_int64), ("out_bytes", ctypes.c_int64),
                ("n_passes", ctypes.c_int32), ("lane_bits", ctypes.c_int32)]


_lib = None


def lib() -> ctypes.CDLL:
    """Load libintfft_b200.so.  Fails loudly — there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(or `make -C intfftk_b200`). intfftk_b200 has no CPU fallback.")
        L = ctypes.CDLL(_SO)
        P, vp = ctypes.POINTER, ctypes.c_void_p
        L.intfft_validate.argtypes = [P(_CGenerics)]
        L.intfft_plan_create.argtypes = [P(vp), P(_CGenerics), ctypes.c_int64, ctypes.c_int]
        L.intfft_plan_destroy.argtypes = [vp]
        L.intfft_query.argtypes = [vp, P(_CLayout)]
        L.intfft_exec.argtypes = [vp, vp, vp, vp]
        L.intfft_exec_host.argtypes = [vp, vp, vp]
        L.intfft_exec_natural.argtypes = [vp, vp, vp, vp]
        L.intfft_twiddles.argtypes = [P(_CGenerics), ctypes.c_int, vp, vp]
        L.intfft_twiddles_device.argtypes = [P(_CGenerics), ctypes.c_int, vp, vp, ctypes.c_int]
        L.intfft_pair_create.argtypes = [P(vp), P(_CGenerics), ctypes.c_int, ctypes.c_int64, ctypes.c_int]
        L.intfft_pair_destroy.argtypes = [vp]
        L.intfft_pair_query.argtypes = [vp, P(_CLayout)]
        L.intfft_pair_exec.argtypes = [vp, vp, vp, vp]
        L.intfft_pair_exec_host.argtypes = [vp, vp, vp]
        L.intfft_host_alloc.argtypes = [P(vp), ctypes.c_size_t]
        L.intfft_host_free.argtypes = [vp]
        L.intfft_multi_create.argtypes = [P(vp), P(_CGenerics), ctypes.c_int64, P(ctypes.c_int), ctypes.c_int]
        L.intfft_multi_destroy.argtypes = [vp]
        L.intfft_multi_devices.argtypes = [vp]
        L.intfft_multi_query.argtypes = [vp, P(_CLayout)]
        L.intfft_multi_shard.argtypes = [vp, ctypes.c_int, P(ctypes.c_int), P(ctypes.c_int64), P(ctypes.c_int64)]
        L.intfft_multi_exec_host.argtypes = [vp, vp, vp]
        L.intfft_multi_exec.argtypes = [vp, P(vp), P(vp), P(vp)]
        L.intfft_bitrev.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, vp, vp, ctypes.c_int, vp]
        L.intfft_fill_random.argtypes = [vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                         ctypes.c_int, vp]
        L.intfft_checksum.argtypes = [vp, ctypes.c_int64, ctypes.c_int, P(ctypes.c_uint64), ctypes.c_int, vp]
        L.intfft_describe.argtypes = [P(_CGenerics), ctypes.c_int64, ctypes.c_char_p, ctypes.c_size_t]
        L.intfft_launch_count.restype = ctypes.c_int64
        L.intfft_strerror.restype = ctypes.c_char_p
        L.intfft_strerror.argtypes = [ctypes.c_int]
        L.intfft_version.restype = ctypes.c_int
        _lib = L
    return _lib


def _strerror(status: int) -> str:
    try:
        return lib().intfft_strerror(status).decode()
    except Exception:  # pragma: no cover
        return "?"


def set_mode(mode: str) -> tuple[int, int]:
    """Legacy MODE string -> (FORMAT, RNDMODE), tb/fft_signle_test.vhd:81-88,109-115."""
    table = {"UNSCALED": (1, 0), "ROUNDING": (0, 1), "TRUNCATE": (0, 0)}
    if mode not in table:
        raise ValueError("MODE must be UNSCALED, ROUNDING or TRUNCATE")
    return table[mode]


@dataclass(frozen=True)
class Generics:
    """Generics of int_fftNk / int_ifftNk.  RAMB_TYPE ("WRAP"/"CONT") and USE_MLT are accepted for
    interface parity; they only select FPGA resources / valid-strobe discipline and never change a
    value (int_fftNk.vhd:22-31, rom_twiddle_int.vhd:225-240), so they do not reach the C-ABI."""
    NFFT: int = 5
    DATA_WIDTH: int = 16
    TWDL_WIDTH: int = 16
    FORMAT: int = 1
    RNDMODE: int = 0
    XSER: str = "NEW"
    USE_FLY: int = 1
    RAMB_TYPE: str = "WRAP"
    USE_MLT: bool = False

    def c_struct(self, direction: int) -> _CGenerics:
        if self.XSER not in ("OLD", "NEW"):
            raise IntfftError(EINVAL, f"XSER={self.XSER!r}")
        if self.RAMB_TYPE not in ("WRAP", "CONT"):
            raise IntfftError(EINVAL, f"RAMB_TYPE={self.RAMB_TYPE!r}")
        return _CGenerics(self.NFFT, self.DATA_WIDTH, self.TWDL_WIDTH, self.FORMAT, self.RNDMODE,
                          1 if self.XSER == "NEW" else 0, self.USE_FLY, direction)

    @property
    def out_width(self) -> int:
        return self.DATA_WIDTH + self.FORMAT * self.NFFT


def scalar_dtype(width: int):
    return np.int16 if width <= 16 else (np.int32 if width <= 32 else np.int64)


def validate(g: Generics, direction: int = 0) -> int:
    c = g.c_struct(direction)
    return lib().intfft_validate(ctypes.byref(c))


def twiddles(g: Generics, stage: int):
    """What rom_twiddle_int(STAGE=stage) streams for these generics (host-only)."""
    c = g.c_struct(0)
    re = np.empty(1 << stage, np.int32)
    im = np.empty(1 << stage, np.int32)
    st = lib().intfft_twiddles(ctypes.byref(c), stage, re.ctypes.data, im.ctypes.data)
    if st:
        raise IntfftError(st, "intfft_twiddles")
    return re, im


def twiddles_device(g: Generics, stage: int, device: int = 0):
    """The same stream recomputed by the kernels' on-device Taylor function (STAGE 11..19; needs a GPU)."""
    c = g.c_struct(0)
    re = np.empty(1 << stage, np.int32)
    im = np.empty(1 << stage, np.int32)
    st = lib().intfft_twiddles_device(ctypes.byref(c), stage, re.ctypes.data, im.ctypes.data, device)
    if st:
        raise IntfftError(st, "intfft_twiddles_device")
    return re, im


from .sharding import shard_range  # noqa: E402,F401  (re-exported)


class Core:
    """An elaborated int_fftNk (direction 0) or int_ifftNk (direction 1) for `batch` frames."""

    def __init__(self, generics: Generics, batch: int, direction: int = 0, device: int = 0):
        self.generics = generics
        self.direction = direction
        self.device = device
        self._h = ctypes.c_void_p()
        c = generics.c_struct(direction)
        st = lib().intfft_plan_create(ctypes.byref(self._h), ctypes.byref(c), batch, device)
        if st:
            self._h = ctypes.c_void_p()
            raise IntfftError(st, "intfft_plan_create")
        lay = _CLayout()
        lib().intfft_query(self._h, ctypes.byref(lay))
        self.layout = lay
        self.n = int(lay.n)
        self.batch = int(lay.batch)
        self.in_dtype = scalar_dtype(lay.in_width)
        self.out_dtype = scalar_dtype(lay.out_width)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().intfft_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    # -- device path: torch tensors are only the memory / stream plumbing ------------------------
    def new_input(self):
        import torch
        return torch.empty((self.batch, self.n, 2), dtype=_torch_dtype(self.in_dtype), device=f"cuda:{self.device}")

    def new_output(self):
        import torch
        return torch.empty((self.batch, self.n, 2), dtype=_torch_dtype(self.out_dtype), device=f"cuda:{self.device}")

    def _check(self, d_in, d_out):
        """Device tensors handed to the C-ABI: right device, contiguous, right element size and count (a wrong
        tensor would otherwise become an out-of-bounds device access inside a kernel)."""
        for t, sb, what in ((d_in, self.layout.in_scalar_bytes, "input"), (d_out, self.layout.out_scalar_bytes, "output")):
            if not (t.is_cuda and t.device.index == self.device and t.is_contiguous()):
                raise IntfftError(EINVAL, f"{what} tensor must be contiguous on cuda:{self.device}")
            if t.numel() != self.batch * self.n * 2 or t.element_size() != sb:
                raise IntfftError(EINVAL, f"{what} tensor must hold {self.batch} x {self.n} x 2 scalars of {sb} bytes")

    def exec(self, d_in, d_out=None, stream=None):
        """Run the batch on device tensors shaped [batch, N, 2] ({re, im} interleaved)."""
        import torch
        if d_out is None:
            d_out = self.new_output()
        self._check(d_in, d_out)
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        st = lib().intfft_exec(self._h, d_in.data_ptr(), d_out.data_ptr(), s)
        if st:
            raise IntfftError(st, "intfft_exec")
        return d_out

    def exec_natural(self, d_in, d_out=None, stream=None):
        """int_fft_single_path semantics: natural order in and out (exec + int_bitrev_order)."""
        import torch
        if d_out is None:
            d_out = self.new_output()
        self._check(d_in, d_out)
        if d_in.data_ptr() == d_out.data_ptr():
            raise IntfftError(EINVAL, "exec_natural is out of place")
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        st = lib().intfft_exec_natural(self._h, d_in.data_ptr(), d_out.data_ptr(), s)
        if st:
            raise IntfftError(st, "intfft_exec_natural")
        return d_out

    # -- host path: what a testbench-style caller uses ------------------------------------------
    def exec_host(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        x = np.ascontiguousarray(x, self.in_dtype).reshape(self.batch, self.n, 2)
        if out is None:
            out = np.empty((self.batch, self.n, 2), self.out_dtype)
        assert out.dtype == self.out_dtype and out.flags.c_contiguous and out.size == x.size
        st = lib().intfft_exec_host(self._h, x.ctypes.data, out.ctypes.data)
        if st:
            raise IntfftError(st, "intfft_exec_host")
        return out

    def exec_host_ptr(self, h_in_ptr: int, h_out_ptr: int):
        st = lib().intfft_exec_host(self._h, h_in_ptr, h_out_ptr)
        if st:
            raise IntfftError(st, "intfft_exec_host")


class Pair:
    """int_fft_ifft_pair on the core lanes: int_fftNk -> int_ifftNk(DATA_WIDTH + FORMAT*NFFT), natural order in
    and out (main/int_fft_ifft_pair.vhd:209-283).  FLY_FWD = generics.USE_FLY, FLY_INV = fly_inv."""

    def __init__(self, generics: Generics, batch: int, fly_inv: int = 1, device: int = 0):
        self.generics, self.device = generics, device
        self._h = ctypes.c_void_p()
        c = generics.c_struct(0)
        st = lib().intfft_pair_create(ctypes.byref(self._h), ctypes.byref(c), fly_inv, batch, device)
        if st:
            self._h = ctypes.c_void_p()
            raise IntfftError(st, "intfft_pair_create")
        lay = _CLayout()
        lib().intfft_pair_query(self._h, ctypes.byref(lay))
        self.layout, self.n, self.batch = lay, int(lay.n), int(lay.batch)
        self.in_dtype, self.out_dtype = scalar_dtype(lay.in_width), scalar_dtype(lay.out_width)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().intfft_pair_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def exec(self, d_in, d_out=None, stream=None):
        import torch
        if d_out is None:
            d_out = torch.empty((self.batch, self.n, 2), dtype=_torch_dtype(self.out_dtype), device=d_in.device)
        for t, sb, what in ((d_in, self.layout.in_scalar_bytes, "input"), (d_out, self.layout.out_scalar_bytes, "output")):
            if not (t.is_cuda and t.device.index == self.device and t.is_contiguous()):
                raise IntfftError(EINVAL, f"{what} tensor must be contiguous on cuda:{self.device}")
            if t.numel() != self.batch * self.n * 2 or t.element_size() != sb:
                raise IntfftError(EINVAL, f"{what} tensor must hold {self.batch} x {self.n} x 2 scalars of {sb} bytes")
        s = stream if stream is not None else torch.cuda.current_stream(self.device).cuda_stream
        st = lib().intfft_pair_exec(self._h, d_in.data_ptr(), d_out.data_ptr(), s)
        if st:
            raise IntfftError(st, "intfft_pair_exec")
        return d_out

    def exec_host(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """tb/fft_double_test.vhd semantics through host buffers; the spectrum never leaves the device."""
        x = np.ascontiguousarray(x, self.in_dtype).reshape(self.batch, self.n, 2)
        if out is None:
            out = np.empty((self.batch, self.n, 2), self.out_dtype)
        assert out.dtype == self.out_dtype and out.flags.c_contiguous and out.size == x.size
        st = lib().intfft_pair_exec_host(self._h, x.ctypes.data, out.ctypes.data)
        if st:
            raise IntfftError(st, "intfft_pair_exec_host")
        return out


class Multi:
    """One process, several devices: the batch is cut into contiguous shards (shard_range), one per device, and
    every device runs the same int_fftNk / int_ifftNk plan on its shard — no exchange step."""

    def __init__(self, generics: Generics, batch: int, direction: int = 0, devices=(0,)):
        self.generics, self.direction, self.devices = generics, direction, list(devices)
        self._h = ctypes.c_void_p()
        c = generics.c_struct(direction)
        arr = (ctypes.c_int * len(self.devices))(*self.devices)
        st = lib().intfft_multi_create(ctypes.byref(self._h), ctypes.byref(c), batch, arr, len(self.devices))
        if st:
            self._h = ctypes.c_void_p()
            raise IntfftError(st, "intfft_multi_create")
        lay = _CLayout()
        lib().intfft_multi_query(self._h, ctypes.byref(lay))
        self.layout, self.n, self.batch = lay, int(lay.n), int(lay.batch)
        self.in_dtype, self.out_dtype = scalar_dtype(lay.in_width), scalar_dtype(lay.out_width)

    def shards(self):
        out = []
        for i in range(lib().intfft_multi_devices(self._h)):
            dev, first, frames = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
            lib().intfft_multi_shard(self._h, i, ctypes.byref(dev), ctypes.byref(first), ctypes.byref(frames))
            out.append((dev.value, first.value, frames.value))
        return out

    def exec_host(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        x = np.ascontiguousarray(x, self.in_dtype).reshape(self.batch, self.n, 2)
        if out is None:
            out = np.empty((self.batch, self.n, 2), self.out_dtype)
        assert out.dtype == self.out_dtype and out.flags.c_contiguous and out.size == x.size
        self.exec_host_ptr(x.ctypes.data, out.ctypes.data)
        return out

    def exec_host_ptr(self, h_in_ptr: int, h_out_ptr: int):
        st = lib().intfft_multi_exec_host(self._h, h_in_ptr, h_out_ptr)
        if st:
            raise IntfftError(st, "intfft_multi_exec_host")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().intfft_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close


class HostBuffer:
    """Page-locked host memory from intfft_host_alloc (portable across devices), viewed as a NumPy array."""

    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = ctypes.c_void_p()
        st = lib().intfft_host_alloc(ctypes.byref(self._p), self.nbytes)
        if st:
            self._p = ctypes.c_void_p()
            raise IntfftError(st, "intfft_host_alloc")
        buf = (ctypes.c_char * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype).reshape(self.shape)

    @property
    def ptr(self) -> int:
        return self._p.value

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            self.array = None
            lib().intfft_host_free(self._p)
            self._p = ctypes.c_void_p()

    __del__ = close


def _torch_dtype(np_dtype):
    import torch
    return {np.int16: torch.int16, np.int32: torch.int32, np.int64: torch.int64}[np_dtype]


def describe(generics: Generics, batch: int = 1, direction: int = 0) -> str:
    """Kernel chain a plan for these generics would run (host-only; no device needed)."""
    buf = ctypes.create_string_buffer(512)
    c = generics.c_struct(direction)
    st = lib().intfft_describe(ctypes.byref(c), batch, buf, len(buf))
    if st:
        raise IntfftError(st, "intfft_describe")
    return buf.value.decode()


def int_fftNk(batch: int, device: int = 0, **generics) -> Core:
    """Forward core: natural in -> bit-reversed out (src/vhdl/fft/int_fftNk.vhd)."""
    return Core(Generics(**generics), batch, 0, device)


def int_ifftNk(batch: int, device: int = 0, **generics) -> Core:
    """Inverse core: bit-reversed in -> natural out (src/vhdl/fft/int_ifftNk.vhd)."""
    return Core(Generics(**generics), batch, 1, device)


def bitrev_order(d_in, nfft_log2: int, d_out=None, stream=None):
    """int_bitrev_order over a batch of frames on the device: out[bitrev(q)] = in[q]."""
    import torch
    if d_out is None:
        d_out = torch.empty_like(d_in)
    n = 1 << nfft_log2
    batch = d_in.numel() // (2 * n)
    s = stream if stream is not None else torch.cuda.current_stream(d_in.device).cuda_stream
    st = lib().intfft_bitrev(nfft_log2, d_in.element_size(), batch, d_in.data_ptr(), d_out.data_ptr(),
                             d_in.device.index or 0, s)
    if st:
        raise IntfftError(st, "intfft_bitrev")
    return d_out


def fill_random(d_buf, width: int, seed: int, stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream(d_buf.device).cuda_stream
    st = lib().intfft_fill_random(d_buf.data_ptr(), d_buf.numel(), d_buf.element_size(), width,
                                  seed & (2 ** 64 - 1), d_buf.device.index or 0, s)
    if st:
        raise IntfftError(st, "intfft_fill_random")
    return d_buf


def checksum(d_buf, stream=None) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream(d_buf.device).cuda_stream
    out = ctypes.c_uint64(0)
    st = lib().intfft_checksum(d_buf.data_ptr(), d_buf.numel(), d_buf.element_size(), ctypes.byref(out),
                               d_buf.device.index or 0, s)
    if st:
        raise IntfftError(st, "intfft_checksum")
    return int(out.value)


def launch_count() -> int:
    return int(lib().intfft_launch_count())
