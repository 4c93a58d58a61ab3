"""intfftk_b200 — B200-native integer FFT/IFFT engine, bit-exact with the intfftk FPGA cores.

Host-side mirror of the reference interface: the entity generics of `int_fftNk` / `int_ifftNk`
(src/vhdl/fft/int_fftNk.vhd:72-103, int_ifftNk.vhd:71-102) become `Generics`; an elaborated entity
becomes a `Core` bound to the C-ABI in include/intfft.h.  PyTorch is used only for device memory and
streams.  There is no CPU fallback: importing `core` loads libintfft_b200.so or raises.
"""
from .core import (Core, Pair, Multi, HostBuffer, Generics, IntfftError, int_fftNk, int_ifftNk, set_mode, lib, twiddles, twiddles_device, validate,
                   bitrev_order, fill_random, checksum, launch_count, shard_range, describe)

__all__ = ["Core", "Pair", "Multi", "HostBuffer", "Generics", "IntfftError", "int_fftNk", "int_ifftNk", "set_mode", "lib", "twiddles", "twiddles_device",
           "validate", "bitrev_order", "fill_random", "checksum", "launch_count", "shard_range", "describe"]
