#!/usr/bin/env python
"""Produce REFERENCE-HELD vectors: simulate the unmodified VHDL of hukenovs/intfftk with GHDL and store the frames it
computes as golden files for tests/test_golden_ghdl.py.  TEST INFRASTRUCTURE ONLY.

    python oracle/rtl/ghdl_tb/run_ghdl.py [--reference /root/reference] [--out tests/golden/ghdl] [--unisim PATH]

Needs `ghdl` (absent from the build image and from the GPU box: profiles/r02/tool_probe_r02.txt — run this wherever
GHDL exists and commit the .npz files it writes).  Without --unisim the behavioural stand-in of this directory
(unisim_standin.vhd) replaces Xilinx's unisim library; with Xilinx's own library compiled for GHDL, pass its directory.

Outputs up to 32 bits wide only (to_integer in the testbench); wider plans are pinned at primitive level by
tests/test_oracle_rtl.py.  What a simulation adds over those tests is the ONE thing they cannot see: the register
alignment between the data path and the twiddle generators (assumption A1 in DESIGN.md §3).
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, ROOT)

# (name, direction, NFFT, DATA_WIDTH, TWDL_WIDTH, FORMAT, RNDMODE, XSER): small frames cover every stage kind (0, 1,
# ROM stages, and with NFFT = 12 / 13 the Taylor stages 11 and 12), the three modes, both device families
CASES = [
    ("fft_n5_16_unscaled_new", 0, 5, 16, 16, 1, 0, "NEW"), ("fft_n7_16_trunc_new", 0, 7, 16, 16, 0, 0, "NEW"),
    ("fft_n7_16_round_new", 0, 7, 16, 16, 0, 1, "NEW"), ("fft_n7_16_trunc_old", 0, 7, 16, 16, 0, 0, "OLD"),
    ("ifft_n7_16_unscaled_new", 1, 7, 16, 16, 1, 0, "NEW"), ("ifft_n7_18_trunc_new", 1, 7, 18, 16, 0, 0, "NEW"),
    ("ifft_n7_16_round_old", 1, 7, 16, 16, 0, 1, "OLD"), ("fft_n10_12_unscaled_new", 0, 10, 12, 18, 1, 0, "NEW"),
    ("fft_n12_16_trunc_new", 0, 12, 16, 16, 0, 0, "NEW"), ("ifft_n13_18_trunc_new", 1, 13, 18, 16, 0, 0, "NEW"),
    ("fft_n12_16_trunc_old", 0, 12, 16, 16, 0, 0, "OLD"), ("fft_n8_20_tw24_trunc_new", 0, 8, 20, 24, 0, 0, "NEW"),
]
FRAMES = 3

REF_FILES = [  # analysis order (leaf entities first); relative to <reference>/src/vhdl
    "math/mults/mlt42x18_dsp48e1.vhd", "math/mults/mlt44x18_dsp48e2.vhd", "math/mults/mlt59x18_dsp48e1.vhd",
    "math/mults/mlt61x18_dsp48e2.vhd", "math/mults/mlt35x25_dsp48e1.vhd", "math/mults/mlt35x27_dsp48e2.vhd",
    "math/mults/mlt52x25_dsp48e1.vhd", "math/mults/mlt52x27_dsp48e2.vhd",
    "math/cmult/int_cmult18x25_dsp48.vhd", "math/cmult/int_cmult_dbl18_dsp48.vhd", "math/cmult/int_cmult_dbl35_dsp48.vhd",
    "math/cmult/int_cmult_trpl18_dsp48.vhd", "math/cmult/int_cmult_trpl52_dsp48.vhd", "math/cmult/int_cmult_dsp48.vhd",
    "math/int_addsub_dsp48.vhd", "twiddle/row_twiddle_tay.vhd", "twiddle/rom_twiddle_int.vhd",
    "delay/int_delay_line.vhd", "delay/int_delay_wrap.vhd", "delay/int_align_fft.vhd", "delay/int_align_ifft.vhd",
    "fft/int_dif2_fly.vhd", "fft/int_dit2_fly.vhd", "fft/int_fftNk.vhd", "fft/int_ifftNk.vhd",
]


def lanes_in(x: np.ndarray, direction: int) -> np.ndarray:
    """flat in-place order [N, 2] -> beats [N/2, 4] = (re0, im0, re1, im1) as the core's input lanes take them."""
    n = x.shape[0]
    a, b = (x[: n // 2], x[n // 2:]) if direction == 0 else (x[0::2], x[1::2])      # int_fftNk.vhd:15-17 / int_ifftNk.vhd:15-17
    return np.concatenate([a, b], axis=1)


def lanes_out(beats: np.ndarray, direction: int) -> np.ndarray:
    """output beats [N/2, 4] -> flat in-place order [N, 2] (FFT: lanes = even / odd; IFFT: lanes = halves)."""
    a, b = beats[:, 0:2], beats[:, 2:4]
    if direction == 0:
        out = np.empty((2 * beats.shape[0], 2), beats.dtype)
        out[0::2], out[1::2] = a, b                                                   # int_fftNk.vhd:19-21
        return out
    return np.concatenate([a, b], axis=0)                                             # int_ifftNk.vhd:19-21


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "ghdl"))
    ap.add_argument("--unisim", default=None, help="directory of a GHDL-compiled Xilinx unisim library (default: the stand-in)")
    ap.add_argument("--std", default="93c")
    a = ap.parse_args()
    ghdl = shutil.which("ghdl")
    if not ghdl:
        raise SystemExit("run_ghdl.py: ghdl not found on PATH (nothing to do here; see the module docstring)")
    from oracle import c_oracle as co
    os.makedirs(a.out, exist_ok=True)
    flags = [f"--std={a.std}", "--ieee=synopsys", "-fexplicit", "-frelaxed-rules"]
    with tempfile.TemporaryDirectory() as wd:
        if a.unisim:
            flags.append(f"-P{a.unisim}")
        else:
            subprocess.check_call([ghdl, "-a", "--work=unisim", *flags, os.path.join(HERE, "unisim_standin.vhd")], cwd=wd)
        src = [os.path.join(a.reference, "src", "vhdl", f) for f in REF_FILES] + [os.path.join(HERE, "tb_intfft_dump.vhd")]
        subprocess.check_call([ghdl, "-a", *flags, *src], cwd=wd)
        for name, direction, nfft, dw, tw, fmt, rnd, xser in CASES:
            n = 1 << nfft
            x = co.fill_random(FRAMES * n * 2, dw, 1000 + nfft * 7 + dw).reshape(FRAMES, n, 2).astype(np.int64)
            stim = np.concatenate([lanes_in(f, direction) for f in x], axis=0)
            np.savetxt(os.path.join(wd, "stim.dat"), stim, fmt="%d")
            gen = dict(DIRECTION=direction, NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd, XSER=xser,
                       FRAMES=FRAMES, STIM="stim.dat", DOUT="dout.dat")
            subprocess.check_call([ghdl, "-r", *flags, "tb_intfft_dump"] + [f"-g{k}={v}" for k, v in gen.items()], cwd=wd)
            beats = np.loadtxt(os.path.join(wd, "dout.dat"), dtype=np.int64).reshape(FRAMES, n // 2, 4)
            y = np.stack([lanes_out(b, direction) for b in beats])
            np.savez_compressed(os.path.join(a.out, name + ".npz"), x=x, y=y,
                                generics=np.array([nfft, dw, tw, fmt, rnd, 1 if xser == "NEW" else 0, 1, direction]),
                                simulator=np.array(subprocess.run([ghdl, "--version"], capture_output=True, text=True).stdout.splitlines()[0]),
                                unisim=np.array("xilinx" if a.unisim else "stand-in (oracle/rtl/ghdl_tb/unisim_standin.vhd)"))
            print("wrote", name)


if __name__ == "__main__":
    main()
