"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/intfft.h declares, validates generics like the reference elaborates them, and its twiddle
generator agrees with the oracle.  No compute call is made (there is no GPU here and no fallback)."""
import ctypes
import itertools
import os
import re

import numpy as np
import pytest

import intfftk_b200 as ib
from oracle import c_oracle as co

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "intfft.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(intfft_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ib.lib()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"libintfft_b200.so does not export {n}"
    assert lib.intfft_version() >= 100
    assert b"elaborate" in lib.intfft_strerror(-1)


def test_header_is_plain_c():
    """The header must compile as C (no torch / C++ types in the signatures)."""
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write('#include "intfft.h"\nint main(void){intfft_generics g; (void)g; return 0;}\n')
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src])


def test_validate_agrees_with_oracle_everywhere():
    n_checked = 0
    for nfft, dw, tw, fmt, rnd, xser, direction in itertools.product(
            (2, 3, 10, 13, 19, 20, 21), (7, 8, 16, 18, 26, 28, 36, 44, 45, 52, 53, 60, 64),
            (7, 8, 16, 18, 19, 25, 26, 27, 28), (0, 1), (0, 1), ("OLD", "NEW"), (0, 1)):
        g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd, XSER=xser)
        want = co.validate(co.generics(nfft, dw, tw, fmt, rnd, 1 if xser == "NEW" else 0, 1, direction))
        assert ib.validate(g, direction) == want, (nfft, dw, tw, fmt, rnd, xser, direction)
        n_checked += 1
    assert n_checked > 5000


@pytest.mark.parametrize("tw,xser", [(8, "NEW"), (16, "NEW"), (16, "OLD"), (17, "NEW"), (18, "OLD"), (24, "NEW"),
                                     (25, "OLD"), (27, "NEW")])
def test_twiddle_readback_matches_oracle(tw, xser):
    """intfft_twiddles == what rom_twiddle_int / row_twiddle_tay stream (oracle restatement)."""
    for stage in range(2, 20):
        re_, im_ = ib.twiddles(ib.Generics(TWDL_WIDTH=tw, XSER=xser), stage)
        r2, i2 = co.twiddle_table(co.generics(12, twdl_width=tw, xser=1 if xser == "NEW" else 0), stage)
        assert np.array_equal(re_, r2) and np.array_equal(im_, i2), stage


def test_plan_create_without_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ib.IntfftError) as ei:
        ib.int_fftNk(4, NFFT=8, FORMAT=0)
    assert ei.value.status == -2          # INTFFT_ECUDA: no fallback path exists


def test_bad_arguments():
    lib = ib.lib()
    assert lib.intfft_validate(None) == -1
    assert lib.intfft_plan_destroy(None) == -1
    assert lib.intfft_query(None, None) == -1
    assert lib.intfft_exec(None, None, None, None) == -1
    assert lib.intfft_twiddles(None, 5, None, None) == -1
    with pytest.raises(ib.IntfftError):
        ib.twiddles(ib.Generics(), 1)
    with pytest.raises(ib.IntfftError):
        ib.Generics(XSER="ULTRA").c_struct(0)


def test_mode_strings_and_generics_mirror():
    assert ib.set_mode("UNSCALED") == (1, 0)
    assert ib.set_mode("ROUNDING") == (0, 1)
    assert ib.set_mode("TRUNCATE") == (0, 0)
    with pytest.raises(ValueError):
        ib.set_mode("SATURATE")
    g = ib.Generics(NFFT=16, DATA_WIDTH=24, FORMAT=1)
    assert g.out_width == 40
    c = g.c_struct(1)
    assert (c.nfft_log2, c.data_width, c.twdl_width, c.format, c.xser, c.direction) == (16, 24, 16, 1, 1, 1)


# ---- host logic: which kernels a plan runs (intfft_describe needs no device) --------------------------------
PLAN_CHAINS = [
    # BASELINE configurations
    (dict(NFFT=12, DATA_WIDTH=16, FORMAT=0), 0, ["fast16[bits 0..11"]),
    (dict(NFFT=16, DATA_WIDTH=24, FORMAT=1), 0, ["fast32_strided[bits 8..15", "fast64[bits 0..7", "instance 1"]),
    (dict(NFFT=20, DATA_WIDTH=16, FORMAT=0), 0, ["fast16_strided[bits 12..19", "fast16[bits 0..11"]),
    (dict(NFFT=13, DATA_WIDTH=18, FORMAT=0), 1, ["fast32_n13[bits 0..12"]),
    (dict(NFFT=13, DATA_WIDTH=18, FORMAT=1), 1, ["fast32_n13[bits 0..12"]),
    # every size class of the 16-bit scaled family, both directions
    (dict(NFFT=3, DATA_WIDTH=16, FORMAT=0), 0, ["fast16[bits 0..2"]),
    (dict(NFFT=7, DATA_WIDTH=12, FORMAT=0, RNDMODE=1), 1, ["fast16[bits 0..6"]),
    (dict(NFFT=13, DATA_WIDTH=16, FORMAT=0), 0, ["fast16_n13[bits 0..12, 4->4 B]"]),            # one pass since round 2
    (dict(NFFT=13, DATA_WIDTH=12, FORMAT=0, RNDMODE=1), 1, ["fast16_n13[bits 0..12, 4->4 B]"]),
    (dict(NFFT=14, DATA_WIDTH=16, FORMAT=0), 0, ["fast16_n14[bits 0..13, 4->4 B]"]),            # one pass since round 2
    (dict(NFFT=15, DATA_WIDTH=16, FORMAT=0), 0, ["fast16_strided[bits 11..14", "fast16[bits 0..10"]),
    (dict(NFFT=16, DATA_WIDTH=16, FORMAT=0), 1, ["fast16[bits 0..11", "fast16_strided[bits 12..15"]),
    (dict(NFFT=17, DATA_WIDTH=14, FORMAT=0), 0, ["fast16_strided[bits 9..16", "fast16[bits 0..8"]),
    # 32-bit lanes: wider data, wider twiddles, unscaled growth that still fits
    (dict(NFFT=5, DATA_WIDTH=18, FORMAT=0), 0, ["fast32[bits 0..4"]),
    (dict(NFFT=12, DATA_WIDTH=16, TWDL_WIDTH=18, FORMAT=0), 0, ["fast32[bits 0..11, 4->4 B"]),
    (dict(NFFT=12, DATA_WIDTH=16, FORMAT=1), 0, ["fast32[bits 0..11, 4->8 B"]),
    (dict(NFFT=18, DATA_WIDTH=20, FORMAT=0), 1, ["fast32[bits 0..9", "fast32_strided[bits 10..17"]),
    # wide plans: every pass on the narrowest lane family its widths allow
    (dict(NFFT=12, DATA_WIDTH=24, FORMAT=1), 0, ["fast32_strided[bits 8..11", "fast64[bits 0..7", "instance 3"]),
    (dict(NFFT=12, DATA_WIDTH=24, FORMAT=1), 1, ["fast32[bits 0..7", "fast64_strided[bits 8..11"]),
    (dict(NFFT=16, DATA_WIDTH=18, FORMAT=1), 0, ["fast32_strided[bits 8..15", "fast64[bits 0..7", "instance 3"]),
    (dict(NFFT=16, DATA_WIDTH=36, FORMAT=0), 0, ["fast64_strided[bits 8..15", "fast64[bits 0..7", "instance 1"]),
    (dict(NFFT=8, DATA_WIDTH=45, FORMAT=0), 0, ["fast64[bits 0..7", "instance 2"]),
    # beyond 64-bit products, or geometries without a specialised kernel: the generic tile kernel
    (dict(NFFT=16, DATA_WIDTH=40, FORMAT=1), 1, ["tile[", "lane 128"]),
    (dict(NFFT=10, DATA_WIDTH=24, FORMAT=1), 0, ["tile[bits 0..9", "lane 64"]),
    (dict(NFFT=12, DATA_WIDTH=16, FORMAT=0, USE_FLY=0), 0, ["bypass"]),
]


@pytest.mark.parametrize("kw,direction,expect", PLAN_CHAINS)
def test_kernel_selection_is_pinned(kw, direction, expect):
    text = ib.describe(ib.Generics(**kw), 8, direction)
    pos = 0
    for piece in expect:                       # the pieces appear in this order
        at = text.find(piece, pos)
        assert at >= 0, f"{kw} dir={direction}: expected {piece!r} in {text!r}"
        pos = at + 1


def test_describe_rejects_what_does_not_elaborate():
    with pytest.raises(ib.IntfftError):
        ib.describe(ib.Generics(NFFT=12, DATA_WIDTH=16, TWDL_WIDTH=30), 1, 0)
    with pytest.raises(ib.IntfftError):
        ib.describe(ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=1, RNDMODE=1), 1, 0)


def _container_bytes(width):
    return 4 if width <= 16 else (8 if width <= 32 else 16)        # bytes per complex sample


def test_every_plan_covers_each_stage_bit_once_and_chains_its_containers():
    """Host logic over the whole grid of generics: whatever kernels are chosen, the passes of a plan partition the
    NFFT stage bits, run them in the direction's order, and hand containers on consistently."""
    pat = re.compile(r"(\w+)\[bits (\d+)\.\.(\d+), (\d+)->(\d+) B")
    checked = 0
    for nfft, dw, tw, fmt, rnd, direction in itertools.product(
            range(3, 21), (8, 12, 16, 18, 24, 27, 32, 36, 40, 48), (12, 16, 18, 24), (0, 1), (0, 1), (0, 1)):
        g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd)
        if ib.validate(g, direction) != 0:
            continue
        text = ib.describe(g, 4, direction)
        passes = [(m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5))) for m in pat.finditer(text)]
        assert passes, text
        bits = sorted((lo, hi) for _, lo, hi, _, _ in passes)
        assert bits[0][0] == 0 and bits[-1][1] == nfft - 1, text
        for (l0, h0), (l1, h1) in zip(bits, bits[1:]):
            assert l1 == h0 + 1, text
        order = [lo for _, lo, _, _, _ in passes]
        assert order == sorted(order, reverse=(direction == 0)), text      # DIF walks the bits downwards
        assert passes[0][3] == _container_bytes(dw), text
        assert passes[-1][4] == _container_bytes(dw + fmt * nfft), text
        for a, b in zip(passes, passes[1:]):
            assert a[4] == b[3], text
        checked += 1
    assert checked > 2000


def test_cpp_host_describes_a_plan_without_a_device():
    import subprocess
    exe = os.path.join(ROOT, "intfftk_b200", "intfft_host")
    if not os.path.exists(exe):
        pytest.skip("intfft_host not built")
    out = subprocess.run([exe, "--describe", "--nfft", "20", "--dw", "16", "--mode", "TRUNCATE"], capture_output=True, text=True)
    assert out.returncode == 0 and "fast16_strided[bits 12..19" in out.stdout and "16-bit out" in out.stdout
    bad = subprocess.run([exe, "--describe", "--tw", "30"], capture_output=True, text=True)
    assert bad.returncode != 0 and "elaborate" in bad.stderr
