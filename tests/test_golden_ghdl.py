"""Reference-held vectors (tests/golden/ghdl/*.npz, produced by oracle/rtl/ghdl_tb/run_ghdl.py from an RTL simulation
of the unmodified reference) against the C oracle and the CUDA path.  Skipped while no vectors are present: neither
this image nor the GPU box has a VHDL simulator (profiles/r02/tool_probe_r02.txt)."""
import glob
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
VECTORS = sorted(glob.glob(os.path.join(HERE, "golden", "ghdl", "*.npz")))


def _load_runner():
    p = os.path.join(os.path.dirname(HERE), "oracle", "rtl", "ghdl_tb", "run_ghdl.py")
    spec = importlib.util.spec_from_file_location("run_ghdl", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_lane_packing_round_trip():
    """The stimulus / dump lane order run_ghdl.py uses is the stream contract of include/intfft.h (FFT: halves in,
    even / odd out; IFFT: even / odd in, halves out) and is its own inverse pairwise."""
    rg = _load_runner()
    x = np.arange(64 * 2).reshape(64, 2)
    b0 = rg.lanes_in(x, 0)
    assert b0.shape == (32, 4) and (b0[5] == [*x[5], *x[37]]).all()
    b1 = rg.lanes_in(x, 1)
    assert (b1[5] == [*x[10], *x[11]]).all()
    assert (rg.lanes_out(b1, 0) == x).all()          # FFT output lanes are the IFFT's input lanes
    assert (rg.lanes_out(b0, 1) == x).all()          # IFFT output lanes are the FFT's input lanes
    assert len({c[0] for c in rg.CASES}) == len(rg.CASES)
    ref = "/root/reference/src/vhdl"
    if os.path.isdir(ref):
        assert all(os.path.exists(os.path.join(ref, f)) for f in rg.REF_FILES)


@pytest.mark.skipif(not VECTORS, reason="no reference-held vectors yet: run oracle/rtl/ghdl_tb/run_ghdl.py where GHDL exists")
@pytest.mark.parametrize("path", VECTORS)
def test_oracle_matches_rtl_simulation(path):
    from oracle import c_oracle as co
    d = np.load(path)
    g = co.generics(*[int(v) for v in d["generics"]])
    x = d["x"].astype(co.scalar_dtype(g.data_width))
    want = d["y"]
    got = co.batch(g, x)
    assert np.array_equal(got.astype(np.int64), want), os.path.basename(path)


@pytest.mark.gpu
@pytest.mark.skipif(not VECTORS, reason="no reference-held vectors yet: run oracle/rtl/ghdl_tb/run_ghdl.py where GHDL exists")
@pytest.mark.parametrize("path", VECTORS)
def test_cuda_matches_rtl_simulation(path):
    import intfftk_b200 as ib
    d = np.load(path)
    nfft, dw, tw, fmt, rnd, xser, fly, direction = [int(v) for v in d["generics"]]
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, FORMAT=fmt, RNDMODE=rnd, XSER="NEW" if xser else "OLD", USE_FLY=fly)
    core = ib.Core(g, d["x"].shape[0], direction)
    got = core.exec_host(d["x"].astype(core.in_dtype))
    core.close()
    assert np.array_equal(got.astype(np.int64), d["y"]), os.path.basename(path)
