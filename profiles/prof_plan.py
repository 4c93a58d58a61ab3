"""Run each named plan twice (for ncu captures: profile the launches of the second run with -s / -k)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import intfftk_b200 as ib

PLANS = {
    "c2": (65536, 0, dict(NFFT=12, DATA_WIDTH=16, FORMAT=0)),
    "c2dit": (65536, 1, dict(NFFT=12, DATA_WIDTH=16, FORMAT=0)),
    "c3": (4096, 0, dict(NFFT=16, DATA_WIDTH=24, FORMAT=1)),
    "c4": (256, 0, dict(NFFT=20, DATA_WIDTH=16, FORMAT=0)),
    "c5": (131072, 1, dict(NFFT=13, DATA_WIDTH=18, FORMAT=0)),
    "c5u": (131072, 1, dict(NFFT=13, DATA_WIDTH=18, FORMAT=1)),
    "u12": (65536, 0, dict(NFFT=12, DATA_WIDTH=16, FORMAT=1)),
    "r12": (65536, 0, dict(NFFT=12, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)),
    "s16": (4096, 0, dict(NFFT=16, DATA_WIDTH=16, FORMAT=0)),
    "d18": (32768, 0, dict(NFFT=12, DATA_WIDTH=18, FORMAT=0)),
    "t18": (32768, 1, dict(NFFT=12, DATA_WIDTH=18, FORMAT=0)),
    "w12": (65536, 0, dict(NFFT=12, DATA_WIDTH=12, FORMAT=0)),
    "n13": (32768, 0, dict(NFFT=13, DATA_WIDTH=16, FORMAT=0)),
    "n13t": (32768, 1, dict(NFFT=13, DATA_WIDTH=16, FORMAT=0)),
    "n10": (262144, 0, dict(NFFT=10, DATA_WIDTH=16, FORMAT=0)),
    "n14": (16384, 0, dict(NFFT=14, DATA_WIDTH=16, FORMAT=0)),
    "n14t": (16384, 1, dict(NFFT=14, DATA_WIDTH=16, FORMAT=0)),
    "d18n13": (16384, 0, dict(NFFT=13, DATA_WIDTH=18, FORMAT=0)),
}
for name in sys.argv[1:]:
    batch, direction, gk = PLANS[name]
    g = ib.Generics(**gk)
    core = ib.Core(g, batch, direction)
    x, y = core.new_input(), core.new_output()
    ib.fill_random(x, g.DATA_WIDTH, 1)
    core.exec(x, y)
    core.exec(x, y)
    torch.cuda.synchronize()
    core.close()
    del x, y
