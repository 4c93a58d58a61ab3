"""Small run of every kernel family for compute-sanitizer (racecheck / memcheck): a handful of frames per plan,
results compared with the oracle so that the run is also a parity check."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import intfftk_b200 as ib
from oracle import c_oracle as co

PLANS = [  # (NFFT, DW, FORMAT, RND, direction, batch)
    (12, 16, 0, 0, 0, 7), (12, 16, 0, 0, 1, 7), (12, 16, 0, 1, 0, 4), (8, 16, 0, 0, 0, 40), (11, 16, 0, 0, 0, 5), (11, 16, 0, 0, 1, 5),
    (9, 12, 0, 0, 1, 9), (20, 16, 0, 0, 0, 1), (16, 16, 0, 0, 1, 2), (13, 16, 0, 0, 0, 3),
    (13, 18, 0, 0, 1, 5), (13, 18, 1, 0, 1, 4), (13, 18, 0, 0, 0, 4), (13, 16, 1, 0, 1, 3),
    (12, 18, 0, 0, 0, 5), (12, 18, 0, 0, 1, 5), (10, 20, 0, 1, 1, 9), (11, 18, 1, 0, 0, 6), (12, 16, 1, 0, 1, 5),
    (14, 18, 0, 0, 0, 2), (16, 24, 1, 0, 0, 2), (8, 40, 0, 0, 1, 9), (7, 16, 1, 0, 0, 6),
    # round 2: TMA-staged strided passes (both geometries, both lane families, several frames per CTA), on-device Taylor
    (17, 16, 0, 0, 0, 3), (17, 16, 0, 0, 1, 3), (15, 16, 0, 0, 0, 5), (18, 18, 0, 0, 1, 2), (14, 18, 0, 0, 1, 5), (16, 24, 1, 0, 1, 2),
    # one-pass 16384-point kernel (several frames per CTA), bulk-TMA input of the 32-bit-lane DIF kernels (two and three rounds)
    (14, 16, 0, 0, 0, 9), (14, 16, 0, 0, 1, 9), (14, 12, 0, 1, 0, 5), (12, 18, 0, 0, 0, 900), (8, 18, 0, 0, 0, 9000), (10, 24, 0, 1, 0, 2500),
]
for nfft, dw, fmt, rnd, direction, batch in PLANS:
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, FORMAT=fmt, RNDMODE=rnd)
    x = co.fill_random(batch * (1 << nfft) * 2, dw, nfft + dw).reshape(batch, 1 << nfft, 2)
    core = ib.Core(g, batch, direction)
    got = core.exec(torch.from_numpy(x).cuda()).cpu().numpy()
    want = co.batch(co.generics(nfft, dw, 16, fmt, rnd, 1, 1, direction), x)
    assert np.array_equal(got, want), (nfft, dw, fmt, rnd, direction)
    if direction == 0 and nfft == 12 and dw == 16:
        nat = core.exec_natural(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(nat, co.bitrev(nfft, want)), "natural order"
    core.close()
print("sanitize_plans: all plans bit-exact")
