"""One-off fuzz of the API surfaces around the core path: for random generics / batches the host path, the natural-order
wrapper, in-place execution and the FFT -> IFFT pair must agree bit for bit with compositions of oracle calls."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import intfftk_b200 as ib
from oracle import c_oracle as co

rng = random.Random(int(os.environ.get("SEED", "3")))
bad, done = 0, {"host": 0, "natural": 0, "inplace": 0, "pair": 0, "pair_host": 0}
for it in range(int(os.environ.get("CASES", "200"))):
    nfft = rng.randint(3, 14)
    dw = rng.choice([16, 16, 12, 18, 24, 9])
    fmt, rnd = rng.choice([(0, 0), (0, 0), (0, 1), (1, 0)])
    d = rng.randint(0, 1)
    if dw + 2 * fmt * nfft > 64: fmt = 0
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, FORMAT=fmt, RNDMODE=rnd)
    if ib.validate(g, d) != 0: continue
    n = 1 << nfft
    batch = rng.choice([1, 3, rng.randint(1, max(1, (1 << 19) >> nfft))])
    x = co.fill_random(batch * n * 2, dw, rng.randint(0, 1 << 30)).reshape(batch, n, 2)
    og = co.generics(nfft, dw, 16, fmt, rnd, 1, 1, d)
    want = co.batch(og, x, 0)
    core = ib.Core(g, batch, d)
    dx = torch.from_numpy(x).cuda()
    what = rng.choice(["host", "natural", "inplace", "pair", "pair_host"])
    ok = True
    if what == "host":
        ok = np.array_equal(core.exec_host(x), want)
    elif what == "natural":
        got = core.exec_natural(dx).cpu().numpy()
        # FFT: natural in -> natural out = bitrev of the core's output; IFFT: natural-order spectrum in = bitrev first
        ref = co.bitrev(nfft, want) if d == 0 else co.batch(og, co.bitrev(nfft, x), 0)
        ok = np.array_equal(got, ref)
    elif what == "inplace":
        if core.new_output().element_size() == dx.element_size():
            y = dx.clone()
            core.exec(y, y)
            ok = np.array_equal(y.cpu().numpy(), want)
        else:
            what = None
    elif what in ("pair", "pair_host") and d == 0:
        try:
            pair = ib.Pair(g, batch)
        except ib.IntfftError:
            what = None
        else:
            spec = co.batch(co.generics(nfft, dw, 16, fmt, rnd, 1, 1, 0), x, 0)
            back = co.batch(co.generics(nfft, dw + fmt * nfft, 16, fmt, rnd, 1, 1, 1), spec, 0)
            got = pair.exec(dx).cpu().numpy() if what == "pair" else pair.exec_host(x)
            ok = np.array_equal(got, back)
            pair.close()
    else:
        what = None
    core.close()
    if what:
        done[what] += 1
        if not ok:
            bad += 1
            print("MISMATCH", what, nfft, dw, fmt, rnd, d, batch, flush=True)
print("done; mismatches:", bad, done)
