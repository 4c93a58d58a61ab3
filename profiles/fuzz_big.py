import os, sys, random
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import intfftk_b200 as ib
from oracle import c_oracle as co
rng = random.Random(int(os.environ.get("SEED", "11")))
bad = 0
fam = {}
for it in range(int(os.environ.get("CASES", "120"))):
    nfft = rng.randint(15, 20)
    xser = rng.choice(["NEW", "OLD"])
    tw = rng.choice([16, 16, 16, rng.randint(8, 18), rng.randint(19, 27 if xser == "NEW" else 25)])
    fmt, rnd = rng.choice([(0, 0), (0, 0), (0, 1), (1, 0)])
    dw = rng.choice([16, 16, 12, 18, 18, 14, 24, rng.randint(8, 32)])
    if dw + fmt * nfft > 64: continue
    d = rng.randint(0, 1)
    g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, TWDL_WIDTH=tw, XSER=xser, FORMAT=fmt, RNDMODE=rnd)
    if ib.validate(g, d) != 0: continue
    batch = rng.randint(1, 3)
    os.environ["INTFFT_STRIDED_TMA"] = rng.choice(["0", "1"])
    n = 1 << nfft
    x = co.fill_random(batch * n * 2, dw, rng.randint(0, 1 << 30)).reshape(batch, n, 2)
    want = co.batch(co.generics(nfft, dw, tw, fmt, rnd, 1 if xser == "NEW" else 0, 1, d), x, 0)
    core = ib.Core(g, batch, d)
    got = core.exec(torch.from_numpy(x).cuda()).cpu().numpy()
    chain = ib.describe(g, batch, d)
    core.close()
    for part in chain.split(" -> "): fam[part.split("[")[0]] = fam.get(part.split("[")[0], 0) + 1
    if not np.array_equal(got, want):
        bad += 1
        print("MISMATCH", nfft, dw, tw, xser, fmt, rnd, d, batch, os.environ["INTFFT_STRIDED_TMA"], chain, flush=True)
print("done; mismatches:", bad, fam)
