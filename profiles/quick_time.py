"""Quick device-only timing of one plan (used while tuning kernels; not the bench contract)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import intfftk_b200 as ib

def time_plan(batch, steps=50, direction=0, **gk):
    steps = int(os.environ.get("QT_STEPS", steps))
    g = ib.Generics(**gk)
    core = ib.Core(g, batch, direction)
    x, y = core.new_input(), core.new_output()
    ib.fill_random(x, g.DATA_WIDTH, 1)
    for _ in range(3):
        core.exec(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        core.exec(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n = 1 << g.NFFT
    byts = batch * n * 2 * (x.element_size() + y.element_size())
    print(f"NFFT={g.NFFT} DW={g.DATA_WIDTH} fmt={g.FORMAT} rnd={g.RNDMODE} dir={direction} batch={batch}: {ms*1e3:9.1f} us  "
          f"{batch*n/ms/1e6:8.1f} Gsamples/s  {byts/ms/1e6:7.0f} GB/s = {byts/ms/1e6/6548.2*100:4.1f}% of HBM peak", flush=True)
    core.close()
    del x, y

def c2():
    for b in (444, 4440, 65536):
        time_plan(b, NFFT=12, DATA_WIDTH=16, FORMAT=0)
    time_plan(65536, direction=1, NFFT=12, DATA_WIDTH=16, FORMAT=0)

def small():
    for n in (8, 9, 10, 11):
        time_plan(65536 << (12 - n), NFFT=n, DATA_WIDTH=16, FORMAT=0)

def others():
    time_plan(131072, steps=5, direction=1, NFFT=13, DATA_WIDTH=18, FORMAT=0)
    time_plan(131072, steps=5, direction=1, NFFT=13, DATA_WIDTH=18, FORMAT=1)
    time_plan(4096, steps=5, NFFT=16, DATA_WIDTH=24, FORMAT=1)
    time_plan(256, steps=5, NFFT=20, DATA_WIDTH=16, FORMAT=0)
    time_plan(4096, steps=5, NFFT=16, DATA_WIDTH=16, FORMAT=0)
    time_plan(65536, steps=5, NFFT=12, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)
    time_plan(65536, steps=5, NFFT=12, DATA_WIDTH=16, FORMAT=1)

def extra():
    time_plan(131072, steps=5, direction=0, NFFT=13, DATA_WIDTH=18, FORMAT=0)
    time_plan(65536, steps=5, direction=1, NFFT=12, DATA_WIDTH=18, FORMAT=0)
    time_plan(65536, steps=5, direction=0, NFFT=12, DATA_WIDTH=18, FORMAT=0)
    time_plan(4096, steps=5, direction=1, NFFT=16, DATA_WIDTH=16, FORMAT=0)
    time_plan(4096, steps=5, direction=1, NFFT=16, DATA_WIDTH=24, FORMAT=0)
    time_plan(4096, steps=5, direction=0, NFFT=16, DATA_WIDTH=24, FORMAT=0)

def timeit(fn, steps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps

def f12():
    """f1 (natural-order wrapper) and f2 (FFT -> IFFT pair) at the c2 shape"""
    for direction in (0, 1):
        g = ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0)
        core = ib.Core(g, 65536, direction)
        x, y = core.new_input(), core.new_output()
        ib.fill_random(x, 16, 1)
        print(f"exec_natural c2 dir={direction}: {timeit(lambda: core.exec_natural(x, y))*1e3:8.1f} us   (exec alone {timeit(lambda: core.exec(x, y))*1e3:8.1f} us)", flush=True)
        core.close()
    z = torch.empty_like(x)
    print(f"bitrev_order alone (2^28 x 4 B): {timeit(lambda: ib.bitrev_order(x, 12, z))*1e3:8.1f} us", flush=True)
    pair = ib.Pair(ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0), 65536)
    print(f"pair c2: {timeit(lambda: pair.exec(x, y))*1e3:8.1f} us", flush=True)
    pair.close()
    g = ib.Generics(NFFT=13, DATA_WIDTH=18, FORMAT=0)
    core = ib.Core(g, 131072, 1)
    x, y = core.new_input(), core.new_output()
    ib.fill_random(x, 18, 1)
    print(f"exec_natural c5: {timeit(lambda: core.exec_natural(x, y), 5)*1e3:8.1f} us   (exec alone {timeit(lambda: core.exec(x, y), 5)*1e3:8.1f} us)", flush=True)
    core.close()

def small32():
    for d in (0, 1):
        for n in (9, 10, 11):
            time_plan(65536 << (12 - n), steps=5, direction=d, NFFT=n, DATA_WIDTH=18, FORMAT=0)
        for n in (14, 15):
            time_plan(65536 >> (n - 12), steps=5, direction=d, NFFT=n, DATA_WIDTH=18, FORMAT=0)

def modes():
    """every mode / width class at 4096 points (and a few two-pass sizes): looks for pathological paths"""
    for dw in (8, 12, 14, 16, 18, 20, 24, 27, 31, 32, 36, 40):
        for fmt, rnd in ((0, 0), (0, 1), (1, 0)):
            for d in (0, 1):
                if d == 0 and fmt == 1 and rnd == 1:
                    continue
                if dw + fmt * 12 > 64:
                    continue
                try:
                    time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=dw, FORMAT=fmt, RNDMODE=rnd)
                except Exception as e:
                    print("skip", dw, fmt, rnd, d, e, flush=True)
    for tw in (10, 18, 24):
        for d in (0, 1):
            time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=16, TWDL_WIDTH=tw, FORMAT=0)

def modes_mini():
    for d in (0, 1):
        time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=12, FORMAT=0)
        time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=12, FORMAT=0, RNDMODE=1)
        time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=16, TWDL_WIDTH=18, FORMAT=0)
        time_plan(256, steps=3, direction=d, NFFT=18, DATA_WIDTH=14, FORMAT=0)

def wide():
    for d in (0, 1):
        time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=24, FORMAT=1)
        time_plan(1024, steps=3, direction=d, NFFT=16, DATA_WIDTH=18, FORMAT=1)
        time_plan(16384, steps=3, direction=d, NFFT=12, DATA_WIDTH=32, FORMAT=0, RNDMODE=1)
        time_plan(1024, steps=3, direction=d, NFFT=16, DATA_WIDTH=24, FORMAT=1)

def tiny():
    for dw in (16, 18):
        for n in (3, 4, 5, 6, 7):
            for d in (0, 1):
                time_plan(1 << (26 - n), steps=3, direction=d, NFFT=n, DATA_WIDTH=dw, FORMAT=0)

def c3():
    time_plan(4096, steps=10, NFFT=16, DATA_WIDTH=24, FORMAT=1)

def c5():
    time_plan(131072, steps=10, direction=1, NFFT=13, DATA_WIDTH=18, FORMAT=0)

def c4():
    time_plan(256, steps=10, NFFT=20, DATA_WIDTH=16, FORMAT=0)

def r12():
    time_plan(65536, steps=3, NFFT=12, DATA_WIDTH=16, FORMAT=0, RNDMODE=1)

if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c2"]):
        globals()[name]()
