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
