import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import intfftk_b200 as ib
g = ib.Generics(NFFT=12, DATA_WIDTH=16, FORMAT=0)
batch = 65536
core = ib.Core(g, batch, 0)
hp_in = torch.empty((batch, 4096, 2), dtype=torch.int16, pin_memory=True)
hp_out = torch.empty((batch, 4096, 2), dtype=torch.int16, pin_memory=True)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("exec_host pinned   : %.2f ms" % t(lambda: core.exec_host_ptr(hp_in.data_ptr(), hp_out.data_ptr())))
np_in = np.zeros((batch, 4096, 2), np.int16); np_out = np.empty_like(np_in)
print("exec_host pageable : %.2f ms" % t(lambda: core.exec_host(np_in, np_out), reps=2))
