"""PCIe probe: can H2D and D2H overlap on this box, and at what rates? (context for the e2e number)"""
import time, torch
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
def chunks(k=32):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
print("H2D 1GiB: %.2f ms  %.1f GB/s" % (t(h2d)*1e3, n/t(h2d)/1e9))
print("D2H 1GiB: %.2f ms  %.1f GB/s" % (t(d2h)*1e3, n/t(d2h)/1e9))
print("both concurrently: %.2f ms" % (t(both)*1e3))
print("both, 32 chunks each: %.2f ms" % (t(chunks)*1e3))
import subprocess
print(subprocess.run(["nvidia-smi","--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max","--format=csv"],capture_output=True,text=True).stdout)
