"""f2: int_fft_ifft_pair at the c2 shape — one fused kernel (packed-16, 2^8..2^12 points) against the two-launch form."""
import sys, os, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child():
    import torch
    import intfftk_b200 as ib
    for nfft, batch in ((12, 65536), (10, 262144), (8, 1 << 20)):
        g = ib.Generics(NFFT=nfft, DATA_WIDTH=16, FORMAT=0)
        pair = ib.Pair(g, batch)
        x = torch.empty((batch, 1 << nfft, 2), dtype=torch.int16, device="cuda")
        y = torch.empty_like(x)
        ib.fill_random(x, 15, 1)
        for _ in range(3):
            pair.exec(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pair.exec(x, y)
        e1.record(); torch.cuda.synchronize()
        print(json.dumps({"nfft": nfft, "batch": batch, "unfused": bool(os.environ.get("INTFFT_PAIR_UNFUSED")),
                          "launches_per_pair": int(pair.layout.n_passes), "ms": round(e0.elapsed_time(e1) / 20, 4),
                          "checksum": f"{ib.checksum(y):016x}"}), flush=True)
        pair.close()

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for env in ({}, {"INTFFT_PAIR_UNFUSED": "1"}):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env))
