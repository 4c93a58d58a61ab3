"""Per-NFFT sweep (north_star: achieved HBM GB/s against the chip's peak for each NFFT): 16-bit scaled
DIF / DIT and 18-bit scaled DIF / DIT for NFFT = 8..20 at 2^28 samples per launch set (2^27 beyond 2^16 points
for the 32-bit containers).  Prints one JSON line per plan; run under
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
with SWEEP_ONCE=1 to get the per-kernel DRAM traffic of the same plans."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import intfftk_b200 as ib

once = os.environ.get("SWEEP_ONCE") == "1"
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
for dw in (16, 18):
    for direction in (0, 1):
        for nfft in range(8, 21):
            total_log2 = 28 if dw == 16 else 27
            batch = 1 << (total_log2 - nfft)
            g = ib.Generics(NFFT=nfft, DATA_WIDTH=dw, FORMAT=0)
            core = ib.Core(g, batch, direction)
            x, y = core.new_input(), core.new_output()
            ib.fill_random(x, dw, nfft)
            steps = 1 if once else 20
            for _ in range(1 if once else 3):
                core.exec(x, y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                core.exec(x, y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            byts = batch * (1 << nfft) * 2 * (x.element_size() + y.element_size())
            print(json.dumps({"nfft": nfft, "data_width": dw, "direction": "DIF" if direction == 0 else "DIT", "batch": batch,
                              "kernels": core.layout.n_passes, "ms": round(ms, 4), "gsamples_s": round(batch * (1 << nfft) / ms / 1e6, 1),
                              "algorithmic_gb_s": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / peak, 3)}), flush=True)
            core.close()
            del x, y
