"""Two-pass plans run group by group (Plan::group_frames): time c3 / c4 / 16-bit NFFT 13..16 for several L2
budgets (INTFFT_GROUP_MB, 0 = whole batch at once = the round-1 behaviour).  One JSON line per (plan, budget)."""
import sys, os, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PLANS = {
    "c4": (dict(NFFT=20, DATA_WIDTH=16, FORMAT=0), 0, 256),
    "c3": (dict(NFFT=16, DATA_WIDTH=24, FORMAT=1), 0, 4096),
    "n13_16b": (dict(NFFT=13, DATA_WIDTH=16, FORMAT=0), 0, 32768),
    "n14_16b": (dict(NFFT=14, DATA_WIDTH=16, FORMAT=0), 0, 16384),
    "n16_16b": (dict(NFFT=16, DATA_WIDTH=16, FORMAT=0), 0, 4096),
    "n16_16b_dit": (dict(NFFT=16, DATA_WIDTH=16, FORMAT=0), 1, 4096),
    "n17_16b": (dict(NFFT=17, DATA_WIDTH=16, FORMAT=0), 0, 2048),
    "n16_18b_dit": (dict(NFFT=16, DATA_WIDTH=18, FORMAT=0), 1, 2048),
}

def child(name):
    import torch
    import intfftk_b200 as ib
    gk, direction, batch = PLANS[name]
    g = ib.Generics(**gk)
    core = ib.Core(g, batch, direction)
    x, y = core.new_input(), core.new_output()
    ib.fill_random(x, g.DATA_WIDTH, 1)
    for _ in range(3):
        core.exec(x, y)
    torch.cuda.synchronize()
    steps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        core.exec(x, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    byts = batch * (1 << g.NFFT) * 2 * (x.element_size() + y.element_size())
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    print(json.dumps({"plan": name, "group_mb": os.environ.get("INTFFT_GROUP_MB"), "ms": round(ms, 4),
                      "frac": round(byts / ms / 1e6 / peak, 3), "checksum": f"{ib.checksum(y):016x}"}), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        for name in PLANS:
            for mb in os.environ.get("SWEEP_MBS", "0 8 16 24 32 48 64").split():
                subprocess.run([sys.executable, __file__, name], env=dict(os.environ, INTFFT_GROUP_MB=mb))
