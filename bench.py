#!/usr/bin/env python
"""bench.py — headline benchmark of intfftk_b200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5|c5u] [--impl ours|reference]

A "step" is one pass of the hot path (one `intfft_exec`) over one batch of synthetic frames that is
already resident in HBM.  Default workload = BASELINE.json configs[1] ("c2"): 4096-pt 16-bit scaled
DIF FFT, batch 65536 per GPU.  Multi-GPU = batch split, one process per GPU, no data-path collective
(frames are independent), so scaling is "weak": every rank runs the full c2 batch.

Output: ONE JSON line on rank 0 (metric Msamples/s, roofline, cpu_baseline, e2e, clocks, ...).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) instead —
the reference itself is VHDL + Octave and cannot run here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 0x696E7466  # "intf"

# name -> (generics kwargs, direction, batch per GPU, description)
CONFIGS = {
    "c2": (dict(NFFT=12, DATA_WIDTH=16, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0, XSER="NEW"), 0, 65536,
           "c2: 4096-pt 16-bit scaled DIF FFT, batch=65536"),
    "c3": (dict(NFFT=16, DATA_WIDTH=24, TWDL_WIDTH=16, FORMAT=1, RNDMODE=0, XSER="NEW"), 0, 4096,
           "c3: 65536-pt 24-bit unscaled FFT, batch=4096"),
    "c4": (dict(NFFT=20, DATA_WIDTH=16, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0, XSER="NEW"), 0, 256,
           "c4: 1048576-pt 16-bit scaled FFT (Taylor twiddles), batch=256"),
    "c5": (dict(NFFT=13, DATA_WIDTH=18, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0, XSER="NEW"), 1, 131072,
           "c5: 8192-pt 18-bit scaled DIT IFFT, batch=131072 per GPU (1M over 8)"),
    "c5u": (dict(NFFT=13, DATA_WIDTH=18, TWDL_WIDTH=16, FORMAT=1, RNDMODE=0, XSER="NEW"), 1, 131072,
            "c5u: 8192-pt 18-bit unscaled DIT IFFT, batch=131072 per GPU"),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def known_traffic(config: str):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(config)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [c.strip() for c in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm restated in C (oracle/), all host threads, bounded sample."""
    if rank != 0:
        return
    import numpy as np
    from oracle import c_oracle as co
    gk, direction, batch, desc = CONFIGS[args.config]
    n = 1 << gk["NFFT"]
    og = co.generics(gk["NFFT"], gk["DATA_WIDTH"], gk["TWDL_WIDTH"], gk["FORMAT"], gk["RNDMODE"],
                     1 if gk["XSER"] == "NEW" else 0, 1, direction)
    cores = os.cpu_count() or 1
    # bounded sample: about 1.5 CPU-seconds of work per step per core-second available
    frames = max(cores, min(batch, int(2.0e7 * cores / 8 / n) or 1))
    x = co.fill_random(frames * n * 2, gk["DATA_WIDTH"], SEED).reshape(frames, n, 2)
    for _ in range(args.warmup):
        co.batch(og, x, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.batch(og, x, 0)
    dt = time.perf_counter() - t0
    value = frames * n * args.steps / dt / 1e6
    sample = f"{frames} frames of {n} points per step (bounded sample of the {batch}-frame batch)"
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (VHDL + Octave) cannot run here; this is its CPU restatement oracle/intfft_oracle.c",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(config: str):
    import numpy as np
    from oracle import c_oracle as co
    gk, direction, batch, _ = CONFIGS[config]
    n = 1 << gk["NFFT"]
    og = co.generics(gk["NFFT"], gk["DATA_WIDTH"], gk["TWDL_WIDTH"], gk["FORMAT"], gk["RNDMODE"],
                     1 if gk["XSER"] == "NEW" else 0, 1, direction)
    cores = os.cpu_count() or 1
    frames = max(cores, min(batch, int(1.6e8 / n)))         # ~10-15 CPU-seconds in total
    x = co.fill_random(frames * n * 2, gk["DATA_WIDTH"], SEED).reshape(frames, n, 2)
    co.batch(og, x[: max(1, frames // 16)], 0)
    t0 = time.perf_counter()
    used = cores
    co.batch(og, x, 0)
    dt = time.perf_counter() - t0
    return {"value": frames * n / dt / 1e6, "unit": "Msamples/s", "cores": used, "kind": "port",
            "sample": f"{frames} frames of {n} points, oracle/intfft_oracle.c on {used} threads, {dt:.2f} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20
        if args.warmup > 3:
            args.warmup = 3
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import intfftk_b200 as ib

    ib.lib()  # fails loudly when the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    gk, direction, batch, desc = CONFIGS[args.config]
    g = ib.Generics(**gk)
    n = 1 << g.NFFT
    core = ib.Core(g, batch, direction, device=local)
    lay = core.layout
    d_in, d_out = core.new_input(), core.new_output()
    ib.fill_random(d_in, g.DATA_WIDTH, SEED + rank)       # full-scale uniform, generated on the device
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-only: inputs resident in HBM ----------------
    for _ in range(args.warmup):
        core.exec(d_in, d_out)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    l0 = ib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        core.exec(d_in, d_out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ib.launch_count() - l0
    checksum = ib.checksum(d_out)

    # ---------------- end to end: pinned host buffers through intfft_exec_host ----------------
    h_in = torch.empty((batch, n, 2), dtype=d_in.dtype, pin_memory=True)
    h_out = torch.empty((batch, n, 2), dtype=d_out.dtype, pin_memory=True)
    h_in.copy_(d_in)
    torch.cuda.synchronize()
    core.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr())          # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        core.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr())      # synchronous: H2D + exec + D2H
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    e2e_ok = bool(torch.equal(h_out[:64], d_out[:64].cpu()))
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        samples_per_step = world * batch * n
        value = samples_per_step * args.steps / (ms * 1e-3) / 1e6
        e2e_value = samples_per_step * args.e2e_steps / (e2e_ms * 1e-3) / 1e6
        peak, peak_src = measured_peak()
        alg_bytes = batch * n * 2 * (lay.in_scalar_bytes + lay.out_scalar_bytes)   # per launch-set of one step
        per_launch_ms = ms / args.steps
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_launch_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32" if lay.lane_bits == 32 else "int64",
            "data": "synthetic",
            "config": {"workload": desc, "generics": gk, "direction": "DIF" if direction == 0 else "DIT",
                       "batch_per_gpu": batch, "parallelism": f"batch-split x{world}",
                       "l2": f"inputs larger than L2 ({lay.in_bytes >> 20} MiB in + {lay.out_bytes >> 20} MiB out per step)",
                       "kernels_per_step": lay.n_passes, "kernel_chain": ib.describe(g, batch, direction)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": known_traffic(args.config), "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg_bytes},
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": int(lay.in_bytes),
                    "d2h_bytes_per_step": int(lay.out_bytes), "steps": args.e2e_steps, "verified": e2e_ok},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "out_checksum": f"{checksum:016x}",
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.config)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
