#!/usr/bin/env python
"""bench.py — headline benchmark of intfftk_b200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c3|c4|c5|c5u] [--impl ours|reference]
                    [--also c3,c4,c5,c5u | none] [--no-cpu-baseline]

A "step" is one pass of the hot path (one `intfft_exec`) over one batch of synthetic frames that is
already resident in HBM.  Headline workload (`value`, `roofline`, `e2e`) = BASELINE.json configs[1]
("c2"): 4096-pt 16-bit scaled DIF FFT, batch 65536 per GPU.  The other BASELINE configurations are
measured in the same run and reported under `also` (c3, c4, c5 and c5's UNSCALED variant), c5 with
the BASELINE multi-GPU batch: 2^20 frames sharded over the ranks (capped at 2^18 frames per rank so
that one GPU holds it comfortably; the cap is stated in the entry).  Multi-GPU = batch split, one process
per GPU, no data-path collective (frames are independent); the only collectives are the barrier, the
max-over-ranks of the elapsed times and the sum of the shard checksums (intfftk_b200/sharding.py).

Every measured configuration is checked against the CPU oracle on sampled frames (first / last frames
of the rank's shard, all samples) OUTSIDE the timed region; a mismatch aborts the run.

Output: ONE JSON line on rank 0 (metric Msamples/s, roofline, cpu_baseline, e2e, clocks, also, ...).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) instead —
the reference itself is VHDL + Octave and cannot run here (no simulator / Octave on the
GPU box either: profiles/r02/tool_probe_r02.txt).
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

# name -> (generics kwargs, direction, frames per GPU (None: sharded job, see job_frames), description)
CONFIGS = {
    "c2": (dict(NFFT=12, DATA_WIDTH=16, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0, XSER="NEW"), 0, 65536,
           "c2: 4096-pt 16-bit scaled DIF FFT, batch=65536"),
    "c3": (dict(NFFT=16, DATA_WIDTH=24, TWDL_WIDTH=16, FORMAT=1, RNDMODE=0, XSER="NEW"), 0, 4096,
           "c3: 65536-pt 24-bit unscaled FFT, batch=4096"),
    "c4": (dict(NFFT=20, DATA_WIDTH=16, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0, XSER="NEW"), 0, 256,
           "c4: 1048576-pt 16-bit scaled FFT (Taylor twiddles), batch=256"),
    "c5": (dict(NFFT=13, DATA_WIDTH=18, TWDL_WIDTH=16, FORMAT=0, RNDMODE=0, XSER="NEW"), 1, None,
           "c5: 8192-pt 18-bit scaled DIT IFFT, batch=1M sharded across the ranks"),
    "c5u": (dict(NFFT=13, DATA_WIDTH=18, TWDL_WIDTH=16, FORMAT=1, RNDMODE=0, XSER="NEW"), 1, None,
            "c5u: 8192-pt 18-bit unscaled DIT IFFT, batch=1M sharded across the ranks"),
}
C5_JOB_FRAMES = 1 << 20       # BASELINE.json configs[4]: batch = 1M over 8 GPUs
C5_RANK_CAP = 1 << 18         # at most 2^18 frames (2 x 17 GB) per rank


def frames_for(config: str, rank: int, world: int):
    """(frames of this rank, note) — fixed per-GPU batches for c2..c4 (weak scaling), the 1M-frame job for c5."""
    from intfftk_b200.sharding import shard_range
    per = CONFIGS[config][2]
    if per is not None:
        return per, f"{per} frames per GPU"
    lo, hi = shard_range(C5_JOB_FRAMES, rank, world)
    n = hi - lo
    if n > C5_RANK_CAP:
        return C5_RANK_CAP, (f"{C5_JOB_FRAMES} frames / {world} rank(s) = {n} per rank, capped at {C5_RANK_CAP} per rank "
                             f"(HBM budget of this bench; frames are independent, so the cap changes the run time only)")
    return n, f"{C5_JOB_FRAMES} frames / {world} ranks = {n} per rank (BASELINE multi-GPU batch)"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def known_traffic(config: str):
    """dram bytes per step of the config's kernels from the committed ncu capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        return d.get(config), d.get("_source", "ncu capture under profiles/")
    except Exception:
        return None, None


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


def oracle_generics(gk, direction):
    from oracle import c_oracle as co
    return co.generics(gk["NFFT"], gk["DATA_WIDTH"], gk["TWDL_WIDTH"], gk["FORMAT"], gk["RNDMODE"],
                       1 if gk["XSER"] == "NEW" else 0, 1, direction)


def config_block(config: str, world: int, frames: int, note: str):
    """The `config` object: identical keys in both arms (`--impl ours` and `--impl reference`)."""
    gk, direction, _, desc = CONFIGS[config]
    n = 1 << gk["NFFT"]
    in_sb = 2 if gk["DATA_WIDTH"] <= 16 else (4 if gk["DATA_WIDTH"] <= 32 else 8)
    ow = gk["DATA_WIDTH"] + gk["FORMAT"] * gk["NFFT"]
    out_sb = 2 if ow <= 16 else (4 if ow <= 32 else 8)
    return {"workload": desc, "generics": gk, "direction": "DIF" if direction == 0 else "DIT",
            "batch_per_gpu": frames, "batch_note": note, "parallelism": f"batch-split x{world}",
            "l2": f"inputs larger than L2 ({frames * n * 2 * in_sb >> 20} MiB in + {frames * n * 2 * out_sb >> 20} MiB out per step)"}


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm restated in C (oracle/), all host threads, bounded sample."""
    if rank != 0:
        return
    from oracle import c_oracle as co
    gk, direction, _, desc = CONFIGS[args.config]
    n = 1 << gk["NFFT"]
    og = oracle_generics(gk, direction)
    cores = os.cpu_count() or 1
    batch, note = frames_for(args.config, 0, world)
    # bounded sample: the whole run (warm-up + timed steps) should take about a minute of wall clock
    probe = max(cores, int(4e6 * cores / 8 / n) or 1)
    x = co.fill_random(probe * n * 2, gk["DATA_WIDTH"], SEED).reshape(probe, n, 2)
    co.batch(og, x, 0)
    t0 = time.perf_counter()
    co.batch(og, x, 0)
    rate = probe * n / (time.perf_counter() - t0)                        # samples / s on this box
    frames = int(rate * 60.0 / (args.steps + args.warmup) / n)
    frames = max(cores, min(batch, frames))
    x = co.fill_random(frames * n * 2, gk["DATA_WIDTH"], SEED).reshape(frames, n, 2)
    for _ in range(args.warmup):
        co.batch(og, x, 0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.batch(og, x, 0)
    dt = time.perf_counter() - t0
    value = frames * n * args.steps / dt / 1e6
    sample = f"{frames} frames of {n} points per step (bounded sample of the {batch}-frame batch), oracle/intfft_oracle.c on {cores} threads"
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": config_block(args.config, world, batch, note),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference (VHDL + Octave) cannot run here or on the GPU box; this is its CPU restatement oracle/intfft_oracle.c",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(config: str):
    from oracle import c_oracle as co
    gk, direction, _, _ = CONFIGS[config]
    n = 1 << gk["NFFT"]
    og = oracle_generics(gk, direction)
    cores = os.cpu_count() or 1
    frames = max(cores, int(1.6e8 / n))                      # ~10-15 CPU-seconds in total
    x = co.fill_random(frames * n * 2, gk["DATA_WIDTH"], SEED).reshape(frames, n, 2)
    co.batch(og, x[: max(1, frames // 16)], 0)
    t0 = time.perf_counter()
    co.batch(og, x, 0)
    dt = time.perf_counter() - t0
    out = {"value": frames * n / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
           "sample": f"{frames} frames of {n} points, oracle/intfft_oracle.c on {cores} threads, {dt:.2f} s"}
    out["c1"] = fn_radix2_baseline()
    return out


def fn_radix2_baseline():
    """BASELINE.json configs[0]: 1024-pt FFT through math/fn_radix2.m.  Octave exists neither in the build image
    nor on the GPU box (profiles/r02/tool_probe_r02.txt), so the line-for-line NumPy restatement is timed."""
    import shutil
    import numpy as np
    from oracle import fn_radix2 as fr
    rng = np.random.default_rng(1)
    din = np.round(rng.uniform(-32767, 32767, 1024)) + 1j * np.round(rng.uniform(-32767, 32767, 1024))
    fr.fn_radix2(din, 1024, "FWD")
    calls = 20
    t0 = time.perf_counter()
    for _ in range(calls):
        fr.fn_radix2(din, 1024, "FWD")
    dt = (time.perf_counter() - t0) / calls
    return {"workload": "c1: 1024-pt FFT via math/fn_radix2.m (double-precision structural model), single transform",
            "ms_per_transform": dt * 1e3, "value": 1024 / dt / 1e6, "unit": "Msamples/s", "cores": 1,
            "kind": "restatement — Octave unavailable" if not shutil.which("octave") else "restatement (octave present but the script needs the signal package)",
            "sample": f"{calls} calls of oracle/fn_radix2.py: fn_radix2(Din, 1024, 'FWD')"}


# ------------------------------------------------------------------------------------------------------------
class Ctx:
    pass


def measure(ctx, config: str, steps: int, warmup: int, sampler=None):
    """One configuration on this rank: timed steps (CUDA events on the launch stream, barrier + synchronize on both
    sides, max over ranks), shard checksums summed over ranks, sampled oracle check outside the timed region."""
    import numpy as np
    import torch
    import intfftk_b200 as ib
    from intfftk_b200.sharding import reduce_report
    from oracle import c_oracle as co
    dist, rank, world, local = ctx.dist, ctx.rank, ctx.world, ctx.local
    gk, direction, _, desc = CONFIGS[config]
    g = ib.Generics(**gk)
    n = 1 << g.NFFT
    frames, note = frames_for(config, rank, world)
    core = ib.Core(g, frames, direction, device=local)
    lay = core.layout
    d_in, d_out = core.new_input(), core.new_output()
    ib.fill_random(d_in, g.DATA_WIDTH, SEED + rank)       # full-scale uniform, generated on the device
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        core.exec(d_in, d_out)
    barrier()
    if sampler is not None:
        sampler.start()
        time.sleep(0.15)
    l0 = ib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        core.exec(d_in, d_out)
    e1.record()
    barrier()
    ms_local = e0.elapsed_time(e1)
    launches = ib.launch_count() - l0
    rep = reduce_report(ms_local, frames * n, ib.checksum(d_out), dist, device=torch.device("cuda", local))

    # ---- parity: sampled frames against the CPU oracle (outside the timed region) ----
    k = min(32, frames)
    idx = list(range(k)) + [f for f in range(frames - k, frames) if f >= k]
    sel = torch.tensor(idx, device=d_in.device)
    x = d_in.index_select(0, sel).cpu().numpy()
    got = d_out.index_select(0, sel).cpu().numpy()
    want = co.batch(oracle_generics(gk, direction), x, 0)
    if not np.array_equal(got, want):
        raise SystemExit(f"bench.py: {config}: GPU result differs from the oracle on the sampled frames (rank {rank})")
    parity_ok = torch.tensor([1], device=d_in.device)
    if dist is not None:
        dist.all_reduce(parity_ok, op=dist.ReduceOp.MIN)

    peak, peak_src = measured_peak()
    alg_bytes = frames * n * 2 * (lay.in_scalar_bytes + lay.out_scalar_bytes)      # this rank, one step
    per_step_ms = rep.ms / steps
    achieved = alg_bytes / (per_step_ms * 1e-3) / 1e9
    traffic, traffic_src = known_traffic(config)
    if traffic is not None and CONFIGS[config][2] is None:
        traffic = int(traffic * frames / (1 << 17))      # the capture was taken with 2^17 frames
    res = {
        "value": rep.samples * steps / (rep.ms * 1e-3) / 1e6, "ms_per_step": per_step_ms, "steps": steps,
        "config": config_block(config, world, frames, note),           # same keys and values in both arms
        "kernels": {"per_step": lay.n_passes, "chain": ib.describe(g, frames, direction)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_step": alg_bytes},
        "parity": f"bit-exact vs oracle/intfft_oracle.c, {len(idx)} sampled frames per rank x {world} rank(s)",
        "checksum_of_shard_checksums": f"{rep.checksum:016x}",
        "samples_per_step_all_ranks": rep.samples,
        "gpu_launches": int(launches),
        "dtype": "int32" if lay.lane_bits == 32 else "int64",
    }
    return res, core, d_in, d_out


def end_to_end(ctx, core, d_in, d_out, e2e_steps: int):
    """Same metric through the host-buffer C-ABI call (intfft_exec_host): pinned host buffers, H2D + kernels + D2H
    inside the timed region; then the bare concurrent H2D + D2H copies of the same bytes as the ceiling of this box."""
    import torch
    dist, rank, world = ctx.dist, ctx.rank, ctx.world
    lay = core.layout
    n, batch = core.n, core.batch

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    h_in = torch.empty((batch, n, 2), dtype=d_in.dtype, pin_memory=True)
    h_out = torch.empty((batch, n, 2), dtype=d_out.dtype, pin_memory=True)
    h_in.copy_(d_in)
    torch.cuda.synchronize()
    core.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr())          # warm-up (allocates the staging ring)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        core.exec_host_ptr(h_in.data_ptr(), h_out.data_ptr())      # synchronous: H2D + exec + D2H
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    ok = bool(torch.equal(h_out[:64], d_out[:64].cpu()) and torch.equal(h_out[-64:], d_out[-64:].cpu()))
    # bare copies, both directions at once on two streams: what the PCIe link (and, with several ranks, the shared
    # host side) allows for these byte counts
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    copy_ms = (time.perf_counter() - t0) * 1e3
    # restore the device input (the probe overwrote it with the identical bytes) — nothing to do
    t = torch.tensor([e2e_ms, copy_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms, copy_ms = float(t[0]), float(t[1])
    samples = world * batch * n
    return {"value": samples * e2e_steps / (e2e_ms * 1e-3) / 1e6, "unit": "Msamples/s",
            "h2d_bytes_per_step": int(lay.in_bytes), "d2h_bytes_per_step": int(lay.out_bytes), "steps": e2e_steps,
            "api": "intfft_exec_host (ring of three 32 MiB staging buffers, three streams)",
            "verified": ok,
            "copy_ceiling": {"value": samples * e2e_steps / (copy_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                             "what": "bare concurrent cudaMemcpyAsync H2D + D2H of the same bytes on every rank at once"},
            "frac_of_copy_ceiling": copy_ms / e2e_ms}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--also", default="c3,c4,c5,c5u", help="other BASELINE configurations measured in the same run ('none' to skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

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
    ctx = Ctx()
    ctx.dist, ctx.rank, ctx.world, ctx.local = dist, rank, world, local

    # ---------------- headline configuration: kernel-only, then end to end ----------------
    sampler = ClockSampler(local) if rank == 0 else None
    res, core, d_in, d_out = measure(ctx, args.config, args.steps, args.warmup, sampler)
    e2e = end_to_end(ctx, core, d_in, d_out, args.e2e_steps)
    clocks = sampler.stop() if sampler is not None else None
    core.close()
    del d_in, d_out, core
    torch.cuda.empty_cache()

    # ---------------- the other BASELINE configurations, same run ----------------
    also = {}
    names = [] if args.also.strip().lower() in ("", "none") else [c.strip() for c in args.also.split(",")]
    for name in names:
        if name == args.config or name not in CONFIGS:
            continue
        r, c, a, b = measure(ctx, name, max(5, min(args.steps, 30)), 3)
        c.close()
        del a, b, c
        torch.cuda.empty_cache()
        also[name] = {"value": r["value"], "unit": "Msamples/s", "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                      "frac": r["roofline"]["frac"], "achieved_gb_s": r["roofline"]["achieved"],
                      "traffic": r["roofline"]["traffic"], "algorithmic_bytes_per_step": r["roofline"]["algorithmic_bytes_per_step"],
                      "workload": r["config"]["workload"], "batch_per_gpu": r["config"]["batch_per_gpu"],
                      "batch_note": r["config"]["batch_note"], "kernel_chain": r["kernels"]["chain"],
                      "parity": r["parity"], "checksum_of_shard_checksums": r["checksum_of_shard_checksums"],
                      "gpu_launches": r["gpu_launches"], "dtype": r["dtype"]}

    if rank == 0:
        line = {
            "metric": "Msamples/s", "value": res["value"], "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": res["dtype"], "data": "synthetic",
            "config": res["config"], "kernels": res["kernels"], "roofline": res["roofline"], "e2e": e2e,
            "gpu_launches": res["gpu_launches"], "clocks": clocks,
            "parity": res["parity"], "checksum_of_shard_checksums": res["checksum_of_shard_checksums"],
            "also": also,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.config)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
