#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from the GPU box in gpurun_out/) into the small text summary that is
committed under profiles/: per-launch DRAM bytes, durations, pipe utilisation, occupancy, stall mix.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01/name.summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# summary of {rep} ({len(data)} launch(es) captured with ncu --set full --clock-control none)"]
    for li, r in enumerate(data):
        lines.append(f"\n## launch {li}: {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
        for k in KEYS:
            if k in col:
                lines.append(f"{k:75s} {r[col[k]]:>16s} {units[col[k]]}")
        if "dram__bytes_read.sum" in col:
            def gb(v, u):
                v = float(v)
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            t = gb(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                gb(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            lines.append(f"{'traffic = dram read + write (bytes per launch)':75s} {t:16.0f} byte")
        lines.append("stall mix (warps per issue-active cycle):")
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[col[h]])
                except ValueError:
                    continue
                if v >= 0.03:
                    lines.append(f"    {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.3f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
