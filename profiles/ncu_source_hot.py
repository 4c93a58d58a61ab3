#!/usr/bin/env python
"""Hot spots of a kernel from an `ncu --set full --import-source on` capture: per-opcode stall samples and the
instructions that collect the most samples.

    ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv ;  python profiles/ncu_source_hot.py x.csv [top]
"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# a report may hold several kernels: split at "Kernel Name" rows
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr) - 2:
                body.append(rows[j])
            j += 1
        col = {h: k for k, h in enumerate(hdr)}
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = collections.Counter()
        byop = collections.defaultdict(collections.Counter)
        total_samples = 0
        execd = 0
        for r in body:
            ins = r[col["Source"]].split()
            op = ins[1] if ins and ins[0].startswith("@") else (ins[0] if ins else "?")
            n = int(r[col["# Samples"]] or 0)
            total_samples += n
            execd += int(r[col["Instructions Executed"]] or 0)
            byop[op]["samples"] += n
            byop[op]["count"] += 1
            byop[op]["exec"] += int(r[col["Instructions Executed"]] or 0)
            for s in stalls:
                v = int(r[col[s]] or 0)
                tot[s] += v
                byop[op][s] += v
        print(f"== {name[:120]}\n   {len(body)} SASS instructions, {execd} warp-instructions executed, {total_samples} stall samples")
        print("   stall reasons: " + ", ".join(f"{k[6:]} {v * 100 // max(1, total_samples)}%" for k, v in tot.most_common(9)))
        print("   by opcode (samples %, static count, exec %, top stall reasons):")
        for op, c in sorted(byop.items(), key=lambda kv: -kv[1]["samples"])[:14]:
            rs = sorted(((s, c[s]) for s in stalls), key=lambda x: -x[1])[:3]
            print(f"     {op:22s} {c['samples'] * 100 / max(1, total_samples):5.1f}%  n={c['count']:4d}  exec {c['exec'] * 100 / max(1, execd):4.1f}%   "
                  + ", ".join(f"{s[6:]} {v * 100 // max(1, c['samples'])}%" for s, v in rs))
        print(f"   top {top} instructions by samples:")
        for r in sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]:
            rs = sorted(((s, int(r[col[s]] or 0)) for s in stalls), key=lambda x: -x[1])[:2]
            print(f"     {r[col['# Samples']]:>6s}  {r[col['Source']].strip()[:70]:70s} " + ", ".join(f"{s[6:]} {v}" for s, v in rs))
        i = j
    else:
        i += 1
