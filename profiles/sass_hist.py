#!/usr/bin/env python
"""Opcode histogram of the SASS of selected kernels (evidence for the per-butterfly instruction
counts quoted in DESIGN.md, and for which copy engines a kernel uses: UBLKCP = 1-D TMA bulk copy,
UTMALDG / UTMASTG = tensor-map TMA load / store, LDGSTS = cp.async).

    python profiles/sass_hist.py <binary-or-.so> <kernel-name-regex> [--butterflies N] [--dump DIR] [--loop]

--butterflies N : also print instructions per butterfly (N = butterflies one thread executes per
                  trip of the kernel's main loop; the whole function body is counted unless --loop)
--loop          : count only the instructions between the LAST backward branch target and that branch
                  (the steady-state frame loop of the persistent kernels)
--dump DIR      : write the full SASS listing of every matching kernel to DIR/<short-name>.sass
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys

PIPE = {  # issue port classes used in DESIGN.md §2.6
    "IMAD": "fma", "IMUL": "fma", "IDP": "fma", "FFMA": "fma", "FMUL": "fma", "FADD": "fma",
    "IADD3": "alu", "IADD": "alu", "LEA": "alu", "SHF": "alu", "LOP3": "alu", "PRMT": "alu", "SGXT": "alu",
    "SEL": "alu", "ISETP": "alu", "IABS": "alu", "IMNMX": "alu", "VIADD": "alu", "VIMNMX": "alu", "BMSK": "alu",
    "FLO": "xu", "POPC": "xu", "BREV": "xu", "MUFU": "xu",
    "LDS": "lsu", "STS": "lsu", "LDG": "lsu", "STG": "lsu", "LDGSTS": "lsu", "LDSM": "lsu", "LD": "lsu", "ST": "lsu",
    "LDL": "lsu", "STL": "lsu", "LDC": "lsu", "ATOMG": "lsu", "RED": "lsu", "ATOMS": "lsu",
    "UBLKCP": "tma", "UTMALDG": "tma", "UTMASTG": "tma", "SYNCS": "tma", "UTMACMDFLUSH": "tma",
    "BAR": "ctl", "BRA": "ctl", "EXIT": "ctl", "WARPSYNC": "ctl", "NOP": "ctl", "BSSY": "ctl", "BSYNC": "ctl",
    "MOV": "alu", "S2R": "xu", "CS2R": "alu", "SHFL": "lsu",
}


def sass_functions(path: str):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    fn, body = None, []
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn:
                yield fn, body
            fn, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", line)
        if m and fn:
            body.append((int(m.group(1), 16), m.group(2)))
    if fn:
        yield fn, body


def demangle(name: str) -> str:
    try:
        dn = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        dn = re.sub(r"\((int|bool)\)", "", dn)                      # fast16_kernel<12, 0, 1, 1, 0, 0>
        return dn.replace("intfft::<unnamed>::", "")
    except FileNotFoundError:
        return name


def opcode(ins: str) -> str:
    t = ins.split()
    if t and t[0].startswith("@"):
        t = t[1:]
    return t[0] if t else "?"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("binary")
    ap.add_argument("pattern")
    ap.add_argument("--butterflies", type=float, default=0)
    ap.add_argument("--loop", action="store_true")
    ap.add_argument("--dump")
    a = ap.parse_args()
    rx = re.compile(a.pattern)
    for fn, body in sass_functions(a.binary):
        dn = demangle(fn)
        if not rx.search(dn):
            continue
        if a.dump:
            os.makedirs(a.dump, exist_ok=True)
            short = re.sub(r"[^A-Za-z0-9_<>,]+", "_", dn.split("(")[0])[-120:]
            with open(os.path.join(a.dump, short + ".sass"), "w") as f:
                f.write(f"// {dn}\n")
                for addr, ins in body:
                    f.write(f"/*{addr:05x}*/ {ins} ;\n")
        sel = body
        if a.loop:
            back = [(addr, ins) for addr, ins in body if opcode(ins) == "BRA" and re.search(r"0x([0-9a-f]+)", ins)
                    and int(re.search(r"0x([0-9a-f]+)", ins).group(1), 16) < addr]
            if back:
                # the widest backward branch = the outer (frame) loop
                addr, ins = max(back, key=lambda x: x[0] - int(re.search(r"0x([0-9a-f]+)", x[1]).group(1), 16))
                tgt = int(re.search(r"0x([0-9a-f]+)", ins).group(1), 16)
                sel = [(ad, i) for ad, i in body if tgt <= ad <= addr]
        full = collections.Counter(opcode(i) for _, i in sel)
        base = collections.Counter()
        for k, v in full.items():
            base[k.split(".")[0]] += v
        pipes = collections.Counter()
        for k, v in base.items():
            pipes[PIPE.get(k, "other")] += v
        total = sum(base.values())
        print(f"== {dn}\n   {total} instructions{' in the main loop' if a.loop else ''}; by port class: "
              + ", ".join(f"{k} {v}" for k, v in pipes.most_common()))
        print("   " + "  ".join(f"{k} {v}" for k, v in full.most_common(28)))
        if a.butterflies:
            print(f"   per butterfly ({a.butterflies:g} per trip): total {total / a.butterflies:.2f}, "
                  + ", ".join(f"{k} {v / a.butterflies:.2f}" for k, v in pipes.most_common()))


if __name__ == "__main__":
    main()
