#!/usr/bin/env python
"""Summarise nvcc -Xptxas -v logs: kernel, registers, spills, stack (run from aggfly_b200/csrc)."""
import glob, re, subprocess, sys
rows = []
for f in sorted(glob.glob("*.ptxas.log")):
    name = None
    txt = open(f).read().splitlines()
    for i, line in enumerate(txt):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
            blob = " ".join(txt[i:i + 6])
            regs = re.search(r"Used (\d+) registers", blob)
            spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blob)
            stack = re.search(r"(\d+) bytes stack frame", blob)
            rows.append((f.replace(".ptxas.log", ""), name, int(regs.group(1)) if regs else -1,
                         int(spill.group(1)) if spill else -1, int(stack.group(1)) if stack else -1))
try:
    dem = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
except Exception:
    dem = [r[1] for r in rows]
for r, d in zip(rows, dem):
    d = re.sub(r"\(agf::K1Params.*", "", d).replace("void agf::", "")
    print(f"{r[0]:26s} {d:60s} regs={r[2]:3d} spill={r[3]:3d} stack={r[4]:3d}")
