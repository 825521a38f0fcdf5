#!/usr/bin/env python
"""Opcode histogram of every kernel in the library's objects (cuobjdump -sass): which instructions prove what the kernels
are made of -- UTMALDG / SYNCS (TMA + mbarrier ring), DADD / DMUL / DFMA (fp64 accumulation), F2F (conversions on the XU pipe),
FSETP / FADD (bin counters), LDS / STS (staging), no tensor-core opcodes (nothing on this path is a contraction worth a GEMM).
usage: tools/sass_hist.py [objects...] > profiles/rN_sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTMALDG", "UTMACCTL", "SYNCS", "BAR", "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "F2F", "MUFU", "FSETP", "FADD",
       "FMNMX", "FMNMX3", "SHFL", "CREDUX", "HSET2", "HADD2", "HFMA2", "PRMT", "IDP", "VOTE", "POPC", "IMAD", "LOP3", "ATOM", "RED", "HMMA", "UTCHMMA", "UTCMMA", "DMMA"]


def main():
    objs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "aggfly_b200", "csrc", "*.o")))
    for obj in objs:
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, hist = None, None
        kernels = []
        for line in out.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                name, hist = m.group(1), collections.Counter()
                kernels.append((name, hist))
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and hist is not None:
                hist[m.group(1).split(".")[0]] += 1
        print(f"== {os.path.basename(obj)}: {len(kernels)} kernels")
        for name, hist in kernels:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            total = sum(hist.values())
            keys = " ".join(f"{k}={hist[k]}" for k in KEY if hist.get(k))
            print(f"  {dem[:110]}\n      instructions={total} {keys}")


if __name__ == "__main__":
    main()
