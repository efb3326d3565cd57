#!/usr/bin/env python
"""Summarises build/obj/*.ptxas.log: demangled kernel name, registers, spills, static smem.
usage: python tools/ptxas_summary.py [substring ...]"""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    pats = sys.argv[1:]
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "build", "obj", "*.ptxas.log"))):
        lines = open(log).read().splitlines()
        name = None
        spill = ""
        for ln in lines:
            m = re.search(r"Compiling entry function '([^']+)'", ln)
            if m:
                name = m.group(1)
                spill = ""
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and name:
                spill = f"stack {m.group(1)} spill {m.group(2)}/{m.group(3)}"
            m = re.search(r"Used (\d+) registers", ln)
            if m and name:
                rows.append((name, int(m.group(1)), spill))
                name = None
    if not rows:
        return
    dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
    for (_, regs, spill), d in zip(rows, dem):
        d = d.replace("nxs::", "")
        if pats and not all(p in d for p in pats):
            continue
        print(f"{regs:4d}  {spill:28s} {d[:200]}")


if __name__ == "__main__":
    main()
