"""Registers / spills per kernel from `nvcc -Xptxas=-v` output (usage: python tools/ptxas_summary.py ptxas.log [filter])."""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cur, sp = None, ("?", "?")
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur:
        sp = (m.group(1), m.group(2))
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
        name = name.replace("srukf::", "").split("(")[0]
        if flt in name:
            print(f"{name[:60]:60s} regs {m.group(1):>4s}  spill st/ld {sp[0]}/{sp[1]}")
