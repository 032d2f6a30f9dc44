"""Stall samples and executed instructions of one kernel per CUDA source line.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <mangled-function-substring> [top]

The report's SASS page (ncu -i ... --page source --csv --print-source sass) is aligned instruction by instruction with
`nvdisasm -g` of the SAME build's cubin (cuobjdump -xelf all libsrukf_b200.so), whose `//## File .. line N` markers give
the source line of every instruction (-lineinfo build)."""
import csv, re, subprocess, sys
from collections import defaultdict

rep, cubin, fn = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[hi]
si, st, ei = H.index('Source'), H.index('# Samples'), H.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
sass = [r for r in rows[hi + 1:] if len(r) > ei]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith('//-----') and '.text.' in l:
        infn = fn in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        lines.append((cur, m.group(2).strip()))
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} in the report (different build?)")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
tot = 0
for (ln, txt), r in zip(lines, sass):
    op_a, op_b = txt.split()[0].lstrip('@!P0123456789 '), r[si].split()[0] if r[si].split() else ''
    smp = int(r[st]) if r[st].isdigit() else 0
    ex = int(r[ei]) if r[ei].isdigit() else 0
    a = agg[ln]
    a[0] += smp; a[1] += ex
    for i, h in stall_cols:
        if r[i].isdigit():
            a[2][h] += int(r[i])
    tot += smp
import os
CS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'cv_monoslam_b200', 'csrc')
srcs = {}
def text(key):
    if not key:
        return ''
    f, ln = key
    if f not in srcs:
        try:
            srcs[f] = open(os.path.join(CS, f)).read().splitlines()
        except OSError:
            srcs[f] = []
    return srcs[f][ln - 1].strip()[:64] if 0 < ln <= len(srcs[f]) else ''
print(f"total samples {tot}")
for ln, (smp, ex, stl) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ' '.join(f"{k[6:]}:{100 * v // max(smp, 1)}" for k, v in sorted(stl.items(), key=lambda kv: -kv[1])[:3])
    where = f"{ln[0][6:-3] if ln[0].startswith('srukf_') else ln[0]}:{ln[1]}" if ln else '?'
    print(f"{where:14s} {100 * smp / tot:5.1f}% exec {ex:>11d}  [{tops}]  {text(ln)}")
