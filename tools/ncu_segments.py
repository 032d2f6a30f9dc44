"""Split a kernel's SASS (ncu --page source --csv) at BAR.SYNC and report executed instructions / stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
si = H.index('Source'); st = H.index('# Samples'); ei = H.index('Instructions Executed')
bars = [i for i,r in enumerate(rows[2:]) if 'BAR.SYNC' in r[si]]
prev = 0
tot = sum(int(r[st]) for r in rows[2:] if r[st].isdigit())
print("total samples", tot)
for bpos in bars + [len(rows)-3]:
    seg = rows[2+prev:2+bpos+1]
    inst = sum(int(r[ei]) for r in seg if r[ei].isdigit())
    smp = sum(int(r[st]) for r in seg if r[st].isdigit())
    if smp * 200 > tot:
        top = sorted(((int(r[st]), r[si].strip()[:46]) for r in seg if r[st].isdigit()), reverse=True)[:4]
        print(f"seg [{prev},{bpos}] static {len(seg)} exec {inst} samples {smp} ({100*smp/tot:.1f}%) top {top}")
    prev = bpos+1
