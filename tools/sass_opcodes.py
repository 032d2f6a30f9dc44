"""Static SASS opcode counts per kernel of the shipped library -> profiles/<tag>_sass_opcodes.txt
usage: python tools/sass_opcodes.py <out.txt>   (cuobjdump -xelf / -sass on cv_monoslam_b200/libsrukf_b200.so)"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cv_monoslam_b200", "libsrukf_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
res = {}
cubins = sorted(f for f in os.listdir(tmp) if f.endswith(".cubin"))
for cb in cubins:
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(tmp, cb)], capture_output=True, text=True).stdout
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("srukf::", "").split("(")[0]
            res[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur:
            res[cur][m.group(1)] += 1
with open(sys.argv[1], "w") as f:
    f.write("# SASS opcode counts of the shipped library (static; cuobjdump -sass on the cubins of libsrukf_b200.so)\n")
    f.write("# cubins: " + " ".join(cubins) + "  (sm_100a only)\n")
    f.write("# DMMA = FP64 tensor-core MMA (mma.sync.m8n8k4.f64); UTMALDG = TMA tensor load (cp.async.bulk.tensor);\n")
    f.write("# SYNCS = mbarrier operations.  No UTC*MMA / LDTM / UTCBAR: tcgen05 has no FP64 kind.\n\n")
    pref = ("DMMA", "UTMA", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "BAR", "SHFL", "LDS", "STS", "STL", "LDL", "UTC", "LDTM", "LDG", "STG")
    for k, c in sorted(res.items()):
        if "k_" not in k:
            continue
        sel = {kk: v for kk, v in c.items() if kk.startswith(pref)}
        f.write(f"{k}: {sum(c.values())} instructions\n    " + ", ".join(f"{a} {b}" for a, b in sorted(sel.items(), key=lambda kv: -kv[1])) + "\n")
print(open(sys.argv[1]).read()[:2500])
