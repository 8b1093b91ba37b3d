#!/usr/bin/env python
"""Per-kernel counts of the Blackwell instructions that show which hardware path a kernel uses (no GPU needed):
UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP / UBLKPF = bulk copy /
bulk L2 prefetch, SYNCS = mbarrier operations.     python scripts/sass_counts.py > profiles/r02_sass_counts.txt"""
import re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "triple_hybrid_rag_b200/lib/libthr.so"
pats = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "UBLKPF", "SYNCS", "LDG.E.64", "LDS", "STS", "ATOMS", "BAR.SYNC"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
rows, name, c, n = [], None, None, 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name:
            rows.append((name, c, n))
        name, c, n = m.group(1), dict.fromkeys(pats, 0), 0
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        n += 1
        for p in pats:
            if p in line:
                c[p] += 1
if name:
    rows.append((name, c, n))
dem = subprocess.run(["c++filt", "-p"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
nv = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {so}   ({re.search(r'release [0-9.]+', nv).group(0)}, -gencode arch=compute_100a,code=sm_100a)")
print(f"{'kernel':56s}" + "".join(f"{p:>9s}" for p in pats) + f"{'insts':>8s}")
for (nm, c, n), d in sorted(zip(rows, dem), key=lambda x: x[1]):
    d = d.replace("(anonymous namespace)::", "")
    print(f"{d[:56]:56s}" + "".join(f"{c[p]:9d}" for p in pats) + f"{n:8d}")
