#!/usr/bin/env python
"""Hot SASS of one kernel from an .ncu-rep (source page): python tools/ncu_hot.py rep kernel-regex [min_frac]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several kernels may follow each other: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ia, isrc, iin, ist, ith = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed",
                                                  "Warp Stall Sampling (All Samples)", "Avg. Threads Executed"))
data = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    try:
        data.append((r[ia], r[isrc], int(r[iin]), int(r[ist]), float(r[ith] or 0)))
    except Exception:
        pass
tot = sum(d[2] for d in data); tots = max(1, sum(d[3] for d in data)); mx = max(d[2] for d in data)
print(f"# {kern}: {len(data)} SASS instrs, {tot/1e6:.1f}M warp-instrs executed, {tots} stall samples")
for d in data:
    if d[2] > frac * mx:
        print(f"{d[0][-5:]} {d[2]/1e6:8.1f}M st{100*d[3]/tots:5.1f}% thr{d[4]:5.1f}  {d[1][:100]}")
