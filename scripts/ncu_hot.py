#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report: python scripts/ncu_hot.py <rep> [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = [r for r in rows[hi + 1:] if len(r) > si and r[si].strip().isdigit()]
S = sum(int(r[si]) for r in data) or 1
print(f"total samples {S}, instructions {len(data)}")
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][si]))[:n]:
    top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print(f"{100*int(r[si])/S:5.1f}%  #{idx:4d} exec={r[ei]:>9}  {r[1][:70]:70s} {top}")
