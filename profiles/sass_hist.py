#!/usr/bin/env python
"""Opcode histogram (executed warp instructions) and top stall sites from an ncu report's source page.
  python profiles/sass_hist.py gpurun_out/x.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the export holds one table per kernel launch, each introduced by a "Kernel Name" row
tables, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}; tables.append(cur); continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = row; continue
    cur["rows"].append(row)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t = tables[k]
h = t["hdr"]
iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
iW = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
ops = collections.Counter(); smp = collections.Counter(); wf = collections.Counter()
tot = 0
for r in t["rows"]:
    op = r[iS].split()[0] if not r[iS].lstrip().startswith("@") else r[iS].split()[1]
    op = op.split(".")[0] if not op.startswith(("LDS", "LDG", "STS", "STG", "HMMA", "MUFU")) else ".".join(op.split(".")[:2])
    n = int(r[iE] or 0); ops[op] += n; tot += n; smp[op] += int(r[iSm] or 0)
    if iW is not None:
        wf[op] += int(r[iW] or 0)
print(t["name"][:100]); print("total warp instructions:", tot)
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:10d} {100.0 * n / tot:5.1f}%   samples {smp[op]:7d}   smem wavefronts {wf[op]:9d}")
print("top stall sites:")
for r in sorted(t["rows"], key=lambda r: -int(r[iSm] or 0))[:14]:
    print(f"  {int(r[iSm] or 0):6d}  {r[iS].strip()[:90]}")
