"""Summarise an ncu --set full report: per-stall-reason totals and the hottest SASS lines.
usage: python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: 0 for h in stall_cols}
lines = []
nsamp = 0
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    try:
        n = int(r[idx["# Samples"]])
    except ValueError:
        continue
    nsamp += n
    for h in stall_cols:
        tot[h] += int(r[idx[h]] or 0)
    lines.append((n, r[idx["Source"]].strip(), {h: int(r[idx[h]] or 0) for h in stall_cols if int(r[idx[h]] or 0)}))
print("kernel:", rows[0][1], " samples:", nsamp)
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:28s} {v:8d}  {100.0 * v / max(nsamp, 1):5.1f}%")
lines.sort(key=lambda x: -x[0])
for n, s, d in lines[:topn]:
    top = sorted(d.items(), key=lambda kv: -kv[1])[:2]
    print(f"{n:7d} {100.0 * n / max(nsamp, 1):5.1f}%  {s[:70]:70s} {top}")
