"""Join the per-instruction samples of an `ncu --set full --import-source on` report with the line table of the
cubin (nvdisasm -g): samples / instructions per device function and per source line.
usage: python tools/ncu_lines.py report.ncu-rep object.o kernel_substring [top_n]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l][0]
ins, cur, fn = [], None, "entry"
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.match(r"\s*(\$?_Z[^:]*):", l)
    if m and "$" in m.group(1):
        fn = m.group(1).split("$")[-1]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((fn, cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
assert len(data) == len(ins), (len(data), len(ins))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
byfn, byline = collections.Counter(), collections.Counter()
fi, li = collections.Counter(), collections.Counter()
fst = collections.defaultdict(collections.Counter)
for (f, c, _), b in zip(ins, data):
    n = int(b[idx["# Samples"]] or 0)
    ie = int(b[idx["Instructions Executed"]] or 0)
    byfn[f] += n; fi[f] += ie; byline[c] += n; li[c] += ie
    for h in stalls:
        fst[f][h] += int(b[idx[h]] or 0)
tot = sum(byfn.values())
print("total samples", tot, "instructions", sum(fi.values()))
for f, n in byfn.most_common(12):
    dem = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()[:100]
    top = ", ".join("%s:%d%%" % (h[6:], 100 * v / max(n, 1)) for h, v in fst[f].most_common(5))
    print("%5.1f%% inst=%10d  %s\n        %s" % (100.0 * n / tot, fi[f], dem, top))
src = {}
for c, n in byline.most_common(topn):
    if not c:
        continue
    if c[0] not in src:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dynaphopy_b200", "csrc", c[0])
        src[c[0]] = open(path).read().split("\n") if os.path.exists(path) else []
    text = src[c[0]][c[1] - 1].strip()[:90] if len(src[c[0]]) >= c[1] else ""
    print("%6d %5.1f%% inst=%9d %s:%d | %s" % (n, 100.0 * n / tot, li[c], c[0], c[1], text))
