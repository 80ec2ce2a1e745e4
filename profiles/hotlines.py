#!/usr/bin/env python
"""Per-source-line warp-stall shares for one kernel: joins the SASS page of an .ncu-rep (address, # samples) with
nvdisasm -g line info of the SAME build's cubin.
usage: python profiles/hotlines.py <file.ncu-rep> <kernel-substring> [top_n] >> profiles/<summary>.md"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "stt_b200", "libstt_b200.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.startswith("engine")][0]
sass = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate the function body
start = None
for i, l in enumerate(sass):
    if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"):
        start = i
        break
assert start is not None, "kernel not found in cubin"
off2line = {}
cur = ("?", 0)
for l in sass[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern.split("ILi")[0]], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r][0]
h = rows[hi]
ai, si, ii = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for r in rows[hi + 1:]:
    try:
        a = int(r[ai], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    key = off2line.get(a - base, ("?", 0))
    agg[key][0] += float(r[si] or 0)
    agg[key][1] += float(r[ii] or 0)
    for i, n in stall_cols:
        try:
            agg[key][2][n] += float(r[i] or 0)
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1.0
src_cache = {}
def src(f, n):
    p = os.path.join(ROOT, "stt_b200", "csrc", f)
    if p not in src_cache:
        src_cache[p] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[p]
    return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
print("\n### hottest source lines of `%s` (share of warp-stall samples; dominant stall reasons)\n\n```" % kern)
for (f, n), (smp, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ",".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(smp, 1)) for k, v in st.most_common(2))
    print("%5.1f%%  %-16s %-28s %s" % (100 * smp / tot, "%s:%d" % (f, n), reasons, src(f, n)))
print("```")
