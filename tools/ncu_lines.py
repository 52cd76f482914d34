#!/usr/bin/env python3
"""tools/ncu_lines.py REPORT.ncu-rep [kernel-substring] -- per-source-line instruction counts / stall samples from an
ncu report captured with --import-source on (needs -lineinfo). Reads `ncu --page source --print-source cuda,sass --csv`."""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, cur_fn, hdr, agg, seen_fn = None, None, None, {}, set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr and r[0].isdigit() and (want in (cur_fn or "")):
        ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:110]])
        num = lambda x: int(x) if x.lstrip("-").isdigit() and x != "-" else 0
        a[0] += num(r[ie]); a[1] += num(r[te]); a[2] += num(r[sm])
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[2] for a in agg.values()) or 1
print("total warp instructions %d, samples %d" % (tot, tots))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%5.1f%% inst %5.1f%% smp  lanes %4.1f  %s:%d  %s" % (100 * a[0] / tot, 100 * a[2] / tots, a[1] / max(a[0], 1), f, ln, a[3]))
