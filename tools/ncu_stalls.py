#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an .ncu-rep (read here, no GPU).
usage: python tools/ncu_stalls.py x.ncu-rep [stall_column] [n]"""
import csv, io, signal, subprocess, sys
signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` ends the listing quietly
rep = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "# Samples"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, iex, ism = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "# Samples"))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
seen, d = set(), []
for r in rows[2:]:
    if len(r) > iex and r[iex].isdigit() and r[ia] not in seen:
        seen.add(r[ia]); d.append(r)
base = int(d[0][ia], 16)
tot = sum(int(r[ism]) for r in d)
print("samples", tot, {k[6:]: round(100.0 * sum(int(r[hdr.index(k)]) for r in d) / tot, 1) for k in stalls})
print("active cycles spread / raw metrics: see ncu_summary.py")
ic = hdr.index(col)
for r in sorted(d, key=lambda r: -int(r[ic]))[:n]:  # (pipe into head: a closed pipe ends the listing)
    st = {k[6:]: int(r[hdr.index(k)]) for k in stalls if int(r[hdr.index(k)]) > 0.15 * max(1, int(r[ism]))}
    print("%#07x %5d %-64s %s" % (int(r[ia], 16) - base, int(r[ism]), r[isrc].strip()[:64], st))
