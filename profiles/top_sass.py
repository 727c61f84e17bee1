#!/usr/bin/env python
"""stdin: `ncu --page source --csv --print-source sass` of one report. stdout: per captured kernel, the totals (stall
samples, instructions, shared-memory wavefronts and the excessive ones) and the 30 SASS instructions with the most stall
samples / the most excessive shared-memory wavefronts -- the part of a 10-20 MB report worth keeping."""
import csv
import io
import sys

txt = sys.stdin.read()
seen = set()
for part in txt.split('"Kernel Name",')[1:]:
    lines = part.split("\n")
    name = lines[0].strip().strip('",')
    if name in seen:
        continue
    seen.add(name)
    rdr = csv.reader(io.StringIO("\n".join(lines[1:])))
    try:
        hdr = next(rdr)
    except StopIteration:
        continue
    rows = [r for r in rdr if len(r) == len(hdr)]
    H = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        try:
            return float(r[H[k]])
        except Exception:
            return 0.0
    print("==", name)
    print("   stall samples %d | warp instructions %d | shared wavefronts %d, excessive %d" % (
        sum(num(r, "# Samples") for r in rows), sum(num(r, "Instructions Executed") for r in rows),
        sum(num(r, "L1 Wavefronts Shared") for r in rows), sum(num(r, "L1 Wavefronts Shared Excessive") for r in rows)))
    print("   -- most stall samples (samples, executions, SASS)")
    for r in sorted(rows, key=lambda r: -num(r, "# Samples"))[:30]:
        print("   %8d %10d  %s" % (num(r, "# Samples"), num(r, "Instructions Executed"), r[H["Source"]].strip()[:110]))
    exc = [r for r in rows if num(r, "L1 Wavefronts Shared Excessive") > 0]
    if exc:
        print("   -- excessive shared-memory wavefronts (excessive, total, SASS)")
        for r in sorted(exc, key=lambda r: -num(r, "L1 Wavefronts Shared Excessive"))[:10]:
            print("   %10d %10d  %s" % (num(r, "L1 Wavefronts Shared Excessive"), num(r, "L1 Wavefronts Shared"), r[H["Source"]].strip()[:110]))
