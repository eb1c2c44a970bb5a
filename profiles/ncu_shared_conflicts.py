#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line from an .ncu-rep: actual, ideal, excess.
usage: ncu_shared_conflicts.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=20):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, cur, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            wx, ix, ex = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("L1 Wavefronts Shared Excessive")
        elif hdr and r and r[0] not in ("", "Function Name"):
            try:
                a = agg.setdefault((cur, r[0], r[1].strip()[:90]), [0, 0, 0])
                a[0] += int(r[wx]); a[1] += int(r[ix]); a[2] += int(r[ex])
            except (ValueError, IndexError):
                pass
    tot = sum(a[0] for a in agg.values()) or 1
    print("kernel %s: %d shared wavefronts, %d excessive" % (kernel, tot, sum(a[2] for a in agg.values())))
    for (f, ln, src), (w, i, e) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        if e:
            print("%10d excess of %10d (ideal %10d)  %s:%s  %s" % (e, w, i, f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 20)
