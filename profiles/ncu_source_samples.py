#!/usr/bin/env python
"""Per-CUDA-source-line stall samples (and instruction counts) from an .ncu-rep, sorted by samples.
usage: ncu_source_samples.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, cur, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ix, sx = hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and r and r[0] not in ("", "Function Name"):
            try:
                a = agg.setdefault((cur, r[0], r[1].strip()[:90]), [0, 0])
                a[0] += int(r[sx])
                a[1] += int(r[ix])
            except ValueError:
                pass
    tot = sum(a[0] for a in agg.values()) or 1
    toti = sum(a[1] for a in agg.values()) or 1
    print("kernel %s: %d warp-instructions, %d samples" % (kernel, toti, tot))
    for (f, ln, src), (smp, ins) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100.0 * smp / tot, 100.0 * ins / toti, f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
