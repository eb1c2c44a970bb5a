#!/usr/bin/env python
"""Per-CUDA-source-line instruction counts and stall samples from an .ncu-rep.
usage: ncu_source_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            ix = hdr.index("Instructions Executed")
            sx = hdr.index("# Samples")
        elif hdr and r and r[0] not in ("", "Function Name"):
            try:
                lines.append((int(r[ix]), int(r[sx]), cur_file, r[0], r[1].strip()[:100]))
            except ValueError:
                pass
    tot = sum(l[0] for l in lines) or 1
    tots = sum(l[1] for l in lines) or 1
    print("kernel %s: %d warp-instructions, %d samples" % (kernel, tot, tots))
    for n, s, f, ln, src in sorted(lines, reverse=True)[:top]:
        print("%11d %5.1f%% inst %5.1f%% smp  %s:%s  %s" % (n, 100.0 * n / tot, 100.0 * s / tots, f, ln, src))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
