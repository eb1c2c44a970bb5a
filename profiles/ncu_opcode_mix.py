#!/usr/bin/env python
"""Executed warp-instructions by SASS opcode (and stall samples) from an .ncu-rep source page.
usage: ncu_opcode_mix.py report.ncu-rep kernel_regex [top_n]"""
import csv
import subprocess
import sys
from collections import Counter


def main(path, kernel, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    inst, smp = Counter(), Counter()
    hdr = None
    done = 0
    for r in rows:
        if r and r[0] == "Kernel Name":
            done += 1
            if done > 1:
                break
        if r and r[0] == "Address":
            hdr = r
            ix, sx, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
        elif hdr and len(r) == len(hdr):
            toks = r[src].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = ".".join(op.split(".")[:2]) if op.startswith(("F2F", "I2F", "F2I", "LDS", "STS", "LDG", "STG", "SHFL")) else op.split(".")[0]
            inst[op] += int(r[ix])
            smp[op] += int(r[sx])
    tot, tots = sum(inst.values()) or 1, sum(smp.values()) or 1
    print("kernel %s (first instance): %d warp-instructions, %d samples" % (kernel, tot, tots))
    for op, n in inst.most_common(top):
        print("%12d %5.1f%% inst %5.1f%% smp  %s" % (n, 100.0 * n / tot, 100.0 * smp[op] / tots, op))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
