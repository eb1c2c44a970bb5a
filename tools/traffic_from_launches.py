#!/usr/bin/env python
"""profiles/rNN_dram_traffic_TAG.json from the ncu launch list of tools/gpu_check.sh.

  python tools/traffic_from_launches.py gpurun_out/launches_TAG.csv profiles/r02_dram_traffic_TAG.json [sound_units]

The launch list holds every kernel of `bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs` with
gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum and smsp__inst_executed_pipe_fp64.sum.  The first
complete encode+decode pass (an encode start up to the next encode start, with a decode in between) is summed per bench.py kernel role.  The file
records the hash of the kernel sources it was taken from; bench.py reports `roofline.traffic` and the FP64
instruction count only while that hash matches the build (tests/test_bench_contract_cpu.py fails otherwise).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_sha  # noqa: E402

ROLES = [("qmf_analysis_kernel", "qmf_analysis"), ("transient_spectrum_kernel", "band_mags"),
         ("transient_modes_kernel", "transient_modes"), ("imdct_kernel", "imdct"), ("mdct_kernel", "mdct"),
         ("alloc_kernel", "alloc"), ("quant_pack_kernel", "quant_pack"), ("unpack_kernel", "unpack_dequant"),
         ("synth_kernel", "synth"), ("encode_fused_kernel", "qmf_mdct"), ("decode_fused_kernel", "imdct_synth")]


def role_of(name):
    for needle, role in ROLES:
        if needle in name:
            return role
    return None


def main():
    src, dst = sys.argv[1], sys.argv[2]
    units = int(sys.argv[3]) if len(sys.argv) > 3 else 620158
    rows = {}
    with open(src, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for rec in csv.DictReader(lines):
        i = int(rec["ID"])
        d = rows.setdefault(i, {"name": rec["Kernel Name"]})
        d[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", ""))
    ids = sorted(rows)
    starts = [i for i in ids if role_of(rows[i]["name"]) in ("qmf_analysis", "qmf_mdct")]
    if not starts:
        raise SystemExit("no encode pass in the launch list")
    # passes start at an encode's first kernel; the first pass that also decodes is a whole device-resident
    # encode+decode pass (the one before it is the untimed set-up encode of the decoder's input)
    bounds = starts + [ids[-1] + 1]
    begin = end = None
    for a, b in zip(bounds[:-1], bounds[1:]):
        if any(role_of(rows[i]["name"]) in ("synth", "imdct_synth") for i in ids if a <= i < b):
            begin, end = a, b
            break
    if begin is None:
        raise SystemExit("no pass with a decode in the launch list")
    kernels = {}
    for i in ids:
        if not (begin <= i < end):
            continue
        role = role_of(rows[i]["name"])
        if role is None:
            continue
        k = kernels.setdefault(role, {"launches": 0, "dram_bytes_read": 0, "dram_bytes_write": 0, "ncu_ns": 0, "fp64_warp_instr": 0})
        k["launches"] += 1
        k["dram_bytes_read"] += int(rows[i].get("dram__bytes_read.sum", 0))
        k["dram_bytes_write"] += int(rows[i].get("dram__bytes_write.sum", 0))
        k["ncu_ns"] += int(rows[i].get("gpu__time_duration.sum", 0))
        k["fp64_warp_instr"] += int(rows[i].get("smsp__inst_executed_pipe_fp64.sum", 0))
    doc = {
        "what": "per bench.py kernel role over one encode+decode pass of the 1 h stereo workload: dram__bytes_read.sum, "
                "dram__bytes_write.sum, gpu__time_duration.sum (ns, under ncu: serialised, cold cache) and "
                "smsp__inst_executed_pipe_fp64.sum (FP64 warp instructions)",
        "command": "tools/gpu_check.sh (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
                   "smsp__inst_executed_pipe_fp64.sum --clock-control none --csv python bench.py --steps 2 --warmup 3 "
                   "--no-cpu-baseline --no-configs)",
        "source": os.path.relpath(src, ROOT) if os.path.isabs(src) else src,
        "source_sha": kernel_source_sha(),
        "sound_units": units,
        "launch_ids": [begin, end - 1],
        "kernels": kernels,
        "total_dram_bytes": sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in kernels.values()),
    }
    with open(dst, "w") as fh:
        json.dump(doc, fh, indent=1)
        fh.write("\n")
    print(json.dumps({"total_dram_GB": doc["total_dram_bytes"] / 1e9, "kernels": {k: round((v["dram_bytes_read"] + v["dram_bytes_write"]) / 1e9, 3) for k, v in kernels.items()}}))


if __name__ == "__main__":
    main()
