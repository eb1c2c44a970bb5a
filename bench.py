#!/usr/bin/env python
"""bench.py -- throughput of the ATRAC1 encode+decode hot path on B200.

Workload (BASELINE.json configs[1]): 1 h of stereo 44.1 kHz synthetic PCM (sine + slow chirp
+ noise), fixedBlockModes [0,0,0], encoded to sound units and decoded back.  One step = one
encode + one decode pass over the whole hour.  With N GPUs every rank processes its own hour
(independent streams, no collective): weak scaling.

  python bench.py [--gpus N --steps K --warmup W]          our CUDA path
  python bench.py --impl reference [...]                    reference algorithm on host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SR = 44100
BYTES_PER_AUDIO_SEC_ENC = 2 * SR * 4 + 2 * (SR / 512) * 212      # 389,320.3
BYTES_PER_AUDIO_SEC = 2 * BYTES_PER_AUDIO_SEC_ENC                # encode + decode
BYTES_PER_SU = 2048 + 212
# FP64 warp-instructions the numerical contract needs per sound unit and direction (counted in the
# SASS of the kernels: QMF 576 DFMA + 24 DADD, transform ~650 incl. its binary32 roundings, quantise
# or dequantise ~50).  An FP64 instruction holds a sub-partition's issue port for 2 clocks on B200
# (profiles/r01_ubench2_butterfly_rounding.txt), so 2 * this is the floor in issue clocks per unit.
FP64_WARP_INSTR_PER_SU = 1300
METRIC = "encoded audio-sec/sec per B200 (stereo 44.1k) at 1/2/4/8 GPU; % of HBM roofline"
UNIT = "audio-s/s"
WORKLOAD = "cfg2: 1 h stereo 44.1 kHz synthetic PCM, encode+decode, fixedBlockModes [0,0,0]"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel, n_su):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r*_dram_traffic_*.json),
    scaled by sound units when this run's workload differs from the captured one; None if absent."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_traffic_*.json")))
    if not files:
        return None, None
    try:
        doc = json.load(open(files[-1]))
        k = doc["kernels"][kernel]
        total = (k["dram_bytes_read"] + k["dram_bytes_write"]) * (n_su / float(doc["sound_units"]))
        return int(total), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


# --------------------------------------------------------------------------------------
# clocks sampler (NVML)
# --------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.ok = index, [], False, False
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, int(reasons)))
            except Exception:
                pass
            time.sleep(0.004)

    def summary(self, t0, t1):
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        sel = [s for s in self.samples if t0 <= s[0] <= t1]
        where = "timed region"
        if not sel:
            sel, where = self.samples, "whole run (timed region shorter than the sampling period)"
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "window": "nvml unavailable"}
        mhz = sorted(s[1] for s in sel)
        bits = 0
        for s in sel:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sel), "window": where}


# --------------------------------------------------------------------------------------
# synthetic input
# --------------------------------------------------------------------------------------
def synth_cfg2_device(torch, seconds, seed, device):
    """sine (440 / 880 Hz) + slow chirp 100 Hz -> 8 kHz + Gaussian noise, f32 planar [2, n]."""
    n = int(round(seconds * SR))
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((2, n), dtype=torch.float32, device=device)
    step = 1 << 24
    k = (8000.0 - 100.0) / max(seconds, 1e-9)
    for a in range(0, n, step):
        b = min(n, a + step)
        t = torch.arange(a, b, device=device, dtype=torch.float64) / SR
        chirp = 0.25 * torch.sin(2 * math.pi * (100.0 * t + 0.5 * k * t * t))
        for c, f in ((0, 440.0), (1, 880.0)):
            x = 0.4 * torch.sin(2 * math.pi * f * t) + chirp
            x = x + 0.05 * torch.randn(b - a, generator=g, device=device, dtype=torch.float64)
            out[c, a:b] = x.to(torch.float32)
    return out


# --------------------------------------------------------------------------------------
# reference arm: the reference's algorithm on the host cores
# --------------------------------------------------------------------------------------
def cpu_pass(O, chans, opts, threads):
    t0 = time.perf_counter()
    su = O.encode_pcm(chans, opts, threads=threads, chunk_frames=256)
    t1 = time.perf_counter()
    pcm = O.decode_su(su, 2, threads=threads, chunk_frames=256)
    t2 = time.perf_counter()
    return su, pcm, t1 - t0, t2 - t1


def run_reference(args, rank):
    if rank != 0:
        return
    import signals as S
    from oracle import oracle as O

    O.build()
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    # bounded sample: ~1-2 s of CPU work per step, whole run within a couple of minutes
    est_rt = 11.0 * threads  # ~11x realtime per host thread for encode+decode of the C port
    sample_s = max(5.0, min(300.0, 90.0 * est_rt / max(total, 1)))
    chans = S.cfg2_stereo(sample_s, seed=0xCA27A2)
    opts = O.make_options(fixed_modes=[0, 0, 0])
    for _ in range(args.warmup):
        cpu_pass(O, chans, opts, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(O, chans, opts, threads)
    dt = time.perf_counter() - t0
    value = args.steps * sample_s / dt
    sample = "%.0f s of the cfg2 stereo workload per step (numpy-generated, same recipe)" % sample_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "threads": threads,
                   "note": "reference = C restatement of carta1's JS algorithm (oracle/): no JavaScript engine "
                           "exists in this image, so the JS reference itself cannot run"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch

    import carta1_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; carta1_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = carta1_b200.Context(local_rank)
    if args.units_per_pass:
        ctx.set_max_units_per_pass(args.units_per_pass)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    seconds = float(args.seconds)
    pcm = synth_cfg2_device(torch, seconds, 0xCA27A2 + rank, dev)
    n = pcm.shape[1]
    frames = (n + 511) // 512
    n_su = frames * 2
    d_su = torch.zeros(n_su * 212, dtype=torch.uint8, device=dev)
    d_out = torch.zeros((2, frames * 512), dtype=torch.float32, device=dev)
    opts = carta1_b200.make_enc_opts(fixed_block_modes=None if args.auto_modes else [0, 0, 0])
    torch.cuda.synchronize()

    def encode():
        ctx.encode_device(pcm.data_ptr(), n, 2, n, 0, frames, opts, d_su.data_ptr(), 2, 1)

    def decode():
        ctx.decode_device(d_su.data_ptr(), 2, 1, n_su, 2, 0, frames, d_out.data_ptr(), frames * 512)

    def timed(fn, steps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)

    def both():
        encode()
        decode()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        both()
    ctx.sync()
    # ---- the timed region: K steps, device-resident inputs (1.27 GB PCM per pass >> 126 MB L2)
    barrier()
    launches0 = ctx.launch_count
    t_start = time.perf_counter()
    ms = timed(both, args.steps)
    t_end = time.perf_counter()
    launches = ctx.launch_count - launches0
    barrier()
    ms_enc = timed(encode, args.steps)
    ms_dec = timed(decode, args.steps)
    # ---- per-kernel durations with CUDA events on the launching stream (separate pass so the
    # event records do not sit inside the headline region)
    ctx.profile(True)
    for _ in range(args.steps):
        both()
    prof = ctx.profile_read()
    ctx.profile(False)

    # ---- end to end through the host-facing C ABI: pinned host buffers, H2D + D2H inside
    pcm_h = torch.empty((2, n), dtype=torch.float32).pin_memory()
    pcm_h.copy_(pcm)
    su_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
    out_h = torch.empty((2, frames * 512), dtype=torch.float32).pin_memory()
    chans_np = [pcm_h[0].numpy(), pcm_h[1].numpy()]
    outs_np = [out_h[0].numpy(), out_h[1].numpy()]
    su_np = su_h.numpy()

    def e2e_step():
        got = ctx.encode_pcm_into(chans_np, su_np, opts)
        assert got == n_su
        ctx.decode_su_into(su_np, n_su, 2, outs_np)

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    ms_e2e_seq = timed(e2e_step, e2e_steps)
    barrier()
    # The same two calls, in flight together: an encoder thread and a decoder thread, each on its
    # own context (a handle is used by one thread at a time, include/carta1_b200.h).  carta1_encode_pcm
    # is H2D-heavy and carta1_decode_su D2H-heavy, so together they use both directions of the PCIe
    # link.  Work per step is unchanged (one hour encoded, one hour decoded); the decoder reads the
    # units of the previous step's encode of the same PCM (identical bytes) from its own pinned buffer.
    ctx2 = carta1_b200.Context(local_rank)
    if args.units_per_pass:
        ctx2.set_max_units_per_pass(args.units_per_pass)
    su2_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
    su2_h.copy_(su_h)
    su2_np = su2_h.numpy()
    errs = []

    def e2e_duplex_step():
        def dec():
            try:
                ctx2.decode_su_into(su2_np, n_su, 2, outs_np)
            except Exception as ex:  # surfaced after the join
                errs.append(ex)

        th = threading.Thread(target=dec)
        th.start()
        got = ctx.encode_pcm_into(chans_np, su_np, opts)
        th.join()
        if errs:
            raise errs[0]
        assert got == n_su

    def wall_ms(fn, steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1000.0

    e2e_duplex_step()
    barrier()
    ms_e2e = wall_ms(e2e_duplex_step, e2e_steps)
    barrier()
    duplex_equal = bool(np.array_equal(su_np, su2_np))
    # WAV-shaped I/O: int16 interleaved PCM in and out, as the reference CLI reads and writes it
    # (bin/cli.js:394-404, processor.js:382-389); the conversions run inside the QMF kernels, so the PCM side of
    # the link carries half the bytes.  Same two calls in flight together.
    wav_h = torch.empty(n * 2, dtype=torch.int16).pin_memory()
    wav_h.copy_((pcm.t().contiguous().clamp(-1.0, 1.0) * 32767.0).to(torch.int16).reshape(-1))
    wav_out_h = torch.empty(frames * 512 * 2, dtype=torch.int16).pin_memory()
    wav_np, wav_out_np = wav_h.numpy(), wav_out_h.numpy()
    su3_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
    su3_np = su3_h.numpy()

    def e2e_s16_step():
        def dec():
            try:
                ctx2.decode_su_s16_into(su2_np, n_su, 2, wav_out_np)
            except Exception as ex:
                errs.append(ex)

        th = threading.Thread(target=dec)
        th.start()
        got = ctx.encode_pcm_s16_into(wav_np, 2, su3_np, opts)
        th.join()
        if errs:
            raise errs[0]
        assert got == n_su

    e2e_s16_step()
    barrier()
    ms_e2e_s16 = wall_ms(e2e_s16_step, e2e_steps)
    barrier()
    ctx2.close()
    # pageable caller arrays (what a host runtime that cannot pin its typed arrays passes): staged through the
    # context's pinned bounce slots by parallel memcpy; one call after the other
    chans_pg = [np.array(c) for c in chans_np]
    outs_pg = [np.empty_like(o) for o in outs_np]
    su_pg = np.empty_like(su_np)

    def e2e_pageable_step():
        got = ctx.encode_pcm_into(chans_pg, su_pg, opts)
        assert got == n_su
        ctx.decode_su_into(su_pg, n_su, 2, outs_pg)

    e2e_pageable_step()
    barrier()
    ms_e2e_pg = wall_ms(e2e_pageable_step, e2e_steps)
    barrier()
    pageable_equal = bool(np.array_equal(su_pg, su_np)) and all(np.array_equal(a.view(np.uint32), b.view(np.uint32))
                                                                 for a, b in zip(outs_pg, outs_np))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if dist is not None:
        t = torch.tensor([ms, ms_enc, ms_dec, ms_e2e, ms_e2e_seq, ms_e2e_s16, ms_e2e_pg], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_enc, ms_dec, ms_e2e, ms_e2e_seq, ms_e2e_s16, ms_e2e_pg = t.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = ms / args.steps
        value = world * seconds / (ms_step / 1000.0)
        # dominant kernel
        dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else ("none", (0.0, 1))
        dom_ms = dom[1][0] / max(dom[1][1], 1)
        step_ms_prof = sum(v[0] for v in prof.values()) / max(args.steps, 1)
        achieved = BYTES_PER_SU * n_su / (dom_ms / 1000.0) / 1e9 if dom_ms > 0 else 0.0
        step_gbs = BYTES_PER_AUDIO_SEC * seconds / (ms_step / 1000.0) / 1e9
        traffic, traffic_src = ncu_traffic(dom[0], n_su)
        props = torch.cuda.get_device_properties(dev)
        clk_mhz = sampler.max_mhz or 1965
        fp64_floor_ms = 2 * n_su * 2.0 * FP64_WARP_INSTR_PER_SU / (props.multi_processor_count * 4 * clk_mhz * 1e6) * 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "audio_seconds_per_gpu": seconds, "sound_units_per_gpu": n_su,
                       "l2": "inputs larger than L2 (1.27 GB PCM + 131 MB units per pass)",
                       "sharding": "independent stereo streams per rank, no collective"},
            "encode_only": {"value": world * seconds / (ms_enc / args.steps / 1000.0), "unit": UNIT},
            "decode_only": {"value": world * seconds / (ms_dec / args.steps / 1000.0), "unit": UNIT},
            "roofline": {
                "bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms": dom_ms, "kernel_share_of_step": (dom_ms * dom[1][1] / max(args.steps, 1)) / step_ms_prof if step_ms_prof else None,
                "algorithmic_bytes_per_launch": BYTES_PER_SU * n_su,
                "step": {"achieved": step_gbs, "frac": step_gbs / peak,
                         "algorithmic_bytes_per_step": BYTES_PER_AUDIO_SEC * seconds},
                "kernels_ms_per_step": {k: v[0] / max(args.steps, 1) for k, v in prof.items()},
                # the bound that actually binds under the bit-exact FP64 contract
                "fp64_issue": {"floor_ms_per_step": fp64_floor_ms, "frac": fp64_floor_ms / ms_step,
                               "fp64_warp_instr_per_unit_per_direction": FP64_WARP_INSTR_PER_SU,
                               "sms": props.multi_processor_count, "sm_mhz": clk_mhz},
            },
            "e2e": {"value": world * seconds / (ms_e2e / e2e_steps / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": int(2 * n * 4 + n_su * 212), "d2h_bytes_per_step": int(n_su * 212 + 2 * frames * 512 * 4),
                    "steps": e2e_steps,
                    "api": "carta1_encode_pcm || carta1_decode_su: both calls in flight (two host threads, one context "
                           "each), pinned host buffers, H2D and D2H inside; host wall clock, max over ranks",
                    "sequential": {"value": world * seconds / (ms_e2e_seq / e2e_steps / 1000.0), "unit": UNIT,
                                   "api": "carta1_encode_pcm then carta1_decode_su on one context"},
                    "wav_int16": {"value": world * seconds / (ms_e2e_s16 / e2e_steps / 1000.0), "unit": UNIT,
                                  "h2d_bytes_per_step": int(2 * n * 2 + n_su * 212), "d2h_bytes_per_step": int(n_su * 212 + 2 * frames * 512 * 2),
                                  "api": "carta1_encode_pcm_s16 || carta1_decode_su_s16 (WAV-shaped int16 PCM in and out)"},
                    "pageable": {"value": world * seconds / (ms_e2e_pg / e2e_steps / 1000.0), "unit": UNIT,
                                 "api": "carta1_encode_pcm then carta1_decode_su with pageable (unpinned) caller arrays",
                                 "identical_to_pinned_run": pageable_equal},
                    "units_identical_across_steps": duplex_equal},
            "gpu_launches": launches,
            "clocks": sampler.summary(t_start, t_end),
        }
        # ---- CPU baseline leg (rank 0, N=1): the oracle on a bounded sample of the same data,
        # doubling as the in-bench parity check of the GPU output on that prefix.
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O

            O.build()
            threads = os.cpu_count() or 1
            sample_s = min(seconds, float(args.cpu_sample_seconds))
            ns = int(sample_s * SR) // 512 * 512
            chans = [np.ascontiguousarray(pcm_h[c, :ns].numpy()) for c in range(2)]
            oopts = O.make_options(fixed_modes=None if args.auto_modes else [0, 0, 0])
            su_ref, pcm_ref, te, td = cpu_pass(O, chans, oopts, threads)
            k = su_ref.shape[0]
            # frame f depends on samples <= 512 f + 511 only, so the prefix must agree exactly
            parity = bool(np.array_equal(su_np.reshape(-1, 212)[:k], su_ref))
            gpu_pcm_ok = all(np.array_equal(outs_np[c][: (k // 2) * 512].view(np.uint32),
                                            pcm_ref[c][: (k // 2) * 512].view(np.uint32)) for c in range(2))
            line["cpu_baseline"] = {
                "value": (ns / SR) / (te + td), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "first %.0f s of the same stereo PCM, encode+decode, %d threads (C restatement; no JS engine in the image)" % (ns / SR, threads),
                "encode_only": (ns / SR) / te, "decode_only": (ns / SR) / td,
                "gpu_output_bit_exact_on_sample": parity and gpu_pcm_ok,
            }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="audio seconds per GPU (default: the 1 h of cfg2)")
    ap.add_argument("--cpu-sample-seconds", type=float, default=3600.0,
                    help="audio seconds of the workload the CPU baseline encodes and decodes (default: the whole hour, ~4 s on 16 threads; it doubles as the bit-exactness check of the GPU output)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--units-per-pass", type=int, default=0,
                    help="development aid: sound units per pipelined pass of the host entry points (default: the library's)")
    ap.add_argument("--auto-modes", action="store_true",
                    help="development aid: transient-driven block modes instead of the headline fixed [0,0,0] (not a bench line)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
