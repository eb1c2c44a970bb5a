#!/usr/bin/env python
"""bench.py -- throughput of the ATRAC1 encode+decode hot path on B200.

Headline workload (BASELINE.json configs[1]): 1 h of stereo 44.1 kHz synthetic PCM (sine + slow chirp
+ noise), fixedBlockModes [0,0,0], encoded to sound units and decoded back.  One step = one encode +
one decode pass over every frame a rank owns.

N = 1: one 1 h stream.  N > 1 (configs[4]): a fixed list of long stereo streams, N hours in total, cut by
carta1_b200.sharding.plan into one contiguous span of the stream-major frame order per rank; a span that
starts inside a stream stages a 2-frame PCM halo (encode) / 1-unit halo (decode).  No collective on the data
path; after the timed region the shards are gathered on rank 0 and compared byte for byte with the
unsharded result ("sharded_output_identical").  Per-GPU work stays 1 h: weak scaling.

The `configs` block carries the other BASELINE configs (cfg1 10 s auto modes, cfg3 1 h transient-heavy auto
modes, cfg4 4096 mono streams through the stateful frame API), each with value, e2e and a parity flag.

  python bench.py [--gpus N --steps K --warmup W]          our CUDA path
  python bench.py --impl reference [...]                    reference algorithm on host cores

Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
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
# FP64 warp-instructions per sound unit and direction when no ncu capture of this build is committed
# (profiles/r*_dram_traffic_*.json carries smsp__inst_executed_pipe_fp64 per kernel; see fp64_per_unit()).
FP64_WARP_INSTR_PER_SU_FALLBACK = 1300
METRIC = "encoded audio-sec/sec per B200 (stereo 44.1k) at 1/2/4/8 GPU; % of HBM roofline"
UNIT = "audio-s/s"
WORKLOAD = "cfg2: 1 h stereo 44.1 kHz synthetic PCM, encode+decode, fixedBlockModes [0,0,0]"
WORKLOAD_N = ("cfg5 at 1 h per GPU: N h of cfg2-recipe stereo streams (1.25 h / 0.75 h alternating), sharded by stream and "
              "contiguous frame range with QMF/MDCT halos, encode+decode, fixedBlockModes [0,0,0]")
ENCODE_KERNELS = ("qmf_analysis", "band_mags", "transient_modes", "mdct", "alloc", "quant_pack")
DECODE_KERNELS = ("unpack_dequant", "imdct", "bands_time", "synth")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_sha():
    """Hash of the kernel sources: an ncu capture describes the build it was taken from and nothing else."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "carta1_b200", "csrc")
    for f in ("c1_common.cuh", "c1_fft.cuh", "c1_fdlibm.cuh", "c1_encode.cu", "c1_decode.cu"):
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_capture():
    """The newest committed per-kernel ncu summary (tools/traffic_from_launches.py over the launch list of
    tools/gpu_check.sh).  Returns (doc, relative path, fresh): fresh is False when the kernel sources changed
    after the capture, and the stale numbers are then NOT reported."""
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_traffic_*.json")))
    if not files:
        return None, None, False
    try:
        doc = json.load(open(files[-1]))
    except Exception:
        return None, None, False
    return doc, os.path.relpath(files[-1], ROOT), doc.get("source_sha") == kernel_source_sha()


def fp64_per_unit(doc, fresh):
    """FP64 warp-instructions per sound unit and direction from smsp__inst_executed_pipe_fp64 of the capture."""
    if not (doc and fresh):
        return float(FP64_WARP_INSTR_PER_SU_FALLBACK), "fallback (hand count; no ncu capture of this build committed)"
    try:
        units = float(doc["sound_units"])
        enc = sum(doc["kernels"][k].get("fp64_warp_instr", 0) for k in ENCODE_KERNELS if k in doc["kernels"])
        dec = sum(doc["kernels"][k].get("fp64_warp_instr", 0) for k in DECODE_KERNELS if k in doc["kernels"])
        if enc <= 0 or dec <= 0:
            raise KeyError("fp64_warp_instr")
        return (enc + dec) / 2.0 / units, "smsp__inst_executed_pipe_fp64.sum of the committed capture"
    except Exception:
        return float(FP64_WARP_INSTR_PER_SU_FALLBACK), "fallback (capture carries no FP64 instruction counts)"


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
# synthetic input: every sample is a pure function of (stream key, sample index), so a rank that only
# holds a frame range of a stream generates exactly the samples the unsharded run sees
# --------------------------------------------------------------------------------------
def _mix64(torch, x):
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic, logical shifts by masking)."""
    x = (x ^ ((x >> 30) & ((1 << 34) - 1))) * -4658895280553007687   # 0xBF58476D1CE4E5B9
    x = (x ^ ((x >> 27) & ((1 << 37) - 1))) * -7723592293110705685   # 0x94D049BB133111EB
    return x ^ ((x >> 31) & ((1 << 33) - 1))


def hashed_uniform(torch, idx, key):
    """idx: int64 sample indices -> uniform in (0, 1), float64."""
    x = _mix64(torch, idx * -7046029254386353131 + int(key))         # 0x9E3779B97F4A7C15
    return (((x >> 11) & ((1 << 53) - 1)).to(torch.float64) + 0.5) * (2.0 ** -53)


def hashed_normal(torch, idx, key):
    u1 = hashed_uniform(torch, idx, 2 * key + 1)
    u2 = hashed_uniform(torch, idx, 2 * key + 2)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos((2.0 * math.pi) * u2)


def synth_cfg2_span(torch, key, stream_seconds, a, b, device):
    """Samples [a, b) of a cfg2-recipe stereo stream: sine (440 / 880 Hz) + slow chirp 100 Hz -> 8 kHz over
    the stream + Gaussian noise (sigma 0.05); f32 planar [2, b - a]."""
    out = torch.empty((2, b - a), dtype=torch.float32, device=device)
    step = 1 << 24
    k = (8000.0 - 100.0) / max(stream_seconds, 1e-9)
    for lo in range(a, b, step):
        hi = min(b, lo + step)
        idx = torch.arange(lo, hi, device=device, dtype=torch.int64)
        t = idx.to(torch.float64) / SR
        chirp = 0.25 * torch.sin(2 * math.pi * (100.0 * t + 0.5 * k * t * t))
        for c, f in ((0, 440.0), (1, 880.0)):
            x = 0.4 * torch.sin(2 * math.pi * f * t) + chirp + 0.05 * hashed_normal(torch, idx, 1000 * key + c)
            out[c, lo - a:hi - a] = x.to(torch.float32)
    return out


def synth_cfg1(torch, seconds, device):
    """configs[0]: 10 s stereo sine + noise."""
    n = int(round(seconds * SR))
    idx = torch.arange(0, n, device=device, dtype=torch.int64)
    t = idx.to(torch.float64) / SR
    out = torch.empty((2, n), dtype=torch.float32, device=device)
    for c, f in ((0, 440.0), (1, 880.0)):
        out[c] = (0.5 * torch.sin(2 * math.pi * f * t) + 0.05 * hashed_normal(torch, idx, 0xC1 + c)).to(torch.float32)
    return out


def synth_cfg3(torch, seconds, device, clicks_per_second=4.0):
    """configs[2]: transient-heavy stereo -- a quiet noise floor (sigma 0.01) with castanet-like clicks, about
    `clicks_per_second` per channel at hashed positions: 2-5 ms bursts of noise under an exponential envelope."""
    import numpy as np

    n = int(round(seconds * SR))
    out = torch.empty((2, n), dtype=torch.float32, device=device)
    step = 1 << 24
    for c in range(2):
        rng = np.random.default_rng(0xCA27A3 + c)
        n_clicks = max(1, int(rng.poisson(clicks_per_second * seconds)))
        pos = np.sort(rng.integers(0, max(n - 1, 1), n_clicks)).astype(np.int64)
        dur = (rng.uniform(0.002, 0.005, n_clicks) * SR).astype(np.int64)
        d_pos = torch.from_numpy(pos).to(device)
        d_dur = torch.from_numpy(dur).to(device)
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            idx = torch.arange(lo, hi, device=device, dtype=torch.int64)
            j = torch.searchsorted(d_pos, idx, right=True) - 1          # the most recent click at or before idx
            jc = j.clamp(min=0)
            since = idx - d_pos[jc]
            du = d_dur[jc]
            inside = (j >= 0) & (since < du)
            env = torch.exp(-since.to(torch.float64) / (0.25 * du.to(torch.float64)))
            burst = torch.where(inside, 0.45 * env * hashed_normal(torch, idx, 0xC300 + c), torch.zeros((), dtype=torch.float64, device=device))
            x = 0.01 * hashed_normal(torch, idx, 0xC310 + c) + burst
            out[c, lo:hi] = x.clamp(-1.0, 1.0).to(torch.float32)
    return out


def synth_cfg4(torch, n_streams, n_frames, device):
    """configs[3]: independent mono streams, three sines of hashed frequency / level / phase + noise; [n_streams, n_frames * 512]."""
    n = n_frames * 512
    s = torch.arange(n_streams, device=device, dtype=torch.int64)
    idx = torch.arange(n, device=device, dtype=torch.int64)
    t = (idx.to(torch.float64) / SR)[None, :]
    flat = (s[:, None] * (1 << 32) + idx[None, :])
    x = 0.02 * hashed_normal(torch, flat, 0xC4)
    for k in range(3):
        f = 80.0 + (9000.0 - 80.0) * hashed_uniform(torch, s, 0xC410 + k)
        a = 0.05 + 0.25 * hashed_uniform(torch, s, 0xC420 + k)
        ph = 6.28 * hashed_uniform(torch, s, 0xC430 + k)
        x = x + a[:, None] * torch.sin(2 * math.pi * f[:, None] * t + ph[:, None])
    return x.to(torch.float32)


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
    sample = ("%.0f s of the cfg2 stereo recipe per step (a bounded sample of the 1 h the GPU arm times: the metric is a rate, "
              "audio-seconds per second, on the same recipe and options)" % sample_s)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "threads": threads,
                   "note": "reference = C restatement of carta1's JS algorithm (oracle/), bit-identical to the output of the "
                           "JavaScript itself on the pinned inputs (tests/golden/ref); the JavaScript cannot travel to "
                           "the GPU box (no reference sources in the repo, no Node there)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------
def stream_seconds(world, seconds):
    """The fixed stream list of the N-GPU run: N * seconds of audio in streams that do NOT line up with the ranks,
    so that the plan has to cut frame ranges (1.25 / 0.75 alternating; an odd tail stream is 1.0)."""
    if world == 1:
        return [seconds]
    lens = [seconds * (1.25 if i % 2 == 0 else 0.75) for i in range(world)]
    if world % 2:
        lens[-1] = seconds
    return lens


def shard_staging(begin, enc_halo):
    """Frame arithmetic of one shard [begin, end) of a stream.  Returns (s0, e0, enc_off_frames, dec_halo):
    PCM is staged from frame s0 (two frames of encode halo, plus one more so that the unit before `begin`, the decode
    halo, can be produced on the same rank); the decode input starts at unit frame e0; the timed encode starts
    enc_off_frames after s0 and skips enc_halo frames; the decode skips dec_halo frames."""
    s0 = begin - 3 if begin >= 5 else 0
    e0 = begin - 1 if s0 > 0 else 0
    return s0, e0, begin - enc_halo - s0, begin - e0


class ShardWork:
    """Device buffers of one (stream, frame range) shard and the launches over them."""

    def __init__(self, torch, ctx, sh, stream_key, stream_s, n_samples, opts, dev):
        self.sh, self.ctx, self.opts = sh, ctx, opts
        begin, end = sh.begin, sh.end
        self.frames = end - begin
        s0, e0, enc_off_frames, dec_halo = shard_staging(begin, sh.enc_halo)
        self.s0, self.e0 = s0, e0
        span = (end - s0) * 512
        self.span = span
        last = min(n_samples, end * 512)
        self.pcm = torch.zeros((2, span), dtype=torch.float32, device=dev)
        self.pcm[:, :last - s0 * 512] = synth_cfg2_span(torch, stream_key, stream_s, s0 * 512, last, dev)
        self.valid = last - s0 * 512
        self.enc_off = enc_off_frames * 512                   # floats from the row start to the encode halo
        self.dec_halo = dec_halo
        self.su_in = torch.zeros((end - e0) * 2 * 212, dtype=torch.uint8, device=dev)
        self.su_out = torch.zeros(self.frames * 2 * 212, dtype=torch.uint8, device=dev)
        self.out = torch.zeros((2, self.frames * 512), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()  # the PCM was written on torch's stream, the library launches on the context's
        # the decode input (with its halo unit): the same encoder over [e0, end), outside the timed region
        ctx.encode_device(self.pcm.data_ptr() + 4 * (e0 - (2 if e0 else 0) - s0) * 512, span, 2,
                          self.valid - (e0 - (2 if e0 else 0) - s0) * 512, 2 if e0 else 0, end - e0, opts,
                          self.su_in.data_ptr(), 2, 1, sync=True)

    def encode(self):
        sh = self.sh
        self.ctx.encode_device(self.pcm.data_ptr() + 4 * self.enc_off, self.span, 2, self.valid - self.enc_off, sh.enc_halo,
                               self.frames, self.opts, self.su_out.data_ptr(), 2, 1)

    def decode(self):
        self.ctx.decode_device(self.su_in.data_ptr(), 2, 1, (self.sh.end - self.e0) * 2, 2, self.dec_halo, self.frames,
                               self.out.data_ptr(), self.frames * 512)


def link_probe(torch, dev, barrier, world, dist):
    """Raw pinned-copy ceiling of the host link, all ranks at once: H2D alone, D2H alone, both (GB/s per direction,
    summed over ranks; the slowest rank's time)."""
    n = 1 << 29
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    out = {}
    for name, h2d, d2h in (("h2d", True, False), ("d2h", False, True), ("duplex", True, True)):
        best = None
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                if h2d:
                    with torch.cuda.stream(s1):
                        d_a.copy_(h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if rep > 0:
                best = dt if best is None else min(best, dt)
        out[name] = world * 2 * n / best / 1e9
    return out


def run_configs(torch, np, carta1_b200, ctx, dev, args):
    """BASELINE configs 1, 3 and 4 on rank 0: value, e2e and a parity flag each (the checker is the CPU oracle)."""
    from oracle import oracle as O

    O.build()
    threads = os.cpu_count() or 1
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    res = {}

    def dev_ms(fn, reps):
        fn()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    def wall_ms(fn, reps):
        for _ in range(3):  # a stateful handle captures and instantiates its CUDA graph on the second call of a shape
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1000.0 / reps

    def stereo_config(pcm, reps, e2e_reps):
        """Device-resident and host-API timing of one auto-block-mode stereo stream + whole-stream parity."""
        n = pcm.shape[1]
        seconds = n / SR
        frames = (n + 511) // 512
        n_su = 2 * frames
        opts = carta1_b200.make_enc_opts()
        d_su = torch.zeros(n_su * 212, dtype=torch.uint8, device=dev)
        d_out = torch.zeros((2, frames * 512), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()  # torch's stream wrote the PCM, the context's stream reads it
        enc = lambda: ctx.encode_device(pcm.data_ptr(), n, 2, n, 0, frames, opts, d_su.data_ptr(), 2, 1)  # noqa: E731
        dec = lambda: ctx.decode_device(d_su.data_ptr(), 2, 1, n_su, 2, 0, frames, d_out.data_ptr(), frames * 512)  # noqa: E731
        ctx.near_threshold(reset=True)
        ms_e, ms_d = dev_ms(enc, reps), dev_ms(dec, reps)
        ctx.sync()
        ctx.near_threshold(reset=True)
        enc()
        near = ctx.near_threshold(reset=True)
        ctx.profile(True)
        enc(); dec()
        prof = ctx.profile_read()
        ctx.profile(False)
        # host API, pinned buffers, the two calls in flight together on two contexts
        pcm_h = torch.empty((2, n), dtype=torch.float32).pin_memory()
        pcm_h.copy_(pcm)
        su_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
        su2_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
        out_h = torch.empty((2, frames * 512), dtype=torch.float32).pin_memory()
        chans = [pcm_h[0].numpy(), pcm_h[1].numpy()]
        outs = [out_h[0].numpy(), out_h[1].numpy()]
        ctx.encode_pcm_into(chans, su_h.numpy(), opts)
        su2_h.copy_(su_h)
        ctx2 = carta1_b200.Context(ctx.device)
        errs = []

        def duplex():
            def d():
                try:
                    ctx2.decode_su_into(su2_h.numpy(), n_su, 2, outs)
                except Exception as ex:
                    errs.append(ex)

            th = threading.Thread(target=d)
            th.start()
            ctx.encode_pcm_into(chans, su_h.numpy(), opts)
            th.join()
            if errs:
                raise errs[0]

        ms_e2e = wall_ms(duplex, e2e_reps)
        ctx2.close()
        # parity: the whole stream against the oracle, sound units and decoded PCM, bit for bit
        want = O.encode_pcm(chans, O.make_options(), threads=threads, chunk_frames=256)
        ref = O.decode_su(want, 2, threads=threads, chunk_frames=256)
        su_dev = d_su.cpu().numpy().reshape(-1, 212)
        ok = bool(np.array_equal(su_dev, want)) and bool(np.array_equal(su_h.numpy().reshape(-1, 212), want))
        out_dev = d_out.cpu().numpy()
        for c in range(2):
            ok = ok and bool(np.array_equal(out_dev[c].view(np.uint32), ref[c].view(np.uint32)))
            ok = ok and bool(np.array_equal(outs[c].view(np.uint32), ref[c].view(np.uint32)))
        hdr = d_su.view(-1, 212)[:, 0].to(torch.int32)
        m0, m1, m2 = 2 - ((hdr >> 6) & 3), 2 - ((hdr >> 4) & 3), 3 - ((hdr >> 2) & 3)
        short = {"any_band": float(((m0 != 0) | (m1 != 0) | (m2 != 0)).float().mean().item()),
                 "low": float((m0 != 0).float().mean().item()), "mid": float((m1 != 0).float().mean().item()),
                 "high": float((m2 != 0).float().mean().item())}
        return {"value": seconds / ((ms_e + ms_d) / 1e3), "unit": UNIT,
                "encode_only": seconds / (ms_e / 1e3), "decode_only": seconds / (ms_d / 1e3),
                "ms_encode": ms_e, "ms_decode": ms_d,
                "e2e": {"value": seconds / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(2 * n * 4 + n_su * 212),
                        "d2h_bytes_per_step": int(n_su * 212 + 2 * frames * 512 * 4),
                        "api": "carta1_encode_pcm || carta1_decode_su, pinned host buffers, host wall clock"},
                "bit_exact_vs_oracle": ok, "parity_span": "whole stream: %d sound units, %d samples per channel" % (n_su, frames * 512),
                "short_block_frames": short, "near_threshold": near,
                "kernels_ms": {k: round(v[0], 4) for k, v in prof.items()}}

    # ---- cfg1
    r = stereo_config(synth_cfg1(torch, 10.0, dev), 20, 20)
    r["workload"] = "cfg1: 10 s stereo sine+noise, default bias, auto block modes"
    res["cfg1"] = r
    # ---- cfg3
    sec3 = float(args.cfg3_seconds)
    r = stereo_config(synth_cfg3(torch, sec3, dev), 5, 2)
    r["workload"] = "cfg3: %.0f s stereo transient-heavy (clicks, ~4 per second per channel), auto block modes" % sec3
    res["cfg3"] = r
    # ---- cfg4: 4096 mono streams through the stateful frame API, host buffers in and out
    ns = 4096
    cfg4 = {"workload": "cfg4: %d independent mono streams through carta1_enc_frames / carta1_dec_frames, pinned host buffers, auto block modes" % ns,
            "unit": UNIT, "frames_per_call": {}}
    total_f = 1 + 8 + 64
    pcm_all = synth_cfg4(torch, ns, total_f, dev).cpu().numpy().reshape(ns, total_f, 512)
    enc = carta1_b200.StreamEncoder(ctx, None, ns)
    dec = carta1_b200.StreamDecoder(ctx, ns)
    su_all = np.zeros((ns, total_f, 212), np.uint8)
    out_all = np.zeros((ns, total_f, 512), np.float32)
    at = 0
    for nf in (1, 8, 64):  # the parity pass: 73 consecutive frames of every stream in calls of 1, 8 and 64 frames
        su_all[:, at:at + nf] = enc.frames(np.ascontiguousarray(pcm_all[:, at:at + nf]))
        out_all[:, at:at + nf] = dec.frames(np.ascontiguousarray(su_all[:, at:at + nf]))
        at += nf
    opt_o = O.make_options()
    from concurrent.futures import ThreadPoolExecutor

    def check(s):
        x = np.ascontiguousarray(pcm_all[s].reshape(-1))
        want = O.encode_pcm([x], opt_o)
        ref = O.decode_su(want, 1)[0]
        return bool(np.array_equal(want, su_all[s])) and bool(np.array_equal(ref.view(np.uint32), out_all[s].reshape(-1).view(np.uint32)))

    with ThreadPoolExecutor(threads) as ex:
        cfg4["bit_exact_vs_oracle"] = all(ex.map(check, range(ns)))
    cfg4["parity_span"] = "%d streams x %d frames, fed in calls of 1, 8 and 64 frames" % (ns, total_f)
    for nf in (1, 8, 64):
        pcm_t = torch.empty((ns, nf, 512), dtype=torch.float32).pin_memory()
        su_t = torch.empty((ns, nf, 212), dtype=torch.uint8).pin_memory()
        out_t = torch.empty((ns, nf, 512), dtype=torch.float32).pin_memory()
        pcm_t.copy_(torch.from_numpy(pcm_all[:, :nf]))
        calls = max(4, 128 // nf)

        def per_call_ms(fn):
            for _ in range(3):  # past the capture of the handle's graph for this call shape
                fn()
            ts = []
            for _ in range(calls):
                t0 = time.perf_counter()
                fn()  # blocking: the result is in host memory when it returns
                ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            return sum(ts) / len(ts), ts[len(ts) // 2], ts[min(len(ts) - 1, (9 * len(ts)) // 10)], ts[-1]

        ms_e, med_e, p90_e, max_e = per_call_ms(lambda: enc.frames(pcm_t.numpy(), su_t.numpy()))
        ms_d, med_d, p90_d, max_d = per_call_ms(lambda: dec.frames(su_t.numpy(), out_t.numpy()))
        audio = ns * nf * 512 / SR
        cfg4["frames_per_call"][str(nf)] = {
            "value": audio / ((ms_e + ms_d) / 1e3), "encode_only": audio / (ms_e / 1e3), "decode_only": audio / (ms_d / 1e3),
            "ms_per_encode_call": ms_e, "ms_per_decode_call": ms_d,
            "per_call_ms": {"encode": {"mean": ms_e, "median": med_e, "p90": p90_e, "max": max_e},
                            "decode": {"mean": ms_d, "median": med_d, "p90": p90_d, "max": max_d}},
            "e2e": {"value": audio / ((ms_e + ms_d) / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(ns * nf * (2048 + 212)),
                    "d2h_bytes_per_step": int(ns * nf * (2048 + 212)),
                    "note": "the stateful API takes host buffers: value and e2e are the same measurement"}}
    cfg4["value"] = cfg4["frames_per_call"]["64"]["value"]
    cfg4["e2e"] = cfg4["frames_per_call"]["64"]["e2e"]
    enc.close()
    dec.close()
    res["cfg4"] = cfg4
    return res


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch

    import carta1_b200
    from carta1_b200 import sharding

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

    if args.lib:  # development aid: a variant build of the library (tools/onchip_experiment.sh)
        from carta1_b200 import _lib as lib_mod

        lib_mod.LIB_PATH = os.path.abspath(args.lib)
    ctx = carta1_b200.Context(local_rank)
    if args.units_per_pass:
        ctx.set_max_units_per_pass(args.units_per_pass)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    seconds = float(args.seconds)
    # ---- the job: a fixed list of stereo streams, `world` x `seconds` of audio in total, cut into one span per rank
    lens = stream_seconds(world, seconds)
    n_samples = [int(round(s * SR)) for s in lens]
    stream_frames = [(n + 511) // 512 for n in n_samples]
    plan = sharding.plan(stream_frames, world)
    sharding.check_plan(plan, stream_frames)
    my = plan[rank]
    opts = carta1_b200.make_enc_opts(fixed_block_modes=None if args.auto_modes else [0, 0, 0])
    # one context (own stream and scratch) per shard of this rank: the launch sequences of a rank's shards are
    # independent, so they run on different streams and one shard's kernels fill the tail of the other's
    ctxs = [ctx] + [carta1_b200.Context(local_rank) for _ in my[1:]]
    streams = [stream] + [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs[1:]]
    works = [ShardWork(torch, c, sh, 0xCA27A2 + sh.stream, lens[sh.stream], n_samples[sh.stream], opts, dev) for c, sh in zip(ctxs, my)]
    my_frames = sum(w.frames for w in works)
    n_su = 2 * my_frames                      # this rank's sound units per pass
    total_seconds = sum(n_samples) / SR       # whole job
    torch.cuda.synchronize()

    def encode():
        for w in works:
            w.encode()

    def decode():
        for w in works:
            w.decode()

    def timed(fn, steps):
        """CUDA events on the first context's stream; the other shards' streams fork after the first event and join
        before the second, so the interval covers every launch of every shard."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for s in streams[1:]:
            s.wait_event(e0)
        for _ in range(steps):
            fn()
        for s in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s)
            stream.wait_event(ev)
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
    for c in ctxs:
        c.sync()
    # ---- the timed region: K steps, device-resident inputs (1.27 GB PCM per pass >> 126 MB L2)
    barrier()
    launches0 = sum(c.launch_count for c in ctxs)
    t_start = time.perf_counter()
    ms = timed(both, args.steps)
    t_end = time.perf_counter()
    launches = sum(c.launch_count for c in ctxs) - launches0
    barrier()
    ms_enc = timed(encode, args.steps)
    ms_dec = timed(decode, args.steps)
    # ---- per-kernel durations with CUDA events on the launching stream (separate pass so the
    # event records do not sit inside the headline region); shard by shard, so that the events of one
    # shard's kernels do not include the other's
    prof = {}
    for c, w in zip(ctxs, works):
        c.profile(True)
        for _ in range(args.steps):
            w.encode()
            w.decode()
        for k, v in c.profile_read().items():
            a = prof.get(k, (0.0, 0))
            prof[k] = (a[0] + v[0], a[1] + v[1])
        c.profile(False)
    # the timed encode reproduces the decode input it was derived from (same encoder, one frame further back)
    local_ok = all(bool(torch.equal(w.su_out, w.su_in[w.dec_halo * 2 * 212:])) for w in works)

    # ---- end to end through the host-facing C ABI: pinned host buffers, H2D + D2H inside.  One encode call and one
    # decode call per shard; a shard that starts a stream goes through carta1_encode_pcm / carta1_decode_su, one that
    # starts inside a stream through carta1_encode_pcm_shard / carta1_decode_su_shard with its halo in the buffer.
    class HostShard:
        def __init__(self, w):
            sh = w.sh
            self.w = w
            self.n = w.valid - w.enc_off                                # PCM samples from the encode halo on
            self.pcm_h = torch.empty((2, self.n), dtype=torch.float32).pin_memory()
            self.pcm_h.copy_(w.pcm[:, w.enc_off:w.enc_off + self.n])
            self.su_h = torch.empty(w.frames * 2 * 212, dtype=torch.uint8).pin_memory()
            self.su_in_h = torch.empty(w.su_in.numel(), dtype=torch.uint8).pin_memory()
            self.su_in_h.copy_(w.su_in)
            self.out_h = torch.empty((2, w.frames * 512), dtype=torch.float32).pin_memory()
            self.chans = [self.pcm_h[0].numpy(), self.pcm_h[1].numpy()]
            self.outs = [self.out_h[0].numpy(), self.out_h[1].numpy()]
            self.enc_halo, self.dec_halo = sh.enc_halo, w.dec_halo

        def encode(self, c):
            if self.enc_halo == 0:
                got = c.encode_pcm_into(self.chans, self.su_h.numpy(), opts)
            else:
                got = c.encode_pcm_shard_into(self.chans, self.enc_halo, self.su_h.numpy(), opts)
            assert got == self.w.frames * 2

        def decode(self, c):
            n_units = self.su_in_h.numel() // 212
            if self.dec_halo == 0:
                c.decode_su_into(self.su_in_h.numpy(), n_units, 2, self.outs)
            else:
                c.decode_su_shard_into(self.su_in_h.numpy(), n_units, 2, self.dec_halo, self.outs)

    hosts = [HostShard(w) for w in works]
    h2d_bytes = sum(2 * h.n * 4 + h.su_in_h.numel() for h in hosts)
    d2h_bytes = sum(h.su_h.numel() + 2 * h.w.frames * 512 * 4 for h in hosts)
    ctx2 = carta1_b200.Context(local_rank)
    if args.units_per_pass:
        ctx2.set_max_units_per_pass(args.units_per_pass)
    errs = []

    def e2e_seq_step():
        for h in hosts:
            h.encode(ctx)
            h.decode(ctx)

    # The same calls, in flight together: an encoder thread and a decoder thread, each on its own context (calls on one
    # context are serialised by the library, include/carta1_b200.h).  carta1_encode_pcm is H2D-heavy and carta1_decode_su
    # D2H-heavy, so together they use both directions of the PCIe link.  Work per step is unchanged.
    def e2e_duplex_step():
        def dec():
            try:
                for h in hosts:
                    h.decode(ctx2)
            except Exception as ex:  # surfaced after the join
                errs.append(ex)

        th = threading.Thread(target=dec)
        th.start()
        for h in hosts:
            h.encode(ctx)
        th.join()
        if errs:
            raise errs[0]

    def wall_ms(fn, steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1000.0

    e2e_steps = max(1, min(args.steps, 5))
    e2e_seq_step()
    barrier()
    ms_e2e_seq = wall_ms(e2e_seq_step, e2e_steps)
    barrier()
    e2e_duplex_step()
    barrier()
    ms_e2e = wall_ms(e2e_duplex_step, e2e_steps)
    barrier()
    host_ok = all(bool(torch.equal(h.su_h, h.w.su_out.cpu())) and bool(torch.equal(h.out_h.view(torch.int32), h.w.out.cpu().view(torch.int32)))
                  for h in hosts)
    link = link_probe(torch, dev, barrier, world, dist)
    extra = {}
    if world == 1:
        h0 = hosts[0]
        n = h0.n
        frames = my_frames
        # WAV-shaped I/O: int16 interleaved PCM in and out, as the reference CLI reads and writes it
        # (bin/cli.js:394-404, processor.js:382-389); the conversions run inside the QMF kernels, so the PCM side of
        # the link carries half the bytes.  Same two calls in flight together.
        wav_h = torch.empty(n * 2, dtype=torch.int16).pin_memory()
        wav_h.copy_((works[0].pcm[:, :n].t().contiguous().clamp(-1.0, 1.0) * 32767.0).to(torch.int16).reshape(-1))
        wav_out_h = torch.empty(frames * 512 * 2, dtype=torch.int16).pin_memory()
        wav_np, wav_out_np = wav_h.numpy(), wav_out_h.numpy()
        su3_h = torch.empty(n_su * 212, dtype=torch.uint8).pin_memory()
        su3_np = su3_h.numpy()
        su2_np = h0.su_in_h.numpy()

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
        ms_e2e_s16 = wall_ms(e2e_s16_step, e2e_steps)
        # pageable caller arrays (what a host runtime that cannot pin its typed arrays passes): staged through the
        # context's pinned bounce slots by parallel memcpy; one call after the other
        chans_pg = [np.array(c) for c in h0.chans]
        outs_pg = [np.empty_like(o) for o in h0.outs]
        su_pg = np.empty_like(h0.su_h.numpy())

        def e2e_pageable_step():
            got = ctx.encode_pcm_into(chans_pg, su_pg, opts)
            assert got == n_su
            ctx.decode_su_into(su_pg, n_su, 2, outs_pg)

        e2e_pageable_step()
        ms_e2e_pg = wall_ms(e2e_pageable_step, e2e_steps)
        pageable_equal = bool(np.array_equal(su_pg, h0.su_h.numpy())) and all(
            np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(outs_pg, h0.outs))
        extra = {
            "wav_int16": {"value": seconds / (ms_e2e_s16 / e2e_steps / 1000.0), "unit": UNIT,
                          "h2d_bytes_per_step": int(2 * n * 2 + n_su * 212), "d2h_bytes_per_step": int(n_su * 212 + 2 * frames * 512 * 2),
                          "api": "carta1_encode_pcm_s16 || carta1_decode_su_s16 (WAV-shaped int16 PCM in and out, the reference CLI's I/O type)"},
            "pageable": {"value": seconds / (ms_e2e_pg / e2e_steps / 1000.0), "unit": UNIT,
                         "api": "carta1_encode_pcm then carta1_decode_su with pageable (unpinned) caller arrays",
                         "identical_to_pinned_run": pageable_equal},
        }
    ctx2.close()
    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    # ---- cross-rank check, outside the timed region: rank 0 runs every stream unsharded on its own GPU and compares the
    # gathered shard outputs with it byte for byte (the gather is verification traffic, not part of the data path)
    sharded_identical = None
    cuts = sum(1 for shards in plan for sh in shards if sh.begin > 0)
    check_detail = None
    if world > 1:
        ok = True
        bad = []
        for si, (sec_i, n_i, fr_i) in enumerate(zip(lens, n_samples, stream_frames)):
            ref_su = ref_out = None
            if rank == 0:
                whole = torch.zeros((2, fr_i * 512), dtype=torch.float32, device=dev)
                whole[:, :n_i] = synth_cfg2_span(torch, 0xCA27A2 + si, sec_i, 0, n_i, dev)
                ref_su = torch.zeros(fr_i * 2 * 212, dtype=torch.uint8, device=dev)
                ref_out = torch.zeros((2, fr_i * 512), dtype=torch.float32, device=dev)
                torch.cuda.synchronize()  # torch's stream wrote the PCM, the context's stream reads it
                ctx.encode_device(whole.data_ptr(), fr_i * 512, 2, n_i, 0, fr_i, opts, ref_su.data_ptr(), 2, 1)
                ctx.decode_device(ref_su.data_ptr(), 2, 1, fr_i * 2, 2, 0, fr_i, ref_out.data_ptr(), fr_i * 512, sync=True)
                del whole
            for r in range(world):
                for sh in plan[r]:
                    if sh.stream != si:
                        continue
                    w = next((x for x in works if x.sh == sh), None) if r == rank else None
                    if rank == 0:
                        if r == 0:
                            su_part, out_part = w.su_out, w.out
                        else:
                            su_part = torch.empty(sh.frames * 2 * 212, dtype=torch.uint8, device=dev)
                            out_part = torch.empty((2, sh.frames * 512), dtype=torch.float32, device=dev)
                            dist.recv(su_part, src=r)
                            dist.recv(out_part, src=r)
                        ok_su = bool(torch.equal(su_part, ref_su[sh.begin * 2 * 212:sh.end * 2 * 212]))
                        ok_pcm = bool(torch.equal(out_part.view(torch.int32), ref_out[:, sh.begin * 512:sh.end * 512].contiguous().view(torch.int32)))
                        if not (ok_su and ok_pcm):
                            ne = (su_part != ref_su[sh.begin * 2 * 212:sh.end * 2 * 212]).nonzero()
                            bad.append({"rank": r, "stream": si, "begin": sh.begin, "end": sh.end, "su": ok_su, "pcm": ok_pcm,
                                        "first_su_byte": int(ne[0].item()) if ne.numel() else None, "su_bytes_differing": int(ne.numel())})
                        ok = ok and ok_su and ok_pcm
                    elif r == rank:
                        torch.cuda.synchronize()
                        dist.send(w.su_out, dst=0)
                        dist.send(w.out, dst=0)
            del ref_su, ref_out
        flags = torch.tensor([1 if local_ok else 0, 1 if host_ok else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        sharded_identical = bool(ok) and bool(flags.min().item() == 1) if rank == 0 else None
        check_detail = {"gathered_shards_equal_unsharded": bool(ok), "timed_encode_equals_decode_input_on_every_rank": bool(flags[0].item() == 1),
                        "host_api_equals_device_on_every_rank": bool(flags[1].item() == 1), "mismatches": bad[:8]}

    if dist is not None:
        t = torch.tensor([ms, ms_enc, ms_dec, ms_e2e, ms_e2e_seq], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_enc, ms_dec, ms_e2e, ms_e2e_seq = t.tolist()
        lt = torch.tensor([launches, h2d_bytes, d2h_bytes, n_su], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches, h2d_all, d2h_all, n_su_all = [int(x) for x in lt.tolist()]
    else:
        h2d_all, d2h_all, n_su_all = h2d_bytes, d2h_bytes, n_su

    line = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_step = ms / args.steps
        value = total_seconds / (ms_step / 1000.0)
        # dominant kernel (rank 0's launches; every rank runs the same kernels over the same number of units)
        per_step = {k: v[0] / max(args.steps, 1) for k, v in prof.items()}
        dom = max(per_step.items(), key=lambda kv: kv[1]) if per_step else ("none", 0.0)
        dom_ms = dom[1]                                       # ms per step of that kernel over this rank's units
        step_ms_prof = sum(per_step.values())
        achieved = BYTES_PER_SU * n_su / (dom_ms / 1000.0) / 1e9 if dom_ms > 0 else 0.0
        step_gbs = BYTES_PER_AUDIO_SEC * total_seconds / world / (ms_step / 1000.0) / 1e9
        cap, cap_src, cap_fresh = ncu_capture()
        traffic = traffic_all = None
        if cap and cap_fresh:
            scale = n_su / float(cap["sound_units"])
            if dom[0] in cap["kernels"]:
                k = cap["kernels"][dom[0]]
                traffic = int((k["dram_bytes_read"] + k["dram_bytes_write"]) * scale)
            traffic_all = {name: int((k["dram_bytes_read"] + k["dram_bytes_write"]) * scale) for name, k in cap["kernels"].items()}
        fp64_su, fp64_src = fp64_per_unit(cap, cap_fresh)
        props = torch.cuda.get_device_properties(dev)
        clk_mhz = sampler.max_mhz or 1965
        fp64_floor_ms = 2 * n_su * 2.0 * fp64_su / (props.multi_processor_count * 4 * clk_mhz * 1e6) * 1e3
        duplex_gbs = max(h2d_all, d2h_all) / (ms_e2e / e2e_steps / 1000.0) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if world == 1 else WORKLOAD_N, "audio_seconds_per_gpu": total_seconds / world,
                       "sound_units_per_gpu": n_su_all // world, "audio_seconds_total": total_seconds,
                       "l2": "inputs larger than L2 (1.27 GB PCM + 131 MB units per pass and GPU)",
                       "streams_hours": [round(s / 3600.0, 4) for s in lens],
                       "sharding": "carta1_b200.sharding.plan: one contiguous span of the stream-major frame order per rank, "
                                   "2-frame PCM halo (encode) / 1-unit halo (decode) where a span starts inside a stream; no collective",
                       "shards_per_rank": [len(p) for p in plan], "frame_range_cuts": cuts,
                       "reference_arm_note": "--impl reference times a bounded sample of the same recipe (a rate, audio-s/s): same metric, smaller sample"},
            "sharded_output_identical": sharded_identical if world > 1 else True,
            "sharded_check": ("rank 0 encodes+decodes every stream unsharded and compares the gathered shard outputs (sound units and PCM) "
                              "byte for byte; host-API shard outputs equal the device ones on every rank") if world > 1
                             else "single shard (N = 1): the plan is the whole stream",
            "sharded_check_detail": check_detail,
            "encode_only": {"value": total_seconds / (ms_enc / args.steps / 1000.0), "unit": UNIT},
            "decode_only": {"value": total_seconds / (ms_dec / args.steps / 1000.0), "unit": UNIT},
            "roofline": {
                "bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": cap_src if (cap and cap_fresh) else None,
                "traffic_stale": bool(cap) and not cap_fresh, "peak_source": peak_src,
                "traffic_all_kernels": traffic_all,
                "kernel_ms": dom_ms, "kernel_share_of_step": dom_ms / step_ms_prof if step_ms_prof else None,
                "algorithmic_bytes_per_launch": BYTES_PER_SU * n_su,
                "step": {"achieved": step_gbs, "frac": step_gbs / peak,
                         "algorithmic_bytes_per_step": BYTES_PER_AUDIO_SEC * total_seconds / world,
                         "traffic": sum(traffic_all.values()) if traffic_all else None},
                "kernels_ms_per_step": per_step,
                # the bound that actually binds under the bit-exact FP64 contract
                "fp64_issue": {"floor_ms_per_step": fp64_floor_ms, "frac": fp64_floor_ms / ms_step,
                               "fp64_warp_instr_per_unit_per_direction": fp64_su, "source": fp64_src,
                               "sms": props.multi_processor_count, "sm_mhz": clk_mhz},
            },
            "e2e": {"value": total_seconds / (ms_e2e / e2e_steps / 1000.0), "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                    "steps": e2e_steps,
                    "api": "carta1_encode_pcm || carta1_decode_su (shards inside a stream: carta1_encode_pcm_shard || carta1_decode_su_shard): "
                           "both calls in flight (two host threads, one context each), pinned host buffers, H2D and D2H inside; "
                           "host wall clock, max over ranks",
                    "sequential": {"value": total_seconds / (ms_e2e_seq / e2e_steps / 1000.0), "unit": UNIT,
                                   "api": "the same calls one after the other on one context"},
                    "host_output_identical_to_device_run": host_ok,
                    # the raw pinned-copy rate of the host link, all ranks copying at once (GB/s per direction, whole box)
                    "link_ceiling_gbs": link, "achieved_gbs_per_direction": duplex_gbs,
                    "frac_of_link_ceiling": duplex_gbs / link["duplex"] if link.get("duplex") else None},
            "gpu_launches": launches,
            "clocks": sampler.summary(t_start, t_end),
        }
        if args.lib:
            line["experiment_library"] = args.lib  # not a bench line: a variant build was timed
        line["e2e"].update(extra)
        # ---- CPU baseline leg (rank 0, N=1): the oracle on the same data, doubling as the in-bench parity check
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O

            O.build()
            threads = os.cpu_count() or 1
            h0 = hosts[0]
            sample_s = min(seconds, float(args.cpu_sample_seconds))
            ns = int(sample_s * SR) // 512 * 512
            chans = [np.ascontiguousarray(h0.pcm_h[c, :ns].numpy()) for c in range(2)]
            oopts = O.make_options(fixed_modes=None if args.auto_modes else [0, 0, 0])
            su_ref, pcm_ref, te, td = cpu_pass(O, chans, oopts, threads)
            k = su_ref.shape[0]
            # frame f depends on samples <= 512 f + 511 only, so the prefix must agree exactly
            parity = bool(np.array_equal(h0.su_h.numpy().reshape(-1, 212)[:k], su_ref))
            gpu_pcm_ok = all(np.array_equal(h0.outs[c][: (k // 2) * 512].view(np.uint32),
                                            pcm_ref[c][: (k // 2) * 512].view(np.uint32)) for c in range(2))
            line["cpu_baseline"] = {
                "value": (ns / SR) / (te + td), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": "first %.0f s of the same stereo PCM, encode+decode, %d threads (C restatement of the reference, itself pinned to the reference's output: reference_pin)" % (ns / SR, threads),
                "encode_only": (ns / SR) / te, "decode_only": (ns / SR) / td,
                "gpu_output_bit_exact_on_sample": parity and gpu_pcm_ok and local_ok and host_ok,
            }
            # ---- and against the reference itself: the bytes carta1's own JavaScript wrote for the golden inputs
            # (tests/golden/ref, produced by tools/ref_run_qjs.py in the build image; /root/reference does not exist
            # on this box, so the committed output is what the CUDA path is held to)
            from oracle import refpin

            if refpin.available():
                from carta1_b200._lib import Tables

                try:
                    line["reference_pin"] = refpin.check_context(
                        lambda t: carta1_b200.Context(torch.cuda.current_device(), t),
                        lambda thr, bias, fixed, bsf: carta1_b200.make_enc_opts(thr, bias, fixed, biased_scale_factors=bsf), Tables)
                except AssertionError as ex:
                    line["reference_pin"] = {"equal_to_reference_output": False, "first_difference": str(ex)}
    del works, hosts
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs, rank 0 only (the other ranks wait at the barrier)
    if rank == 0 and not args.no_configs:
        line["configs"] = run_configs(torch, np, carta1_b200, ctx, dev, args)
        line["configs"]["cfg2"] = {"workload": WORKLOAD, "see": "the top-level keys of this line" if world == 1 else "the N = 1 line"}
        line["configs"]["cfg5"] = {"workload": WORKLOAD_N, "see": "the top-level keys of the N > 1 lines (sharded_output_identical, frame_range_cuts)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    barrier()
    for c in ctxs:
        c.close()
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
    ap.add_argument("--no-configs", action="store_true", help="development aid: skip the cfg1 / cfg3 / cfg4 block")
    ap.add_argument("--cfg3-seconds", type=float, default=3600.0, help="development aid: length of the cfg3 stream (default: its 1 h)")
    ap.add_argument("--units-per-pass", type=int, default=0,
                    help="development aid: sound units per pipelined pass of the host entry points (default: the library's)")
    ap.add_argument("--lib", default="", help="development aid: time a variant build of the library (its output is not checked)")
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
