#!/usr/bin/env python
"""Development aid (not the bench line): the BASELINE.json configs other than configs[1], measured the same
way -- cfg1 (10 s stereo, auto block modes, through the host API), cfg3 (1 h stereo transient-heavy, auto block
modes, device-resident; fraction of short-block frames; first 120 s checked bit for bit against the oracle).
cfg4 is tools/bench_cfg4.py, cfg5 is `bench.py --gpus N` (every rank its own hours).  Prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import carta1_b200  # noqa: E402
import signals as S  # noqa: E402

SR = 44100


def dev_timed(ctx, fn, reps):
    stream = torch.cuda.ExternalStream(ctx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    ctx.sync()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def device_roundtrip(ctx, chans, opts, reps):
    dev = torch.device("cuda", 0)
    n = len(chans[0])
    frames = (n + 511) // 512
    pcm = torch.from_numpy(np.stack(chans)).to(dev)
    d_su = torch.zeros(frames * 2 * 212, dtype=torch.uint8, device=dev)
    d_out = torch.zeros((2, frames * 512), dtype=torch.float32, device=dev)
    enc = lambda: ctx.encode_device(pcm.data_ptr(), n, 2, n, 0, frames, opts, d_su.data_ptr(), 2, 1)  # noqa: E731
    dec = lambda: ctx.decode_device(d_su.data_ptr(), 2, 1, frames * 2, 2, 0, frames, d_out.data_ptr(), frames * 512)  # noqa: E731
    ms_e, ms_d = dev_timed(ctx, enc, reps), dev_timed(ctx, dec, reps)
    ctx.profile(True)
    enc(); dec()
    prof = ctx.profile_read()
    ctx.profile(False)
    return ms_e, ms_d, d_su, d_out, {k: round(v[0], 4) for k, v in prof.items()}


def main():
    from oracle import oracle as O

    O.build()
    ctx = carta1_b200.Context(0)
    threads = os.cpu_count() or 1
    # ---- cfg1
    chans = S.cfg1_stereo(10.0)
    opts = carta1_b200.make_enc_opts()
    su = ctx.encode_pcm(chans, opts)
    t0 = time.perf_counter()
    for _ in range(20):
        su = ctx.encode_pcm(chans, opts)
    t_enc = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for _ in range(20):
        pcm = ctx.decode_su(su, 2)
    t_dec = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    want = O.encode_pcm(chans, O.make_options(), threads=threads, chunk_frames=64)
    t_cpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    want1 = O.encode_pcm(chans, O.make_options(), threads=1)
    t_cpu1 = time.perf_counter() - t0
    ms_e, ms_d, _, _, prof = device_roundtrip(ctx, chans, opts, 20)
    print(json.dumps({"config": "cfg1: 10 s stereo sine+noise, default bias, auto block modes",
                      "host_api_encode_ms": 1e3 * t_enc, "host_api_decode_ms": 1e3 * t_dec,
                      "host_api_encode_audio_s_per_s": 10.0 / t_enc, "device_encode_ms": ms_e, "device_decode_ms": ms_d,
                      "oracle_encode_ms": {"threads_%d" % threads: 1e3 * t_cpu, "threads_1": 1e3 * t_cpu1},
                      "bit_exact": bool(np.array_equal(su, want) and np.array_equal(want, want1)), "kernels_ms": prof}), flush=True)
    # ---- cfg3
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    chans = S.cfg3_transients(seconds)
    ms_e, ms_d, d_su, d_out, prof = device_roundtrip(ctx, chans, opts, 5)
    hdr = d_su.view(-1, 212)[:, 0].to(torch.int32)
    m0, m1, m2 = 2 - ((hdr >> 6) & 3), 2 - ((hdr >> 4) & 3), 3 - ((hdr >> 2) & 3)
    short_any = ((m0 != 0) | (m1 != 0) | (m2 != 0)).float().mean().item()
    per_band = [float((m != 0).float().mean().item()) for m in (m0, m1, m2)]
    ns = min(int(120 * SR) // 512 * 512, len(chans[0]) // 512 * 512)
    head = [np.ascontiguousarray(c[:ns]) for c in chans]
    su_ref = O.encode_pcm(head, O.make_options(), threads=threads, chunk_frames=256)
    pcm_ref = O.decode_su(su_ref, 2, threads=threads, chunk_frames=256)
    k = su_ref.shape[0]
    su_gpu = d_su.view(-1, 212)[:k].cpu().numpy()
    out_gpu = d_out[:, :ns].cpu().numpy()
    ok = bool(np.array_equal(su_gpu, su_ref)) and all(np.array_equal(out_gpu[c].view(np.uint32), pcm_ref[c][:ns].view(np.uint32)) for c in range(2))
    print(json.dumps({"config": "cfg3: %.0f s stereo transient-heavy (clicks, 4 per second per channel), auto block modes" % seconds,
                      "device_encode_ms": ms_e, "device_decode_ms": ms_d,
                      "encode_audio_s_per_s": seconds / ms_e * 1e3, "decode_audio_s_per_s": seconds / ms_d * 1e3,
                      "roundtrip_audio_s_per_s": seconds / (ms_e + ms_d) * 1e3,
                      "short_block_frames": {"any_band": short_any, "low": per_band[0], "mid": per_band[1], "high": per_band[2]},
                      "bit_exact_first_120s": ok, "kernels_ms": prof}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
