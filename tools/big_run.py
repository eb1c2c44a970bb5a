#!/usr/bin/env python
"""Development aid: config 5's per-GPU share at full size (12.5 h of stereo = 5.4 M sound units, 15.9 GB of PCM)
through the device-resident entry points in ONE launch sequence, to catch 32-bit index overflow.  The first and
the last 30 s are checked bit for bit against the oracle (a frame depends on the 2 frames before it only:
SURVEY.md Appendix B, so the tail can be encoded on its own with 2 frames of lead-in; the decoder needs 1 unit)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import carta1_b200  # noqa: E402
from oracle import oracle as O  # noqa: E402

hours = float(sys.argv[1]) if len(sys.argv) > 1 else 12.5
seconds = hours * 3600.0
dev = torch.device("cuda", 0)
O.build()
ctx = carta1_b200.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
pcm = bench.synth_cfg2_device(torch, seconds, 77, dev)
n = pcm.shape[1]
frames = (n + 511) // 512
n_su = 2 * frames
print("%.1f h stereo: %d samples per channel, %d sound units, PCM %.1f GB" % (hours, n, n_su, pcm.numel() * 4 / 1e9), flush=True)
d_su = torch.zeros(n_su * 212, dtype=torch.uint8, device=dev)
d_out = torch.zeros((2, frames * 512), dtype=torch.float32, device=dev)
opts = carta1_b200.make_enc_opts(fixed_block_modes=[0, 0, 0])
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for rep in range(2):
    e0.record(stream)
    ctx.encode_device(pcm.data_ptr(), n, 2, n, 0, frames, opts, d_su.data_ptr(), 2, 1)
    e1.record(stream)
    ctx.decode_device(d_su.data_ptr(), 2, 1, n_su, 2, 0, frames, d_out.data_ptr(), frames * 512)
    e2.record(stream)
    e2.synchronize()
print("encode %.1f ms (%.0f audio-s/s), decode %.1f ms (%.0f audio-s/s)" % (
    e0.elapsed_time(e1), seconds / e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2), seconds / e1.elapsed_time(e2) * 1e3), flush=True)
oo = O.make_options(fixed_modes=[0, 0, 0])
thr = os.cpu_count() or 1
k = int(30 * 44100) // 512  # frames checked at each end
ok = True
# head
head = [np.ascontiguousarray(pcm[c, :k * 512].cpu().numpy()) for c in range(2)]
su_ref = O.encode_pcm(head, oo, threads=thr, chunk_frames=256)
ok &= bool(np.array_equal(d_su.view(-1, 212)[:2 * k].cpu().numpy(), su_ref))
pcm_ref = O.decode_su(su_ref, 2, threads=thr, chunk_frames=256)
ok &= all(np.array_equal(d_out[c, :k * 512].cpu().numpy().view(np.uint32), pcm_ref[c].view(np.uint32)) for c in range(2))
print("head bit-exact:", ok, flush=True)
# tail: 2 frames of lead-in for the encoder state; the oracle starts from silence, so its first 2 frames are dropped
f0 = frames - k - 2
tail = [np.ascontiguousarray(pcm[c, f0 * 512:].cpu().numpy()) for c in range(2)]
su_tail = O.encode_pcm(tail, oo, threads=thr, chunk_frames=256)
got = d_su.view(-1, 212)[2 * (f0 + 2):].cpu().numpy()
ok_t = bool(np.array_equal(got, su_tail[4:]))
# decoder: 1 unit of lead-in per channel
su_dec = d_su.view(-1, 212)[2 * (f0 + 1):].cpu().numpy()
pcm_tail = O.decode_su(np.ascontiguousarray(su_dec), 2, threads=thr, chunk_frames=256)
ok_t &= all(np.array_equal(d_out[c, (f0 + 2) * 512:].cpu().numpy().view(np.uint32), pcm_tail[c][512:].view(np.uint32)) for c in range(2))
print("tail bit-exact:", ok_t, flush=True)
sys.exit(0 if ok and ok_t else 1)
