#!/usr/bin/env python
"""Development aid: latency of the reference's frame closures (encode()(pcm), decode()(frame)): one stream, one
frame per call through carta1_enc_frames / carta1_dec_frames, pinned host buffers."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import carta1_b200  # noqa: E402

ctx = carta1_b200.Context(0)
rng = np.random.default_rng(1)
for ns, nf in ((1, 1), (2, 1), (1, 8), (2, 64)):
    for fixed in (None, [0, 0, 0]):
        enc = carta1_b200.StreamEncoder(ctx, carta1_b200.make_enc_opts(fixed_block_modes=fixed), ns)
        dec = carta1_b200.StreamDecoder(ctx, ns)
        pcm_t = torch.empty((ns, nf, 512), dtype=torch.float32).pin_memory()
        su_t = torch.empty((ns, nf, 212), dtype=torch.uint8).pin_memory()
        out_t = torch.empty((ns, nf, 512), dtype=torch.float32).pin_memory()
        pcm, su_buf, out_buf = pcm_t.numpy(), su_t.numpy(), out_t.numpy()
        pcm[:] = (0.3 * rng.standard_normal((ns, nf, 512))).astype(np.float32)
        for _ in range(20):
            su = enc.frames(pcm, su_buf)
            dec.frames(su, out_buf)
        n = 300
        t0 = time.perf_counter()
        for _ in range(n):
            su = enc.frames(pcm, su_buf)
        t1 = time.perf_counter()
        for _ in range(n):
            dec.frames(su, out_buf)
        t2 = time.perf_counter()
        print("streams %d frames/call %2d %-12s encode %6.1f us/call  decode %6.1f us/call" % (
            ns, nf, "fixed modes" if fixed else "auto modes", 1e6 * (t1 - t0) / n, 1e6 * (t2 - t1) / n), flush=True)
        enc.close()
        dec.close()

enc = carta1_b200.StreamEncoder(ctx, carta1_b200.make_enc_opts(), 1)
dec = carta1_b200.StreamDecoder(ctx, 1)
pcm = (0.3 * rng.standard_normal((1, 1, 512))).astype(np.float32)
for _ in range(10):
    su = enc.frames(pcm)
    dec.frames(su)
ctx.profile(True)
for _ in range(50):
    su = enc.frames(pcm)
    dec.frames(su)
prof = ctx.profile_read()
ctx.profile(False)
print("kernel us per call (1 unit): " + "  ".join("%s %.1f" % (k, 1e3 * v[0] / 50) for k, v in prof.items()))
