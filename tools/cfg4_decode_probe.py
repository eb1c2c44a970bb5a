#!/usr/bin/env python
"""Development aid: where a one-frame decode call of 4096 streams (carta1_dec_frames) spends its wall time, against
the same copies issued through torch on one stream: is the run-to-run spread of the call in the copies or in the call?"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import carta1_b200  # noqa: E402

NS = 4096
ctx = carta1_b200.Context(0)
rng = np.random.default_rng(4)
enc = carta1_b200.StreamEncoder(ctx, None, NS)
dec = carta1_b200.StreamDecoder(ctx, NS)
pcm_t = torch.empty((NS, 1, 512), dtype=torch.float32).pin_memory()
su_t = torch.empty((NS, 1, 212), dtype=torch.uint8).pin_memory()
out_t = torch.empty((NS, 1, 512), dtype=torch.float32).pin_memory()
pcm_t.numpy()[:] = (0.3 * rng.standard_normal((NS, 1, 512))).astype(np.float32)
su = enc.frames(pcm_t.numpy(), su_t.numpy())
dec.frames(su, out_t.numpy())
d_su = torch.empty((NS, 1, 212), dtype=torch.uint8, device="cuda")
d_out = torch.empty((NS, 1, 512), dtype=torch.float32, device="cuda")
s = torch.cuda.Stream()
N = 256


def copies_only():
    with torch.cuda.stream(s):
        d_su.copy_(su_t, non_blocking=True)
        out_t.copy_(d_out, non_blocking=True)
    s.synchronize()


def call():
    dec.frames(su, out_t.numpy())


def enc_call():
    enc.frames(pcm_t.numpy(), su_t.numpy())


for rep in range(3):
    for name, fn in (("copies only (H2D 0.87 MB, D2H 8.4 MB, sync)", copies_only), ("carta1_dec_frames", call), ("carta1_enc_frames", enc_call),
                     ("carta1_dec_frames right after the encode loop", call)):
        fn()
        ts = []
        for _ in range(N):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print("rep %d %-46s mean %.3f ms  median %.3f  p10 %.3f  p90 %.3f  max %.3f" % (
            rep, name, ts.mean(), np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), ts.max()), flush=True)
