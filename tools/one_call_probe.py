import sys, numpy as np, torch
sys.path.insert(0, ".")
import carta1_b200
ctx = carta1_b200.Context(0)
rng = np.random.default_rng(1)
enc = carta1_b200.StreamEncoder(ctx, carta1_b200.make_enc_opts(), 1)
dec = carta1_b200.StreamDecoder(ctx, 1)
pcm = (0.3 * rng.standard_normal((1, 1, 512))).astype(np.float32)
for _ in range(12):
    su = enc.frames(pcm); dec.frames(su)
