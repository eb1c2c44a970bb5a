#!/usr/bin/env python
"""Development aid (not a bench line): BASELINE config 4 -- 4096 independent mono streams advanced
together through the stateful frame API (carta1_enc_frames / carta1_dec_frames), host buffers in
and out, for several frames-per-call values.  Prints audio-seconds per second."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import carta1_b200  # noqa: E402

N_STREAMS = 4096


def main():
    ctx = carta1_b200.Context(0)
    rng = np.random.default_rng(4)
    for nf in (1, 8, 64):
        enc = carta1_b200.StreamEncoder(ctx, None, N_STREAMS)
        dec = carta1_b200.StreamDecoder(ctx, N_STREAMS)
        pcm_t = torch.empty((N_STREAMS, nf, 512), dtype=torch.float32).pin_memory()
        su_t = torch.empty((N_STREAMS, nf, 212), dtype=torch.uint8).pin_memory()
        out_t = torch.empty((N_STREAMS, nf, 512), dtype=torch.float32).pin_memory()
        pcm, su_buf, out_buf = pcm_t.numpy(), su_t.numpy(), out_t.numpy()
        pcm[:] = (0.3 * rng.standard_normal((N_STREAMS, nf, 512))).astype(np.float32)
        calls = max(4, 256 // nf)
        su = enc.frames(pcm, su_buf)
        dec.frames(su, out_buf)
        t0 = time.perf_counter()
        for _ in range(calls):
            su = enc.frames(pcm, su_buf)
        t1 = time.perf_counter()
        for _ in range(calls):
            dec.frames(su, out_buf)
        t2 = time.perf_counter()
        audio = calls * N_STREAMS * nf * 512 / 44100.0
        print("frames/call %3d: encode %9.0f audio-s/s (%.2f ms/call)   decode %9.0f audio-s/s (%.2f ms/call)" % (
            nf, audio / (t1 - t0), 1e3 * (t1 - t0) / calls, audio / (t2 - t1), 1e3 * (t2 - t1) / calls))
        ctx.profile(True)
        for _ in range(20):
            su = enc.frames(pcm, su_buf)
            dec.frames(su, out_buf)
        prof = ctx.profile_read()
        ctx.profile(False)
        print("      kernel ms per call: " + "  ".join("%s %.3f" % (k, v[0] / 20) for k, v in prof.items()))
        enc.close()
        dec.close()
    ctx.close()


if __name__ == "__main__":
    main()
