#!/usr/bin/env python
"""Development aid (not a bench line): BASELINE config 4 -- 4096 independent mono streams advanced
together through the stateful frame API (carta1_enc_frames / carta1_dec_frames), host buffers in
and out, for several frames-per-call values.  Prints audio-seconds per second."""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import carta1_b200  # noqa: E402

N_STREAMS = 4096


class Clocks(threading.Thread):
    """SM / memory clock samples (NVML) while a loop runs: a duty cycle of a few percent lets the clocks idle down."""

    def __init__(self):
        super().__init__(daemon=True)
        import pynvml

        pynvml.nvmlInit()
        self.nv, self.h, self.s, self.stop = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0), [], False

    def run(self):
        while not self.stop:
            self.s.append((time.perf_counter(), self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                           self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_MEM)))
            time.sleep(0.002)

    def span(self, a, b):
        sel = [x for x in self.s if a <= x[0] <= b] or self.s[-1:]
        return "SM %d-%d MHz, mem %d-%d MHz (%d samples)" % (min(x[1] for x in sel), max(x[1] for x in sel),
                                                          min(x[2] for x in sel), max(x[2] for x in sel), len(sel))


def main():
    clocks = Clocks()
    clocks.start()
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
        te, td = [], []
        t0 = time.perf_counter()
        for _ in range(calls):
            a = time.perf_counter()
            su = enc.frames(pcm, su_buf)
            te.append(time.perf_counter() - a)
        t1 = time.perf_counter()
        for _ in range(calls):
            a = time.perf_counter()
            dec.frames(su, out_buf)
            td.append(time.perf_counter() - a)
        t2 = time.perf_counter()
        # rates from the MEDIAN call: the second or third call of a handle with a new call shape captures and instantiates
        # its CUDA graph, once, and that one call takes 1 - 90 ms (it made the mean of this loop look bimodal)
        audio = N_STREAMS * nf * 512 / 44100.0
        print("frames/call %3d: encode %9.0f audio-s/s (%.3f ms/call)   decode %9.0f audio-s/s (%.3f ms/call)" % (
            nf, audio / np.median(te), 1e3 * np.median(te), audio / np.median(td), 1e3 * np.median(td)))
        print("      per call, ms: encode median %.3f p90 %.3f max %.3f; decode median %.3f p90 %.3f max %.3f" % (
            1e3 * np.median(te), 1e3 * np.percentile(te, 90), 1e3 * max(te), 1e3 * np.median(td), 1e3 * np.percentile(td, 90), 1e3 * max(td)))
        print("      clocks during the encode loop: %s; during the decode loop: %s" % (clocks.span(t0, t1), clocks.span(t1, t2)))
        # the raw copy rates of these very buffers (same pinned pages, same link), so that a slow call can be told from
        # a slow link: the decode call is bound by the D2H copy of its PCM from 8 frames per call on
        d_out = torch.empty((N_STREAMS, nf, 512), dtype=torch.float32, device="cuda")
        d_in = torch.empty((N_STREAMS, nf, 512), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        r0 = time.perf_counter()
        for _ in range(calls):
            out_t.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        r1 = time.perf_counter()
        for _ in range(calls):
            d_in.copy_(pcm_t, non_blocking=True)
        torch.cuda.synchronize()
        r2 = time.perf_counter()
        gb = calls * out_t.numel() * 4 / 1e9
        print("      raw pinned copies of the same buffers: D2H %.1f GB/s (%.2f ms/call), H2D %.1f GB/s (%.2f ms/call)" % (
            gb / (r1 - r0), 1e3 * (r1 - r0) / calls, gb / (r2 - r1), 1e3 * (r2 - r1) / calls))
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
