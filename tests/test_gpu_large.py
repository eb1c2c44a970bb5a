"""Full-size indexing (cfg5 puts 15.9 GB of PCM on one GPU): one launch sequence over 10 GB of stereo PCM
(2.5 G samples: flat element indices pass 2^31, byte offsets pass 2^32 in every f32 intermediate), checked bit for bit
against the oracle on three spans -- the head, the span of channel 1 where the flat element index crosses 2^31, and the
tail.  Every frame is a pure function of a short window (2 frames of PCM for the encoder, 1 sound
unit for the decoder), so the oracle can reproduce any span from its own lead-in.  Size-independent properties ride
along: block modes stay the fixed ones, every unit spends its bit budget within 212 bytes, decode of encode keeps the
reference's 266-sample delay and stays close to the input.
"""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SR = 44100


def test_ten_gigabytes_in_one_launch_sequence(oracle):
    import torch

    import bench
    import carta1_b200

    free, _ = torch.cuda.mem_get_info(0)
    if free < 90e9:
        pytest.skip("needs ~70 GB of device memory, %.0f GB free" % (free / 1e9))
    O = oracle
    dev = torch.device("cuda", 0)
    n = 1_250_000_000 // 512 * 512          # samples per channel: 7.9 h
    frames = n // 512
    n_su = 2 * frames
    assert 2 * n > 2 ** 31 and 2 * n * 4 > 2 ** 32
    pcm = bench.synth_cfg2_span(torch, 5, n / SR, 0, n, dev)   # [2, n] f32: sines + chirp + hashed Gaussian noise
    ctx = carta1_b200.Context(0)
    try:
        d_su = torch.zeros(n_su * 212, dtype=torch.uint8, device=dev)
        d_out = torch.zeros((2, n), dtype=torch.float32, device=dev)
        opts = carta1_b200.make_enc_opts(fixed_block_modes=[0, 0, 0])
        torch.cuda.synchronize()
        ctx.encode_device(pcm.data_ptr(), n, 2, n, 0, frames, opts, d_su.data_ptr(), 2, 1)
        ctx.decode_device(d_su.data_ptr(), 2, 1, n_su, 2, 0, frames, d_out.data_ptr(), n)
        ctx.sync()
        su = d_su.view(frames, 2, 212)
        oo = O.make_options(fixed_modes=[0, 0, 0])
        k = 400                               # frames per checked span
        cross = (2 ** 31 - n) // 512          # the frame of channel 1 whose flat element index n + 512 f crosses 2^31
        assert 0 < cross < frames
        for name, f0 in (("head", 0), ("2^31 crossing", cross - k // 2), ("tail", frames - k)):
            lead = min(f0, 2)                 # encoder lead-in frames (state is two frames deep)
            span = [np.ascontiguousarray(pcm[c, (f0 - lead) * 512:(f0 + k) * 512].cpu().numpy()) for c in range(2)]
            want = O.encode_pcm(span, oo, threads=8, chunk_frames=64).reshape(-1, 2, 212)[lead:]
            got = su[f0:f0 + k].cpu().numpy()
            assert np.array_equal(got, want), (name, "sound units")
            dl = min(f0, 1)                   # decoder lead-in: one unit per channel
            units = np.ascontiguousarray(su[f0 - dl:f0 + k].cpu().numpy()).reshape(-1, 212)
            ref = O.decode_su(units, 2, threads=8, chunk_frames=64)
            for c in range(2):
                a = d_out[c, f0 * 512:(f0 + k) * 512].cpu().numpy()
                assert np.array_equal(a.view(np.uint32), ref[c][dl * 512:].view(np.uint32)), (name, "pcm", c)
        # size-independent properties over the whole run, computed on the device
        hdr = su[:, :, 0].to(torch.int32)
        assert int(((hdr >> 2) & 0x3F).ne(0b101011).sum()) == 0, "block-mode fields of fixed long blocks: 2-0, 2-0, 3-0"
        assert int(su[:, :, 209:].to(torch.int32).abs().sum()) == 0, "the last three bytes of a sound unit are zero (serialization.js)"
        delay = 266                           # tests/encoder-decoder round trip of the reference: qmf.test.js / README
        step = 1 << 26
        worst, sq = 0.0, 0.0
        for lo in range(0, n - delay, step):
            hi = min(n - delay, lo + step)
            err = d_out[:, lo + delay:hi + delay] - pcm[:, lo:hi]
            worst = max(worst, err.abs().max().item())
            sq += float((err.double() ** 2).sum().item())
        rms = math.sqrt(sq / (2 * (n - delay)))
        print("round trip over %.1f h per channel: worst |error| %.4f, rms %.5f" % (n / SR / 3600, worst, rms))
        assert math.isfinite(worst) and worst < 0.5 and rms < 0.05, "round trip error: worst %r, rms %r" % (worst, rms)
        assert bool(torch.isfinite(d_out).all())
    finally:
        ctx.close()
