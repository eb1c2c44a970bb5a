"""The reference's own pipeline tests (tests/encoder.test.js, decoder.test.js, processor.test.js)
restated against the GPU path through the host mirror of its JS surface (carta1_b200/codec.py),
plus bit-exactness of the same calls against the CPU oracle."""
import math

import numpy as np
import pytest

import signals as S

pytestmark = pytest.mark.gpu

FRAME_BITS, FRAME_OVERHEAD_BITS, BITS_PER_BFU_METADATA = 1696, 40, 10


@pytest.fixture(scope="module")
def K():
    from carta1_b200 import codec

    return codec


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_frame(K, d, ofr):
    n = ofr.n_bfu
    return (d["nBfu"] == n and list(d["blockModes"]) == list(ofr.modes) and
            np.array_equal(d["scaleFactorIndices"], np.array(ofr.sfi[:n])) and
            np.array_equal(d["wordLengthIndices"], np.array(ofr.wl[:n])) and
            all(np.array_equal(d["quantizedCoefficients"][b], np.array(ofr.q[b][:K.SPECS_PER_BFU[b]])) for b in range(n)))


# ---- tests/encoder.test.js -----------------------------------------------------------------
def test_encoder_full_pipeline(K, oracle):  # :14-24
    encoder = K.encode()
    pcm = S.sine(440)
    result = encoder(pcm)
    assert result["nBfu"] > 0
    for key in ("scaleFactorIndices", "wordLengthIndices", "quantizedCoefficients"):
        assert key in result
    ofr = oracle.FrameEncoder()(pcm)
    assert same_frame(K, result, ofr)


def test_encoder_short_blocks_on_transient(K, oracle):  # :26-91
    opts = K.EncoderOptions({"transientThresholdLow": 1, "transientThresholdMid": 1.5, "transientThresholdHigh": 2})
    encoder = K.encode(opts)
    oenc = oracle.FrameEncoder(oracle.make_options(threshold=1.0))
    silent = np.zeros(512, np.float32)
    i = np.arange(512, dtype=np.float64)
    x = sum(a * np.sin((2 * math.pi * f * i) / 44100) for a, f in
            ((0.8, 60), (0.7, 80), (0.6, 100), (0.5, 120), (0.4, 200), (0.3, 300), (0.3, 400), (0.2, 500)))
    for f in range(600, 5000, 200):
        x = x + 0.1 * np.sin((2 * math.pi * f * i) / 44100)
    burst = ((x * 0.95) / np.abs(x).max()).astype(np.float32)
    results = []
    for fr in (silent, silent, burst, silent):
        results.append(encoder(fr))
        assert same_frame(K, results[-1], oenc(fr))
    assert any(m != 0 for m in results[2]["blockModes"]) or any(m != 0 for m in results[3]["blockModes"])


def test_encoder_bit_budget(K):  # :93-107
    result = K.encode()(S.white_noise(1))
    used = sum(int(K.WORD_LENGTH_BITS[result["wordLengthIndices"][i]]) * len(result["quantizedCoefficients"][i])
               for i in range(result["nBfu"]))
    assert used + FRAME_OVERHEAD_BITS + result["nBfu"] * BITS_PER_BFU_METADATA <= FRAME_BITS


def test_encoder_silence(K):  # :109-120
    result = K.encode()(S.silence())
    total = sum(int(K.WORD_LENGTH_BITS[wl]) * len(result["quantizedCoefficients"][i])
                for i, wl in enumerate(result["wordLengthIndices"]))
    assert total == 0


def test_encoder_options_reach_the_kernels(K, oracle):
    pcm = S.cfg3_transients(0.2, n_ch=1)[0][:512 * 6]
    for kw, okw in (({"allocationBias": 2.5}, dict(bias=2.5)), ({"fixedBlockModes": [2, 0, 3]}, dict(fixed_modes=[2, 0, 3])),
                    ({"transientThresholdLow": 0.2}, dict(threshold=0.2))):
        enc, oenc = K.encode(K.EncoderOptions(kw)), oracle.FrameEncoder(oracle.make_options(**okw))
        for f in range(6):
            assert same_frame(K, enc(pcm[512 * f:512 * f + 512]), oenc(pcm[512 * f:512 * f + 512])), (kw, f)


# ---- tests/decoder.test.js -----------------------------------------------------------------
def test_decoder_full_pipeline(K):  # :8-17
    decoded = K.decode()(K.encode()(S.sine(440)))
    assert decoded.shape == (512,)


def test_codec_delay_266(K, oracle):  # :19-68
    encoder, decoder = K.encode(), K.decode()
    oenc, odec = oracle.FrameEncoder(), oracle.FrameDecoder()
    frames = [S.sine(440) for _ in range(5)]
    decoded = [decoder(encoder(f)) for f in frames]
    ref = [odec(oenc(f)) for f in frames]
    for a, b in zip(decoded, ref):
        assert np.array_equal(bits(a), bits(b))
    orig, dec = np.concatenate(frames), np.concatenate(decoded)
    n = len(orig) - 266
    assert np.abs(dec[266:266 + n] - orig[:n]).sum() / n < 0.1


def test_decoder_all_block_modes_unserialisable_frame(K):  # :70-84
    encoded = {"nBfu": 52, "scaleFactorIndices": np.full(52, 10, np.int32), "wordLengthIndices": np.full(52, 8, np.int32),
               "quantizedCoefficients": [np.ones(10, np.int32) for _ in range(52)], "blockModes": [1, 1, 1]}
    decoded = K.decode()(encoded)
    assert decoded.shape == (512,) and np.isfinite(decoded).all() and np.abs(decoded).max() > 0


def test_decoder_zero_word_lengths(K):  # :86-98
    encoded = {"nBfu": 52, "scaleFactorIndices": np.zeros(52, np.int32), "wordLengthIndices": np.zeros(52, np.int32),
               "quantizedCoefficients": [np.zeros(0, np.int32) for _ in range(52)], "blockModes": [0, 0, 0]}
    assert (K.decode()(encoded) == 0).all()


def test_decode_closure_equals_unit_path(K, oracle):
    """decode() goes through the position-expanded entry; on frames that do fit 212 bytes it must
    give exactly what the sound-unit path (and the oracle) gives."""
    pcm = S.cfg3_transients(0.25, n_ch=1)[0]
    su = oracle.encode_pcm([pcm])
    want = oracle.decode_su(su, 1)[0].reshape(-1, 512)
    decoder = K.decode()
    for f, u in enumerate(su):
        assert np.array_equal(bits(decoder(K.deserializeFrame(u))), bits(want[f])), f


# ---- tests/processor.test.js ----------------------------------------------------------------
def mono_stream(n):
    for _ in range(n):
        yield S.sine(440)


def stereo_stream(n):
    for _ in range(n):
        yield [S.sine(440), S.sine(880)]


def test_encode_stream_counts_and_progress(K):  # :23-47
    assert len(list(K.AudioProcessor.encodeStream(mono_stream(2), {"channelCount": 1}))) == 2
    seen = []
    frames = list(K.AudioProcessor.encodeStream(stereo_stream(2), {"channelCount": 2, "onProgress": seen.append}))
    assert len(frames) == 4 and seen == [0, 1]
    with pytest.raises(ValueError, match="Unsupported channel count: 3"):
        list(K.AudioProcessor.encodeStream(mono_stream(1), {"channelCount": 3}))


def test_streams_match_oracle_across_batches(K, oracle, monkeypatch):
    """Yield order L,R,L,R and results are independent of how many frames each launch carries."""
    chans = S.cfg1_stereo(0.2)
    want = oracle.encode_pcm(chans)
    ref = oracle.decode_su(want, 2)
    for batch in (1, 3, 64):
        monkeypatch.setattr(K.AudioProcessor, "BATCH_FRAMES", batch)
        frames = list(K.AudioProcessor.encodeStream(K.AudioProcessor.frameBufferToFrames(chans), {"channelCount": 2}))
        got = np.stack([K.serializeFrame(f) for f in frames])
        assert np.array_equal(got, want), batch
        seen = []
        pcm = list(K.AudioProcessor.decodeStream(iter(frames), {"channelCount": 2, "onProgress": seen.append}))
        assert seen == list(range(len(frames) // 2))
        for c in range(2):
            assert np.array_equal(bits(np.concatenate([p[c] for p in pcm])), bits(ref[c])), (batch, c)


def test_decode_stream_odd_stereo_tail_uses_dummy_frame(K, oracle):  # processor.js:216-228
    chans = S.cfg1_stereo(0.1)
    su = oracle.encode_pcm(chans)[:-1]
    pcm = list(K.AudioProcessor.decodeStream((K.deserializeFrame(u) for u in su), {"channelCount": 2}))
    ref = oracle.decode_su(su, 2)
    assert len(pcm) == (len(su) + 1) // 2
    for c in range(2):
        assert np.array_equal(bits(np.concatenate([p[c] for p in pcm])), bits(ref[c]))


def test_aea_blob_roundtrip(K):  # :77-92
    blob = K.AudioProcessor.createAeaBlob(K.AudioProcessor.encodeStream(mono_stream(2), {"channelCount": 1}), {"title": "test"})
    parsed = K.AudioProcessor.parseAeaBlob(blob)
    assert parsed["info"]["title"] == "test" and parsed["info"]["frameCount"] == 2 and len(parsed["frameData"]) == 2


def test_complete_aea_helpers(K, oracle):  # :94-117
    channels = [S.sine(440, n=700), S.sine(880, n=700)]
    aea = K.encodeAeaPcm(channels, {"title": "complete helper"})
    assert aea.dtype == np.uint8 and len(aea) == 2048 + 4 * 212
    assert K.AeaFile.parseHeader(aea[:2048]) == {"title": "complete helper", "frameCount": 4, "channelCount": 2}
    assert np.array_equal(aea[2048:].reshape(-1, 212), oracle.encode_pcm(channels))
    decoded = K.decodeAeaPcm(aea)
    assert len(decoded) == 2 and len(decoded[0]) == 1024 and len(decoded[1]) == 1024
    ref = oracle.decode_su(aea[2048:].reshape(-1, 212), 2)
    for a, b in zip(decoded, ref):
        assert np.array_equal(bits(a), bits(b))
    assert np.array_equal(K.encodeAeaPcm(channels, {"allocationBias": 3.0})[2048:].reshape(-1, 212),
                          oracle.encode_pcm(channels, oracle.make_options(bias=3.0)))
    with pytest.raises(ValueError, match="Value for allocationBias must be between 0 and 5"):
        K.encodeAeaPcm(channels, {"allocationBias": 7})


def test_batched_frame_dump_equals_deserialize_frame(K, oracle):
    """carta1_deserialize_units (the batched deserializeFrame of the CLI's JSON dump, bin/cli.js:567-677):
    encoder output, every BFU count, and random bytes whose declared payload runs past the 212 bytes."""
    rng = np.random.default_rng(12)
    x = S.cfg3_transients(0.4, seed=9, n_ch=1)[0]
    units = [oracle.encode_pcm([x]), oracle.encode_pcm([x], oracle.make_options(bias=3.0)),
             rng.integers(0, 256, (40, 212), dtype=np.uint8)]
    K.default_context().set_max_units_per_pass(16)  # several passes
    try:
        for su in units:
            got = K.deserializeFrames(su)
            assert len(got) == len(su)
            for u, g in zip(su, got):
                w = K.deserializeFrame(u)
                assert g["nBfu"] == w["nBfu"] and g["blockModes"] == w["blockModes"]
                assert np.array_equal(g["wordLengthIndices"], w["wordLengthIndices"])
                assert np.array_equal(g["scaleFactorIndices"], w["scaleFactorIndices"])
                assert len(g["quantizedCoefficients"]) == w["nBfu"]
                for a, b in zip(g["quantizedCoefficients"], w["quantizedCoefficients"]):
                    assert a.dtype == np.int32 and np.array_equal(a, b)
    finally:
        K.default_context().set_max_units_per_pass(0)

