"""Host-side mirror of the reference's JS surface (carta1_b200/codec.py): everything that runs
without a GPU.  Restates tests/options.test.js, bitstream.test.js, serialization.test.js and the
framing parts of processor.test.js, and pins the mirror's byte plumbing to the oracle."""
import numpy as np
import pytest

import signals as S
from carta1_b200 import codec as K


# ---- tests/options.test.js ----------------------------------------------------------------
def test_options_defaults():
    o = K.EncoderOptions()
    assert o.getValue("transientThresholdLow") == 1
    assert o.getValue("transientThresholdMid") == 1.5
    assert o.getValue("transientThresholdHigh") == 2.0
    assert o.allocationBias == 1.0 and o.fixedBlockModes is None


def test_options_range_validation():
    o = K.EncoderOptions()
    with pytest.raises(ValueError, match="Value for transientThresholdLow must be between 0.01 and 2, got 10"):
        o.setValue("transientThresholdLow", 10)
    with pytest.raises(ValueError):
        o.setValue("transientThresholdLow", 0.0)
    o.setValue("fixedBlockModes", [7, 7, 7])  # options.js:99: arrays are not range checked
    assert o.fixedBlockModes == [7, 7, 7]


def test_options_unknown_key():
    o = K.EncoderOptions()
    with pytest.raises(ValueError, match="Unknown option: unknownOption"):
        o.setValue("unknownOption", 123)
    with pytest.raises(ValueError, match="Unknown option: nope"):
        o.getValue("nope")
    K.EncoderOptions({"title": "ignored by setOptions"})  # options.js:77-83


def test_options_batch_and_reset():
    o = K.EncoderOptions()
    o.setOptions({"transientThresholdLow": 0.5, "transientThresholdMid": 0.75})
    assert o.getValue("transientThresholdLow") == 0.5 and o.getValue("transientThresholdMid") == 0.75
    o.reset()
    assert o.getValue("transientThresholdLow") == 1.0


def test_options_to_abi():
    a = K.EncoderOptions({"transientThresholdLow": 0.4, "allocationBias": 2.5, "fixedBlockModes": [0, 2, 3]}).to_abi()
    assert (a.transient_threshold_low, a.allocation_bias, a.use_fixed_block_modes) == (0.4, 2.5, 1)
    assert list(a.fixed_block_modes) == [0, 2, 3]
    assert K.EncoderOptions().to_abi().use_fixed_block_modes == 0


# ---- tests/bitstream.test.js ---------------------------------------------------------------
def test_bitstream_kats():
    buf = np.zeros(2, np.uint8)
    K.packBits(buf, 4, 0xF0, 8)  # :13-19
    assert buf.tolist() == [0x0F, 0x00]
    for n in range(1, 31):  # :29-38
        b = np.zeros(8, np.uint8)
        K.packBits(b, 3, (1 << n) - 1, n)
        assert K.unpackBits(b, 3, n) == (1 << n) - 1
    b = np.zeros(2, np.uint8)
    K.packBits(b, 0, 0b1000, 4)  # :41-71
    assert K.unpackSignedBits(b, 0, 4) == -8
    K.packBits(b, 0, 0b0111, 4)
    assert K.unpackSignedBits(b, 0, 4) == 7
    K.packBits(b, 0, 0b1111, 4)
    assert K.unpackSignedBits(b, 0, 4) == -1


# ---- tests/serialization.test.js + oracle pin ----------------------------------------------
def oracle_frame_to_dict(O, fr):
    n = fr.n_bfu
    return {"nBfu": n, "blockModes": list(fr.modes), "scaleFactorIndices": np.array(fr.sfi[:n], np.int32),
            "wordLengthIndices": np.array(fr.wl[:n], np.int32),
            "quantizedCoefficients": [np.array(fr.q[b][:K.SPECS_PER_BFU[b]], np.int32) for b in range(n)]}


def test_frame_roundtrip_and_oracle_bytes(oracle):
    chans = S.cfg3_transients(0.4, n_ch=1)
    su = oracle.encode_pcm(chans)
    assert len(su) > 20
    for u in su:
        d = K.deserializeFrame(u)
        ofr = oracle.deserialize_frame(u)
        want = oracle_frame_to_dict(oracle, ofr)
        assert d["nBfu"] == want["nBfu"] and d["blockModes"] == want["blockModes"]
        assert np.array_equal(d["scaleFactorIndices"], want["scaleFactorIndices"])
        assert np.array_equal(d["wordLengthIndices"], want["wordLengthIndices"])
        for a, b in zip(d["quantizedCoefficients"], want["quantizedCoefficients"]):
            assert np.array_equal(a, b)
        back = K.serializeFrame(d)
        assert back.shape == (212,) and np.array_equal(back, u)


def test_deserialize_random_bytes_matches_oracle(oracle):
    rng = np.random.default_rng(2)
    for _ in range(40):
        u = rng.integers(0, 256, 212, dtype=np.uint8)
        d = K.deserializeFrame(u)
        want = oracle_frame_to_dict(oracle, oracle.deserialize_frame(u))
        assert d["nBfu"] == want["nBfu"]
        assert np.array_equal(d["wordLengthIndices"], want["wordLengthIndices"])
        for a, b in zip(d["quantizedCoefficients"], want["quantizedCoefficients"]):
            assert np.array_equal(a, b)


def test_deserialize_rejects_wrong_size():
    with pytest.raises(ValueError, match="Frame must be 212 bytes"):
        K.deserializeFrame(np.zeros(100, np.uint8))


def test_aea_header():
    h = K.AeaFile.createHeader("Test Title", 100, 2)
    assert h.shape == (2048,) and h[:4].tolist() == [0, 8, 0, 0]
    info = K.AeaFile.parseHeader(h)
    assert info == {"title": "Test Title", "frameCount": 100, "channelCount": 2}
    with pytest.raises(ValueError, match="Header must be 2048 bytes"):
        K.AeaFile.parseHeader(np.zeros(100, np.uint8))
    with pytest.raises(ValueError, match="Invalid AEA file"):
        K.AeaFile.parseHeader(np.full(2048, 1, np.uint8))


# ---- tests/processor.test.js (framing and blob helpers) -------------------------------------
def test_frame_buffer_to_frames():
    frames = list(K.AudioProcessor.frameBufferToFrames([np.zeros(int(512 * 2.5), np.float32)]))
    assert len(frames) == 3 and len(frames[0]) == 512 and len(frames[2]) == 512
    a, b = np.arange(700, dtype=np.float32), np.arange(600, dtype=np.float32)
    st = list(K.AudioProcessor.frameBufferToFrames([a, b]))
    assert len(st) == 2 and st[1][1][87] == 599 and st[1][1][88] == 0 and st[1][0][187] == 699
    with pytest.raises(ValueError, match="Unsupported channel count: 3"):
        list(K.AudioProcessor.frameBufferToFrames([a, a, a]))


def test_aea_blob_roundtrip_on_oracle_frames(oracle):
    su = oracle.encode_pcm([S.sine(440, n=1024)])
    frames = [K.deserializeFrame(u) for u in su]
    blob = K.AudioProcessor.createAeaBlob(iter(frames), {"title": "test"})
    parsed = K.AudioProcessor.parseAeaBlob(blob + b"\x01\x02\x03")  # trailing partial unit is dropped
    assert parsed["info"]["title"] == "test" and parsed["info"]["frameCount"] == 2
    assert len(parsed["frameData"]) == 2 and np.array_equal(parsed["frameData"][1], su[1])


def test_input_validation_without_gpu():
    import asyncio  # noqa: F401  (the reference's helpers are async; the mirror's are plain calls)

    with pytest.raises(TypeError, match="one or two Float32 channels"):
        K.encodeAeaPcm([])
    with pytest.raises(TypeError, match="one or two Float32 channels"):
        K.encodeAeaPcm([np.zeros(4, np.float64)])
    with pytest.raises(TypeError, match="AEA bytes or a Blob"):
        K.decodeAeaPcm("not AEA bytes")


def test_expand_frame_replays_set_sequence():
    fr = {"nBfu": 52, "scaleFactorIndices": np.full(52, 10), "wordLengthIndices": np.full(52, 8),
          "quantizedCoefficients": [np.full(10, b + 1, np.int32) for b in range(52)], "blockModes": [1, 1, 1]}
    q, sfi, bits, modes = K._expand_frame(fr)
    # short-mode BFU 0 sits at 0, BFU 4 at 8: 10-long arrays overlap and later BFUs win
    assert q[0] == 1 and q[8] == 5 and q[9] == 5 and bits.max() == 9 and modes.tolist() == [1, 1, 1]
    d = K._expand_frame(K.AudioProcessor._createDummyFrame())
    assert not d[0].any() and not d[2].any()
