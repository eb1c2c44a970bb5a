"""Host-side mirror of carta1's public surface (codec/index.js:26-47) over the C ABI.

The reference's host language is JavaScript and this image has no Node, so the
host side above libcarta1_b200.so is written in Python with the reference's own names, argument
meaning and error behaviour (the N-API shim a Node host would load instead lives in
carta1_b200/napi/).  Everything numerical happens on the GPU behind the C
ABI; this module only moves bytes, validates arguments and converts between 212-byte sound
units and the reference's frame objects.  There is no CPU fallback.

Names follow the reference verbatim (camelCase) so the parity tests read like its vitest files:

    encode(options)        -> closure (Float32[512]) -> frame object   codec/pipeline/encoder.js:438-450
    decode()               -> closure (frame object) -> Float32[512]   codec/pipeline/decoder.js:408-411
    encodeAeaPcm / decodeAeaPcm                                        codec/io/processor.js:597-654
    AudioProcessor.encodeStream / decodeStream / frameBufferToFrames / createAeaBlob / parseAeaBlob
    EncoderOptions, AeaFile, serializeFrame, deserializeFrame, packBits, unpackBits, unpackSignedBits

A frame object is a dict with the reference's keys: nBfu, scaleFactorIndices, wordLengthIndices,
quantizedCoefficients (list of int32 arrays, one per BFU), blockModes.
"""
from __future__ import annotations

import numpy as np

from . import _lib

SAMPLES_PER_FRAME = 512
SOUND_UNIT_SIZE = 212
AEA_HEADER_SIZE = 2048
AEA_MAGIC = bytes([0x00, 0x08, 0x00, 0x00])
AEA_TITLE_OFFSET, AEA_TITLE_SIZE = 4, 256
AEA_FRAME_COUNT_OFFSET, AEA_CHANNEL_COUNT_OFFSET = 260, 264
NUM_BFUS = 52
# codec/core/constants.js:29-52,141-143
SPECS_PER_BFU = np.array([8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 9, 9,
                          9, 9, 10, 10, 10, 10, 12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20, 20, 20, 20, 20, 20],
                         np.int32)
BFU_AMOUNTS = np.array([20, 28, 32, 36, 40, 44, 48, 52], np.int32)
BFU_START_LONG = np.concatenate([[0], np.cumsum(SPECS_PER_BFU)[:-1]]).astype(np.int32)
WORD_LENGTH_BITS = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16], np.int32)
SCALE_FACTORS = np.array([2.0 ** (i / 3.0 - 21) for i in range(64)])


# --------------------------------------------------------------------------------------
# EncoderOptions (codec/core/options.js:11-164)
# --------------------------------------------------------------------------------------
class EncoderOptions:
    _RANGES = {
        "transientThresholdLow": (1.0, (0.01, 2)),
        "transientThresholdMid": (1.5, (0.01, 3)),
        "transientThresholdHigh": (2.0, (0.01, 4)),
        "allocationBias": (1.0, (0.0, 5.0)),
        "fixedBlockModes": (None, None),  # type 'array': no range validation (options.js:99-107)
    }

    def __init__(self, options=None):
        self.values = {k: v[0] for k, v in self._RANGES.items()}
        self.metadata = {k: ({"default": d, "range": list(r)} if r else {"default": d, "type": "array"})
                         for k, (d, r) in self._RANGES.items()}
        if options:
            self.setOptions(options)

    def setOptions(self, options):  # options.js:77-83: unknown keys are ignored here
        for key, value in dict(options).items():
            if key in self.values:
                self.setValue(key, value)

    def setValue(self, key, value):  # options.js:91-109
        if key not in self.metadata:
            raise ValueError(f"Unknown option: {key}")
        meta = self.metadata[key]
        if meta.get("type") != "array":
            lo, hi = meta["range"]
            if value < lo or value > hi:
                raise ValueError(f"Value for {key} must be between {_js_num(lo)} and {_js_num(hi)}, got {value}")
        self.values[key] = value

    def getValue(self, key):
        if key not in self.values:
            raise ValueError(f"Unknown option: {key}")
        return self.values[key]

    transientThresholdLow = property(lambda s: s.values["transientThresholdLow"])
    transientThresholdMid = property(lambda s: s.values["transientThresholdMid"])
    transientThresholdHigh = property(lambda s: s.values["transientThresholdHigh"])
    allocationBias = property(lambda s: s.values["allocationBias"])
    fixedBlockModes = property(lambda s: s.values["fixedBlockModes"])

    def getMetadata(self, key):
        return self.metadata[key]

    def getAllMetadata(self):
        return dict(self.metadata)

    def reset(self):
        for k, m in self.metadata.items():
            self.values[k] = m["default"]

    def toObject(self):
        return {"values": dict(self.values), "metadata": dict(self.metadata)}

    def to_abi(self) -> _lib.EncOpts:
        """The POD the C ABI takes.  Only transientThresholdLow is read by the hot path
        (codec/pipeline/encoder.js:137-141 uses it for all three bands)."""
        return _lib.make_enc_opts(self.transientThresholdLow, self.allocationBias, self.fixedBlockModes)


def _js_num(v):
    return str(int(v)) if float(v).is_integer() else str(v)


def _as_options(options) -> EncoderOptions:
    if options is None:
        return EncoderOptions()
    if isinstance(options, EncoderOptions):
        return options
    return EncoderOptions(options)


# --------------------------------------------------------------------------------------
# bit stream + frame (de)serialisation: host-side integer plumbing between the 212-byte
# units that cross the C ABI and the reference's frame objects
# (codec/io/bitstream.js:15-82, codec/io/serialization.js:41-176)
# --------------------------------------------------------------------------------------
def packBits(buffer, bitPosition, value, bitCount):
    value = int(value)
    for i in range(bitCount):
        pos = bitPosition + i
        byte = pos >> 3
        if byte >= len(buffer):
            break
        if (value >> (bitCount - 1 - i)) & 1:
            buffer[byte] |= 0x80 >> (pos & 7)
        else:
            buffer[byte] &= ~(0x80 >> (pos & 7)) & 0xFF


def unpackBits(buffer, bitPosition, bitCount):
    """bitstream.js:48-69: stops at the end of the buffer and returns the bits read so far."""
    value = 0
    for i in range(bitCount):
        pos = bitPosition + i
        byte = pos >> 3
        if byte >= len(buffer):
            break
        value = (value << 1) | ((int(buffer[byte]) >> (7 - (pos & 7))) & 1)
    return value


def unpackSignedBits(buffer, bitPosition, bitCount):
    value = unpackBits(buffer, bitPosition, bitCount)
    return value - (1 << bitCount) if value >= (1 << (bitCount - 1)) else value


def serializeFrame(frameData) -> np.ndarray:
    n = int(frameData["nBfu"])
    idx = np.nonzero(BFU_AMOUNTS == n)[0]
    if len(idx) == 0:
        raise ValueError(f"nBfu must be one of {BFU_AMOUNTS.tolist()}, got {n}")
    m = frameData["blockModes"]
    header = (((2 - int(m[0])) << 14) | ((2 - int(m[1])) << 12) | ((3 - int(m[2])) << 10) | (int(idx[0]) << 5)) & 0xFFFF
    wl = np.asarray(frameData["wordLengthIndices"], np.int64)
    sf = np.asarray(frameData["scaleFactorIndices"], np.int64)
    fields = [(header, 16)] + [(int(wl[i]) & 15, 4) for i in range(n)] + [(int(sf[i]) & 63, 6) for i in range(n)]
    for i in range(n):
        bits = int(WORD_LENGTH_BITS[int(wl[i]) & 15])
        if bits > 0:
            for c in np.asarray(frameData["quantizedCoefficients"][i], np.int64):
                fields.append((int(c) & ((1 << bits) - 1), bits))
    acc, nbits = 0, 0
    for v, b in fields:
        acc = (acc << b) | v
        nbits += b
    total = SOUND_UNIT_SIZE * 8
    acc = acc << (total - nbits) if nbits <= total else acc >> (nbits - total)
    out = np.frombuffer(acc.to_bytes(SOUND_UNIT_SIZE, "big"), np.uint8).copy()
    out[-3:] = 0  # serialization.js:92-95
    return out


def deserializeFrame(buffer) -> dict:
    buffer = np.asarray(buffer, np.uint8)
    if buffer.shape != (SOUND_UNIT_SIZE,):
        raise ValueError(f"Frame must be {SOUND_UNIT_SIZE} bytes")
    word = int.from_bytes(buffer.tobytes(), "big")
    total = SOUND_UNIT_SIZE * 8

    def take(pos, bits):  # unpackBits incl. its behaviour past the end of the buffer
        avail = total - pos
        if avail <= 0:
            return 0
        nb = min(bits, avail)
        return (word >> (total - pos - nb)) & ((1 << nb) - 1)

    header = take(0, 16)
    modes = [2 - ((header >> 14) & 3), 2 - ((header >> 12) & 3), 3 - ((header >> 10) & 3)]
    n = int(BFU_AMOUNTS[(header >> 5) & 7])
    pos = 16
    wl = np.array([take(pos + 4 * i, 4) for i in range(n)], np.int32)
    pos += 4 * n
    sf = np.array([take(pos + 6 * i, 6) for i in range(n)], np.int32)
    pos += 6 * n
    coefs = []
    for i in range(n):
        bits = int(WORD_LENGTH_BITS[wl[i]])
        size = int(SPECS_PER_BFU[i])
        q = np.zeros(size, np.int32)
        if bits > 0:
            for j in range(size):
                v = take(pos, bits)
                q[j] = v - (1 << bits) if v >= (1 << (bits - 1)) else v
                pos += bits
        coefs.append(q)
    return {"nBfu": n, "scaleFactorIndices": sf, "wordLengthIndices": wl, "quantizedCoefficients": coefs,
            "blockModes": modes}


BFU_START_SHORT = np.array([0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90, 122, 128,
                            160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182, 214, 246, 256, 288,
                            320, 352, 384, 416, 448, 480, 268, 300, 332, 364, 396, 428, 460, 492], np.int32)


def _expand_frame(frameData):
    """Frame object -> per-position (q, sfi, bits, modes) for carta1_dec_frames_expanded.

    decode() accepts objects the 212-byte layout cannot hold (the dummy frame with nBfu 0,
    processor.js:299-307; arbitrary coefficient-array lengths, tests/decoder.test.js:70-84), so
    this replays dequantizationStage's `coefficients.set(dequantized, position)` loop
    (decoder.js:73-94) on integers; the GPU does the arithmetic."""
    n = int(frameData["nBfu"])
    if n < 0 or n > NUM_BFUS:
        raise ValueError(f"nBfu must be within 0..{NUM_BFUS}, got {n}")
    modes = [int(m) for m in frameData["blockModes"]]
    q = np.zeros(512, np.int32)
    sfi = np.zeros(512, np.uint8)
    bits = np.zeros(512, np.uint8)
    for bfu in range(n):
        wl = int(frameData["wordLengthIndices"][bfu])
        sf = int(frameData["scaleFactorIndices"][bfu])
        if not 0 <= wl < 16 or not 0 <= sf < 64:
            raise ValueError("wordLengthIndices must be within 0..15 and scaleFactorIndices within 0..63")
        b = int(WORD_LENGTH_BITS[wl])
        if b == 0:
            continue
        band = 0 if bfu < 20 else (1 if bfu < 36 else 2)
        pos = int(BFU_START_LONG[bfu] if modes[band] == 0 else BFU_START_SHORT[bfu])
        vals = np.asarray(frameData["quantizedCoefficients"][bfu], np.int32)
        if pos + len(vals) > 512:  # TypedArray.prototype.set throws
            raise ValueError("offset is out of bounds")
        q[pos:pos + len(vals)] = vals
        sfi[pos:pos + len(vals)] = sf
        bits[pos:pos + len(vals)] = b
    return q, sfi, bits, np.array(modes, np.int32)


# --------------------------------------------------------------------------------------
# AeaFile (codec/io/serialization.js:190-253)
# --------------------------------------------------------------------------------------
class AeaFile:
    @staticmethod
    def createHeader(title="", frameCount=0, channelCount=1) -> np.ndarray:
        return _lib.aea_write_header(title, int(frameCount), int(channelCount))

    @staticmethod
    def parseHeader(header) -> dict:
        title, count, n_ch = _lib.aea_parse_header(header)
        return {"title": title, "frameCount": count, "channelCount": n_ch}


# --------------------------------------------------------------------------------------
# frame closures
# --------------------------------------------------------------------------------------
_default_ctx = None


def default_context() -> _lib.Context:
    """One lazily created context on the current CUDA device 0 (raises without a GPU)."""
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None:
        _default_ctx = _lib.Context(0)
    return _default_ctx


def encode(options=None, bufferPool=None, ctx: _lib.Context | None = None):
    """encode(options) -> encoder closure; one closure per mono stream (README.md:108-110).
    `bufferPool` is accepted for signature parity; the stream state lives on the device."""
    opts = _as_options(options)
    enc = _lib.StreamEncoder(ctx or default_context(), opts.to_abi(), 1)

    def encoder(pcm):
        pcm = np.asarray(pcm, np.float32)
        if pcm.shape != (SAMPLES_PER_FRAME,):
            raise ValueError(f"encode() takes {SAMPLES_PER_FRAME} samples per frame")
        return deserializeFrame(enc.frames(pcm)[0, 0])

    encoder.handle = enc
    return encoder


def decode(bufferPool=None, ctx: _lib.Context | None = None):
    dec = _lib.StreamDecoder(ctx or default_context(), 1)

    def decoder(frameData):
        q, sfi, bits, modes = _expand_frame(frameData)
        return dec.frames_expanded(q, sfi, bits, modes)[0, 0].copy()

    decoder.handle = dec
    return decoder


# --------------------------------------------------------------------------------------
# AudioProcessor (codec/io/processor.js:37-585) -- generators instead of async generators
# --------------------------------------------------------------------------------------
class AudioProcessor:
    #: frames gathered per GPU launch by the stream adapters; yields and onProgress calls keep
    #: the reference's per-frame order
    BATCH_FRAMES = 64

    @staticmethod
    def encodeAeaPcm(channels, options=None):
        return encodeAeaPcm(channels, options)

    @staticmethod
    def decodeAeaPcm(data):
        return decodeAeaPcm(data)

    @staticmethod
    def encodeStream(audioFrames, options=None, ctx=None):
        options = dict(options or {})
        channelCount = options.get("channelCount", 1)
        onProgress = options.get("onProgress")
        if channelCount not in (1, 2):
            raise ValueError(f"Unsupported channel count: {channelCount}")
        opts = _as_options(options.get("encoderOptions"))
        enc = _lib.StreamEncoder(ctx or default_context(), opts.to_abi(), channelCount)
        frameIndex = 0
        for batch in _batches(audioFrames, AudioProcessor.BATCH_FRAMES):
            if channelCount == 1:
                pcm = np.stack([np.asarray(f, np.float32) for f in batch])[None]
            else:
                pcm = np.stack([np.stack([np.asarray(f[c], np.float32) for f in batch]) for c in range(2)])
            su = enc.frames(pcm)  # [channel][frame][212]
            for k in range(len(batch)):
                for c in range(channelCount):  # left then right (processor.js:121-130)
                    yield deserializeFrame(su[c, k])
                if onProgress:
                    onProgress(frameIndex)
                frameIndex += 1
        enc.close()

    @staticmethod
    def decodeStream(encodedFrames, options=None, ctx=None):
        options = dict(options or {})
        channelCount = options.get("channelCount", 1)
        onProgress = options.get("onProgress")
        if channelCount not in (1, 2):
            raise ValueError(f"Unsupported channel count: {channelCount}")
        dec = _lib.StreamDecoder(ctx or default_context(), channelCount)
        frameIndex = 0
        for batch in _batches(encodedFrames, AudioProcessor.BATCH_FRAMES * channelCount):
            if channelCount == 2 and len(batch) % 2:  # processor.js:216-228
                batch = batch + [AudioProcessor._createDummyFrame()]
            ex = [_expand_frame(f) for f in batch]
            arr = [np.stack([e[i] for e in ex]) for i in range(4)]  # [frame*channel][...]
            arr = [np.ascontiguousarray(a.reshape(-1, channelCount, a.shape[-1]).swapaxes(0, 1)) for a in arr]
            pcm = dec.frames_expanded(*arr)  # [channel][frame][512]
            for k in range(pcm.shape[1]):
                yield pcm[0, k].copy() if channelCount == 1 else [pcm[0, k].copy(), pcm[1, k].copy()]
                if onProgress:
                    onProgress(frameIndex)
                frameIndex += 1
        dec.close()

    @staticmethod
    def frameBufferToFrames(buffers, frameSize=SAMPLES_PER_FRAME):  # processor.js:246-279
        channelCount = len(buffers)
        if channelCount not in (1, 2):
            raise ValueError(f"Unsupported channel count: {channelCount}")
        maxLength = max(len(b) for b in buffers)
        for i in range(0, maxLength, frameSize):
            frames = []
            for b in buffers:
                frame = np.zeros(frameSize, np.float32)
                part = np.asarray(b[i:i + frameSize], np.float32)
                frame[:len(part)] = part
                frames.append(frame)
            yield frames[0] if channelCount == 1 else frames

    @staticmethod
    def collectFrames(frameStream):
        return list(frameStream)

    @staticmethod
    def _createDummyFrame():  # processor.js:299-307
        return {"nBfu": 0, "blockModes": [0, 0, 0], "scaleFactorIndices": np.zeros(0, np.int32),
                "wordLengthIndices": np.zeros(0, np.int32), "quantizedCoefficients": []}

    @staticmethod
    def createAeaBlob(encodedFrames, options=None) -> bytes:  # processor.js:317-339
        options = dict(options or {})
        frames = [serializeFrame(f) for f in encodedFrames]
        header = AeaFile.createHeader(options.get("title", "encoded by atrac1.js"), len(frames),
                                      options.get("channelCount", 1))
        return header.tobytes() + b"".join(f.tobytes() for f in frames)

    @staticmethod
    def parseAeaBlob(blob):  # processor.js:511-525: a trailing partial unit is dropped
        data = np.frombuffer(bytes(blob), np.uint8)
        info = AeaFile.parseHeader(data[:AEA_HEADER_SIZE])
        n = (len(data) - AEA_HEADER_SIZE) // SOUND_UNIT_SIZE
        body = data[AEA_HEADER_SIZE:AEA_HEADER_SIZE + n * SOUND_UNIT_SIZE].reshape(n, SOUND_UNIT_SIZE)
        return {"info": info, "frameData": [body[i] for i in range(n)]}

    @staticmethod
    def deserializedFrameStream(frameData):
        for frame in frameData:
            yield deserializeFrame(frame)


def _batches(it, size):
    batch = []
    for x in it:
        batch.append(x)
        if len(batch) == size:
            yield batch
            batch = []
    if batch:
        yield batch


# --------------------------------------------------------------------------------------
# whole-buffer helpers (codec/io/processor.js:597-654): one launch sequence for the whole
# signal through carta1_encode_pcm / carta1_decode_su
# --------------------------------------------------------------------------------------
def encodeAeaPcm(channels, options=None, ctx=None) -> np.ndarray:
    if (not isinstance(channels, (list, tuple)) or len(channels) not in (1, 2) or
            any(not (isinstance(c, np.ndarray) and c.dtype == np.float32 and c.ndim == 1) for c in channels)):
        raise TypeError("ATRAC1 encoding requires one or two Float32 channels")
    options = dict(options or {})
    title = options.pop("title", "encoded by carta1")
    opts = EncoderOptions(options)
    su = (ctx or default_context()).encode_pcm(list(channels), opts.to_abi())
    header = AeaFile.createHeader(title, len(su), len(channels))
    return np.concatenate([header, su.reshape(-1)])


def decodeAeaPcm(data, ctx=None):
    if isinstance(data, np.ndarray) and data.dtype == np.uint8:
        raw = data
    elif isinstance(data, (bytes, bytearray, memoryview)):
        raw = np.frombuffer(bytes(data), np.uint8)
    else:
        raise TypeError("ATRAC1 decoding requires AEA bytes or a Blob")
    info = AeaFile.parseHeader(raw[:AEA_HEADER_SIZE])
    n = (len(raw) - AEA_HEADER_SIZE) // SOUND_UNIT_SIZE
    if info["channelCount"] not in (1, 2):
        raise ValueError(f"Unsupported channel count: {info['channelCount']}")
    su = np.ascontiguousarray(raw[AEA_HEADER_SIZE:AEA_HEADER_SIZE + n * SOUND_UNIT_SIZE])
    return (ctx or default_context()).decode_su(su, info["channelCount"])


def encodePcmShard(channels, haloFrames, options=None, ctx=None) -> np.ndarray:
    """One (stream, frame range) shard of encodeAeaPcm's body (no counterpart in the reference, whose loop
    codec/io/processor.js:97-136 is sequential): `channels` start `haloFrames` frames (0, or >= 2) before the first
    frame to encode.  Returns the sound units [n, 212] of the frames after the halo, bit-identical to that span of
    the whole-stream result (carta1_encode_pcm_shard)."""
    if (not isinstance(channels, (list, tuple)) or len(channels) not in (1, 2) or
            any(not (isinstance(c, np.ndarray) and c.dtype == np.float32 and c.ndim == 1) for c in channels)):
        raise TypeError("ATRAC1 encoding requires one or two Float32 channels")
    n = max(len(c) for c in channels)
    chans = [np.ascontiguousarray(c if len(c) == n else np.concatenate([c, np.zeros(n - len(c), np.float32)])) for c in channels]
    frames = max((n + 511) // 512 - int(haloFrames), 0)
    su = np.zeros((frames * len(chans), SOUND_UNIT_SIZE), np.uint8)
    got = (ctx or default_context()).encode_pcm_shard_into(chans, int(haloFrames), su, EncoderOptions(dict(options or {})).to_abi())
    return su[:got]


def decodeUnitsShard(units, channelCount, haloFrames, ctx=None):
    """The decode counterpart (carta1_decode_su_shard): `units` start `haloFrames` frames (0, or >= 1) before the
    first frame to decode; returns one Float32 array per channel for the frames after the halo."""
    if channelCount not in (1, 2):
        raise ValueError(f"Unsupported channel count: {channelCount}")
    units = np.ascontiguousarray(units, np.uint8).reshape(-1, SOUND_UNIT_SIZE)
    frames = max((units.shape[0] + channelCount - 1) // channelCount - int(haloFrames), 0)
    outs = [np.zeros(frames * 512, np.float32) for _ in range(channelCount)]
    if frames:
        (ctx or default_context()).decode_su_shard_into(units, units.shape[0], channelCount, int(haloFrames), outs)
    return outs


def deserializeFrames(units, ctx=None) -> list:
    """deserializeFrame over many sound units at once on the device (carta1_deserialize_units): the frame
    objects the `--json` dump of bin/cli.js:567-677 writes, one per 212-byte unit, equal to
    [deserializeFrame(u) for u in units]."""
    units = np.ascontiguousarray(units, np.uint8).reshape(-1, SOUND_UNIT_SIZE)
    d = (ctx or default_context()).deserialize_units(units)
    starts = np.concatenate([[0], np.cumsum(SPECS_PER_BFU)]).astype(int)
    frames = []
    for i in range(units.shape[0]):
        n = int(d["n_bfu"][i])
        frames.append({
            "nBfu": n,
            "scaleFactorIndices": d["sfi"][i, :n].astype(np.int32),
            "wordLengthIndices": d["wl"][i, :n].astype(np.int32),
            "quantizedCoefficients": [d["q"][i, starts[b]:starts[b + 1]].copy() for b in range(n)],
            "blockModes": [int(m) for m in d["block_modes"][i]],
        })
    return frames
