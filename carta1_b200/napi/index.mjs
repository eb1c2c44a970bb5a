/**
 * carta1-b200: drop-in for carta1's public surface (codec/index.js:26-47) with the ATRAC1
 * encode/decode hot path on a B200 behind the C ABI of include/carta1_b200.h.
 *
 * Non-hot exports (serializeFrame, deserializeFrame, quantize, dequantize, AeaFile, BufferPool,
 * EncoderOptions, FFT, pipe, qmfAnalysisStage, mdctStage and the constant tables) are the
 * reference's own JavaScript, re-exported unchanged from the `carta1` package.  Replaced:
 *   encode / decode                      frame closures, state kept on the device
 *   encodeAeaPcm / decodeAeaPcm          one GPU pass over the whole buffer
 *   AudioProcessor.encodeStream / decodeStream / encodeAeaPcm / decodeAeaPcm
 *
 * NOTE: no Node in the build image.  This file runs there inside Qt's QJSEngine next to the reference, with the addon
 * replaced by tests/js_layer/mock_native.js (tools/ref_run_qjs.py --check-js-layer); the addon itself runs against
 * tests/napi_host on the GPU (tests/test_napi_host.py).  See INTEGRATION.md.
 */
import { createRequire } from 'node:module'
import * as ref from 'carta1'

const native = createRequire(import.meta.url)('./build/Release/carta1_b200.node')

const {
  serializeFrame, deserializeFrame, quantize, dequantize, AeaFile, BufferPool, EncoderOptions, FFT, pipe,
  qmfAnalysisStage, mdctStage, WORD_LENGTH_BITS, SPECS_PER_BFU, SCALE_FACTORS, BFU_START_LONG,
} = ref

const SAMPLES_PER_FRAME = 512
const SOUND_UNIT_SIZE = 212
const AEA_HEADER_SIZE = 2048
const BATCH_FRAMES = 64
const BFU_START_SHORT = Int32Array.from([0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90,
  122, 128, 160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182, 214, 246, 256, 288, 320, 352, 384, 416,
  448, 480, 268, 300, 332, 364, 396, 428, 460, 492])

/** Every libm-derived table, computed by THIS engine's Math.* so the device sees V8's values
 *  (constants.js:60-66,144-150; mdct.js:27-36,215-221; fft.js:36-39). */
function hostTables() {
  const mdctTable = (size, scale) => {
    const t = new Float64Array(size / 2)
    const alpha = (2.0 * Math.PI) / (8.0 * size)
    const omega = (2.0 * Math.PI) / size
    const root = Math.sqrt(scale / size)
    for (let i = 0; i < size / 4; i++) {
      t[2 * i] = root * Math.cos(omega * i + alpha)
      t[2 * i + 1] = root * Math.sin(omega * i + alpha)
    }
    return t
  }
  const fftW = new Float64Array(16)
  for (let k = 0; k < 8; k++) {
    const angle = (-2 * Math.PI) / (2 << k)
    fftW[2 * k] = Math.cos(angle)
    fftW[2 * k + 1] = Math.sin(angle)
  }
  return {
    windowShort: Float64Array.from({ length: 32 }, (_, i) => Math.sin(((i + 0.5) * Math.PI) / 64)),
    scaleFactors: Float64Array.from(SCALE_FACTORS),
    mdctFwd64: mdctTable(64, 0.5), mdctFwd256: mdctTable(256, 0.5), mdctFwd512: mdctTable(512, 1.0),
    mdctInv64: mdctTable(64, 64 * 8), mdctInv256: mdctTable(256, 256 * 8), mdctInv512: mdctTable(512, 512 * 4),
    fftW,
  }
}

let sharedContext = null
function context() {
  if (!sharedContext) sharedContext = native.createContext(Number(process.env.CARTA1_B200_DEVICE ?? 0), hostTables())
  return sharedContext
}

function abiOptions(options) {
  const o = options instanceof EncoderOptions ? options : new EncoderOptions(options ?? {})
  const bias = o.allocationBias
  return {
    transientThresholdLow: o.transientThresholdLow, // read for all three bands (encoder.js:137-141)
    allocationBias: bias,
    fixedBlockModes: o.fixedBlockModes ?? null,
    // bitallocation.js:52-58 with this engine's Math.pow
    biasedScaleFactors: bias === 1 ? Float64Array.from(SCALE_FACTORS) : Float64Array.from(SCALE_FACTORS, (s) => Math.pow(s, bias)),
  }
}

/** decoder.js:73-94 replayed on integers: per-position quantised value / scale-factor index / bits. */
function expandFrame(frame, q, sfi, bits, modes, at) {
  const base = at * 512
  for (let b = 0; b < 3; b++) modes[at * 3 + b] = frame.blockModes[b]
  for (let bfu = 0; bfu < frame.nBfu; bfu++) {
    const width = WORD_LENGTH_BITS[frame.wordLengthIndices[bfu]]
    if (!(width > 0)) continue
    const band = bfu < 20 ? 0 : bfu < 36 ? 1 : 2
    const pos = frame.blockModes[band] === 0 ? BFU_START_LONG[bfu] : BFU_START_SHORT[bfu]
    const values = frame.quantizedCoefficients[bfu]
    if (pos + values.length > 512) throw new RangeError('offset is out of bounds')
    for (let j = 0; j < values.length; j++) {
      q[base + pos + j] = values[j]
      sfi[base + pos + j] = frame.scaleFactorIndices[bfu]
      bits[base + pos + j] = width
    }
  }
}

const dummyFrame = () => ({ nBfu: 0, blockModes: [0, 0, 0], scaleFactorIndices: new Int32Array(0),
  wordLengthIndices: new Int32Array(0), quantizedCoefficients: [] })

/** encode(options, bufferPool) -> (Float32Array[512]) -> frame object (encoder.js:438-450). */
function encode(options = null, _bufferPool = null) {
  const enc = native.createEncoder(context(), abiOptions(options), 1)
  return (pcm) => deserializeFrame(native.encodeFrames(enc, pcm, 1))
}

/** decode(bufferPool) -> (frame object) -> Float32Array[512] (decoder.js:408-411). */
function decode(_bufferPool = null) {
  const dec = native.createDecoder(context(), 1)
  return (frame) => {
    const q = new Int32Array(512), sfi = new Uint8Array(512), bits = new Uint8Array(512), modes = new Int32Array(3)
    expandFrame(frame, q, sfi, bits, modes, 0)
    return native.decodeFramesExpanded(dec, q, sfi, bits, modes, 1)
  }
}

async function* batches(iterable, size) {
  let batch = []
  for await (const item of iterable) {
    batch.push(item)
    if (batch.length === size) { yield batch; batch = [] }
  }
  if (batch.length) yield batch
}

class AudioProcessor extends ref.AudioProcessor {
  static encodeAeaPcm(channels, options = {}) { return encodeAeaPcm(channels, options) }
  static decodeAeaPcm(input) { return decodeAeaPcm(input) }

  /** Same yields, order (L,R,L,R) and onProgress calls as processor.js:69-136; BATCH_FRAMES frames per launch. */
  static async *encodeStream(audioFrames, options = {}) {
    const { channelCount = 1, onProgress, encoderOptions } = options
    if (channelCount !== 1 && channelCount !== 2) throw new Error(`Unsupported channel count: ${channelCount}`)
    const enc = native.createEncoder(context(), abiOptions(encoderOptions), channelCount)
    let frameIndex = 0
    for await (const batch of batches(audioFrames, BATCH_FRAMES)) {
      const n = batch.length
      const pcm = new Float32Array(channelCount * n * SAMPLES_PER_FRAME) // [channel][frame][512]
      batch.forEach((f, k) => {
        if (channelCount === 1) pcm.set(f, k * SAMPLES_PER_FRAME)
        else for (let c = 0; c < 2; c++) pcm.set(f[c], (c * n + k) * SAMPLES_PER_FRAME)
      })
      const su = native.encodeFrames(enc, pcm, n)
      for (let k = 0; k < n; k++) {
        for (let c = 0; c < channelCount; c++) {
          const at = (c * n + k) * SOUND_UNIT_SIZE
          yield deserializeFrame(su.subarray(at, at + SOUND_UNIT_SIZE))
        }
        if (onProgress) onProgress(frameIndex++)
      }
    }
    native.destroy(enc)
  }

  static async *decodeStream(encodedFrames, options = {}) {
    const { channelCount = 1, onProgress } = options
    if (channelCount !== 1 && channelCount !== 2) throw new Error(`Unsupported channel count: ${channelCount}`)
    const dec = native.createDecoder(context(), channelCount)
    let frameIndex = 0
    for await (const batch of batches(encodedFrames, BATCH_FRAMES * channelCount)) {
      if (channelCount === 2 && batch.length % 2) batch.push(dummyFrame()) // processor.js:216-228
      const n = batch.length / channelCount
      const q = new Int32Array(batch.length * 512), sfi = new Uint8Array(batch.length * 512)
      const bits = new Uint8Array(batch.length * 512), modes = new Int32Array(batch.length * 3)
      batch.forEach((f, i) => expandFrame(f, q, sfi, bits, modes, (i % channelCount) * n + Math.floor(i / channelCount)))
      const pcm = native.decodeFramesExpanded(dec, q, sfi, bits, modes, n)
      for (let k = 0; k < n; k++) {
        const ch = (c) => pcm.slice((c * n + k) * SAMPLES_PER_FRAME, (c * n + k + 1) * SAMPLES_PER_FRAME)
        yield channelCount === 1 ? ch(0) : [ch(0), ch(1)]
        if (onProgress) onProgress(frameIndex++)
      }
    }
    native.destroy(dec)
  }
}

/** processor.js:597-617 */
async function encodeAeaPcm(channels, options = {}) {
  if (!Array.isArray(channels) || (channels.length !== 1 && channels.length !== 2) ||
      channels.some((channel) => !(channel instanceof Float32Array))) {
    throw new TypeError('ATRAC1 encoding requires one or two Float32 channels')
  }
  const { title = 'encoded by carta1', ...encoderValues } = options
  const length = Math.max(...channels.map((c) => c.length)) // frameBufferToFrames pads the shorter channel
  const padded = channels.map((c) => { if (c.length === length) return c; const p = new Float32Array(length); p.set(c); return p })
  const su = await native.encodePcm(context(), padded, abiOptions(new EncoderOptions(encoderValues)))
  const out = new Uint8Array(AEA_HEADER_SIZE + su.length)
  out.set(AeaFile.createHeader(title, su.length / SOUND_UNIT_SIZE, channels.length), 0)
  out.set(su, AEA_HEADER_SIZE)
  return out
}

/** processor.js:628-654 */
async function decodeAeaPcm(input) {
  let bytes
  if (typeof Blob !== 'undefined' && input instanceof Blob) bytes = new Uint8Array(await input.arrayBuffer())
  else if (input instanceof Uint8Array) bytes = input
  else if (input instanceof ArrayBuffer) bytes = new Uint8Array(input)
  else throw new TypeError('ATRAC1 decoding requires AEA bytes or a Blob')
  const info = AeaFile.parseHeader(bytes.subarray(0, AEA_HEADER_SIZE))
  const units = Math.floor((bytes.length - AEA_HEADER_SIZE) / SOUND_UNIT_SIZE) // trailing partial unit dropped
  return native.decodeSu(context(), bytes.slice(AEA_HEADER_SIZE, AEA_HEADER_SIZE + units * SOUND_UNIT_SIZE), info.channelCount)
}

/**
 * One shard of a longer stream (no counterpart in the reference, whose loop processor.js:97-136 is sequential): the
 * channels start `haloFrames` frames (0, or >= 2) before the first frame to encode; resolves to the sound units of the
 * frames after the halo, identical to that span of encodeAeaPcm's body.  A host that spreads a long recording over
 * several GPUs (one context per GPU, CARTA1_B200_DEVICE) calls this per (stream, frame range).
 */
async function encodePcmShard(channels, haloFrames, options = {}) {
  return native.encodePcm(context(), channels, abiOptions(new EncoderOptions(options)), haloFrames)
}

/** The decode counterpart: `units` start `haloFrames` frames (0, or >= 1) before the first frame to decode. */
async function decodeUnitsShard(units, channelCount, haloFrames) {
  return native.decodeSu(context(), units, channelCount, haloFrames)
}

/**
 * deserializeFrame over many sound units at once (what the CLI's `--json` dump, bin/cli.js:567-677, runs over a
 * file): frame objects equal to Array.from(units, deserializeFrame).  Not part of the reference's export list.
 * @param {Uint8Array} bytes - n * 212 bytes
 */
function deserializeFrames(bytes) {
  const [nBfu, modes, wl, sfi, q] = native.deserializeUnits(context(), bytes)
  const frames = new Array(nBfu.length)
  for (let i = 0; i < nBfu.length; i++) {
    const n = nBfu[i]
    const quantizedCoefficients = new Array(n)
    for (let b = 0; b < n; b++) {
      const at = 512 * i + BFU_START_LONG[b]
      quantizedCoefficients[b] = q.slice(at, at + SPECS_PER_BFU[b])
    }
    frames[i] = {
      nBfu: n,
      scaleFactorIndices: Int32Array.from(sfi.subarray(52 * i, 52 * i + n)),
      wordLengthIndices: Int32Array.from(wl.subarray(52 * i, 52 * i + n)),
      quantizedCoefficients,
      blockModes: [modes[3 * i], modes[3 * i + 1], modes[3 * i + 2]],
    }
  }
  return frames
}

export {
  deserializeFrames, encodePcmShard, decodeUnitsShard,
  pipe, encode, decode, qmfAnalysisStage, mdctStage, serializeFrame, deserializeFrame, quantize, dequantize, AeaFile,
  BufferPool, EncoderOptions, AudioProcessor, decodeAeaPcm, encodeAeaPcm, FFT, WORD_LENGTH_BITS, SPECS_PER_BFU,
  SCALE_FACTORS, BFU_START_LONG,
}
