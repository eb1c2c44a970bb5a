// Driver evaluated by tools/ref_run_qjs.py inside Qt's QJSEngine after the reference's modules were imported
// and bound to globals: carta1 (codec/index.js), C (constants), M (mdct), ENC (encoder), DEC (decoder), QMF, TR
// (transient), BA (bitallocation), QZ (quantization), BS (bitstream), SER (serialization), BUF (buffers), OPT, FFTM.
// Everything computed here is computed by the reference's own functions; this file only feeds them and
// serialises what they return (ES2016 only: the engine has no object spread, async or padStart).

var HEX = []
for (var i = 0; i < 256; i++) HEX.push((i < 16 ? '0' : '') + i.toString(16))

function hexOfBytes(u8) {
  var parts = new Array(u8.length)
  for (var i = 0; i < u8.length; i++) parts[i] = HEX[u8[i]]
  return parts.join('')
}
function hexOf(arr) {
  return hexOfBytes(new Uint8Array(arr.buffer, arr.byteOffset, arr.byteLength))
}
function fromHex(hex) {
  var out = new Uint8Array(hex.length / 2)
  for (var i = 0; i < out.length; i++) out[i] = parseInt(hex.substr(2 * i, 2), 16)
  return out
}
function f64hex(x) {
  // IEEE-754 bit pattern, big-endian hex (the format of tools/ref_dump.mjs)
  var dv = new DataView(new ArrayBuffer(8))
  dv.setFloat64(0, x, false)
  var s = ''
  for (var i = 0; i < 8; i++) s += HEX[dv.getUint8(i)]
  return s
}
function f64hexes(arr) {
  var out = []
  for (var i = 0; i < arr.length; i++) out.push(f64hex(arr[i]))
  return out
}
function packed(type, shape, arr) {
  return { type: type, shape: shape, hex: hexOf(arr) }
}

// ---- tables (the layout of tools/ref_dump.mjs) ------------------------------------------------------
function dumpTables(biases) {
  var fftW = []
  for (var k = 0; k < 8; k++) {
    // codec/transforms/fft.js:36-39, the same expression
    var stride = 2 << k
    var angle = (-2 * Math.PI) / stride
    fftW.push([f64hex(Math.cos(angle)), f64hex(Math.sin(angle))])
  }
  var biased = {}
  biases.forEach(function (bias) {
    // codec/coding/bitallocation.js:46-61
    var out = new Float64Array(64)
    for (var i = 0; i < 64; i++) out[i] = bias === 1 ? C.SCALE_FACTORS[i] : Math.pow(C.SCALE_FACTORS[i], bias)
    biased[String(bias)] = f64hexes(out)
  })
  return {
    window_short: f64hexes(C.WINDOW_SHORT),
    scale_factors: f64hexes(C.SCALE_FACTORS),
    mdct_fwd64: f64hexes(M.mdct64.sinCosTable),
    mdct_fwd256: f64hexes(M.mdct256.sinCosTable),
    mdct_fwd512: f64hexes(M.mdct512.sinCosTable),
    mdct_inv64: f64hexes(M.imdct64.sinCosTable),
    mdct_inv256: f64hexes(M.imdct256.sinCosTable),
    mdct_inv512: f64hexes(M.imdct512.sinCosTable),
    fft_w: fftW,
    biased_scale_factors: biased,
    log1p_10: f64hex(Math.log1p(10)),
    // libm probes: what this engine's Math returns, so a reader can tell which libm the dump carries
    libm_probe: {
      'log(3)': f64hex(Math.log(3)), 'exp(0.7)': f64hex(Math.exp(0.7)), 'log10(7)': f64hex(Math.log10(7)),
      'pow(1.1,2.5)': f64hex(Math.pow(1.1, 2.5)), 'sin(0.3)': f64hex(Math.sin(0.3)), 'cos(1.3)': f64hex(Math.cos(1.3)),
    },
  }
}

// ---- whole-file API ------------------------------------------------------------------------------------
var channels = []
var lastAea = null

function setInput(hex, nCh) {
  var raw = fromHex(hex)
  var dv = new DataView(raw.buffer)
  var n = raw.length / 2 / nCh
  channels = []
  for (var ch = 0; ch < nCh; ch++) channels.push(new Float32Array(n))
  // bin/cli.js:394-396: readInt16LE / 32768.0 into a Float32Array
  for (var i = 0; i < n; i++) for (var c = 0; c < nCh; c++) channels[c][i] = dv.getInt16((i * nCh + c) * 2, true) / 32768.0
  return n
}
function setInputF32(hexes) {
  channels = hexes.map(function (hx) { var b = fromHex(hx); return new Float32Array(b.buffer) })
  return channels.length
}
function setAea(hex) {
  lastAea = fromHex(hex)
  return lastAea.length
}
function aeaOf(unitsHex, nCh) {
  // a complete AEA image around caller-supplied sound units: the reference's own header writer
  var units = fromHex(unitsHex)
  var header = SER.AeaFile.createHeader('arbitrary', units.length / 212, nCh)
  var all = new Uint8Array(header.length + units.length)
  all.set(header, 0)
  all.set(units, header.length)
  lastAea = all
  return all.length
}
// detectTransient returns only `score > threshold` (transient.js:54) and the score function is not exported;
// the exact binary64 score is recovered from the unmodified function by bisection over the threshold's bit pattern.
function scoreOf(cur, prev) {
  if (!TR.detectTransient(cur, prev, 0)) return TR.detectTransient(cur, prev, -1) ? 0 : -1
  var dv = new DataView(new ArrayBuffer(8))
  function f(hiWord, loWord) { dv.setUint32(0, hiWord, false); dv.setUint32(4, loWord, false); return dv.getFloat64(0, false) }
  if (TR.detectTransient(cur, prev, f(0x7ff00000, 0))) return Infinity
  // the smallest bit pattern t (as a 64-bit integer, positive doubles are ordered like their patterns) with
  // !(score > f(t)) is the score itself; 32 steps for the high word, 32 for the low one
  var lo = 0, hi = 0x7ff00000 // score > f(lo, 0), !(score > f(hi, 0))
  while (hi - lo > 1) {
    var mid = Math.floor((lo + hi) / 2)
    if (TR.detectTransient(cur, prev, f(mid, 0))) lo = mid
    else hi = mid
  }
  // score lies in (f(lo,0), f(hi,0)] : either f(hi,0) itself or f(lo, x) for the smallest x with !(score > f(lo,x))
  if (TR.detectTransient(cur, prev, f(lo, 0xffffffff))) return f(hi, 0)
  var a = 0, b = 0xffffffff // score > f(lo,a), !(score > f(lo,b))
  while (b - a > 1) {
    var m = Math.floor((a + b) / 2)
    if (TR.detectTransient(cur, prev, f(lo, m))) a = m
    else b = m
  }
  return f(lo, b)
}
function runEncode(options) {
  lastAea = carta1.encodeAeaPcm(channels, options) // codec/io/processor.js:597-617
  return hexOfBytes(lastAea)
}
function runDecode() {
  var pcm = carta1.decodeAeaPcm(lastAea) // codec/io/processor.js:628-654
  var total = 0
  pcm.forEach(function (p) { total += p.length })
  var flat = new Float32Array(total)
  var at = 0
  pcm.forEach(function (p) { flat.set(p, at); at += p.length })
  return hexOf(flat)
}

// the CLI's decode-to-WAV route: parseAeaBlob -> deserializedFrameStream -> decodeStream -> collectFrames ->
// createWavBlob (processor.js:147-215, 286-293, 349-447, 511-536); returns the WAV file
function runWav() {
  var AP = carta1.AudioProcessor
  var parsed = AP.parseAeaBlob(new Blob([lastAea]))
  var frames = AP.deserializedFrameStream(parsed.frameData)
  var decoded = AP.collectFrames(AP.decodeStream(frames, { channelCount: parsed.info.channelCount }))
  var wav = AP.createWavBlob(decoded, parsed.info.channelCount)
  return hexOfBytes(new Uint8Array(wav.arrayBuffer()))
}

// EncoderOptions as the reference validates them: for each trial either the resulting values or the error text
function runOptionTrials(trials) {
  return trials.map(function (t) {
    try {
      var o = new OPT.EncoderOptions(t)
      return { values: o.toObject().values }
    } catch (e) {
      return { error: String(e.message), name: e.name }
    }
  })
}

// what the reference throws for bad input: [label, error name, message] per trial
function runErrorTrials() {
  var AP = carta1.AudioProcessor, f32 = new Float32Array(512)
  function drain(g) { var n = 0; for (var x of g) n++; return n }
  var trials = [
    ['encodeAeaPcm: no channels', function () { carta1.encodeAeaPcm([]) }],
    ['encodeAeaPcm: three channels', function () { carta1.encodeAeaPcm([f32, f32, f32]) }],
    ['encodeAeaPcm: Float64Array channel', function () { carta1.encodeAeaPcm([new Float64Array(512)]) }],
    ['encodeAeaPcm: not an array', function () { carta1.encodeAeaPcm(f32) }],
    ['encodeAeaPcm: option out of range', function () { carta1.encodeAeaPcm([f32], { allocationBias: 9 }) }],
    ['decodeAeaPcm: string', function () { carta1.decodeAeaPcm('abc') }],
    ['decodeAeaPcm: short buffer', function () { carta1.decodeAeaPcm(new Uint8Array(100)) }],
    ['decodeAeaPcm: bad magic', function () { carta1.decodeAeaPcm(new Uint8Array(2048 + 212)) }],
    ['deserializeFrame: 211 bytes', function () { SER.deserializeFrame(new Uint8Array(211)) }],
    ['deserializeFrame: 213 bytes', function () { SER.deserializeFrame(new Uint8Array(213)) }],
    ['parseHeader: 2047 bytes', function () { SER.AeaFile.parseHeader(new Uint8Array(2047)) }],
    ['parseHeader: bad magic', function () { SER.AeaFile.parseHeader(new Uint8Array(2048)) }],
    ['setValue: unknown option', function () { new OPT.EncoderOptions().setValue('x', 1) }],
    ['getValue: unknown option', function () { new OPT.EncoderOptions().getValue('x') }],
    ['encodeStream: three channels', function () { drain(AP.encodeStream([f32], { channelCount: 3 })) }],
    ['decodeStream: zero channels', function () { drain(AP.decodeStream([], { channelCount: 0 })) }],
  ]
  return trials.map(function (t) {
    try { t[1](); return [t[0], null, null] } catch (e) { return [t[0], e.name, String(e.message)] }
  })
}

// ---- the same run stage by stage -----------------------------------------------------------------------
// encode(options) is pipe(context, qmfAnalysisStage, blockSelectorStage, mdctStage, quantizationStage)
// (codec/pipeline/encoder.js:438-450) and decode() is pipe(context, dequantizationStage, imdctStage,
// qmfSynthesisStage) (decoder.js:408-411); here the same stage closures are called one after the other on a
// context of their own so that what passes between them can be copied out.
function runStages(optionValues) {
  var nCh = channels.length
  var n = channels[0].length
  var nFrames = Math.ceil(n / 512)
  var fftSizes = [C.FFT_SIZE_LOW, C.FFT_SIZE_MID, C.FFT_SIZE_HIGH]
  var bands = new Float32Array(nCh * nFrames * 512), mags = new Float32Array(nCh * nFrames * 256)
  var modes = new Int32Array(nCh * nFrames * 3), coefs = new Float32Array(nCh * nFrames * 512)
  var nBfu = new Int32Array(nCh * nFrames), sfi = new Int32Array(nCh * nFrames * 52), wl = new Int32Array(nCh * nFrames * 52)
  var q = new Int32Array(nCh * nFrames * 52 * 20), su = new Uint8Array(nCh * nFrames * 212)
  var dcoefs = new Float32Array(nCh * nFrames * 512), dbands = new Float32Array(nCh * nFrames * 512)
  var dpcm = new Float32Array(nCh * nFrames * 512), scores = new Float64Array(nCh * nFrames * 3)
  for (var ch = 0; ch < nCh; ch++) {
    var ectx = { options: new OPT.EncoderOptions(optionValues), bufferPool: new BUF.BufferPool() }
    var s1 = ENC.qmfAnalysisStage(ectx), s2 = ENC.blockSelectorStage(ectx), s3 = ENC.mdctStage(ectx), s4 = ENC.quantizationStage(ectx)
    var dctx = { bufferPool: new BUF.BufferPool() }
    var d1 = DEC.dequantizationStage(dctx), d2 = DEC.imdctStage(dctx), d3 = DEC.qmfSynthesisStage(dctx)
    for (var f = 0; f < nFrames; f++) {
      var u = ch * nFrames + f
      // AudioProcessor.frameBufferToFrames zero-pads the last frame (processor.js)
      var pcm = new Float32Array(512)
      pcm.set(channels[ch].subarray(f * 512, Math.min(n, f * 512 + 512)))
      var a = s1(pcm)
      bands.set(a.bands[0], u * 512); bands.set(a.bands[1], u * 512 + 128); bands.set(a.bands[2], u * 512 + 256)
      if (!ectx.options.fixedBlockModes) {
        // performFFT is a pure function of the band (transient.js:17-35): calling it again changes nothing
        var off = 0
        for (var b = 0; b < 3; b++) {
          var mg = TR.performFFT(a.bands[b], fftSizes[b])
          mags.set(mg, u * 256 + off)
          off += fftSizes[b] / 2
          scores[u * 3 + b] = scoreOf(mg, ectx.bufferPool.transientDetection[b])
        }
      }
      var bsel = s2(a)
      for (var b2 = 0; b2 < 3; b2++) modes[u * 3 + b2] = bsel.blockModes[b2]
      var c = s3(bsel)
      coefs.set(c.coefficients, u * 512)
      var fr = s4(c)
      nBfu[u] = fr.nBfu
      for (var k = 0; k < fr.nBfu; k++) {
        sfi[u * 52 + k] = fr.scaleFactorIndices[k]
        wl[u * 52 + k] = fr.wordLengthIndices[k]
        q.set(fr.quantizedCoefficients[k], (u * 52 + k) * 20)
      }
      var bytes = SER.serializeFrame(fr)
      su.set(bytes, u * 212)
      var back = SER.deserializeFrame(bytes)
      var dq = d1(back)
      dcoefs.set(dq.coefficients, u * 512)
      var tb = d2(dq)
      dbands.set(tb[0], u * 512); dbands.set(tb[1], u * 512 + 128); dbands.set(tb[2], u * 512 + 256)
      dpcm.set(d3(tb), u * 512)
    }
  }
  return {
    enc_bands: packed('f32', [nCh, nFrames, 512], bands), enc_mags: packed('f32', [nCh, nFrames, 256], mags),
    enc_modes: packed('i32', [nCh, nFrames, 3], modes), enc_coefs: packed('f32', [nCh, nFrames, 512], coefs),
    n_bfu: packed('i32', [nCh, nFrames], nBfu), sfi: packed('i32', [nCh, nFrames, 52], sfi), wl: packed('i32', [nCh, nFrames, 52], wl),
    q: packed('i32', [nCh, nFrames, 52, 20], q), su: packed('u8', [nCh, nFrames, 212], su),
    dec_coefs: packed('f32', [nCh, nFrames, 512], dcoefs), dec_bands: packed('f32', [nCh, nFrames, 512], dbands),
    dec_pcm: packed('f32', [nCh, nFrames, 512], dpcm), enc_scores: packed('f64', [nCh, nFrames, 3], scores),
  }
}

// ---- known answers of single functions -----------------------------------------------------------------
// Inputs come from a 32-bit LCG so the Python side can rebuild them; they are also written out.
var seed = 12345
function rnd() { seed = (Math.imul(seed, 1664525) + 1013904223) >>> 0; return seed / 4294967296 }
function randF32(n, amp) {
  var a = new Float32Array(n)
  for (var i = 0; i < n; i++) a[i] = (rnd() * 2 - 1) * amp
  return a
}
function h(arr) { return hexOf(arr) }

function runKats() {
  var out = { fft: [], mdct: [], imdct: [], qmf_analysis: [], qmf_synthesis: [], overlap_add: [], find_scale_factor: [],
              quantize: [], dequantize: [], allocate_bits: [], pack_bits: [], perform_fft: [], detect_transient: [] }
  var pool = new BUF.BufferPool()
  var amps = [1, 1e-3, 30000, 1e-20]
  ;[16, 64, 128, 256].forEach(function (n) {
    amps.forEach(function (amp) {
      var re = randF32(n, amp), im = randF32(n, amp)
      var r0 = h(re), i0 = h(im)
      FFTM.FFT.fft(re, im)
      out.fft.push({ n: n, re_in: r0, im_in: i0, re: h(re), im: h(im) })
    })
  })
  ;[[M.mdct64, M.imdct64, 64], [M.mdct256, M.imdct256, 256], [M.mdct512, M.imdct512, 512]].forEach(function (t) {
    amps.forEach(function (amp) {
      var x = randF32(t[2], amp)
      out.mdct.push({ n: t[2], x: h(x), y: h(t[0].transform(x, pool.mdctBuffers)) })
      var y = randF32(t[2] / 2, amp)
      out.imdct.push({ n: t[2], x: h(y), y: h(t[1].transform(y, pool.mdctBuffers)) })
    })
  })
  ;[512, 256].forEach(function (n) {
    amps.forEach(function (amp) {
      var x = randF32(n, amp), d = randF32(C.QMF_DELAY, amp)
      var r = QMF.qmfAnalysis(x, d, pool.qmfWorkBuffers)
      out.qmf_analysis.push({ x: h(x), delay: h(d), low: h(r.lowBand), high: h(r.highBand), new_delay: h(r.newDelay) })
      var lo = randF32(n / 2, amp), hi = randF32(n / 2, amp), d2 = randF32(C.QMF_DELAY, amp)
      var s = QMF.qmfSynthesis(lo, hi, d2, pool.qmfWorkBuffers)
      out.qmf_synthesis.push({ low: h(lo), high: h(hi), delay: h(d2), out: h(s.output), new_delay: h(s.newDelay) })
    })
  })
  ;[16, 128].forEach(function (n) {
    var p = randF32(n, 1), c = randF32(n, 1), w = new Float64Array(2 * n)
    for (var i = 0; i < 2 * n; i++) w[i] = rnd()
    out.overlap_add.push({ prev: h(p), curr: h(c), window: h(w), out: h(M.overlapAdd(p, c, w)) })
  })
  // findScaleFactor: random magnitudes over the whole table, and every table entry and its two binary32 neighbours
  for (var t = 0; t < 200; t++) {
    var len = [6, 7, 8, 9, 10, 12, 20][t % 7]
    var v = randF32(len, Math.pow(2, -16 + 17 * rnd()))
    out.find_scale_factor.push({ x: h(v), sfi: BA.findScaleFactor(v, len) })
  }
  var one = new Float32Array(1), bits = new Uint32Array(one.buffer)
  for (var s = 0; s < 64; s++) {
    one[0] = C.SCALE_FACTORS[s]
    var centre = bits[0]
    for (var dlt = -1; dlt <= 1; dlt++) {
      var probe = new Float32Array(8)
      bits[0] = centre + dlt
      probe[3] = -one[0]
      out.find_scale_factor.push({ x: h(probe), sfi: BA.findScaleFactor(probe, 8) })
    }
  }
  for (var t2 = 0; t2 < 120; t2++) {
    var sf = Math.floor(rnd() * 64), b = 2 + Math.floor(rnd() * 15), n2 = [6, 8, 10, 12, 20][t2 % 5]
    var cf = randF32(n2, C.SCALE_FACTORS[sf] * (t2 % 9 === 0 ? 1.5 : 1))
    out.quantize.push({ x: h(cf), sfi: sf, bits: b, q: h(QZ.quantize(cf, sf, b)) })
    var qi = new Int32Array(n2), lim = (1 << (b - 1)) - 1
    for (var i2 = 0; i2 < n2; i2++) qi[i2] = Math.round((rnd() * 2 - 1) * lim)
    out.dequantize.push({ q: h(qi), sfi: sf, bits: b, x: h(QZ.dequantize(qi, sf, b)) })
  }
  // allocateBits on whole coefficient rows (groupIntoBFUs first, as quantizationStage does)
  var modeSets = [[0, 0, 0], [2, 2, 3], [0, 2, 0], [2, 0, 3]]
  var biasSet = [1, 0, 0.5, 2.5]
  for (var t3 = 0; t3 < 48; t3++) {
    var row = new Float32Array(512), shape = t3 % 4
    for (var i3 = 0; i3 < 512; i3++) {
      var env = shape === 0 ? 1 : shape === 1 ? Math.exp(-i3 / 60) : shape === 2 ? (i3 % 37 === 0 ? 1 : 0.001) : 1 / (1 + i3)
      row[i3] = (rnd() * 2 - 1) * env * (t3 % 5 === 4 ? 1e-4 : 0.5)
    }
    var md = modeSets[(t3 >> 2) % 4], bias = biasSet[(t3 >> 4) % 4]
    var g = QZ.groupIntoBFUs(row, md)
    var r2 = BA.allocateBits(g.bfuData, g.bfuSizes, g.bfuCount, bias)
    var alloc = new Int32Array(52), sfis = new Int32Array(52)
    for (var k = 0; k < r2.bfuCount; k++) { alloc[k] = r2.allocation[k]; sfis[k] = r2.scaleFactorIndices[k] }
    out.allocate_bits.push({ coefs: h(row), modes: md, bias: bias, n_bfu: r2.bfuCount, wl: h(alloc), sfi: h(sfis) })
  }
  for (var t4 = 0; t4 < 40; t4++) {
    var buf = new Uint8Array(16), pos = 0, ops = []
    while (pos < 100) {
      var cnt = 1 + Math.floor(rnd() * 16), val = Math.floor(rnd() * (1 << cnt))
      BS.packBits(buf, pos, val, cnt)
      ops.push([pos, val, cnt])
      pos += cnt
    }
    out.pack_bits.push({ ops: ops, bytes: hexOfBytes(buf) })
  }
  // performFFT + detectTransient on pairs of consecutive spectra (the decision only: the score is not exported)
  ;[[128, 128], [256, 256]].forEach(function (t) {
    for (var r = 0; r < 24; r++) {
      var x1 = randF32(t[0], 0.3), x2 = randF32(t[0], 0.3)
      if (r % 3 === 1) for (var i = t[0] / 2; i < t[0]; i++) x2[i] *= 8
      if (r % 3 === 2) for (var j = 0; j < t[0]; j++) x2[j] = x1[j] * 1.01
      var m1 = TR.performFFT(x1, t[1]), m2 = TR.performFFT(x2, t[1])
      out.perform_fft.push({ x: h(x1), size: t[1], mag: h(m1) })
      var thr = [0.05, 0.3, 1, 3][r % 4]
      out.detect_transient.push({ prev: h(m1), cur: h(m2), threshold: thr, transient: TR.detectTransient(m2, m1, thr) ? 1 : 0,
                                  score: f64hex(scoreOf(m2, m1)) })
    }
  })
  return out
}
