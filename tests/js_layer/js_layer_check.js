// Runs carta1_b200/napi/index.mjs (global B200) next to the reference (global carta1) inside Qt's QJSEngine and
// compares what the two return for the same calls.  The addon under index.mjs is tests/js_layer/mock_native.js.
// TEST INFRASTRUCTURE ONLY (tools/ref_run_qjs.py --check-js-layer); ES2016 only.
function runJsLayerChecks() {
  var results = []
  function check(name, fn) {
    try { var r = fn(); results.push([name, r === true, r === true ? '' : String(r)]) } catch (e) { results.push([name, false, 'threw ' + e + ' ' + (e.stack || '')]) }
  }
  function sameArr(a, b) {
    if (a.length !== b.length) return false
    for (var i = 0; i < a.length; i++) if (!(a[i] === b[i] || (a[i] !== a[i] && b[i] !== b[i]))) return false
    return true
  }
  function sameBits(a, b) { return sameArr(new Uint32Array(a.buffer, a.byteOffset, a.length), new Uint32Array(b.buffer, b.byteOffset, b.length)) }
  function sameFrame(a, b) {
    if (a.nBfu !== b.nBfu || !sameArr(a.blockModes, b.blockModes)) return 'header'
    if (!sameArr(a.scaleFactorIndices, b.scaleFactorIndices)) return 'scaleFactorIndices'
    if (!sameArr(a.wordLengthIndices, b.wordLengthIndices)) return 'wordLengthIndices'
    if (a.quantizedCoefficients.length !== b.quantizedCoefficients.length) return 'quantizedCoefficients length'
    for (var k = 0; k < a.nBfu; k++) if (!sameArr(a.quantizedCoefficients[k], b.quantizedCoefficients[k])) return 'quantizedCoefficients[' + k + ']'
    return true
  }
  var seed = 99
  function rnd() { seed = (Math.imul(seed, 1664525) + 1013904223) >>> 0; return seed / 4294967296 }
  function signal(n, burstAt) {
    var a = new Float32Array(n)
    for (var i = 0; i < n; i++) a[i] = 0.4 * Math.sin(i * 0.05) + 0.05 * (rnd() * 2 - 1) + (burstAt >= 0 && i >= burstAt && i < burstAt + 150 ? 0.8 * (rnd() * 2 - 1) : 0)
    return a
  }
  function framesOf(ch) { var f = []; for (var i = 0; i + 512 <= ch.length; i += 512) f.push(ch.slice(i, i + 512)); return f }
  function collect(g) { var out = []; for (var x of g) out.push(x); return out }

  check('exports: every name of carta1 is exported', function () {
    var missing = Object.keys(carta1).filter(function (k) { return !(k in B200) })
    return missing.length === 0 ? true : 'missing ' + missing
  })
  check('non-hot exports are the reference objects themselves', function () {
    var names = ['serializeFrame', 'deserializeFrame', 'quantize', 'dequantize', 'AeaFile', 'BufferPool', 'EncoderOptions', 'FFT', 'pipe',
                 'qmfAnalysisStage', 'mdctStage', 'WORD_LENGTH_BITS', 'SPECS_PER_BFU', 'SCALE_FACTORS', 'BFU_START_LONG']
    var bad = names.filter(function (k) { return B200[k] !== carta1[k] })
    return bad.length === 0 ? true : 'differ: ' + bad
  })

  var L = signal(512 * 9 + 77, 2300), R = signal(512 * 7 + 5, -1)
  check('encodeAeaPcm: ragged stereo, title, bias, fixed modes', function () {
    var opts = { title: 'drop-in', allocationBias: 2.5, fixedBlockModes: [0, 2, 0] }
    return sameArr(B200.encodeAeaPcm([L, R], opts), carta1.encodeAeaPcm([L, R], opts))
  })
  var aea = carta1.encodeAeaPcm([L, R], { transientThresholdLow: 0.3 })
  check('encodeAeaPcm: defaults + threshold, auto block modes', function () { return sameArr(B200.encodeAeaPcm([L, R], { transientThresholdLow: 0.3 }), aea) })
  check('decodeAeaPcm: Uint8Array, ArrayBuffer, Blob', function () {
    var want = carta1.decodeAeaPcm(aea)
    var forms = [aea, aea.buffer.slice(aea.byteOffset, aea.byteOffset + aea.length), new Blob([aea])]
    for (var i = 0; i < forms.length; i++) {
      var got = B200.decodeAeaPcm(forms[i])
      if (got.length !== want.length) return 'channel count'
      for (var c = 0; c < want.length; c++) if (!sameBits(got[c], want[c])) return 'form ' + i + ' channel ' + c
    }
    return true
  })
  check('decodeAeaPcm: odd unit count in a stereo file (dummy frame), trailing partial unit', function () {
    var cut = aea.slice(0, aea.length - 212 - 100)
    var hdr = carta1.AeaFile.createHeader('x', (cut.length - 2048 - 112) / 212, 2)
    cut.set(hdr, 0)
    var want = carta1.decodeAeaPcm(cut), got = B200.decodeAeaPcm(cut)
    for (var c = 0; c < 2; c++) if (!sameBits(got[c], want[c])) return 'channel ' + c
    return true
  })
  check('encode() / decode() closures, frame by frame', function () {
    var o = new carta1.EncoderOptions({ transientThresholdLow: 0.3 })
    var e1 = B200.encode(o), e2 = carta1.encode(new carta1.EncoderOptions({ transientThresholdLow: 0.3 }))
    var d1 = B200.decode(), d2 = carta1.decode()
    var fr = framesOf(L)
    for (var i = 0; i < fr.length; i++) {
      var a = e1(fr[i]), b = e2(fr[i])
      var s = sameFrame(a, b)
      if (s !== true) return 'frame ' + i + ': ' + s
      if (!sameArr(carta1.serializeFrame(a), carta1.serializeFrame(b))) return 'frame ' + i + ': bytes'
      if (!sameBits(d1(a), d2(b))) return 'frame ' + i + ': pcm'
    }
    return true
  })
  check('AudioProcessor.encodeStream: mono, 70 frames (two batches), onProgress', function () {
    var x = signal(512 * 70, 512 * 64 + 100), p1 = [], p2 = []
    var a = collect(B200.AudioProcessor.encodeStream(framesOf(x), { channelCount: 1, onProgress: function (i) { p1.push(i) }, encoderOptions: new carta1.EncoderOptions({ allocationBias: 0.5 }) }))
    var b = collect(carta1.AudioProcessor.encodeStream(framesOf(x), { channelCount: 1, onProgress: function (i) { p2.push(i) }, encoderOptions: new carta1.EncoderOptions({ allocationBias: 0.5 }) }))
    if (a.length !== b.length || a.length !== 70) return 'yield count ' + a.length + ' vs ' + b.length
    for (var i = 0; i < a.length; i++) { var s = sameFrame(a[i], b[i]); if (s !== true) return 'yield ' + i + ': ' + s }
    return sameArr(p1, p2) ? true : 'onProgress ' + p1 + ' vs ' + p2
  })
  var stereoFrames = null
  check('AudioProcessor.encodeStream: stereo yields L, R, L, R', function () {
    var fl = framesOf(L), fr = framesOf(R), pairs = []
    for (var i = 0; i < fr.length; i++) pairs.push([fl[i], fr[i]])
    var p1 = [], p2 = []
    var a = collect(B200.AudioProcessor.encodeStream(pairs, { channelCount: 2, onProgress: function (i) { p1.push(i) } }))
    var b = collect(carta1.AudioProcessor.encodeStream(pairs, { channelCount: 2, onProgress: function (i) { p2.push(i) } }))
    stereoFrames = b
    if (a.length !== b.length || a.length !== 2 * fr.length) return 'yield count'
    for (var i2 = 0; i2 < a.length; i2++) { var s = sameFrame(a[i2], b[i2]); if (s !== true) return 'yield ' + i2 + ': ' + s }
    return sameArr(p1, p2) ? true : 'onProgress'
  })
  check('AudioProcessor.decodeStream: stereo with an odd frame count (dummy frame), mono', function () {
    var odd = stereoFrames.slice(0, stereoFrames.length - 1), p1 = [], p2 = []
    var a = collect(B200.AudioProcessor.decodeStream(odd, { channelCount: 2, onProgress: function (i) { p1.push(i) } }))
    var b = collect(carta1.AudioProcessor.decodeStream(odd, { channelCount: 2, onProgress: function (i) { p2.push(i) } }))
    if (a.length !== b.length) return 'yield count ' + a.length + ' vs ' + b.length
    for (var i = 0; i < a.length; i++) for (var c = 0; c < 2; c++) if (!sameBits(a[i][c], b[i][c])) return 'pair ' + i + ' channel ' + c
    if (!sameArr(p1, p2)) return 'onProgress'
    var m1 = collect(B200.AudioProcessor.decodeStream(stereoFrames.slice(0, 5), { channelCount: 1 }))
    var m2 = collect(carta1.AudioProcessor.decodeStream(stereoFrames.slice(0, 5), { channelCount: 1 }))
    for (var k = 0; k < 5; k++) if (!sameBits(m1[k], m2[k])) return 'mono frame ' + k
    return true
  })
  check('deserializeFrames(bytes) equals Array.from(units, deserializeFrame)', function () {
    var body = aea.slice(2048), got = B200.deserializeFrames(body)
    if (got.length !== body.length / 212) return 'count'
    for (var i = 0; i < got.length; i++) {
      var s = sameFrame(got[i], carta1.deserializeFrame(body.slice(i * 212, i * 212 + 212)))
      if (s !== true) return 'unit ' + i + ': ' + s
    }
    return true
  })
  check('encodePcmShard / decodeUnitsShard equal the matching span of the whole-stream result', function () {
    var x = signal(512 * 12, 3000), whole = carta1.encodeAeaPcm([x], {}).slice(2048)
    var su = B200.encodePcmShard([x.slice(512 * 3)], 2, {})
    if (!sameArr(su, whole.slice(5 * 212))) return 'encode shard'
    var pcmWhole = carta1.decodeAeaPcm(carta1.encodeAeaPcm([x], {}))[0]
    var pcm = B200.decodeUnitsShard(whole.slice(4 * 212), 1, 1)
    return sameBits(pcm[0], pcmWhole.slice(5 * 512)) ? true : 'decode shard'
  })
  check('errors: same class and text as the reference', function () {
    var f32 = new Float32Array(512)
    var trials = [
      function (m) { m.encodeAeaPcm([]) }, function (m) { m.encodeAeaPcm([f32, f32, f32]) }, function (m) { m.encodeAeaPcm([new Float64Array(8)]) },
      function (m) { m.encodeAeaPcm([f32], { allocationBias: 9 }) }, function (m) { m.decodeAeaPcm('abc') }, function (m) { m.decodeAeaPcm(new Uint8Array(100)) },
      function (m) { m.decodeAeaPcm(new Uint8Array(2048 + 212)) }, function (m) { collect(m.AudioProcessor.encodeStream([f32], { channelCount: 3 })) },
      function (m) { collect(m.AudioProcessor.decodeStream([], { channelCount: 0 })) },
    ]
    for (var i = 0; i < trials.length; i++) {
      var got = null, want = null
      try { trials[i](B200) } catch (e) { got = e.name + ': ' + e.message }
      try { trials[i](carta1) } catch (e2) { want = e2.name + ': ' + e2.message }
      if (got !== want || want === null) return 'trial ' + i + ': ' + got + ' vs ' + want
    }
    return true
  })
  check('hostTables(): the nine tables handed to createContext equal the reference\'s own, bit for bit', function () {
    var t = __tables
    if (!t) return 'createContext never saw tables'
    function same64(a, b) { return sameArr(new Uint32Array(Float64Array.from(a).buffer), new Uint32Array(Float64Array.from(b).buffer)) }
    var pairs = [[t.windowShort, C.WINDOW_SHORT], [t.scaleFactors, C.SCALE_FACTORS], [t.mdctFwd64, M.mdct64.sinCosTable], [t.mdctFwd256, M.mdct256.sinCosTable],
                 [t.mdctFwd512, M.mdct512.sinCosTable], [t.mdctInv64, M.imdct64.sinCosTable], [t.mdctInv256, M.imdct256.sinCosTable], [t.mdctInv512, M.imdct512.sinCosTable]]
    for (var i = 0; i < pairs.length; i++) if (!same64(pairs[i][0], pairs[i][1])) return 'table ' + i
    for (var k = 0; k < 8; k++) {
      var angle = (-2 * Math.PI) / (2 << k)
      if (t.fftW[2 * k] !== Math.cos(angle) || t.fftW[2 * k + 1] !== Math.sin(angle)) return 'fftW ' + k
    }
    return true
  })
  check('one shared context; the four stream calls destroy their handles, the two closures leave theirs to the finalizer', function () {
    var creates = __calls.filter(function (c) { return c[0] === 'createContext' }).length
    var encs = __calls.filter(function (c) { return c[0] === 'createEncoder' || c[0] === 'createDecoder' }).length
    var destroys = __calls.filter(function (c) { return c[0] === 'destroy' }).length
    if (creates !== 1) return creates + ' contexts'
    return destroys === 4 && encs === 6 ? true : destroys + ' destroys for ' + encs + ' handles'
  })
  return { results: results, calls: __calls.length }
}
