// The addon contract of carta1_b200/napi/carta1_napi.c, implemented in JavaScript over the REFERENCE's own functions,
// so that carta1_b200/napi/index.mjs (the drop-in's JavaScript layer) can run inside Qt's QJSEngine, where no N-API
// exists.  TEST INFRASTRUCTURE ONLY (tools/ref_run_qjs.py --check-js-layer).  What each export takes and returns is the
// header comment of carta1_napi.c; the C shim itself is executed separately, against tests/napi_host, on the GPU.
// Globals provided by the runner: carta1 (codec/index.js), ENC, DEC, SER, OPT, BUF, C.
var __tables = null
var __calls = []   // every call index.mjs makes on the addon: [name, summary of the arguments]
var __carta1_native = (function () {
  function optionsOf(o) {
    // abiOptions() -> EncoderOptions: the addon reads transientThresholdLow, allocationBias, fixedBlockModes only
    var v = { transientThresholdLow: o.transientThresholdLow, allocationBias: o.allocationBias }
    if (o.fixedBlockModes) v.fixedBlockModes = o.fixedBlockModes
    return new OPT.EncoderOptions(v)
  }
  function frameOfExpanded(q, sfi, bits, modes, at) {
    // the inverse of index.mjs expandFrame(): per-position arrays -> a frame object over all 52 BFUs
    var m = [modes[at * 3], modes[at * 3 + 1], modes[at * 3 + 2]]
    var fr = { nBfu: 52, blockModes: m, scaleFactorIndices: new Int32Array(52), wordLengthIndices: new Int32Array(52), quantizedCoefficients: [] }
    for (var b = 0; b < 52; b++) {
      var band = b < 20 ? 0 : b < 36 ? 1 : 2
      var pos = at * 512 + (m[band] === 0 ? C.BFU_START_LONG[b] : C.BFU_START_SHORT[b])
      var n = C.SPECS_PER_BFU[b]
      fr.wordLengthIndices[b] = bits[pos] > 0 ? bits[pos] - 1 : 0   // WORD_LENGTH_BITS[i] = i + 1 for i >= 1
      fr.scaleFactorIndices[b] = sfi[pos]
      fr.quantizedCoefficients.push(Int32Array.from(q.subarray(pos, pos + n)))
    }
    return fr
  }
  return {
    createContext: function (device, tables) { __tables = tables; __calls.push(['createContext', device, tables ? Object.keys(tables).sort().join(',') : null]); return { kind: 'ctx' } },
    createEncoder: function (ctx, opts, nStreams) {
      __calls.push(['createEncoder', nStreams, opts.transientThresholdLow, opts.allocationBias, opts.fixedBlockModes, opts.biasedScaleFactors.length])
      var e = { kind: 'enc', streams: [] }
      for (var s = 0; s < nStreams; s++) e.streams.push(ENC.encode(optionsOf(opts), new BUF.BufferPool()))
      return e
    },
    createDecoder: function (ctx, nStreams) {
      __calls.push(['createDecoder', nStreams])
      var d = { kind: 'dec', streams: [] }
      for (var s = 0; s < nStreams; s++) d.streams.push(DEC.decode(new BUF.BufferPool()))
      return d
    },
    encodeFrames: function (enc, pcm, nFrames) {   // pcm [stream][frame][512] -> Uint8Array [stream][frame][212]
      __calls.push(['encodeFrames', pcm.length, nFrames])
      var nS = enc.streams.length, out = new Uint8Array(nS * nFrames * 212)
      for (var s = 0; s < nS; s++) for (var k = 0; k < nFrames; k++) {
        var at = (s * nFrames + k) * 512
        out.set(SER.serializeFrame(enc.streams[s](pcm.slice(at, at + 512))), (s * nFrames + k) * 212)
      }
      return out
    },
    decodeFrames: function (dec, su, nFrames) {
      __calls.push(['decodeFrames', su.length, nFrames])
      var nS = dec.streams.length, out = new Float32Array(nS * nFrames * 512)
      for (var s = 0; s < nS; s++) for (var k = 0; k < nFrames; k++) {
        var at = (s * nFrames + k) * 212
        out.set(dec.streams[s](SER.deserializeFrame(su.slice(at, at + 212))), (s * nFrames + k) * 512)
      }
      return out
    },
    decodeFramesExpanded: function (dec, q, sfi, bits, modes, nFrames) {
      __calls.push(['decodeFramesExpanded', q.length, nFrames])
      var nS = dec.streams.length, out = new Float32Array(nS * nFrames * 512)
      for (var s = 0; s < nS; s++) for (var k = 0; k < nFrames; k++)
        out.set(dec.streams[s](frameOfExpanded(q, sfi, bits, modes, s * nFrames + k)), (s * nFrames + k) * 512)
      return out
    },
    encodePcm: function (ctx, channels, opts, haloFrames) {   // equal-length channels -> interleaved sound units
      __calls.push(['encodePcm', channels.length, channels[0].length, haloFrames === undefined ? null : haloFrames])
      var nCh = channels.length, n = channels[0].length, frames = Math.ceil(n / 512), halo = haloFrames || 0
      var encs = channels.map(function () { return ENC.encode(optionsOf(opts), new BUF.BufferPool()) })
      var out = new Uint8Array(Math.max(frames - halo, 0) * nCh * 212)
      for (var f = 0; f < frames; f++) for (var c = 0; c < nCh; c++) {
        var pcm = new Float32Array(512)
        pcm.set(channels[c].subarray(f * 512, Math.min(n, f * 512 + 512)))
        var bytes = SER.serializeFrame(encs[c](pcm))
        if (f >= halo) out.set(bytes, ((f - halo) * nCh + c) * 212)
      }
      return out
    },
    decodeSu: function (ctx, su, nCh, haloFrames) {
      __calls.push(['decodeSu', su.length, nCh, haloFrames === undefined ? null : haloFrames])
      var units = su.length / 212, frames = Math.ceil(units / nCh), halo = haloFrames || 0
      var decs = [], outs = []
      for (var c = 0; c < nCh; c++) { decs.push(DEC.decode(new BUF.BufferPool())); outs.push(new Float32Array(Math.max(frames - halo, 0) * 512)) }
      for (var f = 0; f < frames; f++) for (var c2 = 0; c2 < nCh; c2++) {
        var u = f * nCh + c2
        var fr = u < units ? SER.deserializeFrame(su.slice(u * 212, u * 212 + 212)) : carta1.AudioProcessor._createDummyFrame()
        var pcm = decs[c2](fr)
        if (f >= halo) outs[c2].set(pcm, (f - halo) * 512)
      }
      return outs
    },
    deserializeUnits: function (ctx, su) {
      __calls.push(['deserializeUnits', su.length])
      var n = su.length / 212
      var nBfu = new Uint8Array(n), modes = new Int8Array(3 * n), wl = new Uint8Array(52 * n), sfi = new Uint8Array(52 * n), q = new Int32Array(512 * n)
      for (var i = 0; i < n; i++) {
        var fr = SER.deserializeFrame(su.slice(i * 212, i * 212 + 212))
        nBfu[i] = fr.nBfu
        for (var b = 0; b < 3; b++) modes[3 * i + b] = fr.blockModes[b]
        for (var k = 0; k < fr.nBfu; k++) {
          wl[52 * i + k] = fr.wordLengthIndices[k]
          sfi[52 * i + k] = fr.scaleFactorIndices[k]
          q.set(fr.quantizedCoefficients[k], 512 * i + C.BFU_START_LONG[k])   // bitstream order at the long-block positions
        }
      }
      return [nBfu, modes, wl, sfi, q]
    },
    destroy: function (h) { __calls.push(['destroy', h && h.kind]); return undefined },
  }
})()
