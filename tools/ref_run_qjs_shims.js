// Host globals the reference expects from Node / a browser and Qt's QJSEngine does not provide.
// Loaded by tools/ref_run_qjs.py before the reference's modules.  None of this is codec arithmetic.
(function (g) {
  function bytesOf(part) {
    if (part instanceof Uint8Array) return part
    if (part instanceof ArrayBuffer) return new Uint8Array(part)
    if (part && part.buffer instanceof ArrayBuffer) return new Uint8Array(part.buffer, part.byteOffset, part.byteLength)
    throw new TypeError('Blob shim: unsupported part')
  }
  // Blob: only the constructor and arrayBuffer() are used (codec/io/processor.js:338,512,633); arrayBuffer()
  // returns the buffer itself instead of a promise, which is what the downlevelled (await-free) caller expects.
  g.Blob = class Blob {
    constructor(parts, options) {
      let total = 0
      const list = (parts || []).map(bytesOf)
      for (const p of list) total += p.length
      const all = new Uint8Array(total)
      let at = 0
      for (const p of list) { all.set(p, at); at += p.length }
      this._bytes = all
      this.size = total
      this.type = (options && options.type) || ''
    }
    arrayBuffer() { return this._bytes.buffer.slice(this._bytes.byteOffset, this._bytes.byteOffset + this._bytes.length) }
  }
  // TextEncoder / TextDecoder: the AEA title field (codec/io/serialization.js:198,244); ASCII titles only here.
  g.TextEncoder = class TextEncoder {
    encode(s) {
      const out = new Uint8Array(s.length)
      for (let i = 0; i < s.length; i++) {
        const c = s.charCodeAt(i)
        if (c > 127) throw new RangeError('TextEncoder shim: ASCII only')
        out[i] = c
      }
      return out
    }
  }
  g.TextDecoder = class TextDecoder {
    decode(b) {
      let s = ''
      for (let i = 0; i < b.length; i++) s += String.fromCharCode(b[i])
      return s
    }
  }
  if (!g.process) g.process = { env: {} }   // index.mjs reads process.env.CARTA1_B200_DEVICE
})(this)
