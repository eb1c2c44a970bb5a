#!/usr/bin/env node
// ref_dump.mjs -- the reference-side pin (SURVEY.md section 8c; VERDICT round 1, "What's missing" 1).
//
// Runs the REAL carta1 (aynik/carta1, JavaScript) on the inputs of tests/golden/*.npz and writes what its
// own code produces, so that the CPU oracle and the CUDA path can be compared with the reference itself
// instead of with each other.  The build image has no Node (tools/ref_run_qjs.py does the same job there under Qt's
// QJSEngine and writes the same layout, plus stage dumps and known answers); a maintainer
// with Node >= 20.16 (the version the reference's CI pins, .github/workflows/ci.yml:24-26) runs:
//
//     python tests/golden/make_golden.py --export-ref-inputs      # writes tests/golden/ref/inputs/
//     node tools/ref_dump.mjs /path/to/carta1                     # writes tests/golden/ref/
//     python -m pytest tests/test_reference_pin.py                # oracle (CPU) and CUDA path (-m gpu)
//
// Written per case <name> of tests/golden/ref/inputs/cases.json:
//     <name>.aea       encodeAeaPcm(channels, options)            codec/io/processor.js:597-617
//     <name>.pcm.f32   decodeAeaPcm(that)  planar f32 LE, channel after channel   processor.js:628-654
// and once:
//     tables.json      every libm-derived table of the reference as V8 computed it, IEEE-754 bit patterns in hex
//                      (the fields of carta1_tables in include/carta1_b200.h, plus pow(SCALE_FACTORS, bias) for
//                      every bias the cases use: codec/coding/bitallocation.js:46-61), and process.versions.
import { readFileSync, writeFileSync, mkdirSync } from 'node:fs'
import { dirname, join, resolve } from 'node:path'
import { fileURLToPath, pathToFileURL } from 'node:url'

const here = dirname(fileURLToPath(import.meta.url))
const refRoot = resolve(process.argv[2] ?? '')
if (!process.argv[2]) {
  console.error('usage: node tools/ref_dump.mjs /path/to/carta1 [outdir]')
  process.exit(2)
}
const outDir = resolve(process.argv[3] ?? join(here, '..', 'tests', 'golden', 'ref'))
const inDir = join(outDir, 'inputs')
mkdirSync(outDir, { recursive: true })

const load = (rel) => import(pathToFileURL(join(refRoot, rel)).href)
const { encodeAeaPcm, decodeAeaPcm } = await load('codec/index.js')
const C = await load('codec/core/constants.js')
const M = await load('codec/transforms/mdct.js')

// ---- tables -------------------------------------------------------------------------------------
const hex = (x) => {
  const dv = new DataView(new ArrayBuffer(8))
  dv.setFloat64(0, x, false)
  return dv.getBigUint64(0, false).toString(16).padStart(16, '0')
}
const hexes = (arr) => Array.from(arr, hex)

const fftW = []
for (let k = 0; k < 8; k++) {
  // codec/transforms/fft.js:36-39, the same expression
  const stride = 2 << k
  const angle = (-2 * Math.PI) / stride
  fftW.push([hex(Math.cos(angle)), hex(Math.sin(angle))])
}

const cases = JSON.parse(readFileSync(join(inDir, 'cases.json'), 'utf8'))
const biases = [...new Set(cases.map((c) => c.bias))]
const biased = {}
for (const bias of biases) {
  // codec/coding/bitallocation.js:46-61
  const out = new Float64Array(64)
  for (let i = 0; i < 64; i++) out[i] = bias === 1 ? C.SCALE_FACTORS[i] : Math.pow(C.SCALE_FACTORS[i], bias)
  biased[String(bias)] = hexes(out)
}

const tables = {
  versions: process.versions,
  carta1: JSON.parse(readFileSync(join(refRoot, 'package.json'), 'utf8')).version,
  window_short: hexes(C.WINDOW_SHORT),
  scale_factors: hexes(C.SCALE_FACTORS),
  mdct_fwd64: hexes(M.mdct64.sinCosTable),
  mdct_fwd256: hexes(M.mdct256.sinCosTable),
  mdct_fwd512: hexes(M.mdct512.sinCosTable),
  mdct_inv64: hexes(M.imdct64.sinCosTable),
  mdct_inv256: hexes(M.imdct256.sinCosTable),
  mdct_inv512: hexes(M.imdct512.sinCosTable),
  fft_w: fftW,
  biased_scale_factors: biased,
  // the one libm constant the transient score divides by (codec/analysis/transient.js:211)
  log1p_10: hex(Math.log1p(10)),
}
writeFileSync(join(outDir, 'tables.json'), JSON.stringify(tables, null, 1))

// ---- cases --------------------------------------------------------------------------------------
for (const c of cases) {
  const raw = readFileSync(join(inDir, c.name + '.s16'))
  const nCh = c.channels
  const n = raw.length / 2 / nCh
  const channels = []
  for (let ch = 0; ch < nCh; ch++) channels.push(new Float32Array(n))
  for (let i = 0; i < n; i++) {
    // bin/cli.js:394-396: readInt16LE / 32768.0 into a Float32Array
    for (let ch = 0; ch < nCh; ch++) channels[ch][i] = raw.readInt16LE((i * nCh + ch) * 2) / 32768.0
  }
  const options = { transientThresholdLow: c.threshold, allocationBias: c.bias }
  if (c.fixed_modes) options.fixedBlockModes = c.fixed_modes
  const aea = await encodeAeaPcm(channels, options)
  writeFileSync(join(outDir, c.name + '.aea'), aea)
  const pcm = await decodeAeaPcm(aea)
  const total = pcm.reduce((s, p) => s + p.length, 0)
  const flat = new Float32Array(total)
  let at = 0
  for (const p of pcm) {
    flat.set(p, at)
    at += p.length
  }
  writeFileSync(join(outDir, c.name + '.pcm.f32'), Buffer.from(flat.buffer, flat.byteOffset, flat.byteLength))
  console.log(`${c.name}: ${nCh} ch, ${n} samples, ${(aea.length - 2048) / 212} sound units`)
}
console.log(`wrote ${outDir} (node ${process.versions.node}, v8 ${process.versions.v8})`)
