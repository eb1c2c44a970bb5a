#!/usr/bin/env python
"""ref_run_qjs.py -- runs the REAL carta1 (aynik/carta1, JavaScript, /root/reference) in this image.

The image has no Node, but NVIDIA Nsight Compute ships Qt 6.6.3, and libQt6Qml.so.6 contains Qt's own
ECMAScript engine (QJSEngine, "V4": ES2016 + `?.` / `??`, typed arrays, classes, generators, ES modules,
baseline JIT).  This script drives that engine through ctypes (no Qt headers exist here: the six C++ entry
points it needs are called by their mangled names, QString values are laid out by hand) and lets it import
the reference's own modules FROM WHERE THEY LIE.  Nothing of the reference is copied into this repository;
a scratch mirror is made under a temporary directory at run time because two of the reference's files use
syntax newer than the engine:

    codec/core/options.js   three object spreads `{ ...x }`            -> Object.assign({}, x)
    codec/io/processor.js   async / await / `for await` / object rest  -> the synchronous equivalents
                            (every awaited value in that file is computed synchronously; Blob is a host shim)

Every rewritten line is printed and recorded in tests/golden/ref/provenance.json.  All files that hold codec
arithmetic -- constants, buffers, qmf, fft, mdct, transient, bitallocation, quantization, encoder, decoder,
bitstream, serialization, utils -- are loaded byte for byte as the reference ships them (their sha256 is
recorded too).

What it writes (tests/golden/ref/, the layout tools/ref_dump.mjs writes under Node; tests/test_reference_pin.py
reads either):
    tables.json        every libm-derived table as THIS engine computed it (V4 calls the host's glibc)
    <case>.aea         encodeAeaPcm(channels, options)        codec/io/processor.js:597-617
    <case>.pcm.f32     decodeAeaPcm(that)                     codec/io/processor.js:628-654
    stages.npz         per-stage outputs of the reference's own stage functions on the same inputs
                       (qmfAnalysisStage bands, performFFT magnitudes, transient scores, block modes,
                       mdctStage coefficients, the frame object; dequantised coefficients, IMDCT bands, PCM)
    wav.npz, wav.json  the CLI's decode-to-WAV route (createWavBlob: clip, scale by 32767 / 32768, ToInt16): CRC-32 per frame
    api.npz, api.json  whole AEA files (header included) for caller-shaped options (title, per-band thresholds, unknown keys),
                       and what `new EncoderOptions(x)` holds or throws for a list of trials
    kat.json           known answers of single functions (FFT.fft, MDCT/IMDCT transform, qmfAnalysis /
                       qmfSynthesis, findScaleFactor, quantize / dequantize, allocateBits, packBits)

usage:  python tests/golden/make_golden.py --export-ref-inputs
        python tools/ref_run_qjs.py [/root/reference] [outdir]            # writes the dump
        python tools/ref_run_qjs.py --verify [/root/reference] [outdir]   # re-runs the file-level cases, compares, writes nothing
        python tools/ref_run_qjs.py --check-js-layer                      # runs carta1_b200/napi/index.mjs next to the reference
"""
import ctypes
import glob
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

QT_CANDIDATES = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_*-x64"), reverse=True)

GLIB_SYMBOLS = """g_main_context_default g_main_context_iteration g_main_context_new g_main_context_pop_thread_default
g_main_context_push_thread_default g_main_context_ref g_main_context_unref g_main_context_wakeup g_source_add_poll
g_source_attach g_source_destroy g_source_new g_source_remove_poll g_source_set_can_recurse g_source_set_name
g_source_unref""".split()


def find_qt():
    for d in QT_CANDIDATES:
        if os.path.exists(os.path.join(d, "libQt6Qml.so.6")) and os.path.exists(os.path.join(d, "libQt6Core.so.6")):
            return d
    return None


class QString(ctypes.Structure):
    # Qt 6: QArrayDataPointer<char16_t> { Data *d; char16_t *ptr; qsizetype size; }; d == nullptr is a
    # legal non-owning string (QString::fromRawData builds exactly that)
    _fields_ = [("d", ctypes.c_void_p), ("ptr", ctypes.c_void_p), ("size", ctypes.c_longlong)]


class Engine:
    """QCoreApplication + QJSEngine of the Qt that ships with Nsight Compute, by mangled name."""

    def __init__(self, scratch):
        qt = find_qt()
        if qt is None:
            raise RuntimeError("no Qt 6 (libQt6Qml.so.6) found under /opt/nvidia/nsight-compute")
        self.qt_dir = qt
        # libQt6Core wants libglib-2.0.so.0 / libgthread-2.0.so.0 (its optional event dispatcher); the image
        # has neither.  QT_NO_GLIB=1 keeps Qt off them; empty stand-ins satisfy the loader.
        os.environ["QT_NO_GLIB"] = "1"
        stub_c = os.path.join(scratch, "glibstub.c")
        with open(stub_c, "w") as f:
            f.write("#include <stdlib.h>\n" + "".join("void %s(void){abort();}\n" % s for s in GLIB_SYMBOLS))
        for name in ("libglib-2.0.so.0", "libgthread-2.0.so.0"):
            out = os.path.join(scratch, name)
            subprocess.check_call(["gcc", "-shared", "-fPIC", "-o", out, stub_c, "-Wl,-soname," + name])
            ctypes.CDLL(out, mode=ctypes.RTLD_GLOBAL)
        self.core = ctypes.CDLL(os.path.join(qt, "libQt6Core.so.6"), mode=ctypes.RTLD_GLOBAL)
        self.qml = ctypes.CDLL(os.path.join(qt, "libQt6Qml.so.6"), mode=ctypes.RTLD_GLOBAL)
        self._keep = []
        self._argc = ctypes.c_int(1)
        self._argv = (ctypes.c_char_p * 2)(b"ref_run_qjs", None)
        self._app = ctypes.create_string_buffer(64)
        # QCoreApplication::QCoreApplication(int &argc, char **argv, int flags = QT_VERSION)
        self.core._ZN16QCoreApplicationC1ERiPPci(self._app, ctypes.byref(self._argc), self._argv, ctypes.c_int(0x060603))
        self._eng = ctypes.create_string_buffer(64)
        self.qml._ZN9QJSEngineC1Ev(self._eng)  # QJSEngine::QJSEngine()
        self.qml._ZNK8QJSValue7isErrorEv.restype = ctypes.c_bool
        self.version = ctypes.cast(self._call_c("qVersion"), ctypes.c_char_p).value.decode()

    def _call_c(self, name):
        fn = getattr(self.core, name)
        fn.restype = ctypes.c_void_p
        return fn()

    def _qs(self, s):
        b = s.encode("utf-16-le")
        buf = ctypes.create_string_buffer(b + b"\0\0")
        self._keep.append(buf)  # never freed: a QString with d == nullptr is taken for immortal raw data and V4 keeps pointers into it
        return QString(None, ctypes.addressof(buf), len(b) // 2)

    def to_string(self, v):
        r = QString()
        self.qml._ZNK8QJSValue8toStringEv(ctypes.byref(r), v)  # QString QJSValue::toString() const (sret)
        return ctypes.string_at(r.ptr, r.size * 2).decode("utf-16-le") if r.size else ""

    def is_error(self, v):
        return bool(self.qml._ZNK8QJSValue7isErrorEv(v))

    def _check(self, v, what):
        if self.is_error(v):
            p = ctypes.create_string_buffer(16)
            info = []
            for k in ("fileName", "lineNumber", "stack"):
                self.qml._ZNK8QJSValue8propertyERK7QString(p, v, ctypes.byref(self._qs(k)))
                info.append("%s=%s" % (k, self.to_string(p)))
            raise RuntimeError("%s: %s (%s)" % (what, self.to_string(v), ", ".join(info)))
        return v

    def evaluate(self, src, name="<driver>"):
        v = ctypes.create_string_buffer(16)
        # QJSValue QJSEngine::evaluate(const QString &program, const QString &fileName, int line, QStringList *)
        self.qml._ZN9QJSEngine8evaluateERK7QStringS2_iP5QListIS0_E(v, self._eng, ctypes.byref(self._qs(src)),
                                                                 ctypes.byref(self._qs(name)), ctypes.c_int(1), None)
        return self.to_string(self._check(v, name))

    def import_module(self, path, as_global):
        v = ctypes.create_string_buffer(16)
        self.qml._ZN9QJSEngine12importModuleERK7QString(v, self._eng, ctypes.byref(self._qs(path)))
        self._check(v, "import " + path)
        g = ctypes.create_string_buffer(16)
        self.qml._ZNK9QJSEngine12globalObjectEv(g, self._eng)
        self.qml._ZN8QJSValue11setPropertyERK7QStringRKS_(g, ctypes.byref(self._qs(as_global)), v)


# ---------------------------------------------------------------------------------------------------
# The scratch mirror and the two downlevelled files
UNTOUCHED = ["core/constants.js", "core/buffers.js", "transforms/qmf.js", "transforms/fft.js", "transforms/mdct.js",
             "analysis/transient.js", "coding/bitallocation.js", "coding/quantization.js", "pipeline/encoder.js",
             "pipeline/decoder.js", "io/bitstream.js", "io/serialization.js", "utils.js", "index.js"]

OPTIONS_RULES = [(r"\{ \.\.\.(this\.\w+) \}", r"Object.assign({}, \1)")]
PROCESSOR_RULES = [
    (r"\bstatic async \*", "static *"),
    (r"\bstatic async ", "static "),
    (r"\bexport async function ", "export function "),
    (r"= async function\* \(", "= function* ("),
    (r"\bfor await \(", "for ("),
    (r"\bawait ", ""),
    (r"const \{ title = 'encoded by carta1', \.\.\.encoderValues \} = options",
     "const title = options.title === undefined ? 'encoded by carta1' : options.title; "
     "const encoderValues = Object.assign({}, options); delete encoderValues.title"),
]


def downlevel(text, rules, log, rel):
    out = []
    for no, line in enumerate(text.split("\n"), 1):
        new = line
        for pat, rep in rules:
            new = re.sub(pat, rep, new)
        if new != line:
            log.append({"file": rel, "line": no, "from": line.strip(), "to": new.strip()})
        out.append(new)
    return "\n".join(out)


def stage_reference(ref_root, scratch):
    """Mirror <ref_root>/codec (without browser/ and the fs reader) under scratch, downlevelling two files."""
    src = os.path.join(ref_root, "codec")
    dst = os.path.join(scratch, "carta1", "codec")
    log, hashes = [], {}
    for rel in UNTOUCHED + ["core/options.js", "io/processor.js"]:
        s = open(os.path.join(src, rel), encoding="utf-8").read()
        hashes["codec/" + rel] = hashlib.sha256(s.encode()).hexdigest()
        if rel == "core/options.js":
            s = downlevel(s, OPTIONS_RULES, log, "codec/" + rel)
        elif rel == "io/processor.js":
            s = downlevel(s, PROCESSOR_RULES, log, "codec/" + rel)
        os.makedirs(os.path.dirname(os.path.join(dst, rel)), exist_ok=True)
        with open(os.path.join(dst, rel), "w", encoding="utf-8") as f:
            f.write(s)
    return dst, log, hashes


# ---------------------------------------------------------------------------------------------------
def battery_cases():
    """The inputs of tests/test_gpu_parity.py (mono_signals x OPTION_SETS, the non-finite injections, random
    sound units): name -> (channels or None, option dict or None, raw units or None, channel count)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_parity as T

    out = {}
    for si, (name, pcm) in enumerate(T.mono_signals().items()):
        for oi, kw in enumerate(T.OPTION_SETS):
            out["%s|%d" % (name, oi)] = ([pcm], kw, None, 1)
    rng = np.random.default_rng(3)  # test_non_finite_input
    for ii, inject in enumerate(((np.inf,), (-np.inf,), (np.nan,), (3e38, -3e38), (np.inf, np.nan, -3e38, 3e38, -np.inf))):
        pcm = (0.3 * rng.standard_normal(512 * 8)).astype(np.float32)
        for i, v in enumerate(inject):
            pcm[700 + 611 * i] = v
        for oi, kw in enumerate([dict(), dict(fixed_modes=[0, 0, 0]), dict(fixed_modes=[2, 2, 3]), dict(bias=2.5)]):
            out["nonfinite%d|%d" % (ii, oi)] = ([pcm], kw, None, 1)
    rng = np.random.default_rng(5)
    out["random_units_mono"] = (None, None, rng.integers(0, 256, (48, 212), dtype=np.uint8), 1)
    out["random_units_stereo_odd"] = (None, None, rng.integers(0, 256, (33, 212), dtype=np.uint8), 2)
    st = [T.S.cfg1_stereo(0.12)[0], T.S.cfg1_stereo(0.12)[1][:3000]]  # ragged stereo: the shorter channel is padded
    out["stereo_ragged|0"] = (st, dict(), None, 2)
    return out


API_CASES = {  # options exactly as a caller of carta1's index.js passes them (title included): name -> (signal, options)
    "api_title_and_band_thresholds": ("transients", {"title": "hello B200", "transientThresholdLow": 0.3, "transientThresholdMid": 0.01,
                                                     "transientThresholdHigh": 4.0}),
    "api_bias_fixed_modes": ("chirp", {"allocationBias": 2.5, "fixedBlockModes": [0, 2, 0]}),
    "api_defaults": ("sine_noise", {}),
    "api_unknown_keys_ignored": ("white", {"notAnOption": 7, "title": ""}),
}
OPTION_TRIALS = [
    {}, {"transientThresholdLow": 0.01}, {"transientThresholdLow": 2}, {"transientThresholdLow": 2.5}, {"transientThresholdLow": 0.001},
    {"transientThresholdMid": 3.5}, {"transientThresholdHigh": 4.01}, {"allocationBias": -0.1}, {"allocationBias": 5}, {"allocationBias": 5.5},
    {"fixedBlockModes": [2, 2, 3]}, {"fixedBlockModes": None}, {"bogus": 1}, {"transientThresholdLow": "1.5"},
]


def api_cases(eng, out_dir):
    """Whole AEA files (header included) for caller-shaped options, and EncoderOptions validation trials."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_parity as T

    sig = T.mono_signals()
    files, meta = {}, {}
    for name, (signal, options) in API_CASES.items():
        pcm = np.ascontiguousarray(sig[signal][:512 * 9 + 100], np.float32)
        eng.evaluate("setInputF32(%s)" % json.dumps([hex_of(pcm)]))
        aea = np.frombuffer(bytes.fromhex(eng.evaluate("runEncode(%s)" % json.dumps(options))), np.uint8)
        files[name + "/aea"] = aea.copy()
        meta[name] = {"signal": signal, "samples": int(pcm.size), "options": options, "input_sha256": hashlib.sha256(pcm.tobytes()).hexdigest()}
    trials = json.loads(eng.evaluate("JSON.stringify(runOptionTrials(%s))" % json.dumps(OPTION_TRIALS)))
    errors = json.loads(eng.evaluate("JSON.stringify(runErrorTrials())"))
    np.savez_compressed(os.path.join(out_dir, "api.npz"), **files)
    with open(os.path.join(out_dir, "api.json"), "w") as f:
        json.dump({"cases": meta, "option_trials": [{"options": t, "result": r} for t, r in zip(OPTION_TRIALS, trials)],
                   "error_trials": errors}, f, indent=1)
    print("api: %d whole-file cases, %d option trials" % (len(meta), len(trials)))


def js_options(kw):
    o = {}
    if "threshold" in kw:
        o["transientThresholdLow"] = kw["threshold"]
    if "bias" in kw:
        o["allocationBias"] = kw["bias"]
    if kw.get("fixed_modes"):
        o["fixedBlockModes"] = kw["fixed_modes"]
    return o


def battery(eng, out_dir, wavs, wav_meta):
    res, meta = {}, {}
    for name, (chans, kw, units, n_ch) in battery_cases().items():
        if chans is not None:
            chans = [np.ascontiguousarray(c, np.float32) for c in chans]
            eng.evaluate("setInputF32(%s)" % json.dumps([hex_of(c) for c in chans]))
            aea = np.frombuffer(bytes.fromhex(eng.evaluate("runEncode(%s)" % json.dumps(js_options(kw)))), np.uint8)
            res[name + "/su"] = aea[2048:].reshape(-1, 212).copy()
            meta[name] = {"input_sha256": [hashlib.sha256(c.tobytes()).hexdigest() for c in chans], "options": kw, "channels": n_ch}
        else:
            eng.evaluate("aeaOf(%s, %d)" % (json.dumps(hex_of(units)), n_ch))
            meta[name] = {"input_sha256": [hashlib.sha256(units.tobytes()).hexdigest()], "channels": n_ch}
        pcm = np.frombuffer(bytes.fromhex(eng.evaluate("runDecode()")), "<f4").reshape(n_ch, -1)
        res[name + "/pcm_crc"] = frame_crcs(pcm)  # [channel][frame] CRC-32 of the frame's 2048 bytes
        meta[name]["pcm_sha256"] = hashlib.sha256(pcm.tobytes()).hexdigest()
        if name.startswith("random_units") or name.startswith("loud|") or name.startswith("nonfinite4|"):
            # decoded samples beyond +-1 (and whatever non-finite input left behind): the WAV writer's clipping and ToInt16
            wav = bytes.fromhex(eng.evaluate("runWav()"))
            wavs[name + "/crc"] = wav_frame_crcs(wav[44:], n_ch)
            wav_meta[name] = {"header": wav[:44].hex(), "data_sha256": hashlib.sha256(wav[44:]).hexdigest(), "channels": n_ch}
    np.savez_compressed(os.path.join(out_dir, "battery.npz"), **res)
    with open(os.path.join(out_dir, "battery.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("battery: %d runs of the parity suite's inputs through encodeAeaPcm / decodeAeaPcm" % len(meta))


def frame_crcs(pcm):
    """pcm [channels][frames * 512] f32 -> uint32 [channels][frames], CRC-32 of each frame's bytes."""
    import zlib

    pcm = np.ascontiguousarray(pcm, "<f4")
    return np.array([[zlib.crc32(row[f * 512:(f + 1) * 512].tobytes()) for f in range(row.shape[0] // 512)] for row in pcm], np.uint32)


def unit_crcs(su):
    import zlib

    su = np.ascontiguousarray(su, np.uint8).reshape(-1, 212)
    return np.array([zlib.crc32(u.tobytes()) for u in su], np.uint32)


LONG_CASES = {  # golden input (tests/golden/ref/inputs/<name>.s16) repeated to this many seconds
    "cfg1_sine_noise_auto": 10.0,      # BASELINE configs[0]: 10 s stereo, default bias, auto block modes
    "cfg3_transients_auto": 10.0,
    "cfg2_chirp_fixed_long": 5.0,
}


def long_input(in_dir, c, seconds):
    """The case's int16 input tiled to `seconds` (a seam every repetition: a few more transients, nothing else)."""
    s16 = np.fromfile(os.path.join(in_dir, c["name"] + ".s16"), "<i2").reshape(-1, c["channels"])
    reps = int(np.ceil(seconds * 44100 / s16.shape[0]))
    return np.ascontiguousarray(np.tile(s16, (reps, 1))[:int(round(seconds * 44100))])


def long_cases(eng, out_dir, in_dir, cases):
    """Seconds-long runs: only checksums are kept (sha256 of the whole output, CRC-32 per sound unit / PCM frame)."""
    res, meta = {}, {}
    for c in cases:
        if c["name"] not in LONG_CASES:
            continue
        s16 = long_input(in_dir, c, LONG_CASES[c["name"]])
        opts = {"transientThresholdLow": c["threshold"], "allocationBias": c["bias"]}
        if c["fixed_modes"]:
            opts["fixedBlockModes"] = c["fixed_modes"]
        eng.evaluate("setInput(%s, %d)" % (json.dumps(hex_of(s16)), c["channels"]))
        aea = np.frombuffer(bytes.fromhex(eng.evaluate("runEncode(%s)" % json.dumps(opts))), np.uint8)
        pcm = np.frombuffer(bytes.fromhex(eng.evaluate("runDecode()")), "<f4").reshape(c["channels"], -1)
        res[c["name"] + "/su_crc"] = unit_crcs(aea[2048:])
        res[c["name"] + "/pcm_crc"] = frame_crcs(pcm)
        meta[c["name"]] = {"seconds": LONG_CASES[c["name"]], "samples": int(s16.shape[0]), "channels": c["channels"],
                           "input_sha256": hashlib.sha256(s16.tobytes()).hexdigest(), "sound_units": int((len(aea) - 2048) // 212),
                           "aea_sha256": hashlib.sha256(aea.tobytes()).hexdigest(), "pcm_sha256": hashlib.sha256(pcm.tobytes()).hexdigest()}
        print("long %-28s %.0f s, %d sound units" % (c["name"], LONG_CASES[c["name"]], meta[c["name"]]["sound_units"]))
    np.savez_compressed(os.path.join(out_dir, "long.npz"), **res)
    with open(os.path.join(out_dir, "long.json"), "w") as f:
        json.dump(meta, f, indent=1)


def wav_frame_crcs(data_bytes, n_ch):
    """CRC-32 of every 512-sample frame of a WAV data section (interleaved int16)."""
    import zlib

    step = 512 * 2 * n_ch
    return np.array([zlib.crc32(data_bytes[i:i + step]) for i in range(0, len(data_bytes), step)], np.uint32)


def hex_of(a):
    return np.ascontiguousarray(a).tobytes().hex()


INDEX_MJS_RULES = PROCESSOR_RULES + [
    (r"^import \{ createRequire \} from 'node:module'$", "// (no node:module here)"),
    (r"^import \* as ref from 'carta1'$", "import * as ref from './carta1/codec/index.js'"),
    (r"^const native = createRequire\(import\.meta\.url\)\('\./build/Release/carta1_b200\.node'\)$",
     "const native = __carta1_native  // tests/js_layer/mock_native.js: the addon contract over the reference's functions"),
    (r"^async function\* ", "function* "),
    (r"^async function ", "function "),
]


def check_js_layer(eng, scratch, codec_dir):
    """--check-js-layer: run carta1_b200/napi/index.mjs (the drop-in's JavaScript layer) inside the engine, with the
    addon replaced by tests/js_layer/mock_native.js, and compare it call by call with the reference."""
    log = []
    src = open(os.path.join(ROOT, "carta1_b200", "napi", "index.mjs"), encoding="utf-8").read()
    with open(os.path.join(scratch, "index_b200.js"), "w", encoding="utf-8") as f:
        f.write(downlevel(src, INDEX_MJS_RULES, log, "carta1_b200/napi/index.mjs"))
    print("index.mjs: %d lines downlevelled (async / await / object rest / the two Node imports)" % len(log))
    js = os.path.join(ROOT, "tests", "js_layer")
    eng.evaluate(open(os.path.join(js, "mock_native.js")).read(), "mock_native.js")
    eng.import_module(os.path.join(scratch, "index_b200.js"), "B200")
    eng.evaluate(open(os.path.join(js, "js_layer_check.js")).read(), "js_layer_check.js")
    doc = json.loads(eng.evaluate("JSON.stringify(runJsLayerChecks())"))
    bad = 0
    for name, ok, detail in doc["results"]:
        print("%-4s %s%s" % ("ok" if ok else "FAIL", name, "" if ok else "  -- " + detail[:400]))
        bad += not ok
    print("js layer: %d checks, %d failed, %d addon calls" % (len(doc["results"]), bad, doc["calls"]))
    return 1 if bad else 0


def verify(eng, out_dir, in_dir, cases, hashes):
    """--verify: run the reference again on the file-level cases and compare with what is committed; writes nothing."""
    prov = json.load(open(os.path.join(out_dir, "provenance.json")))
    bad = [k for k, v in hashes.items() if prov["reference_sha256"].get(k) != v]
    if bad:
        print("reference sources differ from the ones the committed dump was made from: %s" % bad)
        return 1
    for c in cases:
        s16 = np.fromfile(os.path.join(in_dir, c["name"] + ".s16"), "<i2").reshape(-1, c["channels"])
        opts = {"transientThresholdLow": c["threshold"], "allocationBias": c["bias"]}
        if c["fixed_modes"]:
            opts["fixedBlockModes"] = c["fixed_modes"]
        eng.evaluate("setInput(%s, %d)" % (json.dumps(hex_of(s16)), c["channels"]))
        aea = bytes.fromhex(eng.evaluate("runEncode(%s)" % json.dumps(opts)))
        pcm = bytes.fromhex(eng.evaluate("runDecode()"))
        same = (aea == open(os.path.join(out_dir, c["name"] + ".aea"), "rb").read() and
                pcm == open(os.path.join(out_dir, c["name"] + ".pcm.f32"), "rb").read())
        print("%-32s %s" % (c["name"], "reproduced" if same else "DIFFERS from the committed dump"))
        if not same:
            return 1
    print("verified: carta1 %s under Qt %s QJSEngine reproduces tests/golden/ref" % (prov["carta1"], eng.version))
    return 0


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    only_verify = "--verify" in sys.argv[1:]
    only_js_layer = "--check-js-layer" in sys.argv[1:]
    ref_root = os.path.abspath(args[0] if len(args) > 0 else "/root/reference")
    out_dir = os.path.abspath(args[1] if len(args) > 1 else os.path.join(ROOT, "tests", "golden", "ref"))
    in_dir = os.path.join(out_dir, "inputs")
    if not os.path.exists(os.path.join(in_dir, "cases.json")):
        raise SystemExit("run `python tests/golden/make_golden.py --export-ref-inputs` first")
    scratch = tempfile.mkdtemp(prefix="carta1_qjs_")
    try:
        eng = Engine(scratch)
        codec, log, hashes = stage_reference(ref_root, scratch)
        for e in log:
            print("downlevel %s:%d\n    - %s\n    + %s" % (e["file"], e["line"], e["from"], e["to"]))
        eng.evaluate(open(os.path.join(HERE, "ref_run_qjs_shims.js")).read(), "ref_run_qjs_shims.js")
        mods = {"carta1": "index.js", "C": "core/constants.js", "M": "transforms/mdct.js", "ENC": "pipeline/encoder.js",
                "DEC": "pipeline/decoder.js", "QMF": "transforms/qmf.js", "TR": "analysis/transient.js",
                "BA": "coding/bitallocation.js", "QZ": "coding/quantization.js", "BS": "io/bitstream.js",
                "SER": "io/serialization.js", "BUF": "core/buffers.js", "OPT": "core/options.js", "FFTM": "transforms/fft.js"}
        for g, rel in mods.items():
            eng.import_module(os.path.join(codec, rel), g)
        eng.evaluate(open(os.path.join(HERE, "ref_run_qjs_driver.js")).read(), "ref_run_qjs_driver.js")

        if only_js_layer:
            rc = check_js_layer(eng, scratch, codec)
            shutil.rmtree(scratch, ignore_errors=True)
            sys.stdout.flush()
            os._exit(rc)
        cases = json.load(open(os.path.join(in_dir, "cases.json")))
        if only_verify:
            rc = verify(eng, out_dir, in_dir, cases, hashes)
            shutil.rmtree(scratch, ignore_errors=True)
            sys.stdout.flush()
            os._exit(rc)
        biases = sorted({c["bias"] for c in cases})
        tables = json.loads(eng.evaluate("JSON.stringify(dumpTables(%s))" % json.dumps(biases)))
        pkg = json.load(open(os.path.join(ref_root, "package.json")))
        tables["versions"] = {"engine": "Qt QJSEngine (V4)", "qt": eng.version, "libm": "host glibc (V4 calls std::sin/cos/pow/log/exp)",
                              "from": eng.qt_dir}
        tables["carta1"] = pkg["version"]
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "tables.json"), "w") as f:
            json.dump(tables, f, indent=1)

        stages, wavs, wav_meta = {}, {}, {}
        for c in cases:
            s16 = np.fromfile(os.path.join(in_dir, c["name"] + ".s16"), "<i2").reshape(-1, c["channels"])
            opts = {"transientThresholdLow": c["threshold"], "allocationBias": c["bias"]}
            if c["fixed_modes"]:
                opts["fixedBlockModes"] = c["fixed_modes"]
            # bin/cli.js:394-396: readInt16LE / 32768.0 into a Float32Array, done in the engine
            eng.evaluate("setInput(%s, %d)" % (json.dumps(hex_of(s16)), c["channels"]))
            aea_hex = eng.evaluate("runEncode(%s)" % json.dumps(opts))
            aea = np.frombuffer(bytes.fromhex(aea_hex), np.uint8)
            aea.tofile(os.path.join(out_dir, c["name"] + ".aea"))
            pcm = np.frombuffer(bytes.fromhex(eng.evaluate("runDecode()")), "<f4")
            pcm.tofile(os.path.join(out_dir, c["name"] + ".pcm.f32"))
            wav = bytes.fromhex(eng.evaluate("runWav()"))
            wavs[c["name"] + "/crc"] = wav_frame_crcs(wav[44:], c["channels"])
            wav_meta[c["name"]] = {"header": wav[:44].hex(), "data_sha256": hashlib.sha256(wav[44:]).hexdigest(), "channels": c["channels"]}
            doc = json.loads(eng.evaluate("JSON.stringify(runStages(%s))" % json.dumps(opts)))
            for k, v in doc.items():
                dt = {"f32": "<f4", "f64": "<f8", "u8": np.uint8, "i32": "<i4"}[v["type"]]
                stages[c["name"] + "/" + k] = np.frombuffer(bytes.fromhex(v["hex"]), dt).reshape(v["shape"])
            print("%-32s %d ch, %d samples, %d sound units" % (c["name"], c["channels"], s16.shape[0], (len(aea) - 2048) // 212))
        np.savez_compressed(os.path.join(out_dir, "stages.npz"), **stages)

        battery(eng, out_dir, wavs, wav_meta)
        np.savez_compressed(os.path.join(out_dir, "wav.npz"), **wavs)
        with open(os.path.join(out_dir, "wav.json"), "w") as f:
            json.dump(wav_meta, f, indent=1)
        long_cases(eng, out_dir, in_dir, cases)
        api_cases(eng, out_dir)

        kat = json.loads(eng.evaluate("JSON.stringify(runKats())"))
        with open(os.path.join(out_dir, "kat.json"), "w") as f:
            json.dump(kat, f)
        with open(os.path.join(out_dir, "provenance.json"), "w") as f:
            json.dump({"engine": tables["versions"], "carta1": pkg["version"], "reference_sha256": hashes,
                       "loaded_unmodified": ["codec/" + r for r in UNTOUCHED], "downlevelled_lines": log,
                       "host_shims": ["Blob", "TextEncoder", "TextDecoder"],
                       "generator": "tools/ref_run_qjs.py + ref_run_qjs_driver.js + ref_run_qjs_shims.js"}, f, indent=1)
        print("wrote %s (Qt %s QJSEngine)" % (out_dir, eng.version))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    sys.stdout.flush()
    os._exit(0)  # no QCoreApplication teardown: the objects live in ctypes buffers


if __name__ == "__main__":
    main()
