"""The reference-side pin: oracle (CPU) and CUDA path against what the REAL carta1 produced.

tests/golden/ref/ holds output of aynik/carta1's own JavaScript: tools/ref_run_qjs.py runs it in the build image under
Qt's QJSEngine (shipped inside Nsight Compute; there is no Node), tools/ref_dump.mjs does the file-level part under
Node >= 20.16 for whoever has one.  The dump carries the tables its engine computed; they are injected (the C ABI takes
them as input: include/carta1_b200.h carta1_tables), so the comparison is independent of the engine's libm, and
test_default_tables_against_v8 reports separately whether the library's built-in default tables equal the dump's.

    python tests/golden/make_golden.py --export-ref-inputs
    python tools/ref_run_qjs.py                 # or: node tools/ref_dump.mjs /path/to/carta1
    python -m pytest tests/test_reference_pin.py            # add -m gpu on a B200 box

Every comparison is bit for bit: AEA bytes and decoded PCM of encodeAeaPcm / decodeAeaPcm, every stage's output,
single-function known answers, the parity suite's inputs, seconds-long runs.
"""
import glob
import json
import os
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "ref")
HAVE = os.path.exists(os.path.join(REF, "tables.json")) and os.path.exists(os.path.join(REF, "inputs", "cases.json"))
pytestmark = pytest.mark.skipif(not HAVE, reason="tests/golden/ref/ absent: run tools/ref_run_qjs.py (or tools/ref_dump.mjs under Node)")

TABLE_FIELDS = ["window_short", "scale_factors", "mdct_fwd64", "mdct_fwd256", "mdct_fwd512", "mdct_inv64", "mdct_inv256",
                "mdct_inv512"]


def unhex(values):
    return np.array([int(v, 16) for v in values], np.uint64).view(np.float64)


def ref_tables():
    return json.load(open(os.path.join(REF, "tables.json")))


def fill_tables(t, doc):
    """t: a ctypes struct with the carta1_tables layout (oracle.Tables or carta1_b200 Tables)."""
    for name in TABLE_FIELDS:
        vals = unhex(doc[name])
        arr = getattr(t, name)
        assert len(vals) == len(arr), name
        for i, v in enumerate(vals):
            arr[i] = float(v)
    for k in range(8):
        pair = unhex(doc["fft_w"][k])
        t.fft_w[k][0], t.fft_w[k][1] = float(pair[0]), float(pair[1])
    return t


def cases():
    if not HAVE:
        return []
    return json.load(open(os.path.join(REF, "inputs", "cases.json")))


def load_case(c):
    s16 = np.fromfile(os.path.join(REF, "inputs", c["name"] + ".s16"), "<i2").reshape(-1, c["channels"])
    aea = np.fromfile(os.path.join(REF, c["name"] + ".aea"), np.uint8)
    pcm = np.fromfile(os.path.join(REF, c["name"] + ".pcm.f32"), "<f4").reshape(c["channels"], -1)
    return s16, aea, pcm


def biased(doc, bias):
    key = [k for k in doc["biased_scale_factors"] if float(k) == float(bias)]
    assert key, "tables.json holds no biased scale factors for bias %r" % bias
    return unhex(doc["biased_scale_factors"][key[0]])


def test_default_tables_against_v8(oracle):
    """Informational: do the library's built-in defaults (glibc) equal V8's tables?  Parity below does not
    depend on it (V8's tables are injected); a difference here means NULL tables are not the reference's."""
    doc = ref_tables()
    d = oracle.default_tables()
    diffs = {}
    for name in TABLE_FIELDS:
        a = np.array(list(getattr(d, name)), np.float64).view(np.uint64)
        b = unhex(doc[name]).view(np.uint64)
        n = int((a != b).sum())
        if n:
            diffs[name] = n
    a = np.array([[d.fft_w[k][0], d.fft_w[k][1]] for k in range(8)], np.float64).view(np.uint64)
    b = np.array([unhex(doc["fft_w"][k]) for k in range(8)]).view(np.uint64)
    if (a != b).any():
        diffs["fft_w"] = int((a != b).sum())
    if diffs:
        warnings.warn("default (glibc) tables differ from V8's in these entries: %r -- hosts must upload V8's tables "
                      "(the N-API shim does)" % diffs)


@pytest.mark.parametrize("c", cases(), ids=[c["name"] for c in cases()])
def test_oracle_equals_reference(oracle, c):
    O = oracle
    doc = ref_tables()
    t = fill_tables(O.Tables(), doc)
    s16, aea, pcm_ref = load_case(c)
    chans = [O.int16_to_pcm(s16[:, ch].copy()) for ch in range(c["channels"])]
    opts = O.make_options(threshold=c["threshold"], bias=c["bias"], fixed_modes=c["fixed_modes"], tables=t)
    for i, v in enumerate(biased(doc, c["bias"])):
        opts.biased_sf[i] = float(v)
    su = O.encode_pcm(chans, opts, tables=t)
    title, count, n_ch = O.aea_parse(aea[:2048])
    assert (title, count, n_ch) == ("encoded by carta1", su.shape[0], c["channels"])
    assert np.array_equal(su.reshape(-1), aea[2048:]), "oracle sound units differ from the reference's AEA bytes"
    pcm = np.stack(O.decode_su(aea[2048:].reshape(-1, 212), c["channels"], tables=t))
    assert pcm.shape == pcm_ref.shape
    assert np.array_equal(pcm.view(np.uint32), pcm_ref.view(np.uint32)), "oracle PCM differs from the reference's decodeAeaPcm"


@pytest.mark.gpu
@pytest.mark.parametrize("c", cases(), ids=[c["name"] for c in cases()])
def test_gpu_equals_reference(c):
    import carta1_b200
    from carta1_b200._lib import Tables

    doc = ref_tables()
    t = fill_tables(Tables(), doc)
    s16, aea, pcm_ref = load_case(c)
    ctx = carta1_b200.Context(0, t)
    try:
        opts = carta1_b200.make_enc_opts(c["threshold"], c["bias"], c["fixed_modes"],
                                         biased_scale_factors=biased(doc, c["bias"]))
        su = ctx.encode_pcm_s16(s16, c["channels"], opts)
        assert np.array_equal(su.reshape(-1), aea[2048:]), "GPU sound units differ from the reference's AEA bytes"
        pcm = np.stack(ctx.decode_su(aea[2048:].reshape(-1, 212), c["channels"]))
        assert np.array_equal(pcm.view(np.uint32), pcm_ref.view(np.uint32)), "GPU PCM differs from the reference's decodeAeaPcm"
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------
# Stage by stage: what the reference's own stage closures hand to each other (tools/ref_run_qjs_driver.js
# runStages), and single-function known answers (runKats).  Present when the dump came from tools/ref_run_qjs.py.
HAVE_STAGES = HAVE and os.path.exists(os.path.join(REF, "stages.npz")) and os.path.exists(os.path.join(REF, "kat.json"))
needs_stages = pytest.mark.skipif(not HAVE_STAGES, reason="tests/golden/ref/stages.npz absent (written by tools/ref_run_qjs.py)")


def frame_crcs(pcm):
    from oracle import refpin

    return refpin.frame_crcs(pcm)


def f32bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def unhex_arr(h, dtype):
    return np.frombuffer(bytes.fromhex(h), dtype).copy()


def ulp_distance(a, b):
    a = np.asarray(a, np.float64).view(np.int64)
    b = np.asarray(b, np.float64).view(np.int64)
    return np.abs(a - b)


@needs_stages
@pytest.mark.parametrize("c", cases(), ids=[c["name"] for c in cases()])
def test_oracle_stages_equal_reference(oracle, c):
    O = oracle
    doc = ref_tables()
    t = fill_tables(O.Tables(), doc)
    z = np.load(os.path.join(REF, "stages.npz"))
    ref = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(c["name"] + "/")}
    s16, _, _ = load_case(c)
    n_frames = ref["su"].shape[1]
    worst_score, exact_with_host_libm = 0, True
    for ch in range(c["channels"]):
        x = np.zeros(n_frames * 512, np.float32)
        x[:s16.shape[0]] = O.int16_to_pcm(s16[:, ch].copy())
        opts = O.make_options(threshold=c["threshold"], bias=c["bias"], fixed_modes=c["fixed_modes"], tables=t)
        for i, v in enumerate(biased(doc, c["bias"])):
            opts.biased_sf[i] = float(v)
        enc, dec = O.FrameEncoder(opts, t), O.FrameDecoder(t)
        for f in range(n_frames):
            fr, dbg = enc(x[512 * f:512 * f + 512], debug=True)
            where = (c["name"], ch, f)
            assert np.array_equal(f32bits(dbg.bands), f32bits(ref["enc_bands"][ch, f])), (where, "qmfAnalysisStage bands")
            if not c["fixed_modes"]:
                assert np.array_equal(f32bits(dbg.mags), f32bits(ref["enc_mags"][ch, f])), (where, "performFFT magnitudes")
                # the score passes through log / exp / log10 / log1p: the engine that wrote the dump used the host's
                # glibc, the oracle uses the fdlibm port V8 carries; both are faithfully rounded, not identical
                worst_score = max(worst_score, int(ulp_distance(list(dbg.score), ref["enc_scores"][ch, f]).max()))
                if f:  # with the engine's libm the score itself is bit-identical
                    O.set_host_libm(True)
                    try:
                        off = 0
                        for b, nb in enumerate((64, 64, 128)):
                            s = O.transient_score(np.array(dbg.mags[off:off + nb], np.float32), ref["enc_mags"][ch, f - 1][off:off + nb])
                            exact_with_host_libm &= bool(ulp_distance([s], [ref["enc_scores"][ch, f][b]])[0] == 0)
                            off += nb
                    finally:
                        O.set_host_libm(False)
            assert list(fr.modes) == list(ref["enc_modes"][ch, f]), (where, "block modes")
            assert np.array_equal(f32bits(dbg.coefs), f32bits(ref["enc_coefs"][ch, f])), (where, "mdctStage coefficients")
            n = fr.n_bfu
            assert n == int(ref["n_bfu"][ch, f]), (where, "nBfu")
            assert list(fr.sfi)[:n] == list(ref["sfi"][ch, f][:n]), (where, "scaleFactorIndices")
            assert list(fr.wl)[:n] == list(ref["wl"][ch, f][:n]), (where, "wordLengthIndices")
            q = np.array([list(row) for row in fr.q], np.int32)[:n]
            assert np.array_equal(q, ref["q"][ch, f][:n]), (where, "quantizedCoefficients")
            su = O.serialize_frame(fr)
            assert np.array_equal(su, ref["su"][ch, f]), (where, "serializeFrame")
            pcm, ddbg = dec(O.deserialize_frame(su), debug=True)
            assert np.array_equal(f32bits(ddbg.coefs), f32bits(ref["dec_coefs"][ch, f])), (where, "dequantizationStage")
            assert np.array_equal(f32bits(ddbg.bands), f32bits(ref["dec_bands"][ch, f])), (where, "imdctStage")
            assert np.array_equal(f32bits(pcm), f32bits(ref["dec_pcm"][ch, f])), (where, "qmfSynthesisStage")
    assert exact_with_host_libm, "transient score differs from the engine's even with the engine's libm"
    assert worst_score <= 64, "transient score (fdlibm) differs from the engine's (glibc) by %d ulp" % worst_score


@needs_stages
def test_oracle_function_kats(oracle):
    """FFT.fft, MDCT / IMDCT.transform, qmfAnalysis / qmfSynthesis, overlapAdd, findScaleFactor, quantize, dequantize,
    groupIntoBFUs + allocateBits, packBits, performFFT, detectTransient: each called alone inside the engine."""
    O = oracle
    doc = ref_tables()
    t = fill_tables(O.Tables(), doc)
    kat = json.load(open(os.path.join(REF, "kat.json")))
    f4 = "<f4"
    for k in kat["fft"]:
        re, im = O.fft(unhex_arr(k["re_in"], f4), unhex_arr(k["im_in"], f4), t)
        assert np.array_equal(f32bits(re), f32bits(unhex_arr(k["re"], f4))) and np.array_equal(f32bits(im), f32bits(unhex_arr(k["im"], f4))), ("fft", k["n"])
    for k in kat["mdct"]:
        assert np.array_equal(f32bits(O.mdct(unhex_arr(k["x"], f4), t)), f32bits(unhex_arr(k["y"], f4))), ("mdct", k["n"])
    for k in kat["imdct"]:
        assert np.array_equal(f32bits(O.imdct(unhex_arr(k["x"], f4), t)), f32bits(unhex_arr(k["y"], f4))), ("imdct", k["n"])
    for k in kat["qmf_analysis"]:
        lo, hi, d = O.qmf_analysis(unhex_arr(k["x"], f4), unhex_arr(k["delay"], f4))
        for got, want in ((lo, "low"), (hi, "high"), (d, "new_delay")):
            assert np.array_equal(f32bits(got), f32bits(unhex_arr(k[want], f4))), ("qmfAnalysis", want)
    for k in kat["qmf_synthesis"]:
        out, d = O.qmf_synthesis(unhex_arr(k["low"], f4), unhex_arr(k["high"], f4), unhex_arr(k["delay"], f4))
        assert np.array_equal(f32bits(out), f32bits(unhex_arr(k["out"], f4))) and np.array_equal(f32bits(d), f32bits(unhex_arr(k["new_delay"], f4))), "qmfSynthesis"
    for k in kat["overlap_add"]:
        got = O.overlap_add(unhex_arr(k["prev"], f4), unhex_arr(k["curr"], f4), unhex_arr(k["window"], "<f8"))
        assert np.array_equal(f32bits(got), f32bits(unhex_arr(k["out"], f4))), "overlapAdd"
    for k in kat["find_scale_factor"]:
        x = unhex_arr(k["x"], f4)
        assert O.find_scale_factor(x) == k["sfi"], ("findScaleFactor", x.tolist())
    for k in kat["quantize"]:
        assert np.array_equal(O.quantize(unhex_arr(k["x"], f4), k["sfi"], k["bits"], t), unhex_arr(k["q"], "<i4")), ("quantize", k["sfi"], k["bits"])
    for k in kat["dequantize"]:
        got = O.dequantize(unhex_arr(k["q"], "<i4"), k["sfi"], k["bits"], t)
        assert np.array_equal(f32bits(got), f32bits(unhex_arr(k["x"], f4))), ("dequantize", k["sfi"], k["bits"])
    for k in kat["allocate_bits"]:
        opts = O.make_options(bias=k["bias"], tables=t)
        n, sfi, wl = O.allocate_bits(unhex_arr(k["coefs"], f4), k["modes"], opts)
        assert n == k["n_bfu"], ("allocateBits bfuCount", k["modes"], k["bias"])
        assert np.array_equal(wl[:n], unhex_arr(k["wl"], "<i4")[:n]), ("allocateBits allocation", k["modes"], k["bias"])
        assert np.array_equal(sfi[:n], unhex_arr(k["sfi"], "<i4")[:n]), ("allocateBits scaleFactorIndices", k["modes"], k["bias"])
    for k in kat["perform_fft"]:
        x = unhex_arr(k["x"], f4)
        assert np.array_equal(f32bits(O.perform_fft(x, k["size"], t)), f32bits(unhex_arr(k["mag"], f4))), "performFFT"
    # detectTransient: the decision, and the score itself (recovered from the unmodified function by bisection over the
    # threshold).  With the libm of the engine that wrote the dump the score is bit-identical; with the fdlibm port
    # (V8's libm, the oracle's default and what the CUDA path carries) it is the same to rounding noise, which the
    # sqrt of a near-zero flatness change amplifies (the `cur = 1.01 * prev` cases)
    for k in kat["detect_transient"]:
        cur, prev = unhex_arr(k["cur"], f4), unhex_arr(k["prev"], f4)
        want = np.array([int(k["score"], 16)], np.uint64).view(np.float64)[0]
        O.set_host_libm(True)
        try:
            assert ulp_distance([O.transient_score(cur, prev)], [want])[0] == 0, "transient score with the engine's libm"
            assert int(O.detect_transient(cur, prev, k["threshold"])) == k["transient"], "detectTransient"
        finally:
            O.set_host_libm(False)
        got = O.transient_score(cur, prev)
        assert abs(got - want) <= 1e-10 * max(abs(want), 1e-300), "transient score with fdlibm: %r vs %r" % (got, want)
        if abs(got - k["threshold"]) > 1e-9:
            assert int(O.detect_transient(cur, prev, k["threshold"])) == k["transient"], "detectTransient"


def test_pack_bits_kats(oracle):
    if not HAVE_STAGES:
        pytest.skip("no kat.json")
    import ctypes as C

    kat = json.load(open(os.path.join(REF, "kat.json")))
    L = oracle.lib()
    for k in kat["pack_bits"]:
        buf = np.zeros(16, np.uint8)
        for pos, val, cnt in k["ops"]:
            L.c1o_pack_bits(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(16), C.c_int(pos), C.c_int(val), C.c_int(cnt))
        assert buf.tobytes().hex() == k["bytes"], k["ops"]


# ---------------------------------------------------------------------------------------------------
# The parity suite's own inputs (tests/test_gpu_parity.py: mono_signals x OPTION_SETS, non-finite PCM, random bytes
# as sound units, ragged stereo) through the reference's encodeAeaPcm / decodeAeaPcm.
HAVE_BATTERY = HAVE and os.path.exists(os.path.join(REF, "battery.npz"))
needs_battery = pytest.mark.skipif(not HAVE_BATTERY, reason="tests/golden/ref/battery.npz absent (written by tools/ref_run_qjs.py)")


def ref_tool():
    """tools/ref_run_qjs.py as a module: the input builders and checksum helpers the dump was made with."""
    import importlib.util
    import sys

    if "ref_run_qjs" not in sys.modules:
        spec = importlib.util.spec_from_file_location("ref_run_qjs", os.path.join(HERE, "..", "tools", "ref_run_qjs.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_run_qjs"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["ref_run_qjs"]


def battery_inputs():
    """Rebuilds the inputs tools/ref_run_qjs.py fed to the reference and checks them against the recorded hashes."""
    import hashlib

    mod = ref_tool()
    meta = json.load(open(os.path.join(REF, "battery.json")))
    out = mod.battery_cases()
    assert set(out) == set(meta)
    for name, (chans, kw, units, n_ch) in out.items():
        blobs = [np.ascontiguousarray(c, np.float32).tobytes() for c in chans] if chans is not None else [units.tobytes()]
        assert [hashlib.sha256(b).hexdigest() for b in blobs] == meta[name]["input_sha256"], "input of %s is not what the reference was fed" % name
    return out


@needs_battery
def test_oracle_battery_equals_reference(oracle):
    O = oracle
    doc = ref_tables()
    t = fill_tables(O.Tables(), doc)
    z = np.load(os.path.join(REF, "battery.npz"))
    n_checked = 0
    for name, (chans, kw, units, n_ch) in battery_inputs().items():
        if chans is not None:
            opts = O.make_options(threshold=kw.get("threshold", 1.0), bias=kw.get("bias", 1.0), fixed_modes=kw.get("fixed_modes"), tables=t)
            su = O.encode_pcm(chans, opts, tables=t)
            assert np.array_equal(su, z[name + "/su"]), (name, "sound units")
        else:
            su = units
        pcm = np.stack(O.decode_su(su, n_ch, tables=t))
        assert np.array_equal(frame_crcs(pcm), z[name + "/pcm_crc"]), (name, "pcm")
        n_checked += 1
    assert n_checked >= 100


@pytest.mark.gpu
@needs_battery
def test_gpu_battery_equals_reference():
    import carta1_b200
    from carta1_b200._lib import Tables

    doc = ref_tables()
    t = fill_tables(Tables(), doc)
    z = np.load(os.path.join(REF, "battery.npz"))
    ctx = carta1_b200.Context(0, t)
    try:
        for name, (chans, kw, units, n_ch) in battery_inputs().items():
            if chans is not None:
                opts = carta1_b200.make_enc_opts(kw.get("threshold", 1.0), kw.get("bias", 1.0), kw.get("fixed_modes"))
                su = ctx.encode_pcm(chans, opts)
                assert np.array_equal(su, z[name + "/su"]), (name, "sound units")
            else:
                su = units
            pcm = np.stack(ctx.decode_su(su, n_ch))
            assert np.array_equal(frame_crcs(pcm), z[name + "/pcm_crc"]), (name, "pcm")
    finally:
        ctx.close()


@pytest.mark.gpu
@needs_stages
@pytest.mark.parametrize("c", cases(), ids=[c["name"] for c in cases()])
def test_gpu_stages_equal_reference(c):
    """Every f32 intermediate of the CUDA path against what the reference's stage closures produced."""
    import carta1_b200
    from carta1_b200._lib import Tables

    doc = ref_tables()
    t = fill_tables(Tables(), doc)
    z = np.load(os.path.join(REF, "stages.npz"))
    ref = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(c["name"] + "/")}
    s16, _, _ = load_case(c)
    ctx = carta1_b200.Context(0, t)
    try:
        opts = carta1_b200.make_enc_opts(c["threshold"], c["bias"], c["fixed_modes"], biased_scale_factors=biased(doc, c["bias"]))
        for ch in range(c["channels"]):
            pcm = (s16[:, ch].astype(np.float64) / 32768.0).astype(np.float32)
            got = ctx.debug_encode_stages(pcm, opts)
            assert np.array_equal(f32bits(got["bands"]), f32bits(ref["enc_bands"][ch])), (ch, "bands")
            if not c["fixed_modes"]:
                assert np.array_equal(f32bits(got["mags"]), f32bits(ref["enc_mags"][ch])), (ch, "mags")
                scores = ctx.debug_transient_scores(pcm, opts)
                assert ulp_distance(scores, ref["enc_scores"][ch]).max() <= 64, (ch, "scores")
            assert np.array_equal(got["modes"], ref["enc_modes"][ch]), (ch, "modes")
            assert np.array_equal(f32bits(got["coefs"]), f32bits(ref["enc_coefs"][ch])), (ch, "coefs")
            assert np.array_equal(got["su"], ref["su"][ch]), (ch, "su")
            fr = ctx.deserialize_units(ref["su"][ch])
            n = ref["n_bfu"][ch]
            assert np.array_equal(fr["n_bfu"], n), (ch, "nBfu")
            dec = ctx.debug_decode_stages(ref["su"][ch])
            assert np.array_equal(f32bits(dec["coefs"]), f32bits(ref["dec_coefs"][ch])), (ch, "dequantised")
            assert np.array_equal(f32bits(dec["bands"]), f32bits(ref["dec_bands"][ch])), (ch, "imdct bands")
            assert np.array_equal(f32bits(dec["pcm"]), f32bits(ref["dec_pcm"][ch])), (ch, "pcm")
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------
# Seconds-long runs (BASELINE configs[0] is 10 s stereo with auto block modes): the golden inputs tiled, the
# reference's output kept as checksums (sha256 of the whole image, CRC-32 per sound unit and per PCM frame).
HAVE_LONG = HAVE and os.path.exists(os.path.join(REF, "long.npz"))
needs_long = pytest.mark.skipif(not HAVE_LONG, reason="tests/golden/ref/long.npz absent (written by tools/ref_run_qjs.py)")


def long_list():
    if not HAVE_LONG:
        return []
    meta = json.load(open(os.path.join(REF, "long.json")))
    return [c for c in cases() if c["name"] in meta]


def check_long(c, su, pcm):
    from oracle import refpin

    refpin.check_long(c, su, pcm)


@needs_long
@pytest.mark.parametrize("c", long_list(), ids=[c["name"] for c in long_list()])
def test_oracle_long_runs_equal_reference(oracle, c):
    import hashlib

    O, T = oracle, ref_tool()
    doc = ref_tables()
    t = fill_tables(O.Tables(), doc)
    meta = json.load(open(os.path.join(REF, "long.json")))[c["name"]]
    s16 = T.long_input(os.path.join(REF, "inputs"), c, meta["seconds"])
    assert hashlib.sha256(s16.tobytes()).hexdigest() == meta["input_sha256"]
    chans = [O.int16_to_pcm(s16[:, ch].copy()) for ch in range(c["channels"])]
    opts = O.make_options(threshold=c["threshold"], bias=c["bias"], fixed_modes=c["fixed_modes"], tables=t)
    su = O.encode_pcm(chans, opts, tables=t, threads=4, chunk_frames=64)
    pcm = np.stack(O.decode_su(su, c["channels"], tables=t, threads=4, chunk_frames=64))
    check_long(c, su, pcm)


@pytest.mark.gpu
@needs_long
@pytest.mark.parametrize("c", long_list(), ids=[c["name"] for c in long_list()])
def test_gpu_long_runs_equal_reference(c):
    import carta1_b200
    from carta1_b200._lib import Tables

    T = ref_tool()
    doc = ref_tables()
    t = fill_tables(Tables(), doc)
    meta = json.load(open(os.path.join(REF, "long.json")))[c["name"]]
    s16 = T.long_input(os.path.join(REF, "inputs"), c, meta["seconds"])
    ctx = carta1_b200.Context(0, t)
    try:
        opts = carta1_b200.make_enc_opts(c["threshold"], c["bias"], c["fixed_modes"], biased_scale_factors=biased(doc, c["bias"]))
        su = ctx.encode_pcm_s16(s16, c["channels"], opts)
        pcm = np.stack(ctx.decode_su(su, c["channels"]))
        check_long(c, su, pcm)
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------
# V8's libm.  The dump's engine (Qt's) calls glibc; V8 <= 11.3 (Node 20) carries fdlibm's sin / cos / pow, restated in
# oracle/fdlibm_trig_pow.c.  How much of the output depends on which of the two built the tables?
def test_fdlibm_restatement_selfcheck(oracle):
    import math

    L = oracle.lib()
    assert L.c1o_fd_selfcheck() == 0, "a fdlibm constant's decimal and bit pattern disagree"
    rng = np.random.default_rng(1)
    xs = rng.uniform(-7, 7, 20000)
    for fd, ref in ((L.c1o_fd_sin, math.sin), (L.c1o_fd_cos, math.cos)):
        assert max(ulp_distance([fd(float(x))], [ref(float(x))])[0] for x in xs) <= 1
    for x, y in zip(rng.uniform(1e-7, 2, 20000), rng.uniform(-6, 6, 20000)):
        assert ulp_distance([oracle.fd_pow(float(x), float(y))], [math.pow(float(x), float(y))])[0] <= 1
    for i in range(64):  # Math.pow(2.0, i / 3.0 - 21): fdlibm and glibc agree on every scale factor
        assert oracle.fd_pow(2.0, i / 3.0 - 21) == math.pow(2.0, i / 3.0 - 21)


@needs_long
def test_table_flavours(oracle):
    """glibc-built tables (the library's default, and the dump's engine) against fdlibm-built ones (V8 <= 11.3): a few
    dozen of the ~930 entries differ in the last bit or two, and no emitted byte or decoded sample of the pinned runs
    (4,480 sound units, 2.3 M samples) changes.  (30 minutes of cfg1 / cfg2 / cfg3 material: none either)"""
    from oracle import refpin as R

    O = oracle
    d, f = O.default_tables(), O.fdlibm_tables()
    differing = 0
    for name in R.TABLE_FIELDS:
        a, b = np.array(list(getattr(d, name))), np.array(list(getattr(f, name)))
        dist = ulp_distance(a, b)
        assert dist.max() <= 2, name
        differing += int((dist != 0).sum())
    assert 0 < differing < 64  # the two libms are not identical, and nearly so
    for c in R.cases():
        runs = [R.load_case(c)[0]]
        if c["name"] in R.long_meta():
            runs.append(R.long_input(c, R.long_meta()[c["name"]]["seconds"]))
        for s16 in runs:
            ch = [O.int16_to_pcm(s16[:, k].copy()) for k in range(c["channels"])]
            og = O.make_options(threshold=c["threshold"], bias=c["bias"], fixed_modes=c["fixed_modes"], tables=d)
            of = O.make_options(threshold=c["threshold"], bias=c["bias"], fixed_modes=c["fixed_modes"], tables=f)
            if c["bias"] != 1:
                for i in range(64):
                    of.biased_sf[i] = O.fd_pow(f.scale_factors[i], c["bias"])
            su_g = O.encode_pcm(ch, og, tables=d, threads=4, chunk_frames=64)
            su_f = O.encode_pcm(ch, of, tables=f, threads=4, chunk_frames=64)
            assert np.array_equal(su_g, su_f), (c["name"], "sound units depend on the libm that built the tables")
            p_g = np.stack(O.decode_su(su_g, c["channels"], tables=d, threads=4, chunk_frames=64))
            p_f = np.stack(O.decode_su(su_g, c["channels"], tables=f, threads=4, chunk_frames=64))
            assert np.array_equal(f32bits(p_g), f32bits(p_f)), (c["name"], "PCM depends on the libm that built the tables")


# ---------------------------------------------------------------------------------------------------
# WAV int16 emit: the reference's createWavBlob (processor.js:349-447: clip to [-1, 1], scale by 32768
# below zero and 32767 above, DataView.setInt16 = ToInt16) on decoded frames, incl. samples far beyond +-1.
HAVE_WAV = HAVE and os.path.exists(os.path.join(REF, "wav.npz"))
needs_wav = pytest.mark.skipif(not HAVE_WAV, reason="tests/golden/ref/wav.npz absent (written by tools/ref_run_qjs.py)")


def wav_inputs():
    """name -> (sound units [n][212], channel count) of every run whose WAV the reference wrote."""
    meta = json.load(open(os.path.join(REF, "wav.json")))
    bat = np.load(os.path.join(REF, "battery.npz"))
    units = None
    out = {}
    for name, m in meta.items():
        if os.path.exists(os.path.join(REF, name + ".aea")):
            su = np.fromfile(os.path.join(REF, name + ".aea"), np.uint8)[2048:].reshape(-1, 212)
        elif name + "/su" in bat.files:
            su = bat[name + "/su"]
        else:
            units = units or battery_inputs()
            su = units[name][2]
        out[name] = (np.ascontiguousarray(su), m["channels"])
    return out


def check_wav(name, s16):
    import hashlib

    meta = json.load(open(os.path.join(REF, "wav.json")))[name]
    z = np.load(os.path.join(REF, "wav.npz"))
    data = np.ascontiguousarray(s16, "<i2").tobytes()
    got = ref_tool().wav_frame_crcs(data, meta["channels"])
    bad = np.nonzero(got != z[name + "/crc"])[0]
    assert bad.size == 0, (name, "WAV frames %r differ from the reference's" % bad[:8].tolist())
    assert hashlib.sha256(data).hexdigest() == meta["data_sha256"], name


@needs_wav
def test_oracle_wav_int16_equals_reference(oracle):
    O = oracle
    t = fill_tables(O.Tables(), ref_tables())
    clipped = 0
    for name, (su, n_ch) in wav_inputs().items():
        pcm = O.decode_su(su, n_ch, tables=t)
        s16 = np.stack([O.pcm_to_int16(p) for p in pcm], axis=1)
        clipped += int((np.abs(np.stack(pcm)) > 1).sum())
        check_wav(name, s16)
    assert clipped > 1000  # the clipping branch is exercised (random bytes as sound units, the loud signal)


@pytest.mark.gpu
@needs_wav
def test_gpu_wav_int16_equals_reference():
    import carta1_b200
    from carta1_b200._lib import Tables

    t = fill_tables(Tables(), ref_tables())
    ctx = carta1_b200.Context(0, t)
    try:
        for name, (su, n_ch) in wav_inputs().items():
            check_wav(name, ctx.decode_su_s16(su, n_ch))
    finally:
        ctx.close()


# ---------------------------------------------------------------------------------------------------
# The public surface as a caller uses it: encodeAeaPcm(channels, {title, per-band thresholds, unknown keys ...}) -> a
# whole AEA file, header included; and what `new EncoderOptions(x)` holds or throws (codec/core/options.js:67-109).
HAVE_API = HAVE and os.path.exists(os.path.join(REF, "api.npz"))
needs_api = pytest.mark.skipif(not HAVE_API, reason="tests/golden/ref/api.npz absent (written by tools/ref_run_qjs.py)")


def api_inputs():
    import hashlib

    import test_gpu_parity as T

    doc = json.load(open(os.path.join(REF, "api.json")))
    z = np.load(os.path.join(REF, "api.npz"))
    sig = T.mono_signals()
    out = {}
    for name, m in doc["cases"].items():
        pcm = np.ascontiguousarray(sig[m["signal"]][:m["samples"]], np.float32)
        assert hashlib.sha256(pcm.tobytes()).hexdigest() == m["input_sha256"], name
        out[name] = (pcm, m["options"], z[name + "/aea"])
    return out


@needs_api
def test_option_trials_match_reference():
    from carta1_b200 import codec

    doc = json.load(open(os.path.join(REF, "api.json")))
    checked = 0
    for t in doc["option_trials"]:
        if any(isinstance(v, str) for v in t["options"].values()):
            continue  # JavaScript's string-to-number coercion in `<` is not mirrored
        want = t["result"]
        try:
            got = codec.EncoderOptions(t["options"]).toObject()["values"]
        except ValueError as ex:
            assert "error" in want and str(ex) == want["error"], (t["options"], str(ex), want)
        else:
            assert "values" in want and got == want["values"], (t["options"], got, want)
        checked += 1
    assert checked >= 12


@needs_api
def test_oracle_api_files_equal_reference(oracle):
    from carta1_b200 import codec

    O = oracle
    t = fill_tables(O.Tables(), ref_tables())
    for name, (pcm, options, aea) in api_inputs().items():
        title = options.get("title", "encoded by carta1")
        fixed = options.get("fixedBlockModes")
        opts = O.make_options(threshold=options.get("transientThresholdLow", 1.0), bias=options.get("allocationBias", 1.0),
                              fixed_modes=fixed, tables=t)  # Mid / High thresholds are never read (encoder.js:137-141)
        su = O.encode_pcm([pcm], opts, tables=t)
        assert np.array_equal(O.aea_header(title, su.shape[0], 1), aea[:2048]), (name, "AEA header (oracle)")
        assert np.array_equal(np.asarray(codec.AeaFile.createHeader(title, su.shape[0], 1)), aea[:2048]), (name, "AEA header (C ABI)")
        assert np.array_equal(su.reshape(-1), aea[2048:]), (name, "sound units")
        info = codec.AeaFile.parseHeader(aea[:2048])
        assert (info["title"], info["frameCount"], info["channelCount"]) == (title, su.shape[0], 1), name


@pytest.mark.gpu
@needs_api
def test_gpu_mirror_api_files_equal_reference():
    """carta1_b200.codec (the mirror of carta1's index.js) called the way a user of carta1 calls it."""
    import carta1_b200
    from carta1_b200 import codec
    from carta1_b200._lib import Tables

    t = fill_tables(Tables(), ref_tables())
    ctx = carta1_b200.Context(0, t)
    try:
        for name, (pcm, options, aea) in api_inputs().items():
            got = codec.encodeAeaPcm([pcm], dict(options), ctx=ctx)
            assert np.array_equal(np.asarray(got, np.uint8), aea), (name, "AEA file")
    finally:
        ctx.close()


@needs_api
def test_error_texts_and_classes_match_reference():
    """What carta1 throws for bad input (name and message, recorded from the reference itself) against the Python mirror:
    Error -> ValueError, TypeError -> TypeError.  All of these are raised before any device work."""
    from carta1_b200 import codec

    f32 = np.zeros(512, np.float32)
    calls = {
        "encodeAeaPcm: no channels": lambda: codec.encodeAeaPcm([]),
        "encodeAeaPcm: three channels": lambda: codec.encodeAeaPcm([f32, f32, f32]),
        "encodeAeaPcm: Float64Array channel": lambda: codec.encodeAeaPcm([np.zeros(512, np.float64)]),
        "encodeAeaPcm: not an array": lambda: codec.encodeAeaPcm(f32),
        "encodeAeaPcm: option out of range": lambda: codec.encodeAeaPcm([f32], {"allocationBias": 9}),
        "decodeAeaPcm: string": lambda: codec.decodeAeaPcm("abc"),
        "decodeAeaPcm: short buffer": lambda: codec.decodeAeaPcm(np.zeros(100, np.uint8)),
        "decodeAeaPcm: bad magic": lambda: codec.decodeAeaPcm(np.zeros(2048 + 212, np.uint8)),
        "deserializeFrame: 211 bytes": lambda: codec.deserializeFrame(np.zeros(211, np.uint8)),
        "deserializeFrame: 213 bytes": lambda: codec.deserializeFrame(np.zeros(213, np.uint8)),
        "parseHeader: 2047 bytes": lambda: codec.AeaFile.parseHeader(np.zeros(2047, np.uint8)),
        "parseHeader: bad magic": lambda: codec.AeaFile.parseHeader(np.zeros(2048, np.uint8)),
        "setValue: unknown option": lambda: codec.EncoderOptions().setValue("x", 1),
        "getValue: unknown option": lambda: codec.EncoderOptions().getValue("x"),
        "encodeStream: three channels": lambda: list(codec.AudioProcessor.encodeStream([f32], {"channelCount": 3})),
        "decodeStream: zero channels": lambda: list(codec.AudioProcessor.decodeStream([], {"channelCount": 0})),
    }
    trials = json.load(open(os.path.join(REF, "api.json")))["error_trials"]
    assert {t[0] for t in trials} == set(calls)
    for label, name, message in trials:
        want = {"Error": ValueError, "TypeError": TypeError}[name]
        with pytest.raises(want) as e:
            calls[label]()
        assert type(e.value) is want and str(e.value) == message, (label, type(e.value).__name__, str(e.value), message)


def test_committed_dump_reproduces_from_the_reference():
    """Where the reference checkout and Qt's engine exist (the build container; not the GPU box), run carta1's own
    JavaScript again on the file-level cases and require the committed tests/golden/ref bytes: the fixtures are the
    reference's output, not the oracle's."""
    import subprocess
    import sys

    tool = ref_tool()
    if not os.path.isdir("/root/reference/codec") or tool.find_qt() is None:
        pytest.skip("no /root/reference or no Qt (libQt6Qml) in this environment")
    if not HAVE:
        pytest.skip("tests/golden/ref absent")
    r = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "ref_run_qjs.py"), "--verify"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "verified: carta1" in r.stdout


@pytest.mark.parametrize("c", cases(), ids=[c["name"] for c in cases()])
def test_oracle_generated_goldens_equal_the_reference(c):
    """tests/golden/<case>.npz was generated from the oracle (make_golden.py) before the reference could be run; what it
    stores equals what the reference wrote for the same input."""
    z = np.load(os.path.join(HERE, "golden", c["name"] + ".npz"))
    s16, aea, pcm_ref = load_case(c)
    assert np.array_equal(z["pcm_s16"], s16)
    assert np.array_equal(z["su"].reshape(-1), aea[2048:])
    assert np.array_equal(f32bits(z["pcm_out"]), f32bits(pcm_ref))
