"""The reference-side pin: oracle (CPU) and CUDA path against what the REAL carta1 wrote under Node.

tools/ref_dump.mjs runs aynik/carta1's own encodeAeaPcm / decodeAeaPcm (codec/io/processor.js:597-654) on
the inputs of tests/golden/*.npz and dumps V8's libm-derived tables.  No JavaScript engine exists in the
build image, so the dump cannot be produced there and every test below is SKIPPED until a maintainer
with Node >= 20.16 has run (see README.md, "Pinning against the reference"):

    python tests/golden/make_golden.py --export-ref-inputs
    node tools/ref_dump.mjs /path/to/carta1
    python -m pytest tests/test_reference_pin.py            # add -m gpu on a B200 box

With the dump present the tests demand byte equality of the AEA image and bit equality of the decoded
PCM, with V8's tables injected (the C ABI takes them as input: include/carta1_b200.h carta1_tables) --
and report separately whether the library's built-in default tables (glibc sin/cos/pow) equal V8's.
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
pytestmark = pytest.mark.skipif(not HAVE, reason="tests/golden/ref/ absent: run tools/ref_dump.mjs under Node (unrunnable in the build image)")

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
