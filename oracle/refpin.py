"""Reader of tests/golden/ref: what the REAL carta1 produced (tools/ref_run_qjs.py runs the reference's own
JavaScript under Qt's QJSEngine in the build image; tools/ref_dump.mjs does the same under Node).

TEST INFRASTRUCTURE ONLY, like the rest of oracle/: imported by tests/, __graft_entry__.smoke() and bench.py's
parity block, never by carta1_b200/.  Nothing here computes codec arithmetic: it loads committed bytes and compares.
"""
from __future__ import annotations

import hashlib
import json
import os
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref")

TABLE_FIELDS = ["window_short", "scale_factors", "mdct_fwd64", "mdct_fwd256", "mdct_fwd512", "mdct_inv64", "mdct_inv256",
                "mdct_inv512"]


def available() -> bool:
    return os.path.exists(os.path.join(REF, "tables.json")) and os.path.exists(os.path.join(REF, "inputs", "cases.json"))


def _json(name):
    with open(os.path.join(REF, name)) as f:
        return json.load(f)


def unhex(values) -> np.ndarray:
    return np.array([int(v, 16) for v in values], np.uint64).view(np.float64)


def tables_doc() -> dict:
    return _json("tables.json")


def engine() -> str:
    v = tables_doc().get("versions", {})
    return "%s %s" % (v.get("engine", "node"), v.get("qt", v.get("node", "")))


def fill_tables(t, doc=None):
    """t: a ctypes struct with the carta1_tables layout (oracle.Tables or carta1_b200._lib.Tables)."""
    doc = doc or tables_doc()
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


def biased(doc, bias) -> np.ndarray:
    key = [k for k in doc["biased_scale_factors"] if float(k) == float(bias)]
    assert key, "tables.json holds no biased scale factors for bias %r" % bias
    return unhex(doc["biased_scale_factors"][key[0]])


def cases() -> list:
    return _json(os.path.join("inputs", "cases.json")) if available() else []


def load_case(c):
    """-> (int16 input [samples][channels], the reference's AEA image, the reference's decoded PCM [channels][n])."""
    s16 = np.fromfile(os.path.join(REF, "inputs", c["name"] + ".s16"), "<i2").reshape(-1, c["channels"])
    aea = np.fromfile(os.path.join(REF, c["name"] + ".aea"), np.uint8)
    pcm = np.fromfile(os.path.join(REF, c["name"] + ".pcm.f32"), "<f4").reshape(c["channels"], -1)
    return s16, aea, pcm


def frame_crcs(pcm) -> np.ndarray:
    """pcm [channels][frames * 512] f32 -> uint32 [channels][frames], CRC-32 of each frame's bytes."""
    pcm = np.ascontiguousarray(pcm, "<f4")
    return np.array([[zlib.crc32(row[f * 512:(f + 1) * 512].tobytes()) for f in range(row.shape[0] // 512)] for row in pcm], np.uint32)


def unit_crcs(su) -> np.ndarray:
    su = np.ascontiguousarray(su, np.uint8).reshape(-1, 212)
    return np.array([zlib.crc32(u.tobytes()) for u in su], np.uint32)


def long_meta() -> dict:
    return _json("long.json") if os.path.exists(os.path.join(REF, "long.json")) else {}


def long_input(c, seconds) -> np.ndarray:
    """The case's int16 input tiled to `seconds` (what tools/ref_run_qjs.py fed the reference)."""
    s16 = np.fromfile(os.path.join(REF, "inputs", c["name"] + ".s16"), "<i2").reshape(-1, c["channels"])
    reps = int(np.ceil(seconds * 44100 / s16.shape[0]))
    return np.ascontiguousarray(np.tile(s16, (reps, 1))[:int(round(seconds * 44100))])


def check_long(c, su, pcm) -> None:
    """su [n][212], pcm [channels][frames*512] of a long run against the reference's checksums; raises on a difference."""
    meta = long_meta()[c["name"]]
    z = np.load(os.path.join(REF, "long.npz"))
    su = np.ascontiguousarray(su, np.uint8).reshape(-1, 212)
    assert su.shape[0] == meta["sound_units"], "sound unit count"
    bad = np.nonzero(unit_crcs(su) != z[c["name"] + "/su_crc"])[0]
    assert bad.size == 0, "sound units %r differ from the reference's" % bad[:8].tolist()
    hdr = np.fromfile(os.path.join(REF, c["name"] + ".aea"), np.uint8)[:2048].copy()
    hdr[260:264] = np.frombuffer(np.uint32(su.shape[0]).tobytes(), np.uint8)  # the header's frame count field
    assert hashlib.sha256(hdr.tobytes() + su.tobytes()).hexdigest() == meta["aea_sha256"], "AEA image"
    bad = np.argwhere(frame_crcs(pcm) != z[c["name"] + "/pcm_crc"])
    assert bad.size == 0, "PCM frames %r differ from the reference's" % bad[:8].tolist()
    assert hashlib.sha256(np.ascontiguousarray(pcm, "<f4").tobytes()).hexdigest() == meta["pcm_sha256"], "PCM"


def check_context(make_ctx, make_opts, tables_cls, include_long=True) -> dict:
    """Runs every file-level case (and the seconds-long ones) through a CUDA context and compares with the
    reference's own output.  make_ctx(tables) -> carta1_b200.Context; make_opts(threshold, bias, fixed, biased) ->
    options.  Returns a summary; raises AssertionError on the first difference."""
    doc = tables_doc()
    t = fill_tables(tables_cls(), doc)
    ctx = make_ctx(t)
    units = samples = 0
    try:
        names = []
        for c in cases():
            s16, aea, pcm_ref = load_case(c)
            opts = make_opts(c["threshold"], c["bias"], c["fixed_modes"], biased(doc, c["bias"]))
            su = ctx.encode_pcm_s16(s16, c["channels"], opts)
            assert np.array_equal(su.reshape(-1), aea[2048:]), "%s: sound units differ from the reference's AEA bytes" % c["name"]
            pcm = np.stack(ctx.decode_su(aea[2048:].reshape(-1, 212), c["channels"]))
            assert np.array_equal(pcm.view(np.uint32), pcm_ref.view(np.uint32)), "%s: PCM differs from the reference's decodeAeaPcm" % c["name"]
            units += su.shape[0]
            samples += pcm.size
            names.append(c["name"])
            if include_long and c["name"] in long_meta():
                m = long_meta()[c["name"]]
                big = long_input(c, m["seconds"])
                su = ctx.encode_pcm_s16(big, c["channels"], opts)
                pcm = np.stack(ctx.decode_su(su, c["channels"]))
                check_long(c, su, pcm)
                units += su.shape[0]
                samples += pcm.size
                names.append("%s x %.0f s" % (c["name"], m["seconds"]))
    finally:
        ctx.close()
    return {"equal_to_reference_output": True, "engine": engine(), "carta1": doc.get("carta1"), "cases": names,
            "sound_units": int(units), "pcm_samples": int(samples)}
