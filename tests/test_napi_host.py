"""The N-API shim, linked and executed.

Node is not in the image, but Node-API is a C ABI: tests/napi_host/napi_host.c implements the subset of it the shim
uses (values, typed arrays over caller memory, externals with finalizers, exceptions, promises, async work with the
execute callback on another thread), carta1_b200/napi/carta1_napi.c is compiled UNCHANGED against it, and this file
plays the JavaScript side: it calls the addon's exported functions with the arguments carta1_b200/napi/index.mjs passes
and checks what comes back -- on a GPU against the bytes the reference itself produced (tests/golden/ref).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_DIR = os.path.join(HERE, "napi_host")
HOST_SO = os.path.join(HOST_DIR, "libcarta1_napi_host.so")

# napi_typedarray_type
I8, U8, U8C, I16, U16, I32, U32, F32, F64 = range(9)
NP_TYPE = {np.dtype(np.int8): I8, np.dtype(np.uint8): U8, np.dtype(np.int32): I32, np.dtype(np.float32): F32, np.dtype(np.float64): F64}
TYPE_NP = {v: k for k, v in NP_TYPE.items()}
# value kinds of napi_host.c
V_UNDEFINED, V_NUMBER, V_STRING, V_OBJECT, V_ARRAY, V_ARRAYBUFFER, V_TYPEDARRAY, V_EXTERNAL, V_FUNCTION, V_ERROR, V_PROMISE = range(11)

EXPORTS = {"createContext", "createEncoder", "createDecoder", "encodeFrames", "decodeFrames", "decodeFramesExpanded", "encodePcm",
           "decodeSu", "deserializeUnits", "destroy"}


class JsError(Exception):
    def __init__(self, kind, msg):
        super().__init__(msg)
        self.type_error = kind == 2


class Host:
    """The JavaScript side of the addon boundary."""

    def __init__(self):
        import carta1_b200

        carta1_b200.load()  # the product library first: the shim links against it
        subprocess.check_call(["make", "-C", HOST_DIR, "-s"])
        L = self.L = C.CDLL(HOST_SO)
        vp = C.c_void_p
        for name, res, args in [
            ("host_load", vp, []), ("host_module_name", C.c_char_p, []), ("host_export_names", C.c_int, [C.c_char_p, C.c_size_t]),
            ("host_undefined", vp, []), ("host_number", vp, [C.c_double]), ("host_string", vp, [C.c_char_p]), ("host_object", vp, []),
            ("host_set", None, [vp, C.c_char_p, vp]), ("host_array", vp, [C.c_size_t]), ("host_array_set", None, [vp, C.c_uint32, vp]),
            ("host_typedarray", vp, [C.c_int, vp, C.c_size_t]), ("host_call", vp, [C.c_char_p, C.c_size_t, C.POINTER(vp)]),
            ("host_take_exception", C.c_int, [C.c_char_p, C.c_size_t]), ("host_kind", C.c_int, [vp]), ("host_number_value", C.c_double, [vp]),
            ("host_typedarray_info", C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(vp)]),
            ("host_array_length", C.c_size_t, [vp]), ("host_array_get", vp, [vp, C.c_uint32]), ("host_promise_state", C.c_int, [vp]),
            ("host_promise_value", vp, [vp]), ("host_error_info", C.c_int, [vp, C.c_char_p, C.c_size_t]),
            ("host_release_external", None, [vp]),
        ]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        assert L.host_load(), "the addon did not register itself"
        self._keep = []

    # ---- Python -> "JavaScript" values
    def js(self, x):
        L = self.L
        if x is None:
            return L.host_undefined()
        if isinstance(x, JsHandle):
            return x.v
        if isinstance(x, bool):
            raise TypeError("no booleans cross this boundary")
        if isinstance(x, (int, float)):
            return L.host_number(float(x))
        if isinstance(x, str):
            return L.host_string(x.encode())
        if isinstance(x, np.ndarray):
            assert x.flags["C_CONTIGUOUS"] and x.dtype in NP_TYPE, x.dtype
            self._keep.append(x)
            return L.host_typedarray(NP_TYPE[x.dtype], x.ctypes.data, x.size)
        if isinstance(x, (list, tuple)):
            arr = L.host_array(len(x))
            for i, e in enumerate(x):
                L.host_array_set(arr, i, self.js(e))
            return arr
        if isinstance(x, dict):
            obj = L.host_object()
            for k, v in x.items():
                L.host_set(obj, k.encode(), L.host_undefined() if v is None else self.js(v))
            return obj
        raise TypeError(type(x))

    # ---- "JavaScript" values -> Python
    def py(self, v):
        L = self.L
        kind = L.host_kind(v)
        if kind == V_UNDEFINED:
            return None
        if kind == V_NUMBER:
            return L.host_number_value(v)
        if kind == V_TYPEDARRAY:
            t, n, p = C.c_int(), C.c_size_t(), C.c_void_p()
            assert L.host_typedarray_info(v, C.byref(t), C.byref(n), C.byref(p))
            dt = TYPE_NP[t.value]
            if n.value == 0:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n.value * dt.itemsize,)).view(dt).copy()
        if kind == V_ARRAY:
            return [self.py(L.host_array_get(v, i)) for i in range(L.host_array_length(v))]
        if kind == V_EXTERNAL:
            return JsHandle(v)
        if kind == V_PROMISE:  # `await`: the host settles promises before the call returns
            state = L.host_promise_state(v)
            assert state in (1, 2), "promise still pending"
            inner = L.host_promise_value(v)
            if state == 2:
                buf = C.create_string_buffer(512)
                raise JsError(L.host_error_info(inner, buf, 512), buf.value.decode())
            return self.py(inner)
        raise TypeError("unexpected value kind %d" % kind)

    def call(self, name, *args):
        argv = (C.c_void_p * max(len(args), 1))(*[self.js(a) for a in args])
        r = self.L.host_call(name.encode(), len(args), argv)
        if not r:
            buf = C.create_string_buffer(512)
            kind = self.L.host_take_exception(buf, 512)
            assert kind, "NULL result without a pending exception"
            raise JsError(kind, buf.value.decode())
        return self.py(r)

    def collect(self, handle):
        self.L.host_release_external(handle.v)


class JsHandle:
    def __init__(self, v):
        self.v = v


@pytest.fixture(scope="module")
def host():
    return Host()


def test_addon_registers_and_exports(host):
    assert host.L.host_module_name() == b"carta1_b200"
    buf = C.create_string_buffer(1024)
    n = host.L.host_export_names(buf, 1024)
    assert n == len(EXPORTS) and set(buf.value.decode().split(",")) == EXPORTS
    # every name index.mjs calls on the addon is exported
    src = open(os.path.join(HERE, "..", "carta1_b200", "napi", "index.mjs")).read()
    import re

    assert set(re.findall(r"native\.(\w+)\(", src)) <= EXPORTS


def test_argument_validation_without_a_device(host):
    for name, args in (("encodeFrames", (1.0, np.zeros(512, np.float32), 1)), ("createEncoder", ({}, {})),
                       ("decodeSu", (None, np.zeros(212, np.uint8), 1)), ("encodePcm", ("ctx", [np.zeros(512, np.float32)], {}))):
        with pytest.raises(JsError, match="bad or destroyed handle") as e:
            host.call(name, *args)
        assert e.value.type_error
    assert host.call("destroy", None) is None and host.call("destroy") is None
    with pytest.raises(JsError, match="tables must hold the nine Float64Arrays") as e:
        host.call("createContext", 0, {"windowShort": np.zeros(32, np.float64)})
    assert e.value.type_error


def test_create_context_without_a_device_says_why(host):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(JsError, match="needs a CUDA device"):
        host.call("createContext", 0, None)


# ------------------------------------------------------------------------------------------------- on a B200
def host_tables():
    """What index.mjs hostTables() hands to createContext: the host engine's own libm tables (here the ones of the
    engine that produced tests/golden/ref)."""
    from oracle import refpin as R

    doc = R.tables_doc()
    names = {"windowShort": "window_short", "scaleFactors": "scale_factors", "mdctFwd64": "mdct_fwd64", "mdctFwd256": "mdct_fwd256",
             "mdctFwd512": "mdct_fwd512", "mdctInv64": "mdct_inv64", "mdctInv256": "mdct_inv256", "mdctInv512": "mdct_inv512"}
    t = {js: np.ascontiguousarray(R.unhex(doc[k])) for js, k in names.items()}
    t["fftW"] = np.ascontiguousarray(np.array([R.unhex(doc["fft_w"][k]) for k in range(8)]).reshape(16))
    return t


def abi_options(c):
    """index.mjs abiOptions()."""
    from oracle import refpin as R

    return {"transientThresholdLow": c["threshold"], "allocationBias": c["bias"], "fixedBlockModes": c["fixed_modes"],
            "biasedScaleFactors": np.ascontiguousarray(R.biased(R.tables_doc(), c["bias"]))}


def ref_cases():
    from oracle import refpin as R

    return R.cases()


@pytest.fixture(scope="module")
def addon_ctx(host):
    from oracle import refpin as R

    if not R.available():
        pytest.skip("tests/golden/ref absent")
    ctx = host.call("createContext", 0, host_tables())
    yield ctx
    host.call("destroy", ctx)


@pytest.mark.gpu
@pytest.mark.parametrize("c", ref_cases(), ids=[c["name"] for c in ref_cases()])
def test_whole_buffer_calls_equal_reference(host, addon_ctx, c):
    """encodePcm / decodeSu (promises over napi_async_work) against the reference's own AEA bytes and PCM."""
    from oracle import refpin as R

    s16, aea, pcm_ref = R.load_case(c)
    n = s16.shape[0]
    padded = (n + 511) // 512 * 512
    chans = []
    for ch in range(c["channels"]):  # index.mjs pads the channels to equal length; bin/cli.js:395 scales int16
        x = np.zeros(n, np.float32)
        x[:] = (s16[:, ch].astype(np.float64) / 32768.0).astype(np.float32)
        chans.append(x)
    su = host.call("encodePcm", addon_ctx, chans, abi_options(c))
    assert su.dtype == np.uint8 and su.size == padded // 512 * c["channels"] * 212
    assert np.array_equal(su, aea[2048:]), "addon sound units differ from the reference's AEA bytes"
    pcm = host.call("decodeSu", addon_ctx, np.ascontiguousarray(aea[2048:]), c["channels"])
    assert len(pcm) == c["channels"]
    assert np.array_equal(np.stack(pcm).view(np.uint32), pcm_ref.view(np.uint32)), "addon PCM differs from the reference's"


@pytest.mark.gpu
def test_frame_closures_equal_reference(host, addon_ctx):
    """createEncoder / encodeFrames / createDecoder / decodeFrames / decodeFramesExpanded, one frame per call as the
    reference's encoder(pcm) / decoder(frame) closures are used, against the reference's stage dump."""
    from oracle import refpin as R

    z = np.load(os.path.join(R.REF, "stages.npz"))
    for c in R.cases():
        if c["channels"] != 1:
            continue
        s16, _, _ = R.load_case(c)
        want_su, want_pcm = z[c["name"] + "/su"][0], z[c["name"] + "/dec_pcm"][0]
        x = np.zeros(want_su.shape[0] * 512, np.float32)
        x[:s16.shape[0]] = (s16[:, 0].astype(np.float64) / 32768.0).astype(np.float32)
        enc = host.call("createEncoder", addon_ctx, abi_options(c), 1)
        dec = host.call("createDecoder", addon_ctx, 1)
        dec2 = host.call("createDecoder", addon_ctx, 1)
        for f in range(want_su.shape[0]):
            su = host.call("encodeFrames", enc, np.ascontiguousarray(x[512 * f:512 * f + 512]), 1)
            assert np.array_equal(su, want_su[f]), (c["name"], f, "encodeFrames")
            pcm = host.call("decodeFrames", dec, su, 1)
            assert np.array_equal(pcm.view(np.uint32), want_pcm[f].view(np.uint32)), (c["name"], f, "decodeFrames")
            # decode(frame object): index.mjs expandFrame() turns the frame object into per-position arrays
            q, sfi, bits, modes = expand_frame(z, c["name"], f)
            pcm2 = host.call("decodeFramesExpanded", dec2, q, sfi, bits, modes, 1)
            assert np.array_equal(pcm2.view(np.uint32), want_pcm[f].view(np.uint32)), (c["name"], f, "decodeFramesExpanded")
        for h in (enc, dec, dec2):
            host.call("destroy", h)
        with pytest.raises(JsError, match="bad or destroyed handle"):
            host.call("encodeFrames", enc, np.zeros(512, np.float32), 1)


WORD_LENGTH_BITS = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]


def expand_frame(z, name, f):
    """index.mjs expandFrame(): the frame object the reference's quantizationStage returned -> per-position arrays."""
    from oracle import oracle as O

    start_long = O.const_table("c1o_bfu_start_long", 52, np.int32)
    start_short = O.const_table("c1o_bfu_start_short", 52, np.int32)
    specs = O.const_table("c1o_specs_per_bfu", 52, np.int32)
    q, sfi, bits = np.zeros(512, np.int32), np.zeros(512, np.uint8), np.zeros(512, np.uint8)
    modes = np.ascontiguousarray(z[name + "/enc_modes"][0, f], np.int32)
    for bfu in range(int(z[name + "/n_bfu"][0, f])):
        width = WORD_LENGTH_BITS[int(z[name + "/wl"][0, f, bfu])]
        if width <= 0:
            continue
        band = 0 if bfu < 20 else 1 if bfu < 36 else 2
        pos = int(start_long[bfu] if modes[band] == 0 else start_short[bfu])
        n = int(specs[bfu])
        q[pos:pos + n] = z[name + "/q"][0, f, bfu, :n]
        sfi[pos:pos + n] = z[name + "/sfi"][0, f, bfu]
        bits[pos:pos + n] = width
    return q, sfi, bits, modes


@pytest.mark.gpu
def test_deserialize_units_and_errors(host, addon_ctx):
    from oracle import refpin as R

    c = [c for c in R.cases() if c["name"] == "cfg3_transients_auto"][0]
    z = np.load(os.path.join(R.REF, "stages.npz"))
    su = np.ascontiguousarray(z[c["name"] + "/su"].reshape(-1, 212))
    n_bfu, modes, wl, sfi, q = host.call("deserializeUnits", addon_ctx, su.reshape(-1))
    assert np.array_equal(n_bfu, z[c["name"] + "/n_bfu"].reshape(-1))
    assert np.array_equal(modes.reshape(-1, 3), z[c["name"] + "/enc_modes"].reshape(-1, 3))
    for u in range(su.shape[0]):
        k = int(n_bfu[u])
        assert np.array_equal(wl.reshape(-1, 52)[u, :k], z[c["name"] + "/wl"].reshape(-1, 52)[u, :k])
        assert np.array_equal(sfi.reshape(-1, 52)[u, :k], z[c["name"] + "/sfi"].reshape(-1, 52)[u, :k])
    with pytest.raises(JsError, match="Frame must be 212 bytes"):
        host.call("deserializeUnits", addon_ctx, np.zeros(211, np.uint8))
    with pytest.raises(JsError, match="requires one or two Float32 channels") as e:
        host.call("encodePcm", addon_ctx, [np.zeros(512, np.float32)] * 3, {})
    assert e.value.type_error
    with pytest.raises(JsError, match="requires one or two Float32 channels"):
        host.call("encodePcm", addon_ctx, [np.zeros(512, np.float64)], {})
    with pytest.raises(JsError, match="Unsupported channel count"):
        host.call("decodeSu", addon_ctx, np.zeros(212, np.uint8), 3)
    with pytest.raises(JsError, match="AEA bytes or a Blob") as e:
        host.call("decodeSu", addon_ctx, np.zeros(53, np.float32), 1)
    assert e.value.type_error


@pytest.mark.gpu
def test_shard_calls_and_finalizer(host):
    """haloFrames arguments of encodePcm / decodeSu (index.mjs encodePcmShard / decodeUnitsShard), and a context that is
    never destroyed explicitly: the external's finalizer releases it."""
    from oracle import refpin as R

    if not R.available():
        pytest.skip("tests/golden/ref absent")
    c = [c for c in R.cases() if c["name"] == "cfg1_sine_noise_auto"][0]
    s16, aea, pcm_ref = R.load_case(c)
    ctx = host.call("createContext", 0, host_tables())
    chans = [np.ascontiguousarray((s16[:, ch].astype(np.float64) / 32768.0).astype(np.float32)) for ch in range(2)]
    first, halo = 9, 2   # frames [9, end) with a 2-frame halo
    part = [np.ascontiguousarray(x[(first - halo) * 512:]) for x in chans]
    su = host.call("encodePcm", ctx, part, abi_options(c), halo)
    assert np.array_equal(su, aea[2048 + first * 2 * 212:]), "shard sound units"
    units = np.ascontiguousarray(aea[2048 + (first - 1) * 2 * 212:])
    pcm = host.call("decodeSu", ctx, units, 2, 1)
    assert np.array_equal(np.stack(pcm).view(np.uint32), pcm_ref[:, first * 512:].view(np.uint32)), "shard PCM"
    host.collect(ctx)
    with pytest.raises(JsError, match="bad or destroyed handle"):
        host.call("createDecoder", ctx, 1)


def test_js_layer_runs_next_to_the_reference():
    """carta1_b200/napi/index.mjs executed: inside Qt's QJSEngine (where the reference itself runs, tools/ref_run_qjs.py)
    with the addon replaced by tests/js_layer/mock_native.js, i.e. the addon's documented contract implemented over the
    reference's own functions.  Every public call of the drop-in (encodeAeaPcm, decodeAeaPcm, encode / decode closures,
    AudioProcessor.encodeStream / decodeStream across a batch boundary, deserializeFrames, the shard calls, error
    classes, hostTables) must return what the reference returns.  Needs /root/reference and Qt: skipped elsewhere."""
    import importlib.util
    import sys

    spec = importlib.util.spec_from_file_location("ref_run_qjs_probe", os.path.join(HERE, "..", "tools", "ref_run_qjs.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    if not os.path.isdir("/root/reference/codec") or tool.find_qt() is None:
        pytest.skip("no /root/reference or no Qt (libQt6Qml) in this environment")
    r = subprocess.run([sys.executable, os.path.join(HERE, "..", "tools", "ref_run_qjs.py"), "--check-js-layer"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "js layer: 15 checks, 0 failed" in r.stdout
