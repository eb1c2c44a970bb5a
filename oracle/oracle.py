"""ctypes binding of the CPU oracle (oracle/carta1_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from carta1_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcarta1_oracle.so")

SU_BYTES = 212
FRAME = 512
AEA_HEADER = 2048


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "carta1_oracle.c")
    hdr = os.path.join(_HERE, "carta1_oracle.h")
    fd = os.path.join(_HERE, "fdlibm_trig_pow.c")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr, fd)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class Tables(C.Structure):
    _fields_ = [
        ("window_short", C.c_double * 32),
        ("scale_factors", C.c_double * 64),
        ("mdct_fwd64", C.c_double * 32),
        ("mdct_fwd256", C.c_double * 128),
        ("mdct_fwd512", C.c_double * 256),
        ("mdct_inv64", C.c_double * 32),
        ("mdct_inv256", C.c_double * 128),
        ("mdct_inv512", C.c_double * 256),
        ("fft_w", (C.c_double * 2) * 8),
    ]


class Options(C.Structure):
    _fields_ = [
        ("transient_threshold", C.c_double),
        ("allocation_bias", C.c_double),
        ("use_fixed_modes", C.c_int),
        ("fixed_modes", C.c_int * 3),
        ("biased_sf", C.c_double * 64),
    ]


class Frame(C.Structure):
    _fields_ = [
        ("n_bfu", C.c_int),
        ("modes", C.c_int * 3),
        ("sfi", C.c_int * 52),
        ("wl", C.c_int * 52),
        ("q", (C.c_int * 20) * 52),
    ]


class Encoder(C.Structure):
    _fields_ = [
        ("T", C.POINTER(Tables)),
        ("opt", Options),
        ("delay_low", C.c_float * 46),
        ("delay_mid", C.c_float * 46),
        ("delay_high", C.c_float * 39),
        ("overlap", (C.c_float * 32) * 3),
        ("prev_mag", (C.c_float * 128) * 3),
    ]


class Decoder(C.Structure):
    _fields_ = [
        ("T", C.POINTER(Tables)),
        ("delay_low", C.c_float * 46),
        ("delay_mid", C.c_float * 46),
        ("delay_high", C.c_float * 39),
        ("tail", (C.c_float * 16) * 3),
    ]


class EncDebug(C.Structure):
    _fields_ = [
        ("bands", C.c_float * 512),
        ("mags", C.c_float * 256),
        ("score", C.c_double * 3),
        ("coefs", C.c_float * 512),
    ]


class DecDebug(C.Structure):
    _fields_ = [("coefs", C.c_float * 512), ("bands", C.c_float * 512)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        u8 = C.POINTER(C.c_uint8)
        L.c1o_default_tables.argtypes = [C.POINTER(Tables)]
        L.c1o_fdlibm_tables.argtypes = [C.POINTER(Tables)]
        for name in ("c1o_fd_sin", "c1o_fd_cos"):
            getattr(L, name).argtypes = [C.c_double]
            getattr(L, name).restype = C.c_double
        L.c1o_fd_pow.argtypes = [C.c_double, C.c_double]
        L.c1o_fd_pow.restype = C.c_double
        L.c1o_qmf_even.restype = fp
        L.c1o_qmf_odd.restype = fp
        L.c1o_specs_per_bfu.restype = ip
        L.c1o_bfu_start_long.restype = ip
        L.c1o_bfu_start_short.restype = ip
        L.c1o_sf_thresholds.restype = fp
        L.c1o_options_init.argtypes = [C.POINTER(Options), C.POINTER(Tables), C.c_double, C.c_double, ip]
        L.c1o_qmf_analysis.argtypes = [fp, C.c_int, fp, fp, fp]
        L.c1o_qmf_synthesis.argtypes = [fp, fp, C.c_int, fp, fp]
        L.c1o_fft.argtypes = [fp, fp, C.c_int, C.POINTER(Tables)]
        L.c1o_mdct.argtypes = [C.POINTER(Tables), C.c_int, fp, fp]
        L.c1o_imdct.argtypes = [C.POINTER(Tables), C.c_int, fp, fp]
        L.c1o_overlap_add.argtypes = [fp, fp, C.c_int, dp, fp]
        L.c1o_perform_fft.argtypes = [fp, C.c_int, C.c_int, C.POINTER(Tables), fp]
        L.c1o_transient_score.argtypes = [fp, fp, C.c_int]
        L.c1o_transient_score.restype = C.c_double
        L.c1o_find_scale_factor.argtypes = [fp, C.c_int]
        L.c1o_find_scale_factor_table.argtypes = [C.c_float]
        L.c1o_allocate_bits.argtypes = [fp, ip, C.POINTER(Options), ip, ip, ip]
        L.c1o_quantize.argtypes = [fp, C.c_int, C.c_int, C.c_int, C.POINTER(Tables), ip]
        L.c1o_dequantize.argtypes = [ip, C.c_int, C.c_int, C.c_int, C.POINTER(Tables), fp]
        L.c1o_pack_bits.argtypes = [u8, C.c_size_t, C.c_int, C.c_int, C.c_int]
        L.c1o_unpack_bits.argtypes = [u8, C.c_size_t, C.c_int, C.c_int]
        L.c1o_unpack_signed_bits.argtypes = [u8, C.c_size_t, C.c_int, C.c_int]
        L.c1o_serialize_frame.argtypes = [C.POINTER(Frame), u8]
        L.c1o_deserialize_frame.argtypes = [u8, C.POINTER(Frame)]
        for name in ("c1o_log", "c1o_exp", "c1o_log10", "c1o_log1p"):
            getattr(L, name).argtypes = [C.c_double]
            getattr(L, name).restype = C.c_double
        L.c1o_to_int32.argtypes = [C.c_double]
        L.c1o_to_int32.restype = C.c_int32
        L.c1o_encoder_init.argtypes = [C.POINTER(Encoder), C.POINTER(Tables), C.POINTER(Options)]
        L.c1o_decoder_init.argtypes = [C.POINTER(Decoder), C.POINTER(Tables)]
        L.c1o_encode_frame.argtypes = [C.POINTER(Encoder), fp, C.POINTER(Frame), C.POINTER(EncDebug)]
        L.c1o_decode_frame.argtypes = [C.POINTER(Decoder), C.POINTER(Frame), fp, C.POINTER(DecDebug)]
        L.c1o_frame_count.argtypes = [C.c_size_t]
        L.c1o_frame_count.restype = C.c_size_t
        L.c1o_encode_pcm_range.argtypes = [
            C.POINTER(Tables), C.POINTER(Options), C.POINTER(fp), C.c_int, C.c_size_t,
            C.c_size_t, C.c_size_t, u8]
        L.c1o_decode_su_range.argtypes = [
            C.POINTER(Tables), u8, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.POINTER(fp)]
        L.c1o_aea_header.argtypes = [C.c_char_p, C.c_uint32, C.c_int, u8]
        L.c1o_aea_parse.argtypes = [u8, C.c_size_t, C.c_char_p, C.POINTER(C.c_uint32), ip]
        L.c1o_pcm_to_int16.argtypes = [fp, C.c_size_t, C.POINTER(C.c_int16)]
        L.c1o_int16_to_pcm.argtypes = [C.POINTER(C.c_int16), C.c_size_t, fp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


_default_tables = None


def default_tables() -> Tables:
    global _default_tables
    if _default_tables is None:
        t = Tables()
        lib().c1o_default_tables(C.byref(t))
        _default_tables = t
    return _default_tables


def fdlibm_tables() -> Tables:
    """default_tables() with fdlibm's sin / cos / pow (what V8 <= 11.3 computes) instead of the host libm's."""
    t = Tables()
    lib().c1o_fdlibm_tables(C.byref(t))
    return t


def fd_pow(x: float, y: float) -> float:
    return lib().c1o_fd_pow(C.c_double(x), C.c_double(y))


def make_options(threshold=1.0, bias=1.0, fixed_modes=None, tables=None) -> Options:
    t = tables or default_tables()
    o = Options()
    fm = None
    if fixed_modes is not None:
        fm = (C.c_int * 3)(*[int(x) for x in fixed_modes])
    lib().c1o_options_init(C.byref(o), C.byref(t), float(threshold), float(bias), fm)
    return o


def const_table(name: str, n: int, dtype) -> np.ndarray:
    ptr = getattr(lib(), name)()
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype).copy()


# --------------------------------------------------------------------------- unit functions
def qmf_analysis(x: np.ndarray, delay: np.ndarray):
    x = np.ascontiguousarray(x, np.float32)
    delay = np.ascontiguousarray(delay, np.float32).copy()
    lo = np.zeros(len(x) // 2, np.float32)
    hi = np.zeros(len(x) // 2, np.float32)
    lib().c1o_qmf_analysis(_fp(x), len(x), _fp(delay), _fp(lo), _fp(hi))
    return lo, hi, delay


def qmf_synthesis(lo: np.ndarray, hi: np.ndarray, delay: np.ndarray):
    lo = np.ascontiguousarray(lo, np.float32)
    hi = np.ascontiguousarray(hi, np.float32)
    delay = np.ascontiguousarray(delay, np.float32).copy()
    out = np.zeros(2 * len(lo), np.float32)
    lib().c1o_qmf_synthesis(_fp(lo), _fp(hi), len(lo), _fp(delay), _fp(out))
    return out, delay


def fft(re: np.ndarray, im: np.ndarray, tables=None):
    re = np.ascontiguousarray(re, np.float32).copy()
    im = np.ascontiguousarray(im, np.float32).copy()
    lib().c1o_fft(_fp(re), _fp(im), len(re), C.byref(tables or default_tables()))
    return re, im


def mdct(x: np.ndarray, tables=None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x) // 2, np.float32)
    lib().c1o_mdct(C.byref(tables or default_tables()), len(x), _fp(x), _fp(out))
    return out


def imdct(x: np.ndarray, tables=None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x) * 2, np.float32)
    lib().c1o_imdct(C.byref(tables or default_tables()), len(x) * 2, _fp(x), _fp(out))
    return out


def overlap_add(prev, curr, window) -> np.ndarray:
    prev = np.ascontiguousarray(prev, np.float32)
    curr = np.ascontiguousarray(curr, np.float32)
    window = np.ascontiguousarray(window, np.float64)
    out = np.zeros(2 * len(prev), np.float32)
    lib().c1o_overlap_add(_fp(prev), _fp(curr), len(prev), window.ctypes.data_as(C.POINTER(C.c_double)), _fp(out))
    return out


def perform_fft(samples: np.ndarray, fft_size: int, tables=None) -> np.ndarray:
    samples = np.ascontiguousarray(samples, np.float32)
    mag = np.zeros(fft_size // 2, np.float32)
    lib().c1o_perform_fft(_fp(samples), len(samples), fft_size, C.byref(tables or default_tables()), _fp(mag))
    return mag


def set_host_libm(on: bool) -> None:
    """Transient detector libm: True = the host's (what Qt's QJSEngine calls), False = fdlibm port (V8; default)."""
    lib().c1o_set_host_libm(int(bool(on)))


def transient_score(cur: np.ndarray, prev: np.ndarray) -> float:
    cur = np.ascontiguousarray(cur, np.float32)
    prev = np.ascontiguousarray(prev, np.float32)
    return lib().c1o_transient_score(_fp(cur), _fp(prev), len(cur))


def detect_transient(cur, prev, threshold) -> bool:
    if prev is None or len(prev) == 0:
        return False  # transient.js:46
    return transient_score(cur, prev) > threshold


def find_scale_factor(coefs: np.ndarray) -> int:
    coefs = np.ascontiguousarray(coefs, np.float32)
    return lib().c1o_find_scale_factor(_fp(coefs), len(coefs))


def allocate_bits(coefs: np.ndarray, modes, options: Options):
    coefs = np.ascontiguousarray(coefs, np.float32)
    m = np.asarray(modes, np.int32)
    n = C.c_int()
    sfi = np.zeros(52, np.int32)
    wl = np.zeros(52, np.int32)
    lib().c1o_allocate_bits(_fp(coefs), _ip(m), C.byref(options), C.byref(n), _ip(sfi), _ip(wl))
    return n.value, sfi, wl


def quantize(c: np.ndarray, sfi: int, bits: int, tables=None) -> np.ndarray:
    c = np.ascontiguousarray(c, np.float32)
    out = np.zeros(len(c), np.int32)
    lib().c1o_quantize(_fp(c), len(c), sfi, bits, C.byref(tables or default_tables()), _ip(out))
    return out


def dequantize(q: np.ndarray, sfi: int, bits: int, tables=None) -> np.ndarray:
    q = np.ascontiguousarray(q, np.int32)
    out = np.zeros(len(q), np.float32)
    lib().c1o_dequantize(_ip(q), len(q), sfi, bits, C.byref(tables or default_tables()), _fp(out))
    return out


def serialize_frame(fr: Frame) -> np.ndarray:
    out = np.zeros(SU_BYTES, np.uint8)
    lib().c1o_serialize_frame(C.byref(fr), _u8(out))
    return out


def deserialize_frame(buf: np.ndarray) -> Frame:
    buf = np.ascontiguousarray(buf, np.uint8)
    if len(buf) != SU_BYTES:
        raise ValueError("Frame must be 212 bytes")  # serialization.js:112-114
    fr = Frame()
    lib().c1o_deserialize_frame(_u8(buf), C.byref(fr))
    return fr


# --------------------------------------------------------------------------- closures
class FrameEncoder:
    """encode(options, bufferPool) closure (codec/pipeline/encoder.js:438-450)."""

    def __init__(self, options: Options | None = None, tables: Tables | None = None):
        self.tables = tables or default_tables()
        self.options = options or make_options(tables=self.tables)
        self.state = Encoder()
        lib().c1o_encoder_init(C.byref(self.state), C.byref(self.tables), C.byref(self.options))

    def __call__(self, pcm: np.ndarray, debug: bool = False):
        pcm = np.ascontiguousarray(pcm, np.float32)
        assert len(pcm) == FRAME
        fr = Frame()
        dbg = EncDebug() if debug else None
        lib().c1o_encode_frame(C.byref(self.state), _fp(pcm), C.byref(fr), C.byref(dbg) if debug else None)
        return (fr, dbg) if debug else fr


class FrameDecoder:
    """decode(bufferPool) closure (codec/pipeline/decoder.js:408-411)."""

    def __init__(self, tables: Tables | None = None):
        self.tables = tables or default_tables()
        self.state = Decoder()
        lib().c1o_decoder_init(C.byref(self.state), C.byref(self.tables))

    def __call__(self, fr: Frame, debug: bool = False):
        pcm = np.zeros(FRAME, np.float32)
        dbg = DecDebug() if debug else None
        lib().c1o_decode_frame(C.byref(self.state), C.byref(fr), _fp(pcm), C.byref(dbg) if debug else None)
        return (pcm, dbg) if debug else pcm


# --------------------------------------------------------------------------- whole buffers
def frame_count(n_samples: int) -> int:
    return (n_samples + FRAME - 1) // FRAME


def _pad_channels(channels):
    n = max(len(c) for c in channels)
    out = []
    for c in channels:
        c = np.ascontiguousarray(c, np.float32)
        if len(c) < n:
            c = np.concatenate([c, np.zeros(n - len(c), np.float32)])
        out.append(c)
    return out, n


def encode_pcm(channels, options: Options | None = None, tables: Tables | None = None,
               threads: int = 1, chunk_frames: int = 4096) -> np.ndarray:
    """Planar f32 channels -> interleaved sound units, uint8 [n_su, 212]."""
    t = tables or default_tables()
    o = options or make_options(tables=t)
    chans, n = _pad_channels(channels)
    n_ch = len(chans)
    nf = frame_count(n)
    su = np.zeros((nf * n_ch, SU_BYTES), np.uint8)
    ptrs = (C.POINTER(C.c_float) * n_ch)(*[_fp(c) for c in chans])
    L = lib()

    def run(rng):
        L.c1o_encode_pcm_range(C.byref(t), C.byref(o), ptrs, n_ch, n, rng[0], rng[1], _u8(su))

    if threads <= 1:
        run((0, nf))
    else:
        ranges = [(a, min(a + chunk_frames, nf)) for a in range(0, nf, chunk_frames)]
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, ranges))
    return su


def decode_su(su: np.ndarray, n_ch: int, tables: Tables | None = None, threads: int = 1,
              chunk_frames: int = 4096):
    """Interleaved sound units -> list of planar f32 channels (frames*512 each)."""
    t = tables or default_tables()
    su = np.ascontiguousarray(su, np.uint8).reshape(-1, SU_BYTES)
    n_su = su.shape[0]
    nf = (n_su + n_ch - 1) // n_ch
    outs = [np.zeros(nf * FRAME, np.float32) for _ in range(n_ch)]
    ptrs = (C.POINTER(C.c_float) * n_ch)(*[_fp(c) for c in outs])
    L = lib()

    def run(rng):
        L.c1o_decode_su_range(C.byref(t), _u8(su), n_su, n_ch, rng[0], rng[1], ptrs)

    if threads <= 1:
        run((0, nf))
    else:
        ranges = [(a, min(a + chunk_frames, nf)) for a in range(0, nf, chunk_frames)]
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, ranges))
    return outs


def aea_header(title: str, su_count: int, n_ch: int) -> np.ndarray:
    out = np.zeros(AEA_HEADER, np.uint8)
    lib().c1o_aea_header(title.encode("utf-8"), su_count, n_ch, _u8(out))
    return out


def aea_parse(hdr: np.ndarray):
    hdr = np.ascontiguousarray(hdr, np.uint8)
    title = C.create_string_buffer(257)
    cnt = C.c_uint32()
    nch = C.c_int()
    rc = lib().c1o_aea_parse(_u8(hdr), len(hdr), title, C.byref(cnt), C.byref(nch))
    if rc == -1:
        raise ValueError("Header must be 2048 bytes")
    if rc == -2:
        raise ValueError("Invalid AEA file")
    return title.value.decode("utf-8", "replace"), cnt.value, nch.value


def pcm_to_int16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x), np.int16)
    lib().c1o_pcm_to_int16(_fp(x), len(x), out.ctypes.data_as(C.POINTER(C.c_int16)))
    return out


def int16_to_pcm(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.int16)
    out = np.zeros(len(x), np.float32)
    lib().c1o_int16_to_pcm(x.ctypes.data_as(C.POINTER(C.c_int16)), len(x), _fp(out))
    return out
