"""ctypes binding of libcarta1_b200.so (the C ABI of include/carta1_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is usable,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcarta1_b200.so")

SU_BYTES = 212
FRAME = 512
AEA_HEADER = 2048


class Carta1Error(RuntimeError):
    pass


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


class EncOpts(C.Structure):
    _fields_ = [
        ("transient_threshold_low", C.c_double),
        ("allocation_bias", C.c_double),
        ("use_fixed_block_modes", C.c_int32),
        ("fixed_block_modes", C.c_int32 * 3),
        ("biased_scale_factors", C.POINTER(C.c_double)),
    ]


# every symbol include/carta1_b200.h declares
EXPORTED_SYMBOLS = [
    "carta1_abi_version", "carta1_default_tables", "carta1_default_enc_opts", "carta1_ctx_create",
    "carta1_ctx_destroy", "carta1_last_error", "carta1_device_count", "carta1_frame_count",
    "carta1_encode_pcm", "carta1_decode_su", "carta1_encode_pcm_s16", "carta1_decode_su_s16",
    "carta1_enc_create", "carta1_enc_destroy", "carta1_enc_reset", "carta1_enc_frames",
    "carta1_dec_create", "carta1_dec_destroy", "carta1_dec_reset", "carta1_dec_frames",
    "carta1_dec_frames_expanded",
    "carta1_encode_device", "carta1_decode_device", "carta1_ctx_sync", "carta1_ctx_stream",
    "carta1_ctx_launch_count", "carta1_debug_encode_stages", "carta1_debug_decode_stages",
    "carta1_aea_write_header", "carta1_aea_parse_header", "carta1_kernel_count", "carta1_kernel_name",
    "carta1_ctx_profile", "carta1_ctx_profile_read", "carta1_debug_selftest",
    "carta1_ctx_set_max_units_per_pass", "carta1_host_alloc", "carta1_host_free", "carta1_deserialize_units",
    "carta1_debug_transient_scores", "carta1_ctx_near_threshold",
    "carta1_encode_pcm_shard", "carta1_decode_su_shard",
]

_lib = None


def load():
    """Load the shared library (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Carta1Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(carta1_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, u8 = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint8)
    fpp = C.POINTER(C.POINTER(C.c_float))
    L.carta1_abi_version.restype = C.c_int
    L.carta1_default_tables.argtypes = [C.POINTER(Tables)]
    L.carta1_default_tables.restype = None
    L.carta1_default_enc_opts.argtypes = [C.POINTER(EncOpts)]
    L.carta1_default_enc_opts.restype = None
    L.carta1_ctx_create.argtypes = [C.c_int, C.POINTER(Tables), C.POINTER(vp)]
    L.carta1_ctx_destroy.argtypes = [vp]
    L.carta1_ctx_destroy.restype = None
    L.carta1_last_error.argtypes = [vp]
    L.carta1_last_error.restype = C.c_char_p
    L.carta1_device_count.restype = C.c_int
    L.carta1_frame_count.argtypes = [sz]
    L.carta1_frame_count.restype = sz
    L.carta1_encode_pcm.argtypes = [vp, fpp, C.c_int, sz, C.POINTER(EncOpts), vp, sz, C.POINTER(sz)]
    L.carta1_decode_su.argtypes = [vp, vp, sz, C.c_int, fpp]
    L.carta1_encode_pcm_s16.argtypes = [vp, vp, C.c_int, sz, C.POINTER(EncOpts), vp, sz, C.POINTER(sz)]
    L.carta1_decode_su_s16.argtypes = [vp, vp, sz, C.c_int, vp]
    L.carta1_encode_pcm_shard.argtypes = [vp, fpp, C.c_int, sz, sz, C.POINTER(EncOpts), vp, sz, C.POINTER(sz)]
    L.carta1_decode_su_shard.argtypes = [vp, vp, sz, C.c_int, sz, fpp]
    L.carta1_enc_create.argtypes = [vp, C.POINTER(EncOpts), C.c_int, C.POINTER(vp)]
    L.carta1_enc_destroy.argtypes = [vp]
    L.carta1_enc_destroy.restype = None
    L.carta1_enc_reset.argtypes = [vp]
    L.carta1_enc_frames.argtypes = [vp, vp, C.c_int, vp]
    L.carta1_dec_create.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.carta1_dec_destroy.argtypes = [vp]
    L.carta1_dec_destroy.restype = None
    L.carta1_dec_reset.argtypes = [vp]
    L.carta1_dec_frames.argtypes = [vp, vp, C.c_int, vp]
    L.carta1_dec_frames_expanded.argtypes = [vp, vp, vp, vp, vp, C.c_int, vp]
    L.carta1_encode_device.argtypes = [vp, vp, sz, C.c_int, sz, sz, sz, C.POINTER(EncOpts), vp, sz, sz, C.c_int]
    L.carta1_decode_device.argtypes = [vp, vp, sz, sz, sz, C.c_int, sz, sz, vp, sz, C.c_int]
    L.carta1_ctx_sync.argtypes = [vp]
    L.carta1_ctx_stream.argtypes = [vp]
    L.carta1_ctx_stream.restype = vp
    L.carta1_ctx_launch_count.argtypes = [vp]
    L.carta1_ctx_launch_count.restype = C.c_uint64
    L.carta1_kernel_count.restype = C.c_int
    L.carta1_kernel_name.argtypes = [C.c_int]
    L.carta1_kernel_name.restype = C.c_char_p
    L.carta1_ctx_profile.argtypes = [vp, C.c_int]
    L.carta1_ctx_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.carta1_debug_selftest.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.carta1_ctx_set_max_units_per_pass.argtypes = [vp, sz]
    L.carta1_deserialize_units.argtypes = [vp, vp, sz, vp, vp, vp, vp, vp]
    L.carta1_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.carta1_host_free.argtypes = [vp]
    L.carta1_host_free.restype = None
    L.carta1_debug_encode_stages.argtypes = [vp, vp, sz, C.POINTER(EncOpts), vp, vp, vp, vp, vp]
    L.carta1_debug_decode_stages.argtypes = [vp, vp, sz, vp, vp, vp]
    L.carta1_debug_transient_scores.argtypes = [vp, vp, sz, C.POINTER(EncOpts), vp]
    L.carta1_ctx_near_threshold.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
    L.carta1_aea_write_header.argtypes = [C.c_char_p, C.c_uint32, C.c_int, vp]
    L.carta1_aea_parse_header.argtypes = [vp, sz, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    _lib = L
    return L


def default_tables() -> Tables:
    t = Tables()
    load().carta1_default_tables(C.byref(t))
    return t


def make_enc_opts(transient_threshold_low=1.0, allocation_bias=1.0, fixed_block_modes=None,
                  biased_scale_factors=None) -> EncOpts:
    o = EncOpts()
    load().carta1_default_enc_opts(C.byref(o))
    o.transient_threshold_low = float(transient_threshold_low)
    o.allocation_bias = float(allocation_bias)
    if fixed_block_modes is not None:
        o.use_fixed_block_modes = 1
        for i in range(3):
            o.fixed_block_modes[i] = int(fixed_block_modes[i])
    if biased_scale_factors is not None:
        arr = np.ascontiguousarray(biased_scale_factors, np.float64)
        assert arr.shape == (64,)
        o._keep = arr  # keep the buffer alive with the struct
        o.biased_scale_factors = arr.ctypes.data_as(C.POINTER(C.c_double))
    return o


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


class Context:
    """One context per GPU (carta1_ctx)."""

    def __init__(self, device: int = 0, tables: Tables | None = None):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.carta1_ctx_create(int(device), C.byref(tables) if tables is not None else None, C.byref(h))
        if rc != 0:
            raise Carta1Error(self.L.carta1_last_error(None).decode())
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.carta1_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.L.carta1_last_error(self.h).decode()
            if rc == 1:
                if "requires" in msg:
                    raise TypeError(msg)
                raise ValueError(msg)
            raise Carta1Error(msg)

    # ---- whole buffers, host memory
    def encode_pcm(self, channels, opts: EncOpts | None = None) -> np.ndarray:
        chans = [np.ascontiguousarray(c, np.float32) for c in channels]
        n = max((len(c) for c in chans), default=0)
        chans = [c if len(c) == n else np.concatenate([c, np.zeros(n - len(c), np.float32)]) for c in chans]
        n_ch = len(chans)
        nf = (n + FRAME - 1) // FRAME
        su = np.zeros((nf * max(n_ch, 1), SU_BYTES), np.uint8)
        ptrs = (C.POINTER(C.c_float) * max(n_ch, 1))(*[c.ctypes.data_as(C.POINTER(C.c_float)) for c in chans])
        n_su = C.c_size_t()
        self._check(self.L.carta1_encode_pcm(self.h, ptrs, n_ch, n, C.byref(opts) if opts is not None else None,
                                             _ptr(su), su.nbytes, C.byref(n_su)))
        return su[:n_su.value]

    def decode_su(self, su: np.ndarray, n_ch: int):
        su = np.ascontiguousarray(su, np.uint8).reshape(-1, SU_BYTES)
        n_su = su.shape[0]
        nf = (n_su + max(n_ch, 1) - 1) // max(n_ch, 1)
        outs = [np.zeros(nf * FRAME, np.float32) for _ in range(max(n_ch, 1))]
        ptrs = (C.POINTER(C.c_float) * len(outs))(*[c.ctypes.data_as(C.POINTER(C.c_float)) for c in outs])
        self._check(self.L.carta1_decode_su(self.h, _ptr(su), n_su, n_ch, ptrs))
        return outs[:n_ch]

    def encode_pcm_s16(self, interleaved: np.ndarray, n_ch: int, opts: EncOpts | None = None) -> np.ndarray:
        x = np.ascontiguousarray(interleaved, np.int16).reshape(-1)
        n = len(x) // n_ch
        nf = (n + FRAME - 1) // FRAME
        su = np.zeros((nf * n_ch, SU_BYTES), np.uint8)
        n_su = C.c_size_t()
        self._check(self.L.carta1_encode_pcm_s16(self.h, _ptr(x), n_ch, n, C.byref(opts) if opts is not None else None,
                                                 _ptr(su), su.nbytes, C.byref(n_su)))
        return su[:n_su.value]

    def decode_su_s16(self, su: np.ndarray, n_ch: int) -> np.ndarray:
        su = np.ascontiguousarray(su, np.uint8).reshape(-1, SU_BYTES)
        n_su = su.shape[0]
        nf = (n_su + n_ch - 1) // n_ch
        out = np.zeros(nf * FRAME * n_ch, np.int16)
        self._check(self.L.carta1_decode_su_s16(self.h, _ptr(su), n_su, n_ch, _ptr(out)))
        return out

    # ---- device-resident (raw device pointers as ints)
    def encode_device(self, d_pcm: int, row_stride: int, n_streams: int, valid_samples: int, halo_frames: int,
                      n_frames: int, opts: EncOpts | None, d_su: int, su_frame_stride: int, su_stream_stride: int,
                      sync: bool = False):
        self._check(self.L.carta1_encode_device(self.h, d_pcm, row_stride, n_streams, valid_samples, halo_frames,
                                                n_frames, C.byref(opts) if opts is not None else None, d_su,
                                                su_frame_stride, su_stream_stride, int(sync)))

    def decode_device(self, d_su: int, su_frame_stride: int, su_stream_stride: int, n_su_valid: int, n_streams: int,
                      halo_frames: int, n_frames: int, d_pcm: int, row_stride: int, sync: bool = False):
        self._check(self.L.carta1_decode_device(self.h, d_su, su_frame_stride, su_stream_stride, n_su_valid,
                                                n_streams, halo_frames, n_frames, d_pcm, row_stride, int(sync)))

    def sync(self):
        self._check(self.L.carta1_ctx_sync(self.h))

    def set_max_units_per_pass(self, units: int):
        self._check(self.L.carta1_ctx_set_max_units_per_pass(self.h, int(units)))

    @property
    def stream(self) -> int:
        return self.L.carta1_ctx_stream(self.h) or 0

    @property
    def launch_count(self) -> int:
        return int(self.L.carta1_ctx_launch_count(self.h))

    def profile(self, enable: bool):
        self._check(self.L.carta1_ctx_profile(self.h, int(enable)))

    def profile_read(self) -> dict:
        """{kernel name: (total ms, launches)} since the last read; synchronises."""
        n = self.L.carta1_kernel_count()
        ms = (C.c_double * n)()
        cnt = (C.c_uint64 * n)()
        self._check(self.L.carta1_ctx_profile_read(self.h, ms, cnt, n))
        return {self.L.carta1_kernel_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n) if cnt[i]}

    # ---- whole buffers into caller-provided (e.g. pinned) arrays
    def encode_pcm_into(self, chans, su_out: np.ndarray, opts: EncOpts | None = None) -> int:
        n = len(chans[0])
        ptrs = (C.POINTER(C.c_float) * len(chans))(*[c.ctypes.data_as(C.POINTER(C.c_float)) for c in chans])
        n_su = C.c_size_t()
        self._check(self.L.carta1_encode_pcm(self.h, ptrs, len(chans), n, C.byref(opts) if opts is not None else None,
                                             _ptr(su_out), su_out.nbytes, C.byref(n_su)))
        return n_su.value

    def decode_su_into(self, su: np.ndarray, n_su: int, n_ch: int, outs) -> None:
        ptrs = (C.POINTER(C.c_float) * len(outs))(*[c.ctypes.data_as(C.POINTER(C.c_float)) for c in outs])
        self._check(self.L.carta1_decode_su(self.h, _ptr(su), n_su, n_ch, ptrs))

    def encode_pcm_shard_into(self, chans, halo_frames: int, su_out: np.ndarray, opts: EncOpts | None = None) -> int:
        """One shard of a longer stream (carta1_encode_pcm_shard): chans start halo_frames frames before the first
        emitted frame."""
        n = len(chans[0])
        ptrs = (C.POINTER(C.c_float) * len(chans))(*[c.ctypes.data_as(C.POINTER(C.c_float)) for c in chans])
        n_su = C.c_size_t()
        self._check(self.L.carta1_encode_pcm_shard(self.h, ptrs, len(chans), n, int(halo_frames),
                                                   C.byref(opts) if opts is not None else None, _ptr(su_out), su_out.nbytes,
                                                   C.byref(n_su)))
        return n_su.value

    def decode_su_shard_into(self, su: np.ndarray, n_su: int, n_ch: int, halo_frames: int, outs) -> None:
        """One shard of a longer file (carta1_decode_su_shard): su starts halo_frames frames before the first
        emitted frame."""
        ptrs = (C.POINTER(C.c_float) * len(outs))(*[c.ctypes.data_as(C.POINTER(C.c_float)) for c in outs])
        self._check(self.L.carta1_decode_su_shard(self.h, _ptr(su), n_su, n_ch, int(halo_frames), ptrs))

    def encode_pcm_s16_into(self, interleaved: np.ndarray, n_ch: int, su_out: np.ndarray, opts: EncOpts | None = None) -> int:
        n = interleaved.size // n_ch
        n_su = C.c_size_t()
        self._check(self.L.carta1_encode_pcm_s16(self.h, _ptr(interleaved), n_ch, n, C.byref(opts) if opts is not None else None,
                                                 _ptr(su_out), su_out.nbytes, C.byref(n_su)))
        return n_su.value

    def decode_su_s16_into(self, su: np.ndarray, n_su: int, n_ch: int, out: np.ndarray) -> None:
        self._check(self.L.carta1_decode_su_s16(self.h, _ptr(su), n_su, n_ch, _ptr(out)))

    def deserialize_units(self, su: np.ndarray) -> dict:
        """Batched deserializeFrame (carta1_deserialize_units): arrays per unit, integers in bitstream order."""
        su = np.ascontiguousarray(su, np.uint8).reshape(-1, SU_BYTES)
        n = su.shape[0]
        out = dict(n_bfu=np.zeros(n, np.uint8), block_modes=np.zeros((n, 3), np.int8), wl=np.zeros((n, 52), np.uint8),
                   sfi=np.zeros((n, 52), np.uint8), q=np.zeros((n, 512), np.int32))
        self._check(self.L.carta1_deserialize_units(self.h, _ptr(su), n, _ptr(out["n_bfu"]), _ptr(out["block_modes"]),
                                                    _ptr(out["wl"]), _ptr(out["sfi"]), _ptr(out["q"])))
        return out

    def selftest(self) -> int:
        bad = C.c_uint64()
        self._check(self.L.carta1_debug_selftest(self.h, C.byref(bad)))
        return int(bad.value)

    # ---- stage taps
    def debug_encode_stages(self, pcm: np.ndarray, opts: EncOpts | None = None):
        pcm = np.ascontiguousarray(pcm, np.float32)
        nf = (len(pcm) + FRAME - 1) // FRAME
        bands = np.zeros((nf, 512), np.float32)
        mags = np.zeros((nf, 256), np.float32)
        modes = np.zeros((nf, 3), np.int32)
        coefs = np.zeros((nf, 512), np.float32)
        su = np.zeros((nf, SU_BYTES), np.uint8)
        self._check(self.L.carta1_debug_encode_stages(self.h, _ptr(pcm), len(pcm),
                                                      C.byref(opts) if opts is not None else None, _ptr(bands),
                                                      _ptr(mags), _ptr(modes), _ptr(coefs), _ptr(su)))
        return dict(bands=bands, mags=mags, modes=modes, coefs=coefs, su=su)

    def debug_transient_scores(self, pcm: np.ndarray, opts: EncOpts | None = None) -> np.ndarray:
        """[n_frames][3] transient scores (auto block modes) of one row of PCM."""
        pcm = np.ascontiguousarray(pcm, np.float32)
        nf = (len(pcm) + FRAME - 1) // FRAME
        scores = np.zeros((nf, 3), np.float64)
        self._check(self.L.carta1_debug_transient_scores(self.h, _ptr(pcm), len(pcm),
                                                         C.byref(opts) if opts is not None else None, _ptr(scores)))
        return scores

    def near_threshold(self, reset: bool = False) -> dict:
        """Close calls of `score > threshold` since creation / the last reset (carta1_ctx_near_threshold)."""
        c = (C.c_uint64 * 3)()
        self._check(self.L.carta1_ctx_near_threshold(self.h, c, int(reset)))
        return {"decisions": int(c[0]), "within_1e-9": int(c[1]), "within_1e-12": int(c[2])}

    def debug_decode_stages(self, su: np.ndarray):
        su = np.ascontiguousarray(su, np.uint8).reshape(-1, SU_BYTES)
        n = su.shape[0]
        coefs = np.zeros((n, 512), np.float32)
        bands = np.zeros((n, 512), np.float32)
        pcm = np.zeros((n, 512), np.float32)
        self._check(self.L.carta1_debug_decode_stages(self.h, _ptr(su), n, _ptr(coefs), _ptr(bands), _ptr(pcm)))
        return dict(coefs=coefs, bands=bands, pcm=pcm)


class StreamEncoder:
    """n_streams stateful encode() closures advanced together (carta1_encoder)."""

    def __init__(self, ctx: Context, opts: EncOpts | None = None, n_streams: int = 1):
        self.ctx, self.n_streams, self._opts = ctx, n_streams, opts
        h = C.c_void_p()
        ctx._check(ctx.L.carta1_enc_create(ctx.h, C.byref(opts) if opts is not None else None, n_streams, C.byref(h)))
        self.h = h

    def frames(self, pcm: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """pcm [n_streams][n_frames][512] -> sound units [n_streams][n_frames][212] (into `out` if given,
        e.g. a pinned buffer)."""
        pcm = np.ascontiguousarray(pcm, np.float32).reshape(self.n_streams, -1, FRAME)
        nf = pcm.shape[1]
        su = np.empty((self.n_streams, nf, SU_BYTES), np.uint8) if out is None else out
        assert su.dtype == np.uint8 and su.size == self.n_streams * nf * SU_BYTES and su.flags.c_contiguous
        self.ctx._check(self.ctx.L.carta1_enc_frames(self.h, _ptr(pcm), nf, _ptr(su)))
        return su.reshape(self.n_streams, nf, SU_BYTES)

    def reset(self):
        self.ctx._check(self.ctx.L.carta1_enc_reset(self.h))

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.L.carta1_enc_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamDecoder:
    """n_streams stateful decode() closures advanced together (carta1_decoder)."""

    def __init__(self, ctx: Context, n_streams: int = 1):
        self.ctx, self.n_streams = ctx, n_streams
        h = C.c_void_p()
        ctx._check(ctx.L.carta1_dec_create(ctx.h, n_streams, C.byref(h)))
        self.h = h

    def frames(self, su: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """sound units [n_streams][n_frames][212] -> pcm [n_streams][n_frames][512] (into `out` if given)."""
        su = np.ascontiguousarray(su, np.uint8).reshape(self.n_streams, -1, SU_BYTES)
        nf = su.shape[1]
        pcm = np.empty((self.n_streams, nf, FRAME), np.float32) if out is None else out
        assert pcm.dtype == np.float32 and pcm.size == self.n_streams * nf * FRAME and pcm.flags.c_contiguous
        self.ctx._check(self.ctx.L.carta1_dec_frames(self.h, _ptr(su), nf, _ptr(pcm)))
        return pcm.reshape(self.n_streams, nf, FRAME)

    def frames_expanded(self, q: np.ndarray, sfi: np.ndarray, bits: np.ndarray, modes: np.ndarray) -> np.ndarray:
        """Frame objects in position-expanded form (carta1_dec_frames_expanded)."""
        q = np.ascontiguousarray(q, np.int32).reshape(self.n_streams, -1, FRAME)
        nf = q.shape[1]
        sfi = np.ascontiguousarray(sfi, np.uint8).reshape(self.n_streams, nf, FRAME)
        bits = np.ascontiguousarray(bits, np.uint8).reshape(self.n_streams, nf, FRAME)
        modes = np.ascontiguousarray(modes, np.int32).reshape(self.n_streams, nf, 3)
        pcm = np.zeros((self.n_streams, nf, FRAME), np.float32)
        self.ctx._check(self.ctx.L.carta1_dec_frames_expanded(self.h, _ptr(q), _ptr(sfi), _ptr(bits), _ptr(modes),
                                                              nf, _ptr(pcm)))
        return pcm

    def reset(self):
        self.ctx._check(self.ctx.L.carta1_dec_reset(self.h))

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.L.carta1_dec_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(shape, dtype) -> np.ndarray:
    """A numpy array on page-locked host memory (carta1_host_alloc): the device reads and writes it in place.
    The memory is released when the array (and every view of it) is garbage-collected."""
    L = load()
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    p = C.c_void_p()
    rc = L.carta1_host_alloc(n * dt.itemsize, C.byref(p))
    if rc != 0:
        raise Carta1Error("carta1_host_alloc failed: " + L.carta1_last_error(None).decode())
    buf = (C.c_char * max(n * dt.itemsize, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    import weakref

    weakref.finalize(buf, L.carta1_host_free, p.value)
    return arr


def aea_write_header(title: str, su_count: int, n_ch: int) -> np.ndarray:
    out = np.zeros(AEA_HEADER, np.uint8)
    rc = load().carta1_aea_write_header(title.encode("utf-8"), su_count, n_ch, _ptr(out))
    if rc != 0:
        raise Carta1Error("carta1_aea_write_header failed")
    return out


def aea_parse_header(hdr) -> tuple[str, int, int]:
    hdr = np.ascontiguousarray(np.frombuffer(bytes(hdr), np.uint8) if not isinstance(hdr, np.ndarray) else hdr, np.uint8)
    title = C.create_string_buffer(257)
    cnt = C.c_uint32()
    nch = C.c_int()
    L = load()
    rc = L.carta1_aea_parse_header(_ptr(hdr), len(hdr), title, C.byref(cnt), C.byref(nch))
    if rc != 0:
        raise ValueError(L.carta1_last_error(None).decode())
    return title.value.decode("utf-8", "replace"), cnt.value, nch.value
