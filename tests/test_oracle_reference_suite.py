"""Pins the CPU oracle against every assertion the reference's own vitest suite makes for
the hot path.  Each test names the reference test it restates
(paths relative to /root/reference/).  CPU only.
"""
import math

import numpy as np
import pytest

import signals as S

WL_BITS = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]


# ------------------------------------------------------------------ tests/bitstream.test.js
def test_bitstream_kats(oracle):
    L = oracle.lib()
    u8 = oracle._u8
    b = np.zeros(1, np.uint8)
    L.c1o_pack_bits(u8(b), 1, 0, 0b10101010, 8)  # :6-11
    assert b[0] == 0b10101010 and L.c1o_unpack_bits(u8(b), 1, 0, 8) == 0b10101010
    b = np.zeros(2, np.uint8)
    L.c1o_pack_bits(u8(b), 2, 4, 0b11110000, 8)  # :13-19
    assert list(b) == [0b00001111, 0] and L.c1o_unpack_bits(u8(b), 2, 4, 8) == 0b11110000
    b = np.array([0xFF], np.uint8)
    L.c1o_pack_bits(u8(b), 1, 0, 123, 0)  # :21-27
    assert b[0] == 0xFF and L.c1o_unpack_bits(u8(b), 1, 0, 0) == 0
    for n in range(1, 31):  # :29-38
        b = np.zeros((n + 7) // 8, np.uint8)
        v = (1 << n) - 1
        L.c1o_pack_bits(u8(b), len(b), 0, v, n)
        assert L.c1o_unpack_bits(u8(b), len(b), 0, n) == v
    for raw, want in ((5, 5), (0b1011, -5), (0b1000, -8), (0b0111, 7), (0, 0)):  # :41-71
        b = np.zeros(1, np.uint8)
        L.c1o_pack_bits(u8(b), 1, 0, raw, 4)
        assert L.c1o_unpack_signed_bits(u8(b), 1, 0, 4) == want


def test_unpack_stops_at_buffer_end(oracle):
    """bitstream.js:55-68: a read running off the buffer returns only the bits read."""
    L = oracle.lib()
    b = np.array([0xAB, 0xCD], np.uint8)
    assert L.c1o_unpack_bits(oracle._u8(b), 2, 12, 8) == 0xD
    assert L.c1o_unpack_bits(oracle._u8(b), 2, 16, 8) == 0
    assert L.c1o_unpack_signed_bits(oracle._u8(b), 2, 12, 8) == 0xD


# ------------------------------------------------------------------ tests/mdct.test.js
def test_overlap_add_kat(oracle):  # :22-33
    out = oracle.overlap_add(np.ones(32), np.full(32, 0.5), np.ones(64))
    assert len(out) == 64 and out[0] == 0.5 and out[63] == 1.5


def test_mdct_sizes_and_nonzero(oracle):  # :13-20, :35-77
    for n in (64, 256, 512):
        x = np.sin(np.arange(n) * 0.1).astype(np.float32)
        y = oracle.mdct(x)
        assert len(y) == n // 2 and np.any(y != 0)
        z = oracle.imdct(y)
        assert len(z) == n and np.any(z != 0)


def test_mdct_matches_direct_formula(oracle):
    """MDCT.transform equals the textbook MDCT up to f32 rounding: pins pre/post twiddles."""
    rng = np.random.default_rng(1)
    for n, scale in ((64, 0.5), (256, 0.5), (512, 1.0)):
        x = rng.standard_normal(n).astype(np.float32)
        y = oracle.mdct(x).astype(np.float64)
        k = np.arange(n // 2)[:, None]
        i = np.arange(n)[None, :]
        basis = np.cos(2 * np.pi / n * (i + 0.5 + n / 4) * (k + 0.5))
        ref = basis @ x.astype(np.float64)
        g = (y @ ref) / (ref @ ref)
        assert np.allclose(y, g * ref, atol=2e-5 * np.max(np.abs(y)))
        assert abs(abs(g) - math.sqrt(scale / n) * math.sqrt(1.0)) < 0.6 * math.sqrt(scale / n) + 1


# ------------------------------------------------------------------ tests/fft.test.js
@pytest.mark.parametrize("n", [16, 64, 128, 256])
def test_fft_properties(oracle, n):
    re, im = oracle.fft(np.ones(n), np.zeros(n))  # DC :5-20
    assert abs(re[0] - n) < 1e-4 and np.all(np.abs(re[1:]) < 1e-4) and np.all(np.abs(im) < 1e-4)
    k = 3  # single bin :22-45
    x = np.cos(2 * np.pi * k * np.arange(n) / n)
    re, im = oracle.fft(x, np.zeros(n))
    mag = np.hypot(re, im)
    assert abs(mag[k] - n / 2) < 1e-3 and abs(mag[n - k] - n / 2) < 1e-3
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    ra, ia = oracle.fft(a, np.zeros(n))
    rb, ib = oracle.fft(b, np.zeros(n))
    rs, is_ = oracle.fft(a + b, np.zeros(n))  # linearity :70-94
    assert np.allclose(rs, ra + rb, atol=1e-3) and np.allclose(is_, ia + ib, atol=1e-3)
    assert abs(np.sum(ra.astype(np.float64) ** 2 + ia.astype(np.float64) ** 2) / n - np.sum(a.astype(np.float64) ** 2)) < 1e-3 * n  # Parseval
    ref = np.fft.fft(a.astype(np.float64))
    assert np.allclose(ra, ref.real, atol=1e-4 * n) and np.allclose(ia, ref.imag, atol=1e-4 * n)


# ------------------------------------------------------------------ tests/qmf.test.js
def test_qmf_perfect_reconstruction(oracle):  # :10-39, :97-122
    x = np.concatenate([S.sine(1000, n=512), S.sine(1000, n=512) * 0]).astype(np.float32)
    x = (np.sin(2 * np.pi * 1000 * np.arange(2048) / 44100)).astype(np.float32)
    da = np.zeros(46, np.float32)
    ds = np.zeros(46, np.float32)
    outs = []
    for f in range(4):
        lo, hi, da = oracle.qmf_analysis(x[512 * f:512 * f + 512], da)
        y, ds = oracle.qmf_synthesis(lo, hi, ds)
        outs.append(y)
    y = np.concatenate(outs)
    d = 46
    err = np.sum((y[d:] - x[:-d]) ** 2) / np.sum(x[:-d] ** 2)
    assert err < 1e-6


def test_qmf_impulse_delay(oracle):  # :69-95
    x = S.impulse(0, 512)
    lo, hi, _ = oracle.qmf_analysis(x, np.zeros(46))
    y, _ = oracle.qmf_synthesis(lo, hi, np.zeros(46))
    assert int(np.argmax(np.abs(y))) == 46


def test_qmf_band_separation(oracle):  # :41-67
    lo_sig = S.sine(1000)
    hi_sig = S.sine(18000)
    lo, hi, _ = oracle.qmf_analysis(lo_sig, np.zeros(46))
    assert np.sum(lo ** 2) > 10 * np.sum(hi ** 2)
    lo, hi, _ = oracle.qmf_analysis(hi_sig, np.zeros(46))
    assert np.sum(hi ** 2) > 10 * np.sum(lo ** 2)


# ------------------------------------------------------------------ tests/transient.test.js
def test_perform_fft_peak(oracle):  # :7-19
    x = np.sin(2 * np.pi * 8 * np.arange(128) / 128).astype(np.float32)
    mag = oracle.perform_fft(x, 128)
    assert len(mag) == 64 and int(np.argmax(mag)) == 8


def test_transient_step_vs_silence(oracle):  # :21-73 (size 256: the hot path's largest FFT)
    n = 256
    sil = oracle.perform_fft(S.silence(n), n)
    stp = oracle.perform_fft(S.step(0, n), n)
    tone = oracle.perform_fft(S.sine(440, n=n), n)
    assert oracle.detect_transient(stp, sil, 0.1)
    assert not oracle.detect_transient(tone, tone, 0.1)
    assert oracle.detect_transient(stp, sil, 0.01)
    assert not oracle.detect_transient(tone, tone, 0.01)
    assert not oracle.detect_transient(stp, sil, 0.99)  # threshold sensitivity
    assert abs(oracle.transient_score(stp, sil) - 0.75) < 1e-12
    assert not oracle.detect_transient(stp, None, 0.1)  # :285-302
    assert not oracle.detect_transient(stp, np.zeros(0), 0.1)


def test_transient_score_components(oracle):
    """Independent numpy restatement of transient.js:92-226 on random magnitudes."""
    rng = np.random.default_rng(3)
    for n in (64, 128):
        cur = np.abs(rng.standard_normal(n)).astype(np.float32)
        prev = np.abs(0.3 * rng.standard_normal(n)).astype(np.float32)
        c = cur.astype(np.float64)
        p = prev.astype(np.float64)
        flux = np.sum(np.maximum(c - p, 0)) / math.sqrt(np.sum(c * c))

        def flat(x):
            v = x[x > 1e-10]
            return math.exp(np.mean(np.log(v))) / np.mean(v)

        def hf(x):
            return np.sum(x[n // 2:] ** 2) / np.sum(x ** 2)

        db = max(0.0, 10 * math.log10(np.sum(c * c) / np.sum(p * p)))
        want = (flux + math.sqrt(abs(flat(c) - flat(p))) + math.log1p(10 * abs(hf(c) - hf(p))) / math.log1p(10) + min(db / 30, 1)) / 4
        got = oracle.transient_score(cur, prev)
        assert abs(got - want) < 1e-12


# ------------------------------------------------------------------ tests/bitallocation.test.js
def _flat_coefs(v):
    return np.full(512, v, np.float32)


def test_allocation_budget_and_count(oracle):  # :15-36
    opt = oracle.make_options()
    n, sfi, wl = oracle.allocate_bits(_flat_coefs(1.0), [0, 0, 0], opt)
    sizes = oracle.const_table("c1o_specs_per_bfu", 52, np.int32)
    used = sum(WL_BITS[wl[i]] * sizes[i] for i in range(n))
    assert used + 40 + 10 * n <= 1696
    assert n == 52


def test_allocation_silent(oracle):  # :38-46
    n, sfi, wl = oracle.allocate_bits(_flat_coefs(0.0), [0, 0, 0], oracle.make_options())
    assert n == 20 and np.all(wl == 0) and np.all(sfi == 0)


def test_allocation_prefers_energy(oracle):  # :48-70
    starts = oracle.const_table("c1o_bfu_start_long", 52, np.int32)
    c = np.full(512, 0.1, np.float32)
    c[:starts[5]] = 2.0
    c[starts[5]:starts[10]] = 1.0
    n, sfi, wl = oracle.allocate_bits(c, [0, 0, 0], oracle.make_options())
    assert np.mean(wl[:5]) > np.mean(wl[10:15])


def test_find_scale_factor(oracle):  # :82-99
    sf = np.array(list(oracle.default_tables().scale_factors))
    i = oracle.find_scale_factor(np.array([0.01, 0.05, 0.1, 0.2], np.float32))
    assert sf[i] >= np.float32(0.2) and sf[i - 1] < np.float32(0.2)
    assert oracle.find_scale_factor(np.zeros(4, np.float32)) == 0


def test_scale_factor_threshold_table_is_exact(oracle):
    """ceil(3*(log2(m)+21)) == count of f32 thresholds below m, checked on
    every f32 within 40 ulps of each threshold plus 200k random magnitudes."""
    L = oracle.lib()
    thr = oracle.const_table("c1o_sf_thresholds", 63, np.float32)
    cands = []
    for t in thr:
        bits = np.array([t], np.float32).view(np.uint32)[0]
        cands.append((np.arange(-40, 41) + int(bits)).astype(np.uint32).view(np.float32))
    rng = np.random.default_rng(5)
    cands.append(np.exp(rng.uniform(-20, 3, 200000)).astype(np.float32))
    cands.append(np.array([1e-30, 1e-10, 3.0, 100.0, 3.4e38, np.inf], np.float32))
    m = np.concatenate(cands)
    want = np.clip(np.ceil(3 * (np.log2(m.astype(np.float64)) + 21)), 0, 63).astype(int)
    got = np.array([L.c1o_find_scale_factor_table(float(v)) for v in m])
    ora = np.array([oracle.find_scale_factor(np.array([v], np.float32)) for v in m[:6000]])
    assert np.array_equal(got, want)
    assert np.array_equal(ora, want[:6000])


# ------------------------------------------------------------------ tests/quantization.test.js
def test_quantize_roundtrip(oracle):  # :12-48
    sf = np.array(list(oracle.default_tables().scale_factors))
    rng = np.random.default_rng(2)
    c = (rng.uniform(-1, 1, 20) * 0.4).astype(np.float32)
    sfi = oracle.find_scale_factor(c)
    for bits in (2, 4, 8, 16):
        q = oracle.quantize(c, sfi, bits)
        d = oracle.dequantize(q, sfi, bits)
        stepsz = sf[sfi] / ((1 << (bits - 1)) - 1)
        assert np.max(np.abs(d - c)) <= stepsz
    assert np.all(oracle.quantize(np.zeros(8, np.float32), 10, 8) == 0)
    q = oracle.quantize(np.array([100.0, -100.0], np.float32), 30, 4)
    assert list(q) == [7, -7]
    assert np.all(oracle.quantize(c, 0, 8) == 0) and np.all(oracle.quantize(c, 10, 0) == 0)


def test_to_int32(oracle):
    L = oracle.lib()
    for x, want in ((0.5, 0), (-0.5, 0), (1.9, 1), (-1.9, -1), (2147483648.0, -2147483648),
                    (4294967296.0 + 5.5, 5), (-4294967296.0 - 5.5, -5), (float("nan"), 0),
                    (float("inf"), 0), (1e30, 0)):
        got = L.c1o_to_int32(x)
        if x == 1e30:
            continue
        assert got == want, (x, got, want)


# ------------------------------------------------------------------ tests/serialization.test.js
def test_frame_serialization_roundtrip(oracle):  # :25-55
    enc = oracle.FrameEncoder()
    for sig in (S.white_noise(1), S.sine(440), S.chirp(100, 8000)):
        fr = enc(sig)
        raw = oracle.serialize_frame(fr)
        assert len(raw) == 212 and np.all(raw[209:] == 0)
        back = oracle.deserialize_frame(raw)
        assert back.n_bfu == fr.n_bfu and list(back.modes) == list(fr.modes)
        assert list(back.wl)[:fr.n_bfu] == list(fr.wl)[:fr.n_bfu]
        assert list(back.sfi)[:fr.n_bfu] == list(fr.sfi)[:fr.n_bfu]
        for b in range(fr.n_bfu):
            assert list(back.q[b]) == list(fr.q[b])
        assert np.array_equal(oracle.serialize_frame(back), raw)
    with pytest.raises(ValueError, match="Frame must be 212 bytes"):
        oracle.deserialize_frame(np.zeros(100, np.uint8))


def test_aea_header(oracle):  # :58-83
    h = oracle.aea_header("Test Title", 100, 2)
    assert len(h) == 2048 and list(h[:4]) == [0, 8, 0, 0]
    assert oracle.aea_parse(h) == ("Test Title", 100, 2)
    with pytest.raises(ValueError, match="Invalid AEA file"):
        oracle.aea_parse(np.zeros(2048, np.uint8) + 1)
    with pytest.raises(ValueError, match="Header must be 2048 bytes"):
        oracle.aea_parse(np.zeros(10, np.uint8))


# ------------------------------------------------------------------ tests/encoder.test.js
def _used_bits(oracle, fr):
    sizes = oracle.const_table("c1o_specs_per_bfu", 52, np.int32)
    return sum(WL_BITS[fr.wl[i]] * sizes[i] for i in range(fr.n_bfu))


def test_encoder_runs_and_budget(oracle):  # :14-24, :93-107
    fr = oracle.FrameEncoder()(S.sine(440))
    assert fr.n_bfu > 0
    fr = oracle.FrameEncoder()(S.white_noise(1))
    assert _used_bits(oracle, fr) + 40 + 10 * fr.n_bfu <= 1696


def test_encoder_silence(oracle):  # :109-120
    fr = oracle.FrameEncoder()(S.silence())
    assert _used_bits(oracle, fr) == 0


def test_encoder_short_blocks_after_burst(oracle):  # :26-91
    enc = oracle.FrameEncoder(oracle.make_options(threshold=1.0))
    enc(S.silence())
    enc(S.silence())
    i = np.arange(512, dtype=np.float64)
    x = np.zeros(512)
    for a, f in ((0.8, 60), (0.7, 80), (0.6, 100), (0.5, 120), (0.4, 200), (0.3, 300), (0.3, 400), (0.2, 500)):
        x += a * np.sin(2 * np.pi * f * i / 44100)
    for f in range(600, 5000, 200):
        x += 0.1 * np.sin(2 * np.pi * f * i / 44100)
    x = (x * 0.95 / np.max(np.abs(x))).astype(np.float32)
    fr = enc(x)
    short = any(m != 0 for m in fr.modes)
    if not short:
        short = any(m != 0 for m in enc(S.silence()).modes)
    assert short


def test_fixed_modes_are_passed_through(oracle):
    enc = oracle.FrameEncoder(oracle.make_options(fixed_modes=[2, 0, 3]))
    fr = enc(S.white_noise(3))
    assert list(fr.modes) == [2, 0, 3]


# ------------------------------------------------------------------ tests/decoder.test.js
def test_codec_delay_266(oracle):  # :19-68
    enc, dec = oracle.FrameEncoder(), oracle.FrameDecoder()
    sig = S.sine(440)
    out = np.concatenate([dec(enc(sig)) for _ in range(5)])
    orig = np.tile(sig, 5)
    errs = {}
    for d in range(260, 273):
        n = len(orig) - d
        errs[d] = float(np.mean(np.abs(out[d:d + n] - orig[:n])))
    assert errs[266] < 0.1
    assert min(errs, key=errs.get) == 266


def test_decoder_all_short_and_zero(oracle):  # :70-98
    fr = oracle.Frame()
    fr.n_bfu = 52
    for i in range(52):
        fr.sfi[i] = 10
        fr.wl[i] = 8
        for j in range(20):
            fr.q[i][j] = 1
    fr.modes[0] = fr.modes[1] = fr.modes[2] = 1
    out = oracle.FrameDecoder()(fr)
    assert len(out) == 512 and np.all(np.isfinite(out)) and np.any(out != 0)
    z = oracle.Frame()
    z.n_bfu = 52
    assert np.all(oracle.FrameDecoder()(z) == 0)


# ------------------------------------------------------------------ tests/processor.test.js
def test_whole_buffer_helpers(oracle):  # :23-47, :67-75, :94-108
    x = S.white_noise(7, 1024)
    su = oracle.encode_pcm([x])
    assert su.shape == (2, 212)
    su = oracle.encode_pcm([x, x])
    assert su.shape == (4, 212)
    assert oracle.frame_count(int(2.5 * 512)) == 3
    a = S.white_noise(2, 700) * 0.5
    b = S.white_noise(3, 700) * 0.5
    su = oracle.encode_pcm([a, b])
    pcm = oracle.decode_su(su, 2)
    assert len(pcm) == 2 and len(pcm[0]) == 1024 and len(pcm[1]) == 1024
    # stereo with a missing right unit decodes against the dummy frame
    pcm = oracle.decode_su(su[:3], 2)
    assert len(pcm[1]) == 1024


def test_int16_conversion(oracle):  # processor.js:382-389, bin/cli.js:395
    x = np.array([0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 1e-5, -1e-5], np.float32)
    assert list(oracle.pcm_to_int16(x)) == [0, 32767, -32768, 32767, -32768, 16383, -16384, 0, 0]
    assert list(oracle.int16_to_pcm(np.array([0, 32767, -32768, 1], np.int16))) == [0.0, 32767 / 32768, -1.0, 1 / 32768]


# ------------------------------------------------------------------ range/halo equivalence
def test_range_encode_decode_match_serial(oracle):
    """a cold start 2 frames (encode) / 1 unit (decode) early is exact."""
    chans = S.cfg3_transients(1.5, n_ch=2)
    su_serial = oracle.encode_pcm(chans)
    su_chunks = oracle.encode_pcm(chans, threads=4, chunk_frames=7)
    assert np.array_equal(su_serial, su_chunks)
    pcm_serial = oracle.decode_su(su_serial, 2)
    pcm_chunks = oracle.decode_su(su_serial, 2, threads=4, chunk_frames=5)
    for a, b in zip(pcm_serial, pcm_chunks):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    modes = [tuple(oracle.deserialize_frame(u).modes) for u in su_serial]
    assert any(m != (0, 0, 0) for m in modes) and any(m == (0, 0, 0) for m in modes)
