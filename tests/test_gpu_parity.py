"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs.  Bit-exact for sound units, block modes and every f32 intermediate."""
import numpy as np
import pytest

import signals as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import carta1_b200

    c = carta1_b200.Context(0)
    yield c
    c.close()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def oracle_encode_stages(O, pcm, options):
    enc = O.FrameEncoder(options)
    nf = O.frame_count(len(pcm))
    x = np.zeros(nf * 512, np.float32)
    x[:len(pcm)] = pcm
    out = dict(bands=[], mags=[], modes=[], coefs=[], su=[])
    for f in range(nf):
        fr, dbg = enc(x[512 * f:512 * f + 512], debug=True)
        out["bands"].append(np.array(dbg.bands, np.float32))
        out["mags"].append(np.array(dbg.mags, np.float32))
        out["coefs"].append(np.array(dbg.coefs, np.float32))
        out["modes"].append(list(fr.modes))
        out["su"].append(O.serialize_frame(fr))
    return {k: np.array(v) for k, v in out.items()}


def mono_signals():
    rng = np.random.default_rng(11)
    sigs = {
        "transients": S.cfg3_transients(0.6, n_ch=1)[0],
        "sine_noise": S.cfg1_stereo(0.4)[0],
        "chirp": S.cfg2_stereo(0.4)[1],
        "white": (rng.uniform(-1, 1, 512 * 9)).astype(np.float32),
        "silence": np.zeros(512 * 4, np.float32),
        "tiny": (1e-7 * rng.standard_normal(512 * 5)).astype(np.float32),
        "loud": (4.0 * rng.standard_normal(512 * 5)).astype(np.float32),
        "sparse": np.where(rng.uniform(size=512 * 6) > 0.97, rng.standard_normal(512 * 6), 0).astype(np.float32),
        "ragged": (0.3 * rng.standard_normal(512 * 3 + 77)).astype(np.float32),
        # spectra that straddle the 1e-10 bin threshold of the transient detector (transient.js:126-131):
        # some bins are skipped by the flatness sums, others are not, frame by frame
        "threshold": (np.repeat(10.0 ** rng.uniform(-12.5, -9.5, 8), 512) * rng.standard_normal(512 * 8)).astype(np.float32),
        "denormal": (1e-39 * rng.standard_normal(512 * 4)).astype(np.float32),
    }
    return sigs


OPTION_SETS = [
    dict(),
    dict(fixed_modes=[0, 0, 0]),
    dict(fixed_modes=[2, 2, 3]),
    dict(fixed_modes=[0, 2, 0]),
    dict(threshold=0.3),
    dict(bias=0.0),
    dict(bias=0.5),
    dict(bias=2.5),
    dict(bias=5.0, threshold=0.6),
]


def both_opts(O, kw):
    import carta1_b200

    o = O.make_options(threshold=kw.get("threshold", 1.0), bias=kw.get("bias", 1.0), fixed_modes=kw.get("fixed_modes"))
    g = carta1_b200.make_enc_opts(kw.get("threshold", 1.0), kw.get("bias", 1.0), kw.get("fixed_modes"))
    return o, g


@pytest.mark.parametrize("kw", OPTION_SETS, ids=[str(k) for k in OPTION_SETS])
def test_encode_stages_match_oracle(ctx, oracle, kw):
    o_opt, g_opt = both_opts(oracle, kw)
    for name, pcm in mono_signals().items():
        want = oracle_encode_stages(oracle, pcm, o_opt)
        got = ctx.debug_encode_stages(pcm, g_opt)
        assert np.array_equal(bits(got["bands"]), bits(want["bands"])), (name, "bands")
        if "fixed_modes" not in kw:
            assert np.array_equal(bits(got["mags"]), bits(want["mags"])), (name, "mags")
        assert np.array_equal(got["modes"], want["modes"]), (name, "modes")
        assert np.array_equal(bits(got["coefs"]), bits(want["coefs"])), (name, "coefs")
        assert np.array_equal(got["su"], want["su"]), (name, "sound units")


@pytest.mark.parametrize("kw", [dict(), dict(fixed_modes=[0, 0, 0]), dict(fixed_modes=[2, 2, 3]), dict(bias=2.5)],
                         ids=["auto", "long", "short", "bias"])
def test_non_finite_input(ctx, oracle, kw):
    """+-Infinity, NaN and near-overflow PCM: the reference has no input validation, so whatever its arithmetic
    does with them is the contract (ExactRound path of the transforms, NaN-aware max / min, ToInt32)."""
    rng = np.random.default_rng(3)
    for inject in ((np.inf,), (-np.inf,), (np.nan,), (3e38, -3e38), (np.inf, np.nan, -3e38, 3e38, -np.inf)):
        pcm = (0.3 * rng.standard_normal(512 * 8)).astype(np.float32)
        for i, v in enumerate(inject):
            pcm[700 + 611 * i] = v
        o_opt, g_opt = both_opts(oracle, kw)
        want = oracle.encode_pcm([pcm], o_opt)
        assert np.array_equal(ctx.encode_pcm([pcm], g_opt), want), inject
        assert np.array_equal(bits(ctx.decode_su(want, 1)[0]), bits(oracle.decode_su(want, 1)[0])), inject


def test_decode_stages_match_oracle(ctx, oracle):
    for name, pcm in mono_signals().items():
        su = oracle.encode_pcm([pcm])
        dec = oracle.FrameDecoder()
        coefs, bands, out = [], [], []
        for u in su:
            p, dbg = dec(oracle.deserialize_frame(u), debug=True)
            coefs.append(np.array(dbg.coefs, np.float32))
            bands.append(np.array(dbg.bands, np.float32))
            out.append(p)
        got = ctx.debug_decode_stages(su)
        assert np.array_equal(bits(got["coefs"]), bits(np.array(coefs))), (name, "coefs")
        assert np.array_equal(bits(got["bands"]), bits(np.array(bands))), (name, "bands")
        assert np.array_equal(bits(got["pcm"]), bits(np.array(out))), (name, "pcm")


def test_decode_arbitrary_bytes(ctx, oracle):
    """Malformed input fidelity: random bytes exercise every header value,
    word lengths that run past the 212-byte buffer and scale-factor index 0."""
    rng = np.random.default_rng(5)
    su = rng.integers(0, 256, (64, 212), dtype=np.uint8)
    su[::7, 2:30] = 0xFF  # maximal word lengths -> reads beyond the buffer
    want = oracle.decode_su(su, 1)[0]
    got = ctx.decode_su(su, 1)[0]
    assert np.array_equal(bits(got), bits(want))


@pytest.mark.parametrize("n", [0, 1, 511, 512, 513, 1024, 5000])
def test_whole_buffer_lengths(ctx, oracle, n):
    rng = np.random.default_rng(n)
    a = (0.4 * rng.standard_normal(n)).astype(np.float32)
    b = (0.4 * rng.standard_normal(max(n - 100, 0))).astype(np.float32)
    for chans in ([a], [a, b]):
        if n == 0:
            assert len(ctx.encode_pcm(chans)) == 0
            continue
        su = ctx.encode_pcm(chans)
        want = oracle.encode_pcm(chans)
        assert np.array_equal(su, want)
        pcm = ctx.decode_su(su, len(chans))
        ref = oracle.decode_su(want, len(chans))
        for x, y in zip(pcm, ref):
            assert np.array_equal(bits(x), bits(y))


def test_cfg1_stereo_bitexact(ctx, oracle):
    """BASELINE config 1: 10 s stereo sine+noise, default bias, auto block modes."""
    chans = S.cfg1_stereo(10.0)
    su = ctx.encode_pcm(chans)
    want = oracle.encode_pcm(chans, threads=8, chunk_frames=64)
    assert su.shape == (1724, 212)
    assert np.array_equal(su, want)
    pcm = ctx.decode_su(su, 2)
    ref = oracle.decode_su(want, 2, threads=8, chunk_frames=64)
    for x, y in zip(pcm, ref):
        assert np.array_equal(bits(x), bits(y))
    # 16-bit WAV quantisation is bit-exact as well (processor.js:382-389)
    s16 = ctx.decode_su_s16(su, 2).reshape(-1, 2)
    for c in range(2):
        assert np.array_equal(s16[:, c], oracle.pcm_to_int16(ref[c]))


def test_stereo_odd_unit_count_uses_dummy_frame(ctx, oracle):
    chans = S.cfg1_stereo(0.1)
    su = oracle.encode_pcm(chans)[:-1]
    pcm = ctx.decode_su(su, 2)
    ref = oracle.decode_su(su, 2)
    for x, y in zip(pcm, ref):
        assert np.array_equal(bits(x), bits(y))


def test_s16_ingest(ctx, oracle):
    rng = np.random.default_rng(3)
    x = rng.integers(-32768, 32768, (4000, 2), dtype=np.int16)
    su = ctx.encode_pcm_s16(x, 2)
    chans = [oracle.int16_to_pcm(x[:, 0].copy()), oracle.int16_to_pcm(x[:, 1].copy())]
    assert np.array_equal(su, oracle.encode_pcm(chans))


def test_stateful_streams_match_whole_buffer(ctx, oracle):
    """The batched closure API (config 4 shape): any split into calls gives the whole-stream bytes."""
    import carta1_b200

    x = S.cfg4_mono_streams(6, 0.35)
    nf = x.shape[1] // 512
    x = x[:, :nf * 512]
    want = np.stack([oracle.encode_pcm([x[s]]) for s in range(6)])
    for split in ([nf], [1] * nf, [1, 2, 5, nf - 8]):
        enc = carta1_b200.StreamEncoder(ctx, None, 6)
        dec = carta1_b200.StreamDecoder(ctx, 6)
        outs, pcms, pos = [], [], 0
        for k in split:
            su = enc.frames(x[:, 512 * pos:512 * (pos + k)])
            outs.append(su)
            pcms.append(dec.frames(su))
            pos += k
        got = np.concatenate(outs, axis=1)
        assert np.array_equal(got, want), split
        pcm = np.concatenate(pcms, axis=1).reshape(6, -1)
        for s in range(6):
            assert np.array_equal(bits(pcm[s]), bits(oracle.decode_su(want[s], 1)[0])), split
        enc.close()
        dec.close()


def test_stateful_graph_replay_survives_reallocation_and_profiling(oracle):
    """Same-shaped calls replay a captured CUDA graph (c1_abi.cu GraphCache).  The graph holds raw scratch
    pointers: a larger call on another handle of the same context moves the scratch (new allocation
    generation), a profiling session bypasses graphs; the stream must come out the same through all of it."""
    import carta1_b200

    c = carta1_b200.Context(0)  # a fresh context: its scratch starts small
    try:
        x = S.cfg4_mono_streams(3, 0.5, seed=77)
        nf = x.shape[1] // 512
        x = x[:, :nf * 512]
        want = np.stack([oracle.encode_pcm([x[s]]) for s in range(3)])
        enc, dec = carta1_b200.StreamEncoder(c, None, 3), carta1_b200.StreamDecoder(c, 3)
        big = S.cfg4_mono_streams(40, 0.2, seed=5)
        big = big[:, :(big.shape[1] // 512) * 512]
        outs, pcms = [], []
        for f in range(nf):
            if f == 6:    # grows the context's scratch: the captured graphs must be dropped
                e2 = carta1_b200.StreamEncoder(c, None, 40)
                d2 = carta1_b200.StreamDecoder(c, 40)
                d2.frames(e2.frames(big))
                e2.close(); d2.close()
            if f == 12:
                c.profile(True)
            if f == 15:
                c.profile_read(); c.profile(False)
            su = enc.frames(x[:, 512 * f:512 * (f + 1)])
            outs.append(su)
            pcms.append(dec.frames(su))
        got = np.concatenate(outs, axis=1)
        assert np.array_equal(got, want)
        pcm = np.concatenate(pcms, axis=1).reshape(3, -1)
        for s in range(3):
            assert np.array_equal(bits(pcm[s]), bits(oracle.decode_su(want[s], 1)[0]))
        enc.close(); dec.close()
    finally:
        c.close()


@pytest.mark.parametrize("units", [4, 14, 250])
def test_chunked_host_path(ctx, oracle, units):
    """More frames than one pass holds: the halo logic of the chunked entry points
    (2 frames of PCM for encode, 1 sound unit for decode, at every pass boundary)."""
    chans = S.cfg3_transients(1.5, n_ch=2)
    want = oracle.encode_pcm(chans, threads=8, chunk_frames=32)
    ref = oracle.decode_su(want, 2, threads=8, chunk_frames=32)
    ctx.set_max_units_per_pass(units)
    try:
        su = ctx.encode_pcm(chans)
        assert np.array_equal(su, want)
        pcm = ctx.decode_su(su, 2)
        for x, y in zip(pcm, ref):
            assert np.array_equal(bits(x), bits(y))
        s16 = ctx.decode_su_s16(su, 2).reshape(-1, 2)
        for c in range(2):
            assert np.array_equal(s16[:, c], oracle.pcm_to_int16(ref[c]))
        x16 = np.stack([oracle.pcm_to_int16(c) for c in chans], axis=1)
        back = [oracle.int16_to_pcm(x16[:, c].copy()) for c in range(2)]
        assert np.array_equal(ctx.encode_pcm_s16(x16, 2), oracle.encode_pcm(back, threads=8, chunk_frames=32))
    finally:
        ctx.set_max_units_per_pass(0)


@pytest.mark.parametrize("units", [4, 6, 64, 250])
def test_pinned_host_buffers(ctx, oracle, units):
    """Pinned caller buffers: the encoder writes its units with a copy kernel straight into the mapped
    host buffer (c1_abi.cu small_copy); pass sizes that give 16-byte aligned and unaligned unit
    offsets, more passes than PCM slots (4) and than unit slots (32)."""
    import torch

    chans = S.cfg3_transients(1.7, seed=91, n_ch=2)
    n = len(chans[0])
    want = oracle.encode_pcm(chans, threads=8, chunk_frames=32)
    ref = oracle.decode_su(want, 2, threads=8, chunk_frames=32)
    n_su = want.shape[0]
    pcm_h = torch.empty((2, n), dtype=torch.float32).pin_memory()
    for c in range(2):
        pcm_h[c].copy_(torch.from_numpy(chans[c]))
    su_h = torch.full((n_su * 212 + 64,), 0xEE, dtype=torch.uint8).pin_memory()
    out_h = torch.empty((2, (n_su // 2) * 512), dtype=torch.float32).pin_memory()
    ctx.set_max_units_per_pass(units)
    try:
        for shift in (0, 4):  # a unit buffer that is only 4-byte aligned takes the cudaMemcpyAsync path
            su_np = su_h.numpy()[shift:shift + n_su * 212]
            got = ctx.encode_pcm_into([pcm_h[0].numpy(), pcm_h[1].numpy()], su_np, None)
            assert got == n_su
            assert np.array_equal(su_np.reshape(-1, 212), want), (units, shift)
            assert np.all(su_h.numpy()[shift + n_su * 212:] == 0xEE)
            out_h.zero_()
            ctx.decode_su_into(su_np, n_su, 2, [out_h[0].numpy(), out_h[1].numpy()])
            for c in range(2):
                assert np.array_equal(bits(out_h[c].numpy()), bits(ref[c])), (units, shift, c)
    finally:
        ctx.set_max_units_per_pass(0)


@pytest.mark.parametrize("units", [6, 64, 4096])
def test_pageable_buffers_take_the_bounce_path(ctx, oracle, units, monkeypatch):
    """Pageable caller arrays are staged through the context's pinned bounce slots (parallel memcpy, one pass
    behind); CARTA1_BOUNCE_MIN_BYTES=0 forces that path at test size.  f32 and int16 PCM, both directions."""
    monkeypatch.setenv("CARTA1_BOUNCE_MIN_BYTES", "0")
    chans = S.cfg3_transients(2.3, seed=17, n_ch=2)
    want = oracle.encode_pcm(chans, threads=8, chunk_frames=32)
    ref = oracle.decode_su(want, 2, threads=8, chunk_frames=32)
    ctx.set_max_units_per_pass(units)
    try:
        su = ctx.encode_pcm(chans)
        assert np.array_equal(su, want)
        pcm = ctx.decode_su(su, 2)
        for x, y in zip(pcm, ref):
            assert np.array_equal(bits(x), bits(y))
        s16 = ctx.decode_su_s16(su, 2).reshape(-1, 2)
        for c in range(2):
            assert np.array_equal(s16[:, c], oracle.pcm_to_int16(ref[c]))
        x16 = np.stack([oracle.pcm_to_int16(c) for c in chans], axis=1)
        back = [oracle.int16_to_pcm(x16[:, c].copy()) for c in range(2)]
        assert np.array_equal(ctx.encode_pcm_s16(x16, 2), oracle.encode_pcm(back, threads=8, chunk_frames=32))
        mono = ctx.encode_pcm([chans[0][:70001]])
        assert np.array_equal(mono, oracle.encode_pcm([chans[0][:70001]]))
        assert np.array_equal(bits(ctx.decode_su(mono, 1)[0]), bits(oracle.decode_su(mono, 1)[0]))
    finally:
        ctx.set_max_units_per_pass(0)


def test_encode_and_decode_calls_in_flight_together(oracle):
    """Two contexts, two host threads: carta1_encode_pcm and carta1_decode_su overlap (the e2e leg of
    bench.py); results equal the one-after-the-other results."""
    import threading

    import torch

    import carta1_b200

    chans = S.cfg2_stereo(20.0, seed=5)
    n = len(chans[0])
    want = oracle.encode_pcm(chans, threads=8, chunk_frames=64)
    ref = oracle.decode_su(want, 2, threads=8, chunk_frames=64)
    n_su = want.shape[0]
    pcm_h = torch.empty((2, n), dtype=torch.float32).pin_memory()
    for c in range(2):
        pcm_h[c].copy_(torch.from_numpy(chans[c]))
    su_a = torch.zeros(n_su * 212, dtype=torch.uint8).pin_memory()
    su_b = torch.from_numpy(want.reshape(-1).copy()).pin_memory()
    out_h = torch.zeros((2, (n_su // 2) * 512), dtype=torch.float32).pin_memory()
    c1, c2 = carta1_b200.Context(0), carta1_b200.Context(0)
    try:
        for c in (c1, c2):
            c.set_max_units_per_pass(128)
        opts = carta1_b200.make_enc_opts()
        for _ in range(3):
            su_a.zero_()
            out_h.zero_()
            errs = []

            def dec():
                try:
                    c2.decode_su_into(su_b.numpy(), n_su, 2, [out_h[0].numpy(), out_h[1].numpy()])
                except Exception as ex:
                    errs.append(ex)

            th = threading.Thread(target=dec)
            th.start()
            got = c1.encode_pcm_into([pcm_h[0].numpy(), pcm_h[1].numpy()], su_a.numpy(), opts)
            th.join()
            assert not errs and got == n_su
            assert np.array_equal(su_a.numpy().reshape(-1, 212), want)
            for c in range(2):
                assert np.array_equal(bits(out_h[c].numpy()), bits(ref[c]))
    finally:
        c1.close()
        c2.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_device_path(ctx, oracle, world):
    """Config 5's shape at test size: (stream, frame-range) shards with halos through the
    device-resident entry points; every shard only sees its own slice (+halo) of the input and
    writes a disjoint slice of the output.  All shards run on this one GPU, one after another."""
    import torch

    from carta1_b200 import sharding

    streams = [S.cfg3_transients(sec, seed=40 + i, n_ch=2) for i, sec in enumerate((0.45, 0.2, 0.33))]
    frames = [oracle.frame_count(len(ch[0])) for ch in streams]
    plan = sharding.plan(frames, world)
    sharding.check_plan(plan, frames)
    su_out = [np.zeros((f * 2, 212), np.uint8) for f in frames]
    pcm_out = [np.zeros((2, f * 512), np.float32) for f in frames]
    wants = [oracle.encode_pcm(ch) for ch in streams]
    for shards in plan:
        for sh in shards:
            ch = streams[sh.stream]
            first, last = sh.pcm_span()
            n = last - first
            host = np.zeros((2, n), np.float32)
            for c in range(2):
                part = ch[c][first:last]
                host[c, :len(part)] = part
            d_pcm = torch.from_numpy(host).cuda()
            d_su = torch.zeros(sh.frames * 2 * 212, dtype=torch.uint8, device="cuda")
            ctx.encode_device(d_pcm.data_ptr(), n, 2, n, sh.enc_halo, sh.frames, None, d_su.data_ptr(), 2, 1, sync=True)
            su_out[sh.stream][sh.begin * 2:sh.end * 2] = d_su.cpu().numpy().reshape(-1, 212)
    for i in range(len(streams)):
        assert np.array_equal(su_out[i], wants[i]), i
    for shards in plan:
        for sh in shards:
            ua, ub = sh.unit_span(2)
            d_su = torch.from_numpy(np.ascontiguousarray(wants[sh.stream][ua:ub]).reshape(-1)).cuda()
            d_pcm = torch.zeros((2, sh.frames * 512), dtype=torch.float32, device="cuda")
            ctx.decode_device(d_su.data_ptr(), 2, 1, ub - ua, 2, sh.dec_halo, sh.frames, d_pcm.data_ptr(), sh.frames * 512, sync=True)
            pcm_out[sh.stream][:, sh.begin * 512:sh.end * 512] = d_pcm.cpu().numpy()
    for i in range(len(streams)):
        ref = oracle.decode_su(wants[i], 2)
        for c in range(2):
            assert np.array_equal(bits(pcm_out[i][c]), bits(ref[c])), (i, c)


@pytest.mark.parametrize("shift,stride_pad", [(0, 0), (1, 0), (0, 3), (3, 1)])
@pytest.mark.parametrize("ragged", [0, 1, 137])
def test_device_rows_any_alignment(ctx, oracle, shift, stride_pad, ragged):
    """K1 stages whole frames of 16-byte aligned rows with one TMA bulk copy each, a row's partial last frame with
    zero-filling cp.async chunks, and rows that are only 4-byte aligned with plain loads (qa_prefetch / qa_fill,
    c1_encode.cu); K7 stores float4 where the output row is aligned and scalars otherwise.  All of them have to
    give the reference's bytes: device rows at every alignment (base shifted by `shift` floats, row stride not a
    multiple of four floats) and lengths that end inside a frame."""
    import torch

    ch = S.cfg3_transients(0.31, seed=77, n_ch=2)
    n = len(ch[0]) - ragged
    ch = [c[:n] for c in ch]
    frames = oracle.frame_count(n)
    want = oracle.encode_pcm(ch)
    stride = frames * 512 + stride_pad
    buf = torch.zeros(shift + 2 * stride + 8, dtype=torch.float32, device="cuda")
    for c in range(2):
        buf[shift + c * stride:shift + c * stride + n] = torch.from_numpy(np.ascontiguousarray(ch[c])).cuda()
    d_su = torch.zeros(frames * 2 * 212, dtype=torch.uint8, device="cuda")
    ctx.encode_device(buf.data_ptr() + 4 * shift, stride, 2, n, 0, frames, None, d_su.data_ptr(), 2, 1, sync=True)
    assert np.array_equal(d_su.cpu().numpy().reshape(-1, 212), want)
    out = torch.full((shift + 2 * stride + 8,), 7.0, dtype=torch.float32, device="cuda")
    ctx.decode_device(d_su.data_ptr(), 2, 1, frames * 2, 2, 0, frames, out.data_ptr() + 4 * shift, stride, sync=True)
    ref = oracle.decode_su(want, 2)
    got = out.cpu().numpy()
    for c in range(2):
        a = shift + c * stride
        assert np.array_equal(bits(got[a:a + frames * 512]), bits(ref[c])), c
    assert np.all(got[:shift] == 7.0) and np.all(got[shift + 2 * stride:] == 7.0)  # nothing written outside the rows
    if stride_pad:
        assert np.all(got[shift + frames * 512:shift + stride] == 7.0)


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("units_per_pass", [0, 24])
def test_sharded_host_path(ctx, oracle, world, units_per_pass):
    """The same shards through the host entry points carta1_encode_pcm_shard / carta1_decode_su_shard (what a
    rank of the multi-GPU bench calls with its slice of the PCM / of the AEA file), with whole-call and with
    multi-pass staging: byte-identical to the whole-stream result."""
    from carta1_b200 import sharding

    streams = [S.cfg3_transients(sec, seed=60 + i, n_ch=2) for i, sec in enumerate((0.5, 0.27))]
    frames = [oracle.frame_count(len(ch[0])) for ch in streams]
    plan = sharding.plan(frames, world)
    sharding.check_plan(plan, frames)
    wants = [oracle.encode_pcm(ch) for ch in streams]
    refs = [oracle.decode_su(w, 2) for w in wants]
    ctx.set_max_units_per_pass(units_per_pass)
    try:
        for shards in plan:
            for sh in shards:
                ch = streams[sh.stream]
                first, last = sh.pcm_span()
                part = [np.ascontiguousarray(ch[c][first:last]) for c in range(2)]  # the last shard is ragged
                su = np.zeros((sh.frames * 2, 212), np.uint8)
                got = ctx.encode_pcm_shard_into(part, sh.enc_halo, su, None)
                assert got == sh.frames * 2
                assert np.array_equal(su, wants[sh.stream][sh.begin * 2:sh.end * 2]), sh
                ua, ub = sh.unit_span(2)
                src = np.ascontiguousarray(wants[sh.stream][ua:ub])
                outs = [np.zeros(sh.frames * 512, np.float32) for _ in range(2)]
                ctx.decode_su_shard_into(src, ub - ua, 2, sh.dec_halo, outs)
                for c in range(2):
                    assert np.array_equal(bits(outs[c]), bits(refs[sh.stream][c][sh.begin * 512:sh.end * 512])), (sh, c)
        # the same through the mirror of the JS layer (index.mjs: encodePcmShard / decodeUnitsShard)
        from carta1_b200 import codec

        sh = plan[-1][-1]
        ch = streams[sh.stream]
        first, last = sh.pcm_span()
        su = codec.encodePcmShard([np.ascontiguousarray(ch[c][first:last]) for c in range(2)], sh.enc_halo, ctx=ctx)
        assert np.array_equal(su, wants[sh.stream][sh.begin * 2:sh.end * 2])
        ua, ub = sh.unit_span(2)
        outs = codec.decodeUnitsShard(wants[sh.stream][ua:ub], 2, sh.dec_halo, ctx=ctx)
        for c in range(2):
            assert np.array_equal(bits(outs[c]), bits(refs[sh.stream][c][sh.begin * 512:sh.end * 512]))
        with pytest.raises(ValueError, match="halo_frames must be 0 or >= 2"):
            ctx.encode_pcm_shard_into([np.zeros(2048, np.float32)], 1, np.zeros((3, 212), np.uint8), None)
    finally:
        ctx.set_max_units_per_pass(0)


def test_arithmetic_shortcuts_selftest(ctx):
    """Exhaustive device check: reciprocal division == IEEE division for every (q, R, SF);
    in-FP64 rounding == cvt.rn.f32.f64 on 3e8 values around rounding ties."""
    assert ctx.selftest() == 0


def test_error_messages(ctx):
    with pytest.raises(TypeError, match="one or two Float32 channels"):
        ctx.encode_pcm([np.zeros(10, np.float32)] * 3)
    with pytest.raises(ValueError, match="Unsupported channel count: 3"):
        ctx.decode_su(np.zeros((3, 212), np.uint8), 3)
