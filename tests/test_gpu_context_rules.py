"""GPU tests of context-level behaviour: the near-threshold counter of the transient decision
(codec/analysis/transient.js:44-55), per-device constant tables shared by contexts, and contexts
used from several threads at once (include/carta1_b200.h conventions)."""
import threading

import numpy as np
import pytest

import signals as S

pytestmark = pytest.mark.gpu


def oracle_scores_and_modes(O, pcm, threshold):
    enc = O.FrameEncoder(O.make_options(threshold=threshold))
    nf = O.frame_count(len(pcm))
    x = np.zeros(nf * 512, np.float32)
    x[:len(pcm)] = pcm
    scores, modes = [], []
    for f in range(nf):
        fr, dbg = enc(x[512 * f:512 * f + 512], debug=True)
        scores.append([dbg.score[b] for b in range(3)])
        modes.append(list(fr.modes))
    return np.array(scores), np.array(modes)


def test_scores_match_oracle_and_threshold_equal_to_a_score_is_counted(oracle):
    """A threshold injected at exactly a known score: the decision is `score > threshold` (strict), the
    close call is counted, and one ulp below the score the decision flips -- on the GPU as in the oracle."""
    import carta1_b200

    O = oracle
    pcm = S.cfg3_transients(0.5, n_ch=1)[0]
    ctx = carta1_b200.Context(0)
    try:
        scores = ctx.debug_transient_scores(pcm)
        want, _ = oracle_scores_and_modes(O, pcm, 1.0)
        assert scores.shape == want.shape
        assert np.array_equal(scores.view(np.uint64), want.view(np.uint64)), "transient scores differ from the oracle"
        # a frame / band whose score is an ordinary positive number
        f, b = np.argwhere((scores > 0.05) & (scores < 5.0))[3]
        s = float(scores[f, b])
        for thr, expect_transient in ((s, False), (float(np.nextafter(s, -np.inf)), True)):
            ctx.near_threshold(reset=True)
            got = ctx.debug_encode_stages(pcm, carta1_b200.make_enc_opts(transient_threshold_low=thr))
            _, omodes = oracle_scores_and_modes(O, pcm, thr)
            assert np.array_equal(got["modes"], omodes)
            assert bool(got["modes"][f, b] != 0) == expect_transient
            c = ctx.near_threshold()
            assert c["decisions"] == 3 * scores.shape[0]
            assert c["within_1e-12"] >= 1 and c["within_1e-9"] >= c["within_1e-12"]
        # far from every score: nothing is close
        ctx.near_threshold(reset=True)
        ctx.debug_encode_stages(pcm, carta1_b200.make_enc_opts(transient_threshold_low=1e6))
        c = ctx.near_threshold()
        assert c == {"decisions": 3 * scores.shape[0], "within_1e-9": 0, "within_1e-12": 0}
        # fixed block modes take no decision
        ctx.near_threshold(reset=True)
        ctx.debug_encode_stages(pcm, carta1_b200.make_enc_opts(fixed_block_modes=[0, 0, 0]))
        assert ctx.near_threshold()["decisions"] == 0
    finally:
        ctx.close()


def test_second_context_with_other_fft_tables_is_refused(oracle):
    """FFT twiddles of stages 0..2 live in per-device constant memory: a second live context on the same
    device must carry the same fft_w, otherwise creation fails instead of silently mixing two tables."""
    import carta1_b200
    from carta1_b200._lib import Carta1Error, default_tables

    O = oracle
    pcm = S.cfg1_stereo(0.2)[0]
    a = carta1_b200.Context(0)
    try:
        t = default_tables()
        t.fft_w[3][0] = float(np.nextafter(t.fft_w[3][0], 0.0))
        with pytest.raises((Carta1Error, ValueError)):
            carta1_b200.Context(0, t)
        # other tables may differ per context: they live in the context's own device memory
        t2 = default_tables()
        t2.window_short[5] = float(np.nextafter(t2.window_short[5], 0.0))
        b = carta1_b200.Context(0, t2)
        try:
            su_a = a.encode_pcm([pcm], carta1_b200.make_enc_opts(fixed_block_modes=[0, 0, 0]))
            su_b = b.encode_pcm([pcm], carta1_b200.make_enc_opts(fixed_block_modes=[0, 0, 0]))
            want_a = O.encode_pcm([pcm], O.make_options(fixed_modes=[0, 0, 0]))
            ot = O.Tables.from_buffer_copy(O.default_tables())
            ot.window_short[5] = t2.window_short[5]
            want_b = O.encode_pcm([pcm], O.make_options(fixed_modes=[0, 0, 0], tables=ot), tables=ot)
            assert np.array_equal(su_a, want_a)
            assert np.array_equal(su_b, want_b)
        finally:
            b.close()
    finally:
        a.close()
    # once no context is alive the device takes new twiddles
    c = carta1_b200.Context(0, t)
    try:
        ot = O.Tables.from_buffer_copy(O.default_tables())
        ot.fft_w[3][0] = t.fft_w[3][0]
        got = c.encode_pcm([pcm])
        want = O.encode_pcm([pcm], O.make_options(tables=ot), tables=ot)
        assert np.array_equal(got, want)
    finally:
        c.close()
    d = carta1_b200.Context(0)  # and back to the defaults
    try:
        assert np.array_equal(d.encode_pcm([pcm]), O.encode_pcm([pcm]))
    finally:
        d.close()


def test_contexts_on_threads_from_cold(oracle):
    """Fresh contexts used from four threads at once, no single-threaded warm-up before them (launch-side
    caches are filled concurrently): every thread's output equals the oracle's."""
    import carta1_b200

    O = oracle
    sigs = [S.cfg1_stereo(0.3, seed=100 + i) for i in range(4)]
    want = [O.encode_pcm(list(s)) for s in sigs]
    want_pcm = [O.decode_su(w, 2) for w in want]
    errs, got = [], [None] * 4

    def work(i):
        try:
            c = carta1_b200.Context(0)
            try:
                su = c.encode_pcm(list(sigs[i]))
                pcm = c.decode_su(su, 2)
                got[i] = (su, pcm)
            finally:
                c.close()
        except Exception as ex:  # surfaced after the joins
            errs.append(ex)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(4):
        assert np.array_equal(got[i][0], want[i])
        for a, b in zip(got[i][1], want_pcm[i]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_one_context_from_several_threads(oracle):
    """One context, used from three threads at once: an encoder handle fed frame by frame, a decoder handle fed
    unit by unit, and whole-buffer calls on the context itself.  They share the context's stream and scratch buffers
    (the encode and the decode path both write `coefs` and `modes`), so the library runs the calls one after the other
    (include/carta1_b200.h conventions); every result has to be the oracle's."""
    import carta1_b200

    O = oracle
    x = np.ascontiguousarray(S.cfg3_transients(0.6, seed=31, n_ch=1)[0])
    frames = O.frame_count(len(x))
    padded = np.zeros(frames * 512, np.float32)
    padded[:len(x)] = x
    want = O.encode_pcm([x])                      # [frames][212]
    want_pcm = O.decode_su(want, 1)[0]
    y = list(S.cfg1_stereo(0.25, seed=32))
    want_y = O.encode_pcm(y)
    ctx = carta1_b200.Context(0)
    errs, out = [], {}

    def enc_thread():
        try:
            enc = carta1_b200.StreamEncoder(ctx, None, 1)
            got = [enc.frames(padded[512 * f:512 * (f + 1)].reshape(1, 1, 512)).reshape(212).copy() for f in range(frames)]
            enc.close()
            out["su"] = np.stack(got)
        except Exception as ex:
            errs.append(ex)

    def dec_thread():
        try:
            dec = carta1_b200.StreamDecoder(ctx, 1)
            got = [dec.frames(np.ascontiguousarray(want[f]).reshape(1, 1, 212)).reshape(512).copy() for f in range(frames)]
            dec.close()
            out["pcm"] = np.concatenate(got)
        except Exception as ex:
            errs.append(ex)

    def whole_thread():
        try:
            out["y"] = [ctx.encode_pcm(y) for _ in range(6)]
        except Exception as ex:
            errs.append(ex)

    th = [threading.Thread(target=f) for f in (enc_thread, dec_thread, whole_thread)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    try:
        assert not errs, errs
        assert np.array_equal(out["su"], want)
        assert np.array_equal(out["pcm"].view(np.uint32), want_pcm[:frames * 512].view(np.uint32))
        for g in out["y"]:
            assert np.array_equal(g, want_y)
    finally:
        ctx.close()
