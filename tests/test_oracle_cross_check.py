"""The C oracle against the second, independent restatement (oracle/oracle_np.py): two transcriptions of the
reference's JavaScript must agree bit for bit on sound units, block modes and decoded PCM.  (The pin against
the reference itself is tests/test_reference_pin.py -- DESIGN.md section 3.)"""
import numpy as np
import pytest

import signals as S
from oracle import oracle_np as N


# float32 stores of values beyond the binary32 range round to infinity, as a Float32Array store does
pytestmark = pytest.mark.filterwarnings("ignore:overflow encountered in cast:RuntimeWarning",
                                        "ignore:invalid value encountered:RuntimeWarning")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _non_finite(rng):
    """+-Infinity, NaN and near-overflow samples: where JavaScript number semantics (Math.max / Math.min with
    NaN, `| 0` of non-finite values, NaN never winning a `>`) are easiest to get wrong in a restatement."""
    x = (0.3 * rng.standard_normal(512 * 8)).astype(np.float32)
    x[700], x[1500], x[2600], x[2601], x[3300] = np.inf, np.nan, -3e38, 3e38, -np.inf
    return x


def cases():
    rng = np.random.default_rng(23)
    t = np.arange(512 * 6) / 44100.0
    return {
        "sine_noise": S.cfg1_stereo(0.25)[0],
        "chirp": S.cfg2_stereo(0.25)[1],
        "clicks": S.cfg3_transients(0.6, seed=3, n_ch=1)[0],
        "white": rng.uniform(-1, 1, 512 * 12).astype(np.float32),
        "loud": (4.0 * rng.standard_normal(512 * 4)).astype(np.float32),
        "tiny": (1e-7 * rng.standard_normal(512 * 4)).astype(np.float32),
        "step": np.concatenate([np.zeros(700, np.float32), 0.8 * np.ones(1348, np.float32)]),
        "ragged_tone": (0.6 * np.sin(2 * np.pi * 3000 * t[:512 * 3 + 77])).astype(np.float32),
        "silence_then_burst": np.concatenate([np.zeros(1024, np.float32), rng.standard_normal(1024).astype(np.float32)]),
        "non_finite": _non_finite(rng),
    }


CASES = cases()
OPTS = [dict(), dict(fixed_modes=[0, 0, 0]), dict(fixed_modes=[2, 2, 3]), dict(bias=2.5), dict(threshold=0.3, bias=0.5)]


@pytest.mark.parametrize("name", sorted(CASES))
def test_encode_and_decode_agree(oracle, name):
    pcm = CASES[name]
    for kw in (OPTS if name in ("clicks", "white", "non_finite") else OPTS[:2]):
        want = oracle.encode_pcm([pcm], oracle.make_options(**kw))
        got, modes = N.encode_mono(pcm, **kw)
        assert got.shape == want.shape, (name, kw)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, "%s %s: sound units differ in frames %s (block modes there: %s)" % (
            name, kw, bad[:8], [modes[i] for i in bad[:8]])
        want_pcm = oracle.decode_su(want, 1)[0]
        got_pcm = N.decode_mono(want)
        assert np.array_equal(bits(got_pcm), bits(want_pcm)), (name, kw)
        if name == "clicks" and not kw:
            assert any(any(m) for m in modes) and not all(all(m) for m in modes), "short and long blocks both occur"


def test_decode_of_arbitrary_units_agrees(oracle):
    """Units that no encoder produced (random bytes with a valid BFU-count field and in-range word lengths
    are not required: every byte pattern deserialises) exercise word lengths and scale factors the
    encoder rarely picks."""
    rng = np.random.default_rng(5)
    units = rng.integers(0, 256, (24, 212), dtype=np.uint8)
    units[:, 209:] = 0
    # keep the declared payload inside the unit so that coefficients.set() stays in range in both
    units[:, 0] &= 0xFC  # block modes stay random: long and short IMDCT paths
    units[:, 1] = 0      # 20 BFUs
    for u in units:
        for i in range(20):
            N.pack_bits(u, 16 + 4 * i, int(rng.integers(0, 7)), 4)
    want = oracle.decode_su(units, 1)[0]
    got = N.decode_mono(units)
    assert np.array_equal(bits(got), bits(want))


def test_stage_functions_agree(oracle):
    rng = np.random.default_rng(9)
    x = rng.standard_normal(512).astype(np.float32)
    d = rng.standard_normal(46).astype(np.float32)
    lo, hi, nd = N.qmf_analysis(x, d)
    lo2, hi2, nd2 = oracle.qmf_analysis(x, d)
    assert np.array_equal(bits(lo), bits(lo2)) and np.array_equal(bits(hi), bits(hi2)) and np.array_equal(bits(nd), bits(nd2))
    y, nd = N.qmf_synthesis(lo, hi, d)
    y2, nd2 = oracle.qmf_synthesis(lo, hi, d)
    assert np.array_equal(bits(y), bits(y2)) and np.array_equal(bits(nd), bits(nd2))
    for n in (16, 64, 128, 256):
        re, im = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
        r1, i1 = re.copy(), im.copy()
        N.fft(r1, i1)
        r2, i2 = oracle.fft(re, im)
        assert np.array_equal(bits(r1), bits(r2)) and np.array_equal(bits(i1), bits(i2)), n
    for n, fwd, inv in ((64, N.mdct64, N.imdct64), (256, N.mdct256, N.imdct256), (512, N.mdct512, N.imdct512)):
        v = rng.standard_normal(n).astype(np.float32)
        assert np.array_equal(bits(fwd.transform(v)), bits(oracle.mdct(v))), n
        c = rng.standard_normal(n // 2).astype(np.float32)
        assert np.array_equal(bits(inv.transform(c)), bits(oracle.imdct(c))), n
    for n in (128, 256):
        cur = N.perform_fft(rng.standard_normal(n).astype(np.float32), n)
        prev = N.perform_fft(rng.standard_normal(n).astype(np.float32), n)
        a, b = N.transient_score(cur, prev), oracle.transient_score(cur, prev)
        assert a == b or abs(a - b) <= 4e-16 * abs(b), (a, b)  # libm (glibc) here, fdlibm there
