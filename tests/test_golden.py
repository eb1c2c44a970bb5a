"""Golden vectors (tests/golden/*.npz, generated from the oracle by make_golden.py; the same
inputs run through the reference itself are tests/golden/ref, see test_reference_pin.py): the oracle must keep reproducing them on CPU, the
CUDA path must reproduce them on the GPU."""
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load(path):
    z = np.load(path)
    fixed = None if z["fixed_modes"][0] < 0 else [int(v) for v in z["fixed_modes"]]
    return z, dict(threshold=float(z["threshold"]), bias=float(z["bias"]), fixed_modes=fixed)


def test_fixtures_exist():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_golden(oracle, path):
    z, kw = load(path)
    chans = [oracle.int16_to_pcm(z["pcm_s16"][:, c].copy()) for c in range(z["pcm_s16"].shape[1])]
    su = oracle.encode_pcm(chans, oracle.make_options(**kw))
    assert np.array_equal(su, z["su"])
    pcm = np.stack(oracle.decode_su(su, len(chans)))
    assert np.array_equal(pcm.view(np.uint32), z["pcm_out"].view(np.uint32))
    assert np.array_equal(np.stack([oracle.pcm_to_int16(p) for p in pcm], axis=1), z["pcm_out_s16"])
    assert np.array_equal(np.array([list(oracle.deserialize_frame(u).modes) for u in su]), z["modes"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_gpu_reproduces_golden(path):
    import carta1_b200

    z, kw = load(path)
    ctx = carta1_b200.Context(0)
    n_ch = z["pcm_s16"].shape[1]
    opts = carta1_b200.make_enc_opts(kw["threshold"], kw["bias"], kw["fixed_modes"])
    su = ctx.encode_pcm_s16(z["pcm_s16"], n_ch, opts)  # WAV int16 ingest fused into the QMF kernel
    assert np.array_equal(su, z["su"])
    pcm = np.stack(ctx.decode_su(z["su"], n_ch))
    assert np.array_equal(pcm.view(np.uint32), z["pcm_out"].view(np.uint32))
    assert np.array_equal(ctx.decode_su_s16(z["su"], n_ch).reshape(-1, n_ch), z["pcm_out_s16"])
    ctx.close()
