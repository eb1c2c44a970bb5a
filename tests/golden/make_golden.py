#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle.

PROVENANCE: these vectors come from oracle/ (the C restatement), NOT from carta1 itself.  They pin the
oracle and the CUDA path against drift.  Their inputs are also what the reference itself is run on
(--export-ref-inputs, then tools/ref_run_qjs.py under Qt's QJSEngine in the build image, or
tools/ref_dump.mjs under Node): tests/golden/ref holds the reference's output, which the `su` and
`pcm_out` stored here equal byte for byte (tests/test_reference_pin.py).

    python tests/golden/make_golden.py                       # rewrites the fixtures in place
    python tests/golden/make_golden.py --export-ref-inputs   # writes tests/golden/ref/inputs/ for tools/ref_dump.mjs

The second form does not touch the fixtures: it unpacks their `pcm_s16` inputs as raw little-endian int16
files plus a cases.json (channels, options), the form tools/ref_run_qjs.py / tools/ref_dump.mjs feed to the real carta1;
tests/test_reference_pin.py then compares oracle and CUDA path with what the reference wrote.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import signals as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (channels as float32, option kwargs)
    "cfg1_sine_noise_auto": (lambda: S.cfg1_stereo(0.25), dict()),
    "cfg2_chirp_fixed_long": (lambda: S.cfg2_stereo(0.25), dict(fixed_modes=[0, 0, 0])),
    "cfg3_transients_auto": (lambda: S.cfg3_transients(0.3, n_ch=2), dict()),
    "cfg3_transients_bias2_thr03": (lambda: S.cfg3_transients(0.2, n_ch=1), dict(bias=2.0, threshold=0.3)),
    "mono_ragged_fixed_short": (lambda: [S.cfg4_mono_streams(1, 0.2)[0][:512 * 11 + 77]], dict(fixed_modes=[2, 2, 3])),
}


def export_ref_inputs():
    import glob
    import json

    out = os.path.join(HERE, "ref", "inputs")
    os.makedirs(out, exist_ok=True)
    cases = []
    for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
        z = np.load(path)
        name = os.path.basename(path)[:-4]
        s16 = np.ascontiguousarray(z["pcm_s16"], "<i2")  # [samples][channels]: interleaved
        s16.tofile(os.path.join(out, name + ".s16"))
        fixed = None if z["fixed_modes"][0] < 0 else [int(v) for v in z["fixed_modes"]]
        cases.append(dict(name=name, channels=int(s16.shape[1]), samples=int(s16.shape[0]),
                          threshold=float(z["threshold"]), bias=float(z["bias"]), fixed_modes=fixed))
    with open(os.path.join(out, "cases.json"), "w") as f:
        json.dump(cases, f, indent=1)
    print("wrote %d cases to %s" % (len(cases), out))


def main():
    if "--export-ref-inputs" in sys.argv[1:]:
        export_ref_inputs()
        return
    O.build()
    for name, (make, kw) in CASES.items():
        chans = make()
        # store the input as int16 (compact, and the CLI's own ingest format): the float input
        # of the case is exactly int16_to_pcm(pcm_s16)
        s16 = np.stack([O.pcm_to_int16(np.clip(c, -1, 1)) for c in chans], axis=1)
        chans = [O.int16_to_pcm(s16[:, c].copy()) for c in range(s16.shape[1])]
        opts = O.make_options(**kw)
        su = O.encode_pcm(chans, opts)
        pcm = np.stack(O.decode_su(su, len(chans)))
        out16 = np.stack([O.pcm_to_int16(p) for p in pcm], axis=1)
        modes = np.array([list(O.deserialize_frame(u).modes) for u in su], np.int8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pcm_s16=s16, su=su, pcm_out=pcm, pcm_out_s16=out16,
                            modes=modes, threshold=kw.get("threshold", 1.0), bias=kw.get("bias", 1.0),
                            fixed_modes=np.array(kw.get("fixed_modes") or [-1, -1, -1], np.int32))
        print("%-32s %d ch, %d units, short-block units: %d" % (name, len(chans), len(su), int((modes != 0).any(axis=1).sum())))


if __name__ == "__main__":
    main()
