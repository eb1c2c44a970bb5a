"""Deterministic test signals.

The first block restates the reference's tests/testSignals.js:2-46 generators; the
second block holds the seeded synthetic inputs of BASELINE.md section 5.
"""
import math

import numpy as np

SR = 44100


def silence(n=512):
    return np.zeros(n, np.float32)


def dc(v=1.0, n=512):
    return np.full(n, v, np.float32)


def sine(freq, sr=SR, n=512):
    i = np.arange(n, dtype=np.float64)
    return np.sin((2 * math.pi * freq * i) / sr).astype(np.float32)


def impulse(pos=0, n=512):
    a = np.zeros(n, np.float32)
    a[pos] = 1.0
    return a


def white_noise(seed=1, n=512):
    a = np.zeros(n, np.float32)
    x = float(seed)
    for i in range(n):
        x = math.sin(x) * 10000
        a[i] = x - math.floor(x)
    return a


def chirp(f0, f1, n=512, sr=SR):
    i = np.arange(n, dtype=np.float64)
    t = i / sr
    phase = 2 * math.pi * (f0 * t + ((f1 - f0) * t * t) / ((2 * n) / sr))
    return np.sin(phase).astype(np.float32)


def step(pos=256, n=512):
    a = np.zeros(n, np.float32)
    a[pos:] = 1.0
    return a


# ---- BASELINE.md section 5 synthetic inputs (seeded) ----
def cfg1_stereo(seconds=10.0, seed=0xCA27A1):
    n = int(round(seconds * SR))
    t = np.arange(n, dtype=np.float64) / SR
    rng = np.random.default_rng(seed)
    left = 0.5 * np.sin(2 * math.pi * 440 * t) + 0.05 * rng.standard_normal(n)
    right = 0.5 * np.sin(2 * math.pi * 880 * t) + 0.05 * rng.standard_normal(n)
    return [left.astype(np.float32), right.astype(np.float32)]


def cfg2_stereo(seconds, seed=0xCA27A2):
    n = int(round(seconds * SR))
    t = np.arange(n, dtype=np.float64) / SR
    rng = np.random.default_rng(seed)
    dur = max(seconds, 1e-9)
    k = (8000.0 - 100.0) / dur
    ch = 0.25 * np.sin(2 * math.pi * (100.0 * t + 0.5 * k * t * t))
    left = 0.4 * np.sin(2 * math.pi * 440 * t) + ch + 0.05 * rng.standard_normal(n)
    right = 0.4 * np.sin(2 * math.pi * 880 * t) + ch + 0.05 * rng.standard_normal(n)
    return [left.astype(np.float32), right.astype(np.float32)]


def cfg3_transients(seconds, seed=0xCA27A3, n_ch=2):
    n = int(round(seconds * SR))
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_ch):
        x = 0.01 * rng.standard_normal(n)
        n_clicks = rng.poisson(4.0 * seconds)
        for _ in range(n_clicks):
            pos = int(rng.integers(0, max(n - 1, 1)))
            dur = int(rng.uniform(0.002, 0.005) * SR)
            m = min(dur, n - pos)
            if m <= 0:
                continue
            burst = rng.standard_normal(m)
            burst = np.convolve(burst, np.ones(3) / 3.0, mode="same")
            env = np.exp(-np.arange(m) / (0.25 * dur))
            x[pos:pos + m] += 0.9 * burst * env / max(np.max(np.abs(burst)), 1e-9)
        out.append(np.clip(x, -1, 1).astype(np.float32))
    return out


def cfg4_mono_streams(n_streams, seconds, seed=0xCA27A4):
    n = int(round(seconds * SR))
    t = np.arange(n, dtype=np.float64) / SR
    out = np.zeros((n_streams, n), np.float32)
    for s in range(n_streams):
        rng = np.random.default_rng(seed + s)
        x = 0.02 * rng.standard_normal(n)
        for _ in range(3):
            f = rng.uniform(80, 9000)
            a = rng.uniform(0.05, 0.3)
            x += a * np.sin(2 * math.pi * f * t + rng.uniform(0, 6.28))
        out[s] = x.astype(np.float32)
    return out
