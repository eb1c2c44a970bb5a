"""oracle_np.py -- a SECOND, independent restatement of carta1's ATRAC1 encode/decode hot path.

TEST INFRASTRUCTURE ONLY (tests/ may import it; the product never does).  It predates the reference pin
(tests/golden/ref, tools/ref_run_qjs.py: the reference's own JavaScript under Qt's QJSEngine), which now pins
oracle/carta1_oracle.c directly; it stays as a cross-check and as the glibc-libm twin of the transient score
(its scores equal the engine's bit for bit).

Why it exists: oracle/carta1_oracle.c is what every GPU parity test compares against, and it was written
by reading the reference.  This module was written separately, again from the reference's JavaScript, in
plain Python + numpy with the reference's own structure (BufferPool state, one function per reference
function, frame objects).  tests/test_oracle_cross_check.py requires the two to agree bit for bit on sound
units, block modes and decoded PCM: a transcription slip in either one shows up as a mismatch.

Semantics kept from JavaScript: numbers are IEEE doubles (Python float); a Float32Array store rounds to
binary32 (numpy float32 assignment); `x | 0` is ToInt32; Math.* are the platform libm.  Slow (pure-Python
loops): a few frames per second, which is what the cross-check needs.
"""
import math

import numpy as np

F32 = np.float32

# ---------------------------------------------------------------- codec/core/constants.js
SPECS_PER_BFU = [8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7, 7, 7, 9, 9, 9, 9,
                 10, 10, 10, 10, 12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20, 20, 20, 20, 20, 20]         # :28-32
BFU_AMOUNTS = [20, 28, 32, 36, 40, 44, 48, 52]                                                       # :35
BFU_BAND_BOUNDARIES = [20, 36, 52]                                                                   # :36
BFU_START_LONG = [0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116, 122, 128, 134,
                  140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226, 236, 246, 256, 268, 280, 292, 304,
                  316, 328, 340, 352, 372, 392, 412, 432, 452, 472, 492]                              # :39-44
BFU_START_SHORT = [0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90, 122, 128, 160,
                   192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182, 214, 246, 256, 288, 320, 352, 384,
                   416, 448, 480, 268, 300, 332, 364, 396, 428, 460, 492]                             # :46-51
WINDOW_SHORT = [math.sin(((i + 0.5) * math.pi) / 64) for i in range(32)]                             # :60-66
QMF_COEFFS = np.array([-0.00001461907, -0.00009205479, -0.000056157569, 0.00030117269, 0.0002422519,
                       -0.00085293897, -0.0005205574, 0.0020340169, 0.00078333891, -0.0042153862,
                       -0.00075614988, 0.0078402944, -0.000061169922, -0.01344162, 0.0024626821, 0.021736089,
                       -0.007801671, -0.034090221, 0.01880949, 0.054326009, -0.043596379, -0.099384367,
                       0.13207909, 0.46424159], F32)                                                  # :74-80
QMF_WINDOW = np.zeros(48, F32)                                                                       # :83-90
for _i in range(24):
    QMF_WINDOW[_i] = float(QMF_COEFFS[_i]) * 2.0
    QMF_WINDOW[47 - _i] = float(QMF_COEFFS[_i]) * 2.0
QMF_EVEN = [float(QMF_WINDOW[2 * i]) for i in range(24)]                                             # :93-99
QMF_ODD = [float(QMF_WINDOW[2 * i + 1]) for i in range(24)]                                          # :101-107
MDCT_BAND_CONFIGS = [(128, 48), (128, 48), (256, 112)]  # (size, windowStart)                        # :115-119
WORD_LENGTH_BITS = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16]                           # :141-143
MAX_WORD_LENGTH_INDEX = 15
SCALE_FACTORS = [math.pow(2.0, i / 3.0 - 21) for i in range(64)]                                     # :144-150
INV_POWER_OF_TWO = [math.pow(2, -b) for b in range(17)]                                              # :153-160
WORD_LENGTH_DELTA_BITS = [WORD_LENGTH_BITS[i + 1] - WORD_LENGTH_BITS[i] for i in range(15)]          # :162-168
DISTORTION_DELTA_FACTORS = [2.0 - INV_POWER_OF_TWO[WORD_LENGTH_BITS[1]]] + [
    INV_POWER_OF_TWO[WORD_LENGTH_BITS[i]] - INV_POWER_OF_TWO[WORD_LENGTH_BITS[i + 1]] for i in range(1, 15)]  # :170-179
FRAME_BITS, FRAME_OVERHEAD_BITS, BITS_PER_BFU_METADATA = 212 * 8, 40, 10


def to_int32(x):
    """ECMAScript ToInt32 of a double (the `| 0` of quantization.js:51)."""
    if x != x or x in (math.inf, -math.inf):
        return 0
    v = int(math.trunc(x)) & 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


# ---------------------------------------------------------------- codec/core/buffers.js
class BufferPool:  # only the state that survives a frame (:31-42, :44-49, :67-79)
    def __init__(self):
        self.qmf_low = np.zeros(46, F32)
        self.qmf_mid = np.zeros(46, F32)
        self.qmf_high = np.zeros(39, F32)
        self.transient = [np.zeros(64, F32), np.zeros(64, F32), np.zeros(128, F32)]
        self.mdct_overlap = [np.zeros(32, F32) for _ in range(3)]
        self.imdct_overlap = [np.zeros(256, F32), np.zeros(256, F32), np.zeros(512, F32)]


# ---------------------------------------------------------------- codec/transforms/qmf.js
def qmf_analysis(inp, delay):  # :19-50
    n_out = len(inp) >> 1
    work = np.concatenate([delay, inp]).astype(F32)
    w = [float(v) for v in work]
    low, high = np.zeros(n_out, F32), np.zeros(n_out, F32)
    for i in range(n_out):
        even = odd = 0.0
        off = 2 * i
        for j in range(24):
            even += w[off + 47 - 2 * j] * QMF_EVEN[j]
            odd += w[off + 46 - 2 * j] * QMF_ODD[j]
        low[i] = even + odd
        high[i] = even - odd
    return low, high, work[-46:].copy()


def qmf_synthesis(low, high, delay):  # :60-105
    n = len(low)
    work = np.zeros(46 + 2 * n, F32)
    work[:46] = delay
    for i in range(n):
        lo, hi = float(low[i]), float(high[i])
        work[46 + 2 * i] = 0.5 * (lo + hi)
        work[46 + 2 * i + 1] = 0.5 * (lo - hi)
    w = [float(v) for v in work]
    out = np.zeros(2 * n, F32)
    for i in range(n):
        s0 = s1 = 0.0
        for j in range(24):
            idx = 2 * i + 2 * j
            s0 += w[idx] * QMF_EVEN[j]
            s1 += w[idx + 1] * QMF_ODD[j]
        out[2 * i] = s1
        out[2 * i + 1] = s0
    return out, work[-46:].copy()


# ---------------------------------------------------------------- codec/transforms/fft.js
def fft(real, imag):  # :14-68, in place on float32 arrays
    size = len(real)
    if size == 1:
        return
    bits = int(round(math.log2(size)))
    for i in range(size):
        rev, t = 0, i
        for _ in range(bits):
            rev = (rev << 1) | (t & 1)
            t >>= 1
        if rev > i:
            real[i], real[rev] = real[rev], real[i]
            imag[i], imag[rev] = imag[rev], imag[i]
    stride = 2
    while stride <= size:
        half = stride >> 1
        angle = (-2 * math.pi) / stride
        w_re, w_im = math.cos(angle), math.sin(angle)
        for start in range(0, size, stride):
            t_re, t_im = 1.0, 0.0
            for k in range(half):
                e, o = start + k, start + k + half
                e_re, e_im, o_re, o_im = float(real[e]), float(imag[e]), float(real[o]), float(imag[o])
                x_re = o_re * t_re - o_im * t_im
                x_im = o_re * t_im + o_im * t_re
                real[e] = e_re + x_re
                imag[e] = e_im + x_im
                real[o] = e_re - x_re
                imag[o] = e_im - x_im
                nxt = t_re * w_re - t_im * w_im
                t_im = t_re * w_im + t_im * w_re
                t_re = nxt
        stride <<= 1


# ---------------------------------------------------------------- codec/transforms/mdct.js
class _Base:  # :17-37
    def __init__(self, size, scale):
        self.size, self.half, self.quarter = size, size >> 1, size >> 2
        self.fft_size = self.half >> 1
        alpha = (2.0 * math.pi) / (8.0 * size)
        omega = (2.0 * math.pi) / size
        root = math.sqrt(scale / size)
        self.tab = [0.0] * self.half
        for i in range(self.quarter):
            a = omega * i + alpha
            self.tab[2 * i] = root * math.cos(a)
            self.tab[2 * i + 1] = root * math.sin(a)


class MDCT(_Base):
    def transform(self, inp):  # :54-122
        n4, n34 = self.quarter, 3 * self.quarter
        real, imag = np.zeros(self.fft_size, F32), np.zeros(self.fft_size, F32)
        x = [float(v) for v in inp]
        for i in range(0, n4, 2):
            r = x[n34 - 1 - i] + x[n34 + i]
            m = x[n4 + i] - x[n4 - 1 - i]
            c, s = self.tab[i], self.tab[i + 1]
            real[i >> 1] = r * c + m * s
            imag[i >> 1] = m * c - r * s
        for i in range(n4, self.half, 2):
            r = x[n34 - 1 - i] - x[i - n4]
            m = x[n4 + i] + x[5 * n4 - 1 - i]
            c, s = self.tab[i], self.tab[i + 1]
            real[i >> 1] = r * c + m * s
            imag[i >> 1] = m * c - r * s
        fft(real, imag)
        out = np.zeros(self.half, F32)
        for i in range(self.fft_size):
            c, s = self.tab[2 * i], self.tab[2 * i + 1]
            re, im = float(real[i]), float(imag[i])
            out[2 * i] = -re * c - im * s
            out[self.half - 1 - 2 * i] = -re * s + im * c
        return out


class IMDCT(_Base):
    def transform(self, inp):  # :139-205
        n4, n34 = self.quarter, 3 * self.quarter
        real, imag = np.zeros(self.fft_size, F32), np.zeros(self.fft_size, F32)
        x = [float(v) for v in inp]
        for i in range(self.fft_size):
            i2 = 2 * i
            r, m = -x[i2], -x[self.half - 1 - i2]
            c, s = self.tab[i2], self.tab[i2 + 1]
            real[i] = m * s + r * c
            imag[i] = m * c - r * s
        fft(real, imag)
        out = np.zeros(self.size, F32)
        hf = self.fft_size // 2
        for i in range(hf):
            i2 = 2 * i
            c, s = self.tab[i2], self.tab[i2 + 1]
            re, im = float(real[i]), float(imag[i])
            r1 = re * c + im * s
            i1 = re * s - im * c
            out[n34 - 1 - i2] = r1
            out[n34 + i2] = r1
            out[n4 + i2] = i1
            out[n4 - 1 - i2] = -i1
        for i in range(hf, self.fft_size):
            idx = (i - hf) * 2 + n4
            i2 = 2 * i
            c, s = self.tab[i2], self.tab[i2 + 1]
            re, im = float(real[i]), float(imag[i])
            r1 = re * c + im * s
            i1 = re * s - im * c
            out[n34 - 1 - idx] = r1
            out[idx - n4] = -r1
            out[n4 + idx] = i1
            out[5 * n4 - 1 - idx] = i1
        return out


mdct64, mdct256, mdct512 = MDCT(64, 0.5), MDCT(256, 0.5), MDCT(512, 1.0)            # :215-217
imdct64, imdct256, imdct512 = IMDCT(64, 64 * 8), IMDCT(256, 256 * 8), IMDCT(512, 512 * 4)  # :219-221


def overlap_add(prev, curr, window):  # :230-245
    size = len(prev)
    out = np.zeros(2 * size, F32)
    for i in range(size):
        w1, w2 = window[i], window[2 * size - 1 - i]
        p, c = float(prev[i]), float(curr[size - 1 - i])
        out[i] = p * w2 - c * w1
        out[2 * size - 1 - i] = p * w1 + c * w2
    return out


# ---------------------------------------------------------------- codec/analysis/transient.js
def perform_fft(samples, fft_size):  # :17-35
    real, imag = np.zeros(fft_size, F32), np.zeros(fft_size, F32)
    n = min(len(samples), fft_size)
    real[:n] = samples[:n]
    fft(real, imag)
    mag = np.zeros(fft_size // 2, F32)
    for i in range(fft_size // 2):
        re, im = float(real[i]), float(imag[i])
        mag[i] = math.sqrt(re * re + im * im)
    return mag


def _flux(cur, prev):  # :92-112
    flux = energy = 0.0
    for i in range(len(cur)):
        c, p = abs(float(cur[i])), abs(float(prev[i]))
        d = c - p
        if d > 0:
            flux += d
        energy += c * c
    norm = math.sqrt(energy)
    if norm == 0 or norm != norm:  # `|| 1e-6`: 0 and NaN are falsy
        norm = 1e-6
    return flux / norm


def _flatness(co):  # :116-141
    eps = 1e-10
    s_log = s_lin = 0.0
    valid = 0
    for v in co:
        m = abs(float(v))
        if m > eps:
            s_log += math.log(m)
            s_lin += m
            valid += 1
    if valid == 0:
        return 0.0
    geo = math.exp(s_log / valid)
    ari = s_lin / valid
    return geo / ari if ari > eps else 0.0


def _hf_ratio(co):  # :145-163
    mid = len(co) // 2
    lo = hi = 0.0
    for i in range(mid):
        lo += float(co[i]) * float(co[i])
    for i in range(mid, len(co)):
        hi += float(co[i]) * float(co[i])
    tot = lo + hi
    return hi / tot if tot > 0 else 0.0


def _js_max(a, b):
    if a != a or b != b:
        return math.nan
    return a if a > b else b


def _energy_change(cur, prev):  # :167-189
    ce = pe = 0.0
    for i in range(len(cur)):
        ce += float(cur[i]) * float(cur[i])
        pe += float(prev[i]) * float(prev[i])
    ce, pe = _js_max(ce, 1e-10), _js_max(pe, 1e-10)
    db = 10 * math.log10(ce / pe)
    return _js_max(0.0, db)


def transient_score(cur, prev):  # :57-88, :197-226
    flux = _flux(cur, prev)
    flat = abs(_flatness(cur) - _flatness(prev))
    hf = abs(_hf_ratio(cur) - _hf_ratio(prev))
    en = _energy_change(cur, prev)
    e_c = en / 30
    if e_c != e_c:
        e_c = math.nan
    elif e_c > 1:
        e_c = 1
    return (flux + math.sqrt(flat) + math.log1p(hf * 10) / math.log1p(10) + e_c) / 4


# ---------------------------------------------------------------- codec/coding/bitallocation.js
def find_scale_factor(co, length):  # :290-299
    mx = 0.0
    for i in range(length):
        a = abs(float(co[i]))
        if a > mx:
            mx = a
    if mx == 0:
        return 0
    if mx == math.inf:
        return 63
    idx = math.ceil(3 * (math.log2(mx) + 21))
    return max(0, min(63, idx))


def _sift_down(hi, hp, start, size):  # :314-341
    i = start
    iv, pv = hi[i], hp[i]
    while True:
        left = 2 * i + 1
        right = left + 1
        mi, mp = i, pv
        if left < size and hp[left] > mp:
            mi, mp = left, hp[left]
        if right < size and hp[right] > mp:
            mi = right
        if mi == i:
            break
        hi[i], hp[i] = hi[mi], hp[mi]
        i = mi
    hi[i], hp[i] = iv, pv


def _distribute(active, sizes, remaining, bsf, sfi):  # :203-281
    wl = [0] * active
    hi, hp = [0] * active, np.zeros(active, F32)
    n = 0
    for b in range(active):
        if sizes[b] == 0 or sfi[b] == 0:
            continue
        hi[n] = b
        hp[n] = (bsf[sfi[b]] * DISTORTION_DELTA_FACTORS[0]) / WORD_LENGTH_DELTA_BITS[0]
        n += 1
    if n == 0:
        return wl
    for i in range((n >> 1) - 1, -1, -1):
        _sift_down(hi, hp, i, n)
    while remaining > 0 and n > 0:
        b = hi[0]
        cur = wl[b]
        cost = WORD_LENGTH_DELTA_BITS[cur] * sizes[b]
        if cost > remaining or cost <= 0:
            hi[0], hp[0] = hi[n - 1], hp[n - 1]
            n -= 1
            if n > 0:
                _sift_down(hi, hp, 0, n)
            continue
        remaining -= cost
        nxt = cur + 1
        wl[b] = nxt
        if nxt < MAX_WORD_LENGTH_INDEX and WORD_LENGTH_DELTA_BITS[nxt] > 0:
            hp[0] = (bsf[sfi[b]] * DISTORTION_DELTA_FACTORS[nxt]) / WORD_LENGTH_DELTA_BITS[nxt]
            _sift_down(hi, hp, 0, n)
        else:
            hi[0], hp[0] = hi[n - 1], hp[n - 1]
            n -= 1
            if n > 0:
                _sift_down(hi, hp, 0, n)
    return wl


def _total_distortion(active, max_count, sizes, wl, sfi, bsf, zero_bit):  # :157-190
    total = 0.0
    for i in range(active):
        bits = WORD_LENGTH_BITS[wl[i]]
        if bits == 0:
            total += float(zero_bit[i])
            continue
        if sfi[i] == 0:
            continue
        total += bsf[sfi[i]] * INV_POWER_OF_TWO[bits] * sizes[i]
    for i in range(active, max_count):
        total += float(zero_bit[i])
    return total


def allocate_bits(bfu_data, sizes, max_count, bias):  # :74-155
    bsf = list(SCALE_FACTORS) if bias == 1 else [math.pow(SCALE_FACTORS[i], bias) for i in range(64)]  # :52-58
    sfi = [0] * max_count
    zero_bit = np.zeros(max_count, F32)
    for i in range(max_count):
        if sizes[i] == 0:
            continue
        sfi[i] = find_scale_factor(bfu_data[i], sizes[i])
        if sfi[i] > 0:
            zero_bit[i] = bsf[sfi[i]] * 2.0 * sizes[i]
    best, min_total = None, math.inf
    for cand in BFU_AMOUNTS:
        if cand > max_count:
            continue
        avail = FRAME_BITS - FRAME_OVERHEAD_BITS - cand * BITS_PER_BFU_METADATA
        if avail < 0:
            continue
        wl = _distribute(cand, sizes, avail, bsf, sfi)
        total = _total_distortion(cand, max_count, sizes, wl, sfi, bsf, zero_bit)
        if total < min_total:
            min_total, best = total, (cand, wl, sfi)
    if best is None:
        return BFU_AMOUNTS[0], [0] * BFU_AMOUNTS[0], [0] * 52
    return best


# ---------------------------------------------------------------- codec/coding/quantization.js
def quantize(co, sfi, bits):  # :34-56
    if bits == 0 or sfi == 0:
        return [0] * len(co)
    rng = (1 << (bits - 1)) - 1
    norm = rng / SCALE_FACTORS[sfi]
    out = []
    for v in co:
        x = float(v) * norm
        y = to_int32(x + (0.5 if x >= 0 else -0.5))
        out.append(rng if y > rng else -rng if y < -rng else y)
    return out


def dequantize(q, sfi, bits):  # :65-78
    out = np.zeros(len(q), F32)
    if bits == 0 or sfi == 0:
        return out
    rng = (1 << (bits - 1)) - 1
    sf = SCALE_FACTORS[sfi]
    for i, v in enumerate(q):
        out[i] = (v * sf) / rng
    return out


def group_into_bfus(co, modes):  # :106-149
    data, sizes = [], []
    idx = bfu = 0
    for band in range(3):
        band_start, band_size = idx, 256 if band == 2 else 128
        band_end = BFU_BAND_BOUNDARIES[band] if band < 2 else len(SPECS_PER_BFU)
        starts = BFU_START_LONG if modes[band] == 0 else BFU_START_SHORT
        while bfu < band_end:
            size = SPECS_PER_BFU[bfu]
            sp = starts[bfu] - band_start
            ep = sp + size
            blk = np.zeros(size, F32)
            if sp >= 0 and ep <= band_size:
                blk[:] = co[band_start + sp:band_start + ep]
            elif sp < band_size and ep > 0:
                a, b = max(0, sp), min(band_size, ep)
                d = max(0, -sp)
                blk[d:d + (b - a)] = co[band_start + a:band_start + b]
            data.append(blk)
            sizes.append(size)
            bfu += 1
        idx += band_size
    return data, sizes, bfu


# ---------------------------------------------------------------- codec/io/bitstream.js, serialization.js
def pack_bits(buf, pos, value, count):  # bitstream.js:15-39
    if count == 0:
        return
    byte, off = pos // 8, pos % 8
    value &= (1 << count) - 1
    written = 0
    while written < count and byte < len(buf):
        avail = 8 - off
        n = min(count - written, avail)
        shift = count - written - n
        bits = (value >> shift) & ((1 << n) - 1)
        mask = ((1 << n) - 1) << (avail - n)
        buf[byte] = (int(buf[byte]) & ~mask & 0xFF) | (bits << (avail - n))
        written += n
        byte += 1
        off = 0


def unpack_bits(buf, pos, count):  # bitstream.js:49-68
    if count == 0:
        return 0
    byte, off = pos // 8, pos % 8
    value = 0
    read = 0
    while read < count and byte < len(buf):
        avail = 8 - off
        n = min(count - read, avail)
        bits = (int(buf[byte]) >> (avail - n)) & ((1 << n) - 1)
        value = (value << n) | bits
        read += n
        byte += 1
        off = 0
    return value


def unpack_signed_bits(buf, pos, count):  # bitstream.js:77-81
    v = unpack_bits(buf, pos, count)
    sign = 1 << (count - 1)
    return v - (1 << count) if v >= sign else v


def serialize_frame(fr):  # serialization.js:41-98
    buf = np.zeros(212, np.uint8)
    idx = BFU_AMOUNTS.index(fr["nBfu"])
    m = fr["blockModes"]
    header = (((2 - m[0]) << 14) | ((2 - m[1]) << 12) | ((3 - m[2]) << 10) | (idx << 5)) & 0xFFFF
    buf[0], buf[1] = header >> 8, header & 0xFF
    pos = 16
    for i in range(fr["nBfu"]):
        pack_bits(buf, pos, fr["wordLengthIndices"][i], 4)
        pos += 4
    for i in range(fr["nBfu"]):
        pack_bits(buf, pos, fr["scaleFactorIndices"][i], 6)
        pos += 6
    for i in range(fr["nBfu"]):
        bits = WORD_LENGTH_BITS[fr["wordLengthIndices"][i]]
        if bits > 0:
            for c in fr["quantizedCoefficients"][i]:
                pack_bits(buf, pos, c + (1 << bits) if c < 0 else c, bits)
                pos += bits
    buf[209] = buf[210] = buf[211] = 0
    return buf


def deserialize_frame(buf):  # serialization.js:111-176
    header = (int(buf[0]) << 8) | int(buf[1])
    modes = [2 - ((header >> 14) & 3), 2 - ((header >> 12) & 3), 3 - ((header >> 10) & 3)]
    n = BFU_AMOUNTS[(header >> 5) & 7]
    pos = 16
    wl, sf = [], []
    for _ in range(n):
        wl.append(unpack_bits(buf, pos, 4))
        pos += 4
    for _ in range(n):
        sf.append(unpack_bits(buf, pos, 6))
        pos += 6
    q = []
    for i in range(n):
        bits = WORD_LENGTH_BITS[wl[i]]
        co = [0] * SPECS_PER_BFU[i]
        if bits > 0:
            for j in range(SPECS_PER_BFU[i]):
                co[j] = unpack_signed_bits(buf, pos, bits)
                pos += bits
        q.append(co)
    return {"nBfu": n, "scaleFactorIndices": sf, "wordLengthIndices": wl, "quantizedCoefficients": q, "blockModes": modes}


# ---------------------------------------------------------------- codec/pipeline/encoder.js
def _tail_window(samples, overlap, block_size):  # :309-316
    t0 = block_size - 32
    for i in range(32):
        v = float(samples[t0 + i])
        overlap[i] = WINDOW_SHORT[i] * v
        samples[t0 + i] = v * WINDOW_SHORT[31 - i]


def make_encoder(threshold=1.0, bias=1.0, fixed_modes=None):
    """encode(options, bufferPool), encoder.js:438-450: returns the per-frame closure."""
    pool = BufferPool()

    def encode_frame(pcm):
        pcm = np.asarray(pcm, F32)
        # qmfAnalysisStage :52-93
        s1_lo, s1_hi, pool.qmf_low = qmf_analysis(pcm, pool.qmf_low)
        lo, mid, pool.qmf_mid = qmf_analysis(s1_lo, pool.qmf_mid)
        delayed = np.concatenate([pool.qmf_high, s1_hi]).astype(F32)
        high = delayed[:len(s1_hi)].copy()
        pool.qmf_high = delayed[len(s1_hi):].copy()
        bands = [lo, mid, high]
        # blockSelectorStage :104-146
        if fixed_modes is not None:
            modes = list(fixed_modes)
        else:
            modes = []
            for b, size in enumerate((128, 128, 256)):
                co = perform_fft(bands[b], size)
                score = transient_score(co, pool.transient[b])
                pool.transient[b] = co
                modes.append((1 if score > threshold else 0) * max(b + 1, 2))
        # mdctStage :170-349
        coefs = np.zeros(512, F32)
        for b in range(3):
            size, ws = MDCT_BAND_CONFIGS[b]
            samples, ov = bands[b], pool.mdct_overlap[b]
            if modes[b] == 0:
                n = 512 if b == 2 else 256
                inp = np.zeros(n, F32)
                inp[ws:ws + 32] = ov
                _tail_window(samples, ov, size)
                inp[ws + 32:ws + 32 + size] = samples
                spec = (mdct512 if b == 2 else mdct256).transform(inp)
                if b > 0:
                    spec = spec[::-1].copy()
            else:
                spec = np.zeros(size, F32)
                for blk in range(size // 32):
                    seg = samples[32 * blk:32 * blk + 32]  # a view: the windowing writes through, as subarray does
                    inp = np.zeros(64, F32)
                    inp[:32] = ov
                    _tail_window(seg, ov, 32)
                    inp[32:] = seg
                    sp = mdct64.transform(inp)
                    if b > 0:
                        sp = sp[::-1].copy()
                    spec[32 * blk:32 * blk + 32] = sp
            off = 0 if b == 0 else 128 if b == 1 else 256
            coefs[off:off + size] = spec
        # quantizationStage :361-403
        data, sizes, count = group_into_bfus(coefs, modes)
        n_bfu, wl, sfi = allocate_bits(data, sizes, count, bias)
        q = [quantize(data[b][:sizes[b]], sfi[b], WORD_LENGTH_BITS[wl[b]]) for b in range(n_bfu)]
        return {"nBfu": n_bfu, "scaleFactorIndices": list(sfi[:n_bfu]), "wordLengthIndices": list(wl[:n_bfu]),
                "quantizedCoefficients": q, "blockModes": modes}

    return encode_frame


# ---------------------------------------------------------------- codec/pipeline/decoder.js
def make_decoder():
    """decode(bufferPool), decoder.js:408-411: returns the per-frame closure."""
    pool = BufferPool()

    def decode_frame(fr):
        # dequantizationStage :52-104
        coefs = np.zeros(512, F32)
        modes = fr["blockModes"]
        for b in range(fr["nBfu"]):
            bits = WORD_LENGTH_BITS[fr["wordLengthIndices"][b]]
            band = 0 if b < 20 else 1 if b < 36 else 2
            pos = BFU_START_LONG[b] if modes[band] == 0 else BFU_START_SHORT[b]
            if bits > 0:
                d = dequantize(fr["quantizedCoefficients"][b], fr["scaleFactorIndices"][b], bits)
                coefs[pos:pos + len(d)] = d
        # imdctStage :116-330
        bands = []
        for b in range(3):
            size, _ = MDCT_BAND_CONFIGS[b]
            off = 0 if b == 0 else 128 if b == 1 else 256
            co, ov = coefs[off:off + size], pool.imdct_overlap[b]
            inv_buf = np.zeros(512, F32)
            prev = ov[2 * size - 16:2 * size].copy()
            if modes[b] == 0:
                spec = co[::-1].copy() if b > 0 else co
                inv = (imdct512 if b == 2 else imdct256).transform(spec)
                st = len(inv) // 4
                inv_buf[:size] = inv[st:st + size]
                ov[0:32] = overlap_add(prev, inv_buf[:16], WINDOW_SHORT)
                n_copy = 240 if b == 2 else 112
                ov[32:32 + n_copy] = inv_buf[16:16 + n_copy]
            else:
                start = 0
                for blk in range(size // 32):
                    spec = co[32 * blk:32 * blk + 32].copy()
                    if b > 0:
                        spec = spec[::-1].copy()
                    inv = imdct64.transform(spec)
                    st = len(inv) // 4
                    inv_buf[start:start + 32] = inv[st:st + 32]
                    ov[start:start + 32] = overlap_add(prev, inv_buf[start:start + 16], WINDOW_SHORT)
                    prev = inv_buf[start + 16:start + 32].copy()
                    start += 32
            ov[2 * size - 16:2 * size] = inv_buf[size - 16:size]
            bands.append(ov[:size].copy())
        # qmfSynthesisStage :360-388
        n_hi = 2 * len(bands[0])
        delayed = np.concatenate([pool.qmf_high, bands[2]]).astype(F32)
        high = delayed[:n_hi].copy()
        pool.qmf_high = delayed[n_hi:].copy()
        s2, pool.qmf_mid = qmf_synthesis(bands[0], bands[1], pool.qmf_mid)
        out, pool.qmf_low = qmf_synthesis(s2, high, pool.qmf_low)
        return out

    return decode_frame


# ---------------------------------------------------------------- whole signals (processor.js:246-279, 317-339)
def encode_mono(pcm, **opts):
    """Frames of 512 (zero padded), one closure: [n_frames, 212] sound units and the block modes."""
    pcm = np.asarray(pcm, F32)
    n_frames = (len(pcm) + 511) // 512
    padded = np.zeros(n_frames * 512, F32)
    padded[:len(pcm)] = pcm
    enc = make_encoder(**opts)
    units, modes = np.zeros((n_frames, 212), np.uint8), []
    for f in range(n_frames):
        fr = enc(padded[512 * f:512 * f + 512].copy())
        units[f] = serialize_frame(fr)
        modes.append(fr["blockModes"])
    return units, modes


def decode_mono(units):
    dec = make_decoder()
    return np.concatenate([dec(deserialize_frame(u)) for u in units]) if len(units) else np.zeros(0, F32)
