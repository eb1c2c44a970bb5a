/*
 * carta1_oracle.c -- CPU ORACLE (test infrastructure, see carta1_oracle.h).
 *
 * Plain-C restatement of the ATRAC1 hot path of aynik/carta1 (JavaScript).  Every
 * function cites the reference lines it follows (paths relative to /root/reference/).
 * "f32(...)" in comments marks a Float32Array store == round-to-nearest-even to
 * binary32; everything else is binary64, one rounding per operator, JS source order.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 */
#define _GNU_SOURCE
#include "carta1_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------
 * Format constants (codec/core/constants.js)
 * ---------------------------------------------------------------------------------- */
#define FRAME_BITS (C1O_SU_BYTES * 8)     /* constants.js:20 */
#define FRAME_OVERHEAD_BITS 40            /* constants.js:21 */
#define BITS_PER_BFU_METADATA 10          /* constants.js:27 */
#define MAX_WL_INDEX 15                   /* constants.js:140 */

static const int SPECS_PER_BFU[52] = { /* constants.js:29-33 */
    8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7,
    7, 7, 9, 9, 9, 9, 10, 10, 10, 10, 12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20,
    20, 20, 20, 20, 20};
static const int BFU_AMOUNTS[8] = {20, 28, 32, 36, 40, 44, 48, 52}; /* constants.js:36 */
static const int BFU_START_LONG[52] = {                             /* constants.js:40-45 */
    0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116,
    122, 128, 134, 140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226,
    236, 246, 256, 268, 280, 292, 304, 316, 328, 340, 352, 372, 392, 412, 432,
    452, 472, 492};
static const int BFU_START_SHORT[52] = { /* constants.js:47-52 */
    0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90,
    122, 128, 160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182,
    214, 246, 256, 288, 320, 352, 384, 416, 448, 480, 268, 300, 332, 364, 396,
    428, 460, 492};
static const int WL_BITS[16] = {0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}; /* :141-143 */

/* constants.js:74-80 -- decimal literals parsed as double, then stored to Float32Array. */
static const double QMF_COEFF_LITERALS[24] = {
    -0.00001461907, -0.00009205479, -0.000056157569, 0.00030117269, 0.0002422519,
    -0.00085293897, -0.0005205574,  0.0020340169,    0.00078333891, -0.0042153862,
    -0.00075614988, 0.0078402944,   -0.000061169922, -0.01344162,   0.0024626821,
    0.021736089,    -0.007801671,   -0.034090221,    0.01880949,    0.054326009,
    -0.043596379,   -0.099384367,   0.13207909,      0.46424159};

static float g_qmf_even[24], g_qmf_odd[24];
static int g_qmf_ready = 0;

static void qmf_tables_init(void) {
  if (g_qmf_ready) return;
  float window[48];
  for (int i = 0; i < 24; i++) { /* constants.js:83-90 */
    float c = (float)QMF_COEFF_LITERALS[i];
    window[i] = (float)((double)c * 2.0);
    window[47 - i] = (float)((double)c * 2.0);
  }
  for (int i = 0; i < 24; i++) { /* constants.js:93-107 */
    g_qmf_even[i] = window[2 * i];
    g_qmf_odd[i] = window[2 * i + 1];
  }
  g_qmf_ready = 1;
}

const float *c1o_qmf_even(void) { qmf_tables_init(); return g_qmf_even; }
const float *c1o_qmf_odd(void) { qmf_tables_init(); return g_qmf_odd; }
const int *c1o_specs_per_bfu(void) { return SPECS_PER_BFU; }
const int *c1o_bfu_start_long(void) { return BFU_START_LONG; }
const int *c1o_bfu_start_short(void) { return BFU_START_SHORT; }

/* DISTORTION_DELTA_FACTORS / WORD_LENGTH_DELTA_BITS / INV_POWER_OF_TWO
 * (constants.js:153-179): all exact powers of two, no libm involved. */
static double inv_pow2(int b) { return ldexp(1.0, -b); }
static int wl_delta_bits(int i) { return WL_BITS[i + 1] - WL_BITS[i]; }
static double distortion_delta_factor(int i) {
  if (i == 0) return 2.0 - inv_pow2(WL_BITS[1]);
  return inv_pow2(WL_BITS[i]) - inv_pow2(WL_BITS[i + 1]);
}

static void mdct_table(double *tab, int size, double scale) { /* mdct.js:21-37 */
  const double alpha = (2.0 * M_PI) / (8.0 * size);
  const double omega = (2.0 * M_PI) / size;
  const double scale_root = sqrt(scale / size);
  for (int i = 0; i < size / 4; i++) {
    const double angle = omega * i + alpha;
    tab[2 * i] = scale_root * cos(angle);
    tab[2 * i + 1] = scale_root * sin(angle);
  }
}

void c1o_default_tables(c1o_tables *t) {
  qmf_tables_init();
  for (int i = 0; i < 32; i++) /* constants.js:60-66 */
    t->window_short[i] = sin(((i + 0.5) * M_PI) / 64);
  for (int i = 0; i < 64; i++) /* constants.js:144-150 */
    t->scale_factors[i] = pow(2.0, i / 3.0 - 21);
  mdct_table(t->mdct_fwd64, 64, 0.5); /* mdct.js:215-221 */
  mdct_table(t->mdct_fwd256, 256, 0.5);
  mdct_table(t->mdct_fwd512, 512, 1.0);
  mdct_table(t->mdct_inv64, 64, 64 * 8);
  mdct_table(t->mdct_inv256, 256, 256 * 8);
  mdct_table(t->mdct_inv512, 512, 512 * 4);
  for (int k = 0; k < 8; k++) { /* fft.js:36-39 */
    const int stride = 2 << k;
    const double angle = (-2 * M_PI) / stride;
    t->fft_w[k][0] = cos(angle);
    t->fft_w[k][1] = sin(angle);
  }
}

void c1o_options_init(c1o_options *o, const c1o_tables *t, double threshold, double bias,
                      const int *fixed_modes) {
  o->transient_threshold = threshold;
  o->allocation_bias = bias;
  o->use_fixed_modes = fixed_modes != NULL;
  for (int i = 0; i < 3; i++) o->fixed_modes[i] = fixed_modes ? fixed_modes[i] : 0;
  for (int i = 0; i < 64; i++) /* bitallocation.js:51-58 */
    o->biased_sf[i] = (bias == 1) ? t->scale_factors[i] : pow(t->scale_factors[i], bias);
}

/* ------------------------------------------------------------------------------------
 * V8's Math.log / exp / log10 / log1p are ports of Sun fdlibm (v8/src/base/ieee754.cc,
 * not part of /root/reference; Node 20.16 pins V8 11.3).  The published fdlibm 5.3
 * algorithms (e_log.c, e_exp.c, e_log10.c, s_log1p.c) are restated here so the
 * transient score (transient.js:129,137,185,211) follows the engine the reference's CI
 * runs on rather than glibc.
 * ---------------------------------------------------------------------------------- */
static int32_t hi_word(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
static uint32_t lo_word(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u; }
static double set_hi_word(double x, int32_t hi) {
  uint64_t u; memcpy(&u, &x, 8);
  u = (u & 0xffffffffull) | ((uint64_t)(uint32_t)hi << 32);
  memcpy(&x, &u, 8); return x;
}
static double make_double(uint32_t hi, uint32_t lo) {
  uint64_t u = ((uint64_t)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x;
}

static const double LN2_HI = 6.93147180369123816490e-01, /* 3fe62e42 fee00000 */
    LN2_LO = 1.90821492927058770002e-10,                  /* 3dea39ef 35793c76 */
    TWO54 = 1.80143985094819840000e+16,                   /* 43500000 00000000 */
    LG1 = 6.666666666666735130e-01,                       /* 3FE55555 55555593 */
    LG2 = 3.999999999940941908e-01,                       /* 3FD99999 9997FA04 */
    LG3 = 2.857142874366239149e-01,                       /* 3FD24924 94229359 */
    LG4 = 2.222219843214978396e-01,                       /* 3FCC71C5 1D8E78AF */
    LG5 = 1.818357216161805012e-01,                       /* 3FC74664 96CB03DE */
    LG6 = 1.531383769920937332e-01,                       /* 3FC39A09 D078C69F */
    LG7 = 1.479819860511658591e-01;                       /* 3FC2F112 DF3E5244 */

double c1o_log(double x) { /* fdlibm e_log.c */
  double hfsq, f, s, z, R, w, t1, t2, dk;
  int32_t k, hx, i, j;
  uint32_t lx;
  hx = hi_word(x);
  lx = lo_word(x);
  k = 0;
  if (hx < 0x00100000) {
    if (((hx & 0x7fffffff) | lx) == 0) return -TWO54 / 0.0;
    if (hx < 0) return (x - x) / 0.0;
    k -= 54;
    x *= TWO54;
    hx = hi_word(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  i = (hx + 0x95f64) & 0x100000;
  x = set_hi_word(x, hx | (i ^ 0x3ff00000));
  k += (i >> 20);
  f = x - 1.0;
  if ((0x000fffff & (2 + hx)) < 3) {
    if (f == 0.0) {
      if (k == 0) return 0.0;
      dk = (double)k;
      return dk * LN2_HI + dk * LN2_LO;
    }
    R = f * f * (0.5 - 0.33333333333333333 * f);
    if (k == 0) return f - R;
    dk = (double)k;
    return dk * LN2_HI - ((R - dk * LN2_LO) - f);
  }
  s = f / (2.0 + f);
  dk = (double)k;
  z = s * s;
  i = hx - 0x6147a;
  w = z * z;
  j = 0x6b851 - hx;
  t1 = w * (LG2 + w * (LG4 + w * LG6));
  t2 = z * (LG1 + w * (LG3 + w * (LG5 + w * LG7)));
  i |= j;
  R = t2 + t1;
  if (i > 0) {
    hfsq = 0.5 * f * f;
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return dk * LN2_HI - ((hfsq - (s * (hfsq + R) + dk * LN2_LO)) - f);
  }
  if (k == 0) return f - s * (f - R);
  return dk * LN2_HI - ((s * (f - R) - dk * LN2_LO) - f);
}

double c1o_exp(double x) { /* fdlibm e_exp.c as carried by V8 (exp(1) special case) */
  static const double half[2] = {0.5, -0.5}, huge = 1.0e+300,
                      twom1000 = 9.33263618503218878990e-302,
                      two1023 = 8.988465674311579539e307,
                      o_threshold = 7.09782712893383973096e+02,
                      u_threshold = -7.45133219101941108420e+02,
                      ln2HI[2] = {6.93147180369123816490e-01, -6.93147180369123816490e-01},
                      ln2LO[2] = {1.90821492927058770002e-10, -1.90821492927058770002e-10},
                      invln2 = 1.44269504088896338700e+00,
                      P1 = 1.66666666666666019037e-01,  /* 3FC55555 5555553E */
                      P2 = -2.77777777770155933842e-03, /* BF66C16C 16BEBD93 */
                      P3 = 6.61375632143793436117e-05,  /* 3F11566A AF25DE2C */
                      P4 = -1.65339022054652515390e-06, /* BEBBBD41 C5D26BF1 */
                      P5 = 4.13813679705723846039e-08,  /* 3E663769 72BEA4D0 */
                      E = 2.718281828459045;
  double y, hi = 0.0, lo = 0.0, c, t, twopk;
  int32_t k = 0, xsb;
  uint32_t hx;
  hx = (uint32_t)hi_word(x);
  xsb = (hx >> 31) & 1;
  hx &= 0x7fffffff;
  if (hx >= 0x40862E42) {
    if (hx >= 0x7ff00000) {
      if (((hx & 0xfffff) | lo_word(x)) != 0) return x + x;
      return (xsb == 0) ? x : 0.0;
    }
    if (x > o_threshold) return huge * huge;
    if (x < u_threshold) return twom1000 * twom1000;
  }
  if (hx > 0x3fd62e42) {
    if (hx < 0x3FF0A2B2) {
      if (x == 1.0) return E;
      hi = x - ln2HI[xsb];
      lo = ln2LO[xsb];
      k = 1 - xsb - xsb;
    } else {
      k = (int32_t)(invln2 * x + half[xsb]);
      t = k;
      hi = x - t * ln2HI[0];
      lo = t * ln2LO[0];
    }
    x = hi - lo;
  } else if (hx < 0x3e300000) {
    if (huge + x > 1.0) return 1.0 + x;
  } else {
    k = 0;
  }
  t = x * x;
  if (k >= -1021)
    twopk = make_double(0x3ff00000u + ((uint32_t)k << 20), 0);
  else
    twopk = make_double(0x3ff00000u + ((uint32_t)(k + 1000) << 20), 0);
  c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
  y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  if (k >= -1021) {
    if (k == 1024) return y * 2.0 * two1023;
    return y * twopk;
  }
  return y * twopk * twom1000;
}

double c1o_log10(double x) { /* fdlibm e_log10.c */
  static const double ivln10 = 4.34294481903251816668e-01,
                      log10_2hi = 3.01029995663611771306e-01, /* 3FD34413 509F6000 */
                      log10_2lo = 3.69423907715893078616e-13; /* 3D59FEF3 11F12B36 */
  double y, z;
  int32_t i, k, hx;
  uint32_t lx;
  hx = hi_word(x);
  lx = lo_word(x);
  k = 0;
  if (hx < 0x00100000) {
    if (((hx & 0x7fffffff) | lx) == 0) return -TWO54 / 0.0;
    if (hx < 0) return (x - x) / 0.0;
    k -= 54;
    x *= TWO54;
    hx = hi_word(x);
    lx = lo_word(x);
  }
  if (hx >= 0x7ff00000) return x + x;
  if (hx == 0x3ff00000 && lx == 0) return 0.0;
  k += (hx >> 20) - 1023;
  i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
  hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
  y = (double)(k + i);
  x = make_double((uint32_t)hx, lx);
  z = y * log10_2lo + ivln10 * c1o_log(x);
  return z + y * log10_2hi;
}

double c1o_log1p(double x) { /* fdlibm s_log1p.c */
  static const double Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01,
                      Lp3 = 2.857142874366239149e-01, Lp4 = 2.222219843214978396e-01,
                      Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
                      Lp7 = 1.479819860511658591e-01;
  double hfsq, f = 0.0, c = 0.0, s, z, R, u;
  int32_t k, hx, hu = 0, ax;
  hx = hi_word(x);
  ax = hx & 0x7fffffff;
  k = 1;
  if (hx < 0x3FDA827A) {
    if (ax >= 0x3ff00000) {
      if (x == -1.0) return -TWO54 / 0.0;
      return (x - x) / (x - x);
    }
    if (ax < 0x3e200000) {
      if (TWO54 + x > 0.0 && ax < 0x3c900000) return x;
      return x - x * x * 0.5;
    }
    if (hx > 0 || hx <= ((int32_t)0xbfd2bec4)) {
      k = 0;
      f = x;
      hu = 1;
    }
  }
  if (hx >= 0x7ff00000) return x + x;
  if (k != 0) {
    if (hx < 0x43400000) {
      u = 1.0 + x;
      hu = hi_word(u);
      k = (hu >> 20) - 1023;
      c = (k > 0) ? 1.0 - (u - x) : x - (u - 1.0);
      c /= u;
    } else {
      u = x;
      hu = hi_word(u);
      k = (hu >> 20) - 1023;
      c = 0;
    }
    hu &= 0x000fffff;
    if (hu < 0x6a09e) {
      u = set_hi_word(u, hu | 0x3ff00000);
    } else {
      k += 1;
      u = set_hi_word(u, hu | 0x3fe00000);
      hu = (0x00100000 - hu) >> 2;
    }
    f = u - 1.0;
  }
  hfsq = 0.5 * f * f;
  if (hu == 0) {
    if (f == 0.0) {
      if (k == 0) return 0.0;
      c += k * LN2_LO;
      return k * LN2_HI + c;
    }
    R = hfsq * (1.0 - 0.66666666666666666 * f);
    if (k == 0) return f - R;
    return k * LN2_HI - ((R - (k * LN2_LO + c)) - f);
  }
  s = f / (2.0 + f);
  z = s * s;
  R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
  if (k == 0) return f - (hfsq - s * (hfsq + R));
  return k * LN2_HI - ((hfsq - (s * (hfsq + R) + (k * LN2_LO + c))) - f);
}

/* ECMAScript ToInt32 (the `| 0` of quantization.js:51). */
int32_t c1o_to_int32(double x) {
  if (!isfinite(x)) return 0;
  double t = trunc(x);
  double m = fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int32_t)(uint32_t)(uint64_t)m;
}

/* ------------------------------------------------------------------------------------
 * QMF (codec/transforms/qmf.js)
 * ---------------------------------------------------------------------------------- */
void c1o_qmf_analysis(const float *in, int n, float *delay46, float *lo, float *hi) {
  /* qmf.js:19-50 */
  qmf_tables_init();
  float work[46 + 512];
  memcpy(work, delay46, 46 * sizeof(float));
  memcpy(work + 46, in, (size_t)n * sizeof(float));
  const int n_out = n >> 1;
  for (int i = 0; i < n_out; i++) {
    double even_sum = 0, odd_sum = 0;
    const int off = i * 2;
    for (int j = 0; j < 24; j++) { /* qmf.js:38-41 */
      even_sum += (double)work[off + 47 - j * 2] * (double)g_qmf_even[j];
      odd_sum += (double)work[off + 46 - j * 2] * (double)g_qmf_odd[j];
    }
    lo[i] = (float)(even_sum + odd_sum); /* f32 */
    hi[i] = (float)(even_sum - odd_sum); /* f32 */
  }
  memcpy(delay46, work + n, 46 * sizeof(float)); /* qmf.js:48 slice(-46) */
}

void c1o_qmf_synthesis(const float *lo, const float *hi, int n_sub, float *delay46, float *out) {
  /* qmf.js:60-105 */
  qmf_tables_init();
  float work[46 + 512];
  const int n_out = n_sub * 2;
  memcpy(work, delay46, 46 * sizeof(float));
  for (int i = 0; i < n_sub; i++) { /* qmf.js:77-83 */
    const double l = lo[i], h = hi[i];
    work[46 + 2 * i] = (float)(0.5 * (l + h));
    work[46 + 2 * i + 1] = (float)(0.5 * (l - h));
  }
  for (int i = 0; i < n_sub; i++) { /* qmf.js:88-101 */
    const int off = i * 2;
    double s0 = 0, s1 = 0;
    for (int j = 0; j < 24; j++) {
      const int idx = off + j * 2;
      s0 += (double)work[idx] * (double)g_qmf_even[j];
      s1 += (double)work[idx + 1] * (double)g_qmf_odd[j];
    }
    out[2 * i] = (float)s1;
    out[2 * i + 1] = (float)s0;
  }
  memcpy(delay46, work + n_out, 46 * sizeof(float));
}

/* ------------------------------------------------------------------------------------
 * FFT (codec/transforms/fft.js:14-68): in-place radix-2 DIT on f32 arrays, binary64
 * arithmetic, f32 store after every butterfly, twiddles by recurrence per group.
 * ---------------------------------------------------------------------------------- */
void c1o_fft(float *re, float *im, int n, const c1o_tables *t) {
  if (n == 1) return;
  int bits = 0;
  while ((1 << bits) < n) bits++;
  for (int i = 0; i < n; i++) { /* fft.js:21-32 */
    int r = 0, tmp = i;
    for (int b = 0; b < bits; b++) { r = (r << 1) | (tmp & 1); tmp >>= 1; }
    if (r > i) {
      float a = re[i]; re[i] = re[r]; re[r] = a;
      a = im[i]; im[i] = im[r]; im[r] = a;
    }
  }
  int level = 0;
  for (int stride = 2; stride <= n; stride <<= 1, level++) { /* fft.js:35-66 */
    const int half = stride >> 1;
    const double w_re = t->fft_w[level][0];
    const double w_im = t->fft_w[level][1];
    for (int start = 0; start < n; start += stride) {
      double tw_re = 1, tw_im = 0;
      for (int k = 0; k < half; k++) {
        const int e = start + k, o = e + half;
        const double e_re = re[e], e_im = im[e], o_re = re[o], o_im = im[o];
        const double t_re = o_re * tw_re - o_im * tw_im;
        const double t_im = o_re * tw_im + o_im * tw_re;
        re[e] = (float)(e_re + t_re);
        im[e] = (float)(e_im + t_im);
        re[o] = (float)(e_re - t_re);
        im[o] = (float)(e_im - t_im);
        const double next_re = tw_re * w_re - tw_im * w_im;
        tw_im = tw_re * w_im + tw_im * w_re;
        tw_re = next_re;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * MDCT / IMDCT (codec/transforms/mdct.js)
 * ---------------------------------------------------------------------------------- */
static const double *mdct_tab(const c1o_tables *t, int size, int inverse) {
  if (size == 64) return inverse ? t->mdct_inv64 : t->mdct_fwd64;
  if (size == 256) return inverse ? t->mdct_inv256 : t->mdct_fwd256;
  return inverse ? t->mdct_inv512 : t->mdct_fwd512;
}

void c1o_mdct(const c1o_tables *t, int size, const float *in, float *out) {
  /* mdct.js:54-122 */
  const double *tab = mdct_tab(t, size, 0);
  const int half = size >> 1, n4 = size >> 2, n34 = 3 * n4, fft_n = half >> 1;
  float re[128], im[128];
  memset(re, 0, sizeof re);
  memset(im, 0, sizeof im);
  for (int i = 0; i < n4; i += 2) { /* mdct.js:76-89 */
    const double r = (double)in[n34 - 1 - i] + (double)in[n34 + i];
    const double m = (double)in[n4 + i] - (double)in[n4 - 1 - i];
    const double c = tab[i], s = tab[i + 1];
    re[i >> 1] = (float)(r * c + m * s);
    im[i >> 1] = (float)(m * c - r * s);
  }
  for (int i = n4; i < half; i += 2) { /* mdct.js:91-105 */
    const double r = (double)in[n34 - 1 - i] - (double)in[i - n4];
    const double m = (double)in[n4 + i] + (double)in[5 * n4 - 1 - i];
    const double c = tab[i], s = tab[i + 1];
    re[i >> 1] = (float)(r * c + m * s);
    im[i >> 1] = (float)(m * c - r * s);
  }
  c1o_fft(re, im, fft_n, t);
  for (int i = 0; i < fft_n; i++) { /* mdct.js:111-119 */
    const double c = tab[i * 2], s = tab[i * 2 + 1];
    const double r = re[i], m = im[i];
    out[i * 2] = (float)(-r * c - m * s);
    out[half - 1 - i * 2] = (float)(-r * s + m * c);
  }
}

void c1o_imdct(const c1o_tables *t, int size, const float *in, float *out) {
  /* mdct.js:139-211 */
  const double *tab = mdct_tab(t, size, 1);
  const int half = size >> 1, n4 = size >> 2, n34 = 3 * n4, fft_n = half >> 1;
  float re[128], im[128];
  for (int i = 0; i < fft_n; i++) { /* mdct.js:161-170 */
    const int i2 = i * 2;
    const double r = -(double)in[i2];
    const double m = -(double)in[half - 1 - i2];
    const double c = tab[i2], s = tab[i2 + 1];
    re[i] = (float)(m * s + r * c);
    im[i] = (float)(m * c - r * s);
  }
  c1o_fft(re, im, fft_n, t);
  for (int i = 0; i < fft_n / 2; i++) { /* mdct.js:177-191 */
    const int i2 = i * 2;
    const double c = tab[i2], s = tab[i2 + 1];
    const double r = re[i], m = im[i];
    const double r1 = r * c + m * s;
    const double i1 = r * s - m * c;
    out[n34 - 1 - i2] = (float)r1;
    out[n34 + i2] = (float)r1;
    out[n4 + i2] = (float)i1;
    out[n4 - 1 - i2] = (float)(-i1);
  }
  for (int i = fft_n / 2; i < fft_n; i++) { /* mdct.js:193-208 */
    const int idx = (i - fft_n / 2) * 2 + n4;
    const int i2 = i * 2;
    const double c = tab[i2], s = tab[i2 + 1];
    const double r = re[i], m = im[i];
    const double r1 = r * c + m * s;
    const double i1 = r * s - m * c;
    out[n34 - 1 - idx] = (float)r1;
    out[idx - n4] = (float)(-r1);
    out[n4 + idx] = (float)i1;
    out[5 * n4 - 1 - idx] = (float)i1;
  }
}

void c1o_overlap_add(const float *prev, const float *curr, int size, const double *window,
                     float *out) {
  /* mdct.js:230-245 */
  for (int i = 0; i < size; i++) {
    const double w1 = window[i], w2 = window[2 * size - 1 - i];
    const double p = prev[i], c = curr[size - 1 - i];
    out[i] = (float)(p * w2 - c * w1);
    out[2 * size - 1 - i] = (float)(p * w1 + c * w2);
  }
}

/* ------------------------------------------------------------------------------------
 * Transient detection (codec/analysis/transient.js)
 * ---------------------------------------------------------------------------------- */
void c1o_perform_fft(const float *samples, int n_samples, int fft_size, const c1o_tables *t,
                     float *mag) {
  /* transient.js:17-35 */
  float re[256], im[256];
  memset(re, 0, sizeof re);
  memset(im, 0, sizeof im);
  const int copy = n_samples < fft_size ? n_samples : fft_size;
  memcpy(re, samples, (size_t)copy * sizeof(float));
  c1o_fft(re, im, fft_size, t);
  for (int i = 0; i < fft_size / 2; i++) {
    const double r = re[i], m = im[i];
    mag[i] = (float)sqrt(r * r + m * m);
  }
}

static double js_max(double a, double b) { /* Math.max */
  if (isnan(a) || isnan(b)) return NAN;
  if (a == 0 && b == 0) return signbit(a) ? b : a;
  return a > b ? a : b;
}
static double js_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  if (a == 0 && b == 0) return signbit(a) ? a : b;
  return a < b ? a : b;
}

static double spectral_flux(const float *cur, const float *prev, int n) { /* :92-112 */
  double flux = 0, energy = 0;
  for (int i = 0; i < n; i++) {
    const double c = fabs((double)cur[i]), p = fabs((double)prev[i]);
    const double diff = c - p;
    if (diff > 0) flux += diff;
    energy += c * c;
  }
  double norm = sqrt(energy);
  if (norm == 0 || isnan(norm)) norm = 1e-6; /* `|| 1e-6` */
  return flux / norm;
}

/* Which libm the transient detector's log / exp / log10 / log1p use: 0 (default) the fdlibm port above, which is
 * what V8 carries; 1 the host's libm, which is what Qt's QJSEngine (the engine that wrote tests/golden/ref) calls.
 * Test-only switch, process-wide. */
static int g_host_libm = 0;
void c1o_set_host_libm(int on) { g_host_libm = on; }
static double m_log(double x) { return g_host_libm ? log(x) : c1o_log(x); }
static double m_exp(double x) { return g_host_libm ? exp(x) : c1o_exp(x); }
static double m_log10(double x) { return g_host_libm ? log10(x) : c1o_log10(x); }
static double m_log1p(double x) { return g_host_libm ? log1p(x) : c1o_log1p(x); }

static double spectral_flatness(const float *x, int n) { /* :120-141 */
  const double EPS = 1e-10;
  double sum_log = 0, sum_lin = 0;
  int valid = 0;
  for (int i = 0; i < n; i++) {
    const double m = fabs((double)x[i]);
    if (m > EPS) {
      sum_log += m_log(m);
      sum_lin += m;
      valid++;
    }
  }
  if (valid == 0) return 0;
  const double geo = m_exp(sum_log / valid);
  const double arith = sum_lin / valid;
  return arith > EPS ? geo / arith : 0;
}

static double hf_ratio(const float *x, int n) { /* :149-164 */
  const int mid = n / 2;
  double lo = 0, hi = 0;
  for (int i = 0; i < mid; i++) lo += (double)x[i] * (double)x[i];
  for (int i = mid; i < n; i++) hi += (double)x[i] * (double)x[i];
  const double total = lo + hi;
  return total > 0 ? hi / total : 0;
}

static double energy_change(const float *cur, const float *prev, int n) { /* :172-189 */
  double ce = 0, pe = 0;
  for (int i = 0; i < n; i++) {
    ce += (double)cur[i] * (double)cur[i];
    pe += (double)prev[i] * (double)prev[i];
  }
  ce = js_max(ce, 1e-10);
  pe = js_max(pe, 1e-10);
  const double db = 10 * m_log10(ce / pe);
  return js_max(0, db);
}

double c1o_transient_score(const float *cur, const float *prev, int n) {
  /* transient.js:63-86 (features) and :197-226 (score) */
  const double flux = spectral_flux(cur, prev, n);
  const double flat_change = fabs(spectral_flatness(cur, n) - spectral_flatness(prev, n));
  const double hf_change = fabs(hf_ratio(cur, n) - hf_ratio(prev, n));
  const double e_change = energy_change(cur, prev, n);
  const double flat_c = sqrt(flat_change);
  const double hf_c = m_log1p(hf_change * 10) / m_log1p(10);
  const double e_c = js_min(e_change / 30, 1);
  return (flux + flat_c + hf_c + e_c) / 4;
}

/* ------------------------------------------------------------------------------------
 * Bit allocation (codec/coding/bitallocation.js)
 * ---------------------------------------------------------------------------------- */
int c1o_find_scale_factor(const float *coefs, int n) { /* bitallocation.js:290-299 */
  double max_amp = 0.0;
  for (int i = 0; i < n; i++) {
    const double a = fabs((double)coefs[i]);
    if (a > max_amp) max_amp = a;
  }
  if (max_amp == 0) return 0;
  const double idx = ceil(3 * (log2(max_amp) + 21));
  double r = js_min(63, idx);
  r = js_max(0, r);
  return (int)r;
}

/* Exact replacement of the log2 expression by 63 f32 thresholds:
 * index = #{k in 0..62 : max_abs > thr[k]}, thr[k] = largest f32 <= 2^(k/3 - 21).
 * Cube roots of 2 are irrational, so only k % 3 == 0 hits an f32 exactly.  The table is
 * derived with exact integer arithmetic (no libm): m^3 is compared against 2 or 4. */
static float g_sf_thr[63];
static int g_sf_thr_ready = 0;
static void sf_thr_init(void) {
  if (g_sf_thr_ready) return;
  /* largest 24-bit integer m with m^3 <= r * 2^69  (r = 2 or 4): m / 2^23 = floor cube root */
  uint32_t root[3];
  root[0] = 1u << 23;
  for (int r = 1; r <= 2; r++) {
    uint32_t lo = 1u << 23, hi = (1u << 24) - 1;
    const unsigned __int128 target = (unsigned __int128)(r == 1 ? 2 : 4) << 69;
    while (lo < hi) {
      uint32_t mid = lo + (hi - lo + 1) / 2;
      unsigned __int128 cube = (unsigned __int128)mid * mid * mid;
      if (cube <= target) lo = mid; else hi = mid - 1;
    }
    root[r] = lo;
  }
  for (int k = 0; k < 63; k++) {
    const int e = k / 3 - 21;
    g_sf_thr[k] = (float)ldexp((double)root[k % 3], e - 23);
  }
  g_sf_thr_ready = 1;
}
const float *c1o_sf_thresholds(void) { sf_thr_init(); return g_sf_thr; }
int c1o_find_scale_factor_table(float max_abs) {
  sf_thr_init();
  if (!(max_abs > 0)) return 0;
  int idx = 0;
  for (int k = 0; k < 63; k++) idx += (max_abs > g_sf_thr[k]);
  return idx;
}

static void sift_down(int *h_idx, float *h_pri, int start, int size) { /* :314-341 */
  int i = start;
  const int idx_val = h_idx[i];
  const float pr_val = h_pri[i];
  for (;;) {
    const int l = (i << 1) + 1, r = l + 1;
    int max_i = i;
    float max_p = pr_val;
    if (l < size && h_pri[l] > max_p) { max_i = l; max_p = h_pri[l]; }
    if (r < size && h_pri[r] > max_p) { max_i = r; }
    if (max_i == i) break;
    h_idx[i] = h_idx[max_i];
    h_pri[i] = h_pri[max_i];
    i = max_i;
  }
  h_idx[i] = idx_val;
  h_pri[i] = pr_val;
}

static void distribute_bits_rdo(int active, const int *sizes, int remaining, const double *bsf,
                                const int *sfi, int *wl) { /* :203-281 */
  int h_idx[52];
  float h_pri[52];
  int h_size = 0;
  memset(wl, 0, sizeof(int) * (size_t)active);
  for (int b = 0; b < active; b++) {
    if (sizes[b] == 0) continue;
    if (sfi[b] == 0) continue;
    const int db = wl_delta_bits(0);
    if (db <= 0) continue;
    const double dd = bsf[sfi[b]] * distortion_delta_factor(0);
    h_idx[h_size] = b;
    h_pri[h_size] = (float)(dd / db); /* f32 */
    h_size++;
  }
  if (h_size == 0) return;
  for (int i = (h_size >> 1) - 1; i >= 0; i--) sift_down(h_idx, h_pri, i, h_size);
  while (remaining > 0 && h_size > 0) {
    const int b = h_idx[0];
    const int cur = wl[b];
    const int sz = sizes[b];
    /* WORD_LENGTH_DELTA_BITS has 15 entries; cur never reaches 15 inside the heap. */
    const int db = wl_delta_bits(cur);
    const int cost = db * sz;
    if (cost > remaining || cost <= 0) {
      const int last = h_size - 1;
      h_idx[0] = h_idx[last];
      h_pri[0] = h_pri[last];
      h_size--;
      if (h_size > 0) sift_down(h_idx, h_pri, 0, h_size);
      continue;
    }
    remaining -= cost;
    const int nxt = cur + 1;
    wl[b] = nxt;
    if (nxt < MAX_WL_INDEX && wl_delta_bits(nxt) > 0) {
      const double dd = bsf[sfi[b]] * distortion_delta_factor(nxt);
      h_pri[0] = (float)(dd / wl_delta_bits(nxt)); /* f32 */
      sift_down(h_idx, h_pri, 0, h_size);
    } else {
      const int last = h_size - 1;
      h_idx[0] = h_idx[last];
      h_pri[0] = h_pri[last];
      h_size--;
      if (h_size > 0) sift_down(h_idx, h_pri, 0, h_size);
    }
  }
}

static double total_distortion(int active, int max_bfu, const int *sizes, const int *wl,
                               const int *sfi, const double *bsf, const float *zero_bit) {
  /* :157-190 */
  double total = 0.0;
  for (int i = 0; i < active; i++) {
    const int bits = WL_BITS[wl[i]];
    if (bits == 0) { total += (double)zero_bit[i]; continue; }
    if (sfi[i] == 0) continue;
    total += bsf[sfi[i]] * inv_pow2(bits) * sizes[i];
  }
  for (int i = active; i < max_bfu; i++) total += (double)zero_bit[i];
  return total;
}

static const int *bfu_starts(const int *modes, int bfu) { /* quantization.js:113-118 */
  const int band = bfu < 20 ? 0 : bfu < 36 ? 1 : 2;
  return modes[band] == 0 ? BFU_START_LONG : BFU_START_SHORT;
}

static double g_dbg_totals[8];
const double *c1o_debug_last_totals(void) { return g_dbg_totals; } /* not thread safe: tests only */

void c1o_allocate_bits(const float *coefs, const int *modes, const c1o_options *o, int *n_bfu,
                       int *sfi52, int *wl52) {
  /* quantization.js:106-149 (groupIntoBFUs: every BFU slice lies inside its band, so the
   * gather reduces to a start offset) + bitallocation.js:74-142 */
  float zero_bit[52];
  memset(zero_bit, 0, sizeof zero_bit);
  for (int i = 0; i < 52; i++) {
    const int sz = SPECS_PER_BFU[i];
    const int sfi = c1o_find_scale_factor(coefs + bfu_starts(modes, i)[i], sz);
    sfi52[i] = sfi;
    if (sfi > 0) zero_bit[i] = (float)(o->biased_sf[sfi] * 2.0 * sz); /* f32 */
  }
  int best = -1;
  double min_total = INFINITY;
  int best_wl[52];
  memset(best_wl, 0, sizeof best_wl);
  for (int c = 0; c < 8; c++) {
    const int cand = BFU_AMOUNTS[c];
    const int avail = FRAME_BITS - FRAME_OVERHEAD_BITS - cand * BITS_PER_BFU_METADATA;
    if (avail < 0) continue;
    int wl[52];
    distribute_bits_rdo(cand, SPECS_PER_BFU, avail, o->biased_sf, sfi52, wl);
    const double total = total_distortion(cand, 52, SPECS_PER_BFU, wl, sfi52, o->biased_sf, zero_bit);
    g_dbg_totals[c] = total;
    if (total < min_total) {
      min_total = total;
      best = cand;
      memset(best_wl, 0, sizeof best_wl);
      memcpy(best_wl, wl, sizeof(int) * (size_t)cand);
    }
  }
  if (best < 0) { /* bitallocation.js:132-139 (only reachable with NaN distortion) */
    *n_bfu = BFU_AMOUNTS[0];
    memset(wl52, 0, sizeof(int) * 52);
    memset(sfi52, 0, sizeof(int) * 52);
    return;
  }
  *n_bfu = best;
  memcpy(wl52, best_wl, sizeof best_wl);
}

/* ------------------------------------------------------------------------------------
 * Quantisation (codec/coding/quantization.js)
 * ---------------------------------------------------------------------------------- */
void c1o_quantize(const float *c, int n, int sfi, int bits, const c1o_tables *t, int *out) {
  /* quantization.js:34-56 */
  if (bits == 0 || sfi == 0) { memset(out, 0, sizeof(int) * (size_t)n); return; }
  const double sf = t->scale_factors[sfi];
  const int range = (1 << (bits - 1)) - 1;
  const double norm = range / sf;
  for (int i = 0; i < n; i++) {
    const double x = (double)c[i] * norm;
    const int32_t y = c1o_to_int32(x + (x >= 0 ? 0.5 : -0.5));
    out[i] = y > range ? range : y < -range ? -range : y;
  }
}

void c1o_dequantize(const int *q, int n, int sfi, int bits, const c1o_tables *t, float *out) {
  /* quantization.js:65-78 */
  if (bits == 0 || sfi == 0) { memset(out, 0, sizeof(float) * (size_t)n); return; }
  const double sf = t->scale_factors[sfi];
  const int range = (1 << (bits - 1)) - 1;
  for (int i = 0; i < n; i++) out[i] = (float)(((double)q[i] * sf) / range); /* f32 */
}

/* ------------------------------------------------------------------------------------
 * Bitstream + sound-unit layout (codec/io/bitstream.js, codec/io/serialization.js)
 * ---------------------------------------------------------------------------------- */
void c1o_pack_bits(uint8_t *buf, size_t buf_len, int bit_pos, int value, int bit_count) {
  /* bitstream.js:15-39 */
  if (bit_count == 0) return;
  size_t byte = (size_t)(bit_pos / 8);
  int off = bit_pos % 8;
  uint32_t v = (uint32_t)value & (uint32_t)((1u << bit_count) - 1u);
  int written = 0;
  while (written < bit_count && byte < buf_len) {
    const int avail = 8 - off;
    const int n = (bit_count - written) < avail ? (bit_count - written) : avail;
    const int shift = bit_count - written - n;
    const uint32_t bits = (v >> shift) & ((1u << n) - 1u);
    const uint32_t mask = ((1u << n) - 1u) << (avail - n);
    buf[byte] = (uint8_t)((buf[byte] & ~mask) | (bits << (avail - n)));
    written += n;
    byte++;
    off = 0;
  }
}

int c1o_unpack_bits(const uint8_t *buf, size_t buf_len, int bit_pos, int bit_count) {
  /* bitstream.js:48-69 -- stops at the end of the buffer, returning the bits read so far */
  if (bit_count == 0) return 0;
  size_t byte = (size_t)(bit_pos / 8);
  int off = bit_pos % 8;
  uint32_t value = 0;
  for (int read = 0; read < bit_count && byte < buf_len;) {
    const int avail = 8 - off;
    const int n = (bit_count - read) < avail ? (bit_count - read) : avail;
    const uint32_t mask = (1u << n) - 1u;
    const uint32_t bits = ((uint32_t)buf[byte] >> (avail - n)) & mask;
    value = (value << n) | bits;
    read += n;
    byte++;
    off = 0;
  }
  return (int)value;
}

int c1o_unpack_signed_bits(const uint8_t *buf, size_t buf_len, int bit_pos, int bit_count) {
  /* bitstream.js:78-82 */
  const int v = c1o_unpack_bits(buf, buf_len, bit_pos, bit_count);
  const int sign = 1 << (bit_count - 1);
  return v >= sign ? v - (1 << bit_count) : v;
}

static int bfu_amount_index(int n_bfu) {
  for (int i = 0; i < 8; i++) if (BFU_AMOUNTS[i] == n_bfu) return i;
  return -1; /* Array.indexOf */
}

void c1o_serialize_frame(const c1o_frame *f, uint8_t out[C1O_SU_BYTES]) {
  /* serialization.js:41-98 */
  memset(out, 0, C1O_SU_BYTES);
  const int idx = bfu_amount_index(f->n_bfu);
  const uint32_t header = ((uint32_t)(2 - f->modes[0]) << 14) | ((uint32_t)(2 - f->modes[1]) << 12) |
                          ((uint32_t)(3 - f->modes[2]) << 10) | ((uint32_t)idx << 5);
  out[0] = (uint8_t)(header >> 8); /* setUint16 big-endian, value taken mod 2^16 */
  out[1] = (uint8_t)header;
  int pos = 16;
  for (int i = 0; i < f->n_bfu; i++) { c1o_pack_bits(out, C1O_SU_BYTES, pos, f->wl[i], 4); pos += 4; }
  for (int i = 0; i < f->n_bfu; i++) { c1o_pack_bits(out, C1O_SU_BYTES, pos, f->sfi[i], 6); pos += 6; }
  for (int i = 0; i < f->n_bfu; i++) {
    const int bits = WL_BITS[f->wl[i]];
    if (bits > 0) {
      for (int j = 0; j < SPECS_PER_BFU[i]; j++) {
        const int c = f->q[i][j];
        const int v = c < 0 ? c + (1 << bits) : c;
        c1o_pack_bits(out, C1O_SU_BYTES, pos, v, bits);
        pos += bits;
      }
    }
  }
  out[C1O_SU_BYTES - 3] = 0;
  out[C1O_SU_BYTES - 2] = 0;
  out[C1O_SU_BYTES - 1] = 0;
}

void c1o_deserialize_frame(const uint8_t in[C1O_SU_BYTES], c1o_frame *f) {
  /* serialization.js:111-176 */
  memset(f, 0, sizeof *f);
  const int header = (in[0] << 8) | in[1];
  f->modes[0] = 2 - ((header >> 14) & 3);
  f->modes[1] = 2 - ((header >> 12) & 3);
  f->modes[2] = 3 - ((header >> 10) & 3);
  f->n_bfu = BFU_AMOUNTS[(header >> 5) & 7];
  int pos = 16;
  for (int i = 0; i < f->n_bfu; i++) { f->wl[i] = c1o_unpack_bits(in, C1O_SU_BYTES, pos, 4); pos += 4; }
  for (int i = 0; i < f->n_bfu; i++) { f->sfi[i] = c1o_unpack_bits(in, C1O_SU_BYTES, pos, 6); pos += 6; }
  for (int i = 0; i < f->n_bfu; i++) {
    const int bits = WL_BITS[f->wl[i]];
    if (bits > 0) {
      for (int j = 0; j < SPECS_PER_BFU[i]; j++) {
        f->q[i][j] = c1o_unpack_signed_bits(in, C1O_SU_BYTES, pos, bits);
        pos += bits;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * Encoder pipeline (codec/pipeline/encoder.js)
 * ---------------------------------------------------------------------------------- */
void c1o_encoder_init(c1o_encoder *e, const c1o_tables *t, const c1o_options *o) {
  memset(e, 0, sizeof *e);
  e->T = t;
  e->opt = *o;
}

static void tail_window(const c1o_tables *t, float *samples, float *overlap, int block) {
  /* encoder.js:309-316 */
  const int start = block - 32;
  for (int i = 0; i < 32; i++) {
    const double v = samples[start + i];
    overlap[i] = (float)(t->window_short[i] * v);
    samples[start + i] = (float)(v * t->window_short[31 - i]);
  }
}

static void reverse_f32(float *x, int n) { /* utils.js:42-48 */
  for (int i = 0; i < n / 2; i++) { float a = x[i]; x[i] = x[n - 1 - i]; x[n - 1 - i] = a; }
}

static void mdct_band(const c1o_tables *t, float *samples, int band, int mode, float *overlap,
                      float *out) {
  const int size = band == 2 ? 256 : 128;        /* constants.js:115-119 */
  const int window_start = band == 2 ? 112 : 48;
  if (mode == 0) { /* encoder.js:228-258 */
    const int n = band == 2 ? 512 : 256;
    float buf[512];
    memset(buf, 0, sizeof buf);
    memcpy(buf + window_start, overlap, 32 * sizeof(float));
    tail_window(t, samples, overlap, size);
    memcpy(buf + window_start + 32, samples, (size_t)size * sizeof(float));
    c1o_mdct(t, n, buf, out);
    if (band > 0) reverse_f32(out, size);
  } else { /* encoder.js:269-307 */
    const int blocks = size == 256 ? 8 : 4;
    for (int b = 0; b < blocks; b++) {
      float buf[64];
      memcpy(buf, overlap, 32 * sizeof(float));
      tail_window(t, samples + b * 32, overlap, 32);
      memcpy(buf + 32, samples + b * 32, 32 * sizeof(float));
      c1o_mdct(t, 64, buf, out + b * 32);
      if (band > 0) reverse_f32(out + b * 32, 32);
    }
  }
}

void c1o_encode_frame(c1o_encoder *e, const float pcm[C1O_FRAME], c1o_frame *out,
                      c1o_enc_debug *dbg) {
  const c1o_tables *t = e->T;
  /* qmfAnalysisStage, encoder.js:69-95 */
  float lo1[256], hi1[256], bands[512];
  float *low = bands, *mid = bands + 128, *high = bands + 256;
  c1o_qmf_analysis(pcm, 512, e->delay_low, lo1, hi1);
  c1o_qmf_analysis(lo1, 256, e->delay_mid, low, mid);
  {
    float delayed[39 + 256];
    memcpy(delayed, e->delay_high, 39 * sizeof(float));
    memcpy(delayed + 39, hi1, 256 * sizeof(float));
    memcpy(high, delayed, 256 * sizeof(float));
    memcpy(e->delay_high, delayed + 256, 39 * sizeof(float));
  }
  if (dbg) { memcpy(dbg->bands, bands, sizeof bands); memset(dbg->mags, 0, sizeof dbg->mags); }

  /* blockSelectorStage, encoder.js:126-151 */
  int modes[3];
  if (e->opt.use_fixed_modes) {
    for (int b = 0; b < 3; b++) modes[b] = e->opt.fixed_modes[b];
  } else {
    static const int fft_sizes[3] = {128, 128, 256};
    static const int offs[3] = {0, 128, 256};
    static const int mag_offs[3] = {0, 64, 128};
    for (int b = 0; b < 3; b++) {
      float mag[128];
      const int n = fft_sizes[b] / 2;
      c1o_perform_fft(bands + offs[b], b == 2 ? 256 : 128, fft_sizes[b], t, mag);
      /* every band compares against transientThresholdLow (encoder.js:137-141) */
      const double score = c1o_transient_score(mag, e->prev_mag[b], n);
      const int transient = score > e->opt.transient_threshold;
      memcpy(e->prev_mag[b], mag, (size_t)n * sizeof(float));
      modes[b] = transient * ((b + 1) > 2 ? (b + 1) : 2);
      if (dbg) { memcpy(dbg->mags + mag_offs[b], mag, (size_t)n * sizeof(float)); dbg->score[b] = score; }
    }
  }

  /* mdctStage, encoder.js:330-348 */
  float coefs[512];
  mdct_band(t, low, 0, modes[0], e->overlap[0], coefs);
  mdct_band(t, mid, 1, modes[1], e->overlap[1], coefs + 128);
  mdct_band(t, high, 2, modes[2], e->overlap[2], coefs + 256);
  if (dbg) memcpy(dbg->coefs, coefs, sizeof coefs);

  /* quantizationStage, encoder.js:381-417 */
  memset(out, 0, sizeof *out);
  int sfi[52], wl[52], n_bfu;
  c1o_allocate_bits(coefs, modes, &e->opt, &n_bfu, sfi, wl);
  out->n_bfu = n_bfu;
  for (int b = 0; b < 3; b++) out->modes[b] = modes[b];
  for (int i = 0; i < n_bfu; i++) {
    out->sfi[i] = sfi[i];
    out->wl[i] = wl[i];
    c1o_quantize(coefs + bfu_starts(modes, i)[i], SPECS_PER_BFU[i], sfi[i], WL_BITS[wl[i]], t,
                 out->q[i]);
  }
}

/* ------------------------------------------------------------------------------------
 * Decoder pipeline (codec/pipeline/decoder.js)
 * ---------------------------------------------------------------------------------- */
void c1o_decoder_init(c1o_decoder *d, const c1o_tables *t) {
  memset(d, 0, sizeof *d);
  d->T = t;
}

static void imdct_band(const c1o_tables *t, const float *coefs, int band, int mode, float *tail,
                       float *out) {
  const int size = band == 2 ? 256 : 128;
  float inv_buf[256];
  if (mode == 0) { /* decoder.js:175-233 */
    const int n = band == 2 ? 512 : 256;
    float specs[256], inv[512];
    memcpy(specs, coefs, (size_t)size * sizeof(float));
    if (band > 0) reverse_f32(specs, size);
    c1o_imdct(t, n, specs, inv);
    for (int i = 0; i < size; i++) inv_buf[i] = inv[n / 4 + i];
    c1o_overlap_add(tail, inv_buf, 16, t->window_short, out);
    for (int i = 0; i < size - 32; i++) out[32 + i] = inv_buf[16 + i];
  } else { /* decoder.js:244-306 */
    const int blocks = size == 256 ? 8 : 4;
    float prev[16];
    memcpy(prev, tail, sizeof prev);
    for (int b = 0; b < blocks; b++) {
      float specs[32], inv[64];
      memcpy(specs, coefs + b * 32, sizeof specs);
      if (band > 0) reverse_f32(specs, 32);
      c1o_imdct(t, 64, specs, inv);
      for (int i = 0; i < 32; i++) inv_buf[b * 32 + i] = inv[16 + i];
      c1o_overlap_add(prev, inv_buf + b * 32, 16, t->window_short, out + b * 32);
      memcpy(prev, inv_buf + b * 32 + 16, sizeof prev);
    }
  }
  memcpy(tail, inv_buf + size - 16, 16 * sizeof(float));
}

void c1o_decode_frame(c1o_decoder *d, const c1o_frame *in, float pcm[C1O_FRAME],
                      c1o_dec_debug *dbg) {
  const c1o_tables *t = d->T;
  /* dequantizationStage, decoder.js:52-98 */
  float coefs[512];
  memset(coefs, 0, sizeof coefs);
  for (int i = 0; i < in->n_bfu; i++) {
    const int bits = WL_BITS[in->wl[i]];
    if (bits > 0)
      c1o_dequantize(in->q[i], SPECS_PER_BFU[i], in->sfi[i], bits, t,
                     coefs + bfu_starts(in->modes, i)[i]);
  }
  if (dbg) memcpy(dbg->coefs, coefs, sizeof coefs);
  /* imdctStage, decoder.js:315-329 */
  float bands[512];
  imdct_band(t, coefs, 0, in->modes[0], d->tail[0], bands);
  imdct_band(t, coefs + 128, 1, in->modes[1], d->tail[1], bands + 128);
  imdct_band(t, coefs + 256, 2, in->modes[2], d->tail[2], bands + 256);
  if (dbg) memcpy(dbg->bands, bands, sizeof bands);
  /* qmfSynthesisStage, decoder.js:360-388 */
  float delayed[39 + 256], high[256], stage2[256];
  memcpy(delayed, d->delay_high, 39 * sizeof(float));
  memcpy(delayed + 39, bands + 256, 256 * sizeof(float));
  memcpy(high, delayed, sizeof high);
  memcpy(d->delay_high, delayed + 256, 39 * sizeof(float));
  c1o_qmf_synthesis(bands, bands + 128, 128, d->delay_mid, stage2);
  c1o_qmf_synthesis(stage2, high, 256, d->delay_low, pcm);
}

/* ------------------------------------------------------------------------------------
 * Whole-buffer helpers (codec/io/processor.js)
 * ---------------------------------------------------------------------------------- */
size_t c1o_frame_count(size_t n_samples) { return (n_samples + 511) / 512; } /* :246-279 */

static void load_frame(const float *ch, size_t n_samples, size_t frame, float *dst) {
  /* frameBufferToFrames: zero-padded tail; stereo pads the shorter channel (:246-279) */
  for (size_t j = 0; j < 512; j++) {
    const size_t s = frame * 512 + j;
    dst[j] = s < n_samples ? ch[s] : 0.0f;
  }
}

void c1o_encode_pcm_range(const c1o_tables *t, const c1o_options *o, const float *const *ch,
                          int n_ch, size_t n_samples, size_t frame_begin, size_t frame_end,
                          uint8_t *su_out) {
  /* processor.js:97-136 (one encoder per channel, L then R per frame) + :317-339 */
  const size_t warm = frame_begin >= 2 ? 2 : frame_begin;
  for (int c = 0; c < n_ch; c++) {
    c1o_encoder enc;
    c1o_encoder_init(&enc, t, o);
    float pcm[512];
    c1o_frame fr;
    for (size_t f = frame_begin - warm; f < frame_end; f++) {
      load_frame(ch[c], n_samples, f, pcm);
      c1o_encode_frame(&enc, pcm, &fr, NULL);
      if (f >= frame_begin) c1o_serialize_frame(&fr, su_out + (f * (size_t)n_ch + (size_t)c) * C1O_SU_BYTES);
    }
  }
}

void c1o_decode_su_range(const c1o_tables *t, const uint8_t *su, size_t n_su, int n_ch,
                         size_t frame_begin, size_t frame_end, float *const *ch_out) {
  /* processor.js:167-237: de-interleave L,R; a missing right unit is replaced by the
   * dummy frame {nBfu:0, blockModes:[0,0,0]} (:299-307). */
  const size_t warm = frame_begin >= 1 ? 1 : 0;
  for (int c = 0; c < n_ch; c++) {
    c1o_decoder dec;
    c1o_decoder_init(&dec, t);
    c1o_frame fr;
    float pcm[512];
    for (size_t f = frame_begin - warm; f < frame_end; f++) {
      const size_t idx = f * (size_t)n_ch + (size_t)c;
      if (idx < n_su) c1o_deserialize_frame(su + idx * C1O_SU_BYTES, &fr);
      else memset(&fr, 0, sizeof fr);
      c1o_decode_frame(&dec, &fr, pcm, NULL);
      if (f >= frame_begin) memcpy(ch_out[c] + f * 512, pcm, sizeof pcm);
    }
  }
}

void c1o_aea_header(const char *title, uint32_t su_count, int n_ch, uint8_t out[C1O_AEA_HEADER]) {
  /* serialization.js:190-211 */
  memset(out, 0, C1O_AEA_HEADER);
  out[0] = 0x00; out[1] = 0x08; out[2] = 0x00; out[3] = 0x00;
  size_t n = strlen(title);
  if (n > 255) n = 255;
  memcpy(out + 4, title, n);
  out[260] = (uint8_t)su_count;
  out[261] = (uint8_t)(su_count >> 8);
  out[262] = (uint8_t)(su_count >> 16);
  out[263] = (uint8_t)(su_count >> 24);
  out[264] = (uint8_t)n_ch;
}

int c1o_aea_parse(const uint8_t *hdr, size_t len, char title[257], uint32_t *su_count, int *n_ch) {
  /* serialization.js:222-253 */
  if (len != C1O_AEA_HEADER) return -1; /* 'Header must be 2048 bytes' */
  static const uint8_t magic[4] = {0, 8, 0, 0};
  if (memcmp(hdr, magic, 4) != 0) return -2; /* 'Invalid AEA file' */
  size_t end = 4;
  while (end < C1O_AEA_HEADER && hdr[end] != 0) end++;
  size_t tl = (end == C1O_AEA_HEADER) ? 256 : end - 4;
  if (tl > 256) tl = 256;
  memcpy(title, hdr + 4, tl);
  title[tl] = 0;
  *su_count = (uint32_t)hdr[260] | ((uint32_t)hdr[261] << 8) | ((uint32_t)hdr[262] << 16) |
              ((uint32_t)hdr[263] << 24);
  *n_ch = hdr[264];
  return 0;
}

void c1o_pcm_to_int16(const float *in, size_t n, int16_t *out) {
  /* processor.js:382-389: clamp to [-1,1], x32768 below zero / x32767 otherwise,
   * DataView.setInt16 == ToInt16 (truncate toward zero). */
  for (size_t i = 0; i < n; i++) {
    double s = in[i];
    s = js_max(-1, js_min(1, s));
    const double v = s < 0 ? s * 32768 : s * 32767;
    out[i] = (int16_t)(uint16_t)(uint32_t)c1o_to_int32(v);
  }
}

void c1o_int16_to_pcm(const int16_t *in, size_t n, float *out) {
  /* bin/cli.js:395: readInt16LE / 32768.0 stored into a Float32Array */
  for (size_t i = 0; i < n; i++) out[i] = (float)((double)in[i] / 32768.0);
}
