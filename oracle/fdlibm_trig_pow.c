/*
 * fdlibm_trig_pow.c -- sin, cos and pow as V8 <= 11.3 (Node 20) computes them, for the oracle's
 * "V8 flavoured" tables.  TEST INFRASTRUCTURE ONLY (see carta1_oracle.h).
 *
 * V8's Math.sin / Math.cos / Math.pow are ports of Sun fdlibm 5.3 (v8/src/base/ieee754.cc: s_sin.c, s_cos.c,
 * k_sin.c, k_cos.c, e_rem_pio2.c, e_pow.c; not part of /root/reference).  The published algorithms are restated
 * here.  They matter for the ~900 table entries the reference derives from libm (codec/core/constants.js:63,147;
 * codec/transforms/mdct.js:29-35; fft.js:38-39; codec/coding/bitallocation.js:56): c1o_default_tables uses the
 * host's glibc (which is also what Qt's QJSEngine, the engine of the reference pin, calls), c1o_fdlibm_tables the
 * functions below.  tests/test_reference_pin.py::test_table_flavours reports how many entries differ between
 * the two and whether any emitted byte or decoded sample of the pinned runs changes with them.
 *
 * Every constant is given as its IEEE-754 bit pattern next to the decimal fdlibm prints; fd_selfcheck() verifies
 * that the two agree, so a transcription slip in either shows.  Arguments beyond 2^19 * pi/2 (the Payne-Hanek
 * path of e_rem_pio2.c) do not occur in the reference's tables and are refused.
 */
#define _DEFAULT_SOURCE /* M_PI, scalbn under -std=c11 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "carta1_oracle.h"

static int32_t HI(double x) { uint64_t u; memcpy(&u, &x, 8); return (int32_t)(u >> 32); }
static uint32_t LO(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u; }
static double MK(uint32_t hi, uint32_t lo) { uint64_t u = ((uint64_t)hi << 32) | lo; double x; memcpy(&x, &u, 8); return x; }
static double SET_HI(double x, int32_t hi) { return MK((uint32_t)hi, LO(x)); }
static double SET_LO(double x, uint32_t lo) { return MK((uint32_t)HI(x), lo); }

/* ---- k_sin.c / k_cos.c ---- */
#define S1 MK(0xBFC55555u, 0x55555549u) /* -1.66666666666666324348e-01 */
#define S2 MK(0x3F811111u, 0x1110F8A6u) /*  8.33333333332248946124e-03 */
#define S3 MK(0xBF2A01A0u, 0x19C161D5u) /* -1.98412698298579493134e-04 */
#define S4 MK(0x3EC71DE3u, 0x57B1FE7Du) /*  2.75573137070700676789e-06 */
#define S5 MK(0xBE5AE5E6u, 0x8A2B9CEBu) /* -2.50507602534068634195e-08 */
#define S6 MK(0x3DE5D93Au, 0x5ACFD57Cu) /*  1.58969099521155010221e-10 */
#define C1 MK(0x3FA55555u, 0x5555554Cu) /*  4.16666666666666019037e-02 */
#define C2 MK(0xBF56C16Cu, 0x16C15177u) /* -1.38888888888741095749e-03 */
#define C3 MK(0x3EFA01A0u, 0x19CB1590u) /*  2.48015872894767294178e-05 */
#define C4 MK(0xBE927E4Fu, 0x809C52ADu) /* -2.75573143513906633035e-07 */
#define C5 MK(0x3E21EE9Eu, 0xBDB4B1C4u) /*  2.08757232129817482790e-09 */
#define C6 MK(0xBDA8FAE9u, 0xBE8838D4u) /* -1.13596475577881948265e-11 */

static double k_sin(double x, double y, int iy) {
  const int32_t ix = HI(x) & 0x7fffffff;
  if (ix < 0x3e400000) { /* |x| < 2^-27 */
    if ((int)x == 0) return x;
  }
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  if (iy == 0) return x + v * (S1 + z * r);
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static double k_cos(double x, double y) {
  double qx;
  const int32_t ix = HI(x) & 0x7fffffff;
  if (ix < 0x3e400000) {
    if ((int)x == 0) return 1.0;
  }
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  if (ix < 0x3FD33333) return 1.0 - (0.5 * z - (z * r - x * y));
  if (ix > 0x3fe90000) qx = 0.28125;
  else qx = MK((uint32_t)(ix - 0x00200000), 0);
  const double hz = 0.5 * z - qx;
  const double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

/* ---- e_rem_pio2.c (without the Payne-Hanek tail) ---- */
#define INVPIO2 MK(0x3FE45F30u, 0x6DC9C883u) /* 6.36619772367581382433e-01 */
#define PIO2_1 MK(0x3FF921FBu, 0x54400000u)  /* 1.57079632673412561417e+00 */
#define PIO2_1T MK(0x3DD0B461u, 0x1A626331u) /* 6.07710050650619224932e-11 */
#define PIO2_2 MK(0x3DD0B461u, 0x1A600000u)  /* 6.07710050630396597660e-11 */
#define PIO2_2T MK(0x3BA3198Au, 0x2E037073u) /* 2.02226624879595063154e-21 */
#define PIO2_3 MK(0x3BA3198Au, 0x2E000000u)  /* 2.02226624871116645580e-21 */
#define PIO2_3T MK(0x397B839Au, 0x252049C1u) /* 8.47842766036889956997e-32 */

static const int32_t npio2_hw[32] = {
    0x3FF921FB, 0x400921FB, 0x4012D97C, 0x401921FB, 0x401F6A7A, 0x4022D97C, 0x4025FDBB, 0x402921FB,
    0x402C463A, 0x402F6A7A, 0x4031475C, 0x4032D97C, 0x40346B9C, 0x4035FDBB, 0x40378FDB, 0x403921FB,
    0x403AB41B, 0x403C463A, 0x403DD85A, 0x403F6A7A, 0x40407E4C, 0x4041475C, 0x4042106C, 0x4042D97C,
    0x4043A28C, 0x40446B9C, 0x404534AC, 0x4045FDBB, 0x4046C6CB, 0x40478FDB, 0x404858EB, 0x404921FB,
};

static int rem_pio2(double x, double y[2]) {
  double z, w, t, r, fn;
  int i, j, n;
  const int32_t hx = HI(x);
  const int32_t ix = hx & 0x7fffffff;
  if (ix <= 0x3fe921fb) { y[0] = x; y[1] = 0; return 0; }
  if (ix < 0x4002d97c) { /* |x| < 3pi/4 */
    if (hx > 0) {
      z = x - PIO2_1;
      if (ix != 0x3ff921fb) { y[0] = z - PIO2_1T; y[1] = (z - y[0]) - PIO2_1T; }
      else { z -= PIO2_2; y[0] = z - PIO2_2T; y[1] = (z - y[0]) - PIO2_2T; }
      return 1;
    }
    z = x + PIO2_1;
    if (ix != 0x3ff921fb) { y[0] = z + PIO2_1T; y[1] = (z - y[0]) + PIO2_1T; }
    else { z += PIO2_2; y[0] = z + PIO2_2T; y[1] = (z - y[0]) + PIO2_2T; }
    return -1;
  }
  if (ix <= 0x413921fb) { /* |x| ~<= 2^19 * (pi/2) */
    t = fabs(x);
    n = (int)(t * INVPIO2 + 0.5);
    fn = (double)n;
    r = t - fn * PIO2_1;
    w = fn * PIO2_1T;
    if (n < 32 && ix != npio2_hw[n - 1]) {
      y[0] = r - w;
    } else {
      j = ix >> 20;
      y[0] = r - w;
      i = j - ((HI(y[0]) >> 20) & 0x7ff);
      if (i > 16) {
        t = r;
        w = fn * PIO2_2;
        r = t - w;
        w = fn * PIO2_2T - ((t - r) - w);
        y[0] = r - w;
        i = j - ((HI(y[0]) >> 20) & 0x7ff);
        if (i > 49) {
          t = r;
          w = fn * PIO2_3;
          r = t - w;
          w = fn * PIO2_3T - ((t - r) - w);
          y[0] = r - w;
        }
      }
    }
    y[1] = (r - y[0]) - w;
    if (hx < 0) { y[0] = -y[0]; y[1] = -y[1]; return -n; }
    return n;
  }
  abort(); /* Payne-Hanek range: not reachable from the reference's table arguments */
}

double c1o_fd_sin(double x) { /* s_sin.c */
  double y[2];
  const int32_t ix = HI(x) & 0x7fffffff;
  if (ix <= 0x3fe921fb) return k_sin(x, 0.0, 0);
  if (ix >= 0x7ff00000) return x - x;
  const int n = rem_pio2(x, y);
  switch (n & 3) {
    case 0: return k_sin(y[0], y[1], 1);
    case 1: return k_cos(y[0], y[1]);
    case 2: return -k_sin(y[0], y[1], 1);
    default: return -k_cos(y[0], y[1]);
  }
}

double c1o_fd_cos(double x) { /* s_cos.c */
  double y[2];
  const int32_t ix = HI(x) & 0x7fffffff;
  if (ix <= 0x3fe921fb) return k_cos(x, 0.0);
  if (ix >= 0x7ff00000) return x - x;
  const int n = rem_pio2(x, y);
  switch (n & 3) {
    case 0: return k_cos(y[0], y[1]);
    case 1: return -k_sin(y[0], y[1], 1);
    case 2: return -k_cos(y[0], y[1]);
    default: return k_sin(y[0], y[1], 1);
  }
}

/* ---- e_pow.c ---- */
#define DP_H1 MK(0x3FE2B803u, 0x40000000u)  /* 5.84962487220764160156e-01 */
#define DP_L1 MK(0x3E4CFDEBu, 0x43CFD006u)  /* 1.35003920212974897128e-08 */
#define PL1 MK(0x3FE33333u, 0x33333303u)    /* 5.99999999999994648725e-01 */
#define PL2 MK(0x3FDB6DB6u, 0xDB6FABFFu)    /* 4.28571428578550184252e-01 */
#define PL3 MK(0x3FD55555u, 0x518F264Du)    /* 3.33333329818377432918e-01 */
#define PL4 MK(0x3FD17460u, 0xA91D4101u)    /* 2.72728123808534006489e-01 */
#define PL5 MK(0x3FCD864Au, 0x93C9DB65u)    /* 2.30660745775561754067e-01 */
#define PL6 MK(0x3FCA7E28u, 0x4A454EEFu)    /* 2.06975017800338417784e-01 */
#define PP1 MK(0x3FC55555u, 0x5555553Eu)    /* 1.66666666666666019037e-01 */
#define PP2 MK(0xBF66C16Cu, 0x16BEBD93u)    /* -2.77777777770155933842e-03 */
#define PP3 MK(0x3F11566Au, 0xAF25DE2Cu)    /* 6.61375632143793436117e-05 */
#define PP4 MK(0xBEBBBD41u, 0xC5D26BF1u)    /* -1.65339022054652515390e-06 */
#define PP5 MK(0x3E663769u, 0x72BEA4D0u)    /* 4.13813679705723846039e-08 */
#define LG2 MK(0x3FE62E42u, 0xFEFA39EFu)    /* 6.93147180559945286227e-01 */
#define LG2_H MK(0x3FE62E43u, 0x00000000u)  /* 6.93147182464599609375e-01 */
#define LG2_L MK(0xBE205C61u, 0x0CA86C39u)  /* -1.90465429995776804525e-09 */
#define OVT 8.0085662595372944372e-0017     /* -(1024-log2(ovfl+.5ulp)) */
#define CP MK(0x3FEEC709u, 0xDC3A03FDu)     /* 9.61796693925975554329e-01 = 2/(3 ln2) */
#define CP_H MK(0x3FEEC709u, 0xE0000000u)   /* 9.61796700954437255859e-01 */
#define CP_L MK(0xBE3E2FE0u, 0x145B01F5u)   /* -7.02846165095275826516e-09 */
#define IVLN2 MK(0x3FF71547u, 0x652B82FEu)  /* 1.44269504088896338700e+00 */
#define IVLN2_H MK(0x3FF71547u, 0x60000000u) /* 1.44269502162933349609e+00 */
#define IVLN2_L MK(0x3E54AE0Bu, 0xF85DDF44u) /* 1.92596299112661746887e-08 */

double c1o_fd_pow(double x, double y) {
  static const double bp[2] = {1.0, 1.5};
  const double dp_h[2] = {0.0, DP_H1}, dp_l[2] = {0.0, DP_L1};
  const double two53 = 9007199254740992.0, huge = 1.0e300, tiny = 1.0e-300;
  double z, ax, z_h, z_l, p_h, p_l;
  double y1, t1, t2, r, s, t, u, v, w;
  int32_t i, j, k, yisint, n;
  int32_t hx, hy, ix, iy;
  uint32_t lx, ly;

  hx = HI(x); lx = LO(x);
  hy = HI(y); ly = LO(y);
  ix = hx & 0x7fffffff; iy = hy & 0x7fffffff;

  if ((iy | ly) == 0) return 1.0; /* x**0 = 1 */
  if (ix > 0x7ff00000 || ((ix == 0x7ff00000) && (lx != 0)) || iy > 0x7ff00000 || ((iy == 0x7ff00000) && (ly != 0)))
    return x + y; /* NaN */

  /* y an odd integer when x < 0? yisint = 0 no, 1 odd, 2 even */
  yisint = 0;
  if (hx < 0) {
    if (iy >= 0x43400000) yisint = 2;
    else if (iy >= 0x3ff00000) {
      k = (iy >> 20) - 0x3ff;
      if (k > 20) {
        j = (int32_t)(ly >> (52 - k));
        if (((uint32_t)j << (52 - k)) == ly) yisint = 2 - (j & 1);
      } else if (ly == 0) {
        j = iy >> (20 - k);
        if ((j << (20 - k)) == iy) yisint = 2 - (j & 1);
      }
    }
  }

  if (ly == 0) { /* special values of y */
    if (iy == 0x7ff00000) {
      if (((ix - 0x3ff00000) | (int32_t)lx) == 0) return y - y; /* (+-1)**inf is NaN */
      if (ix >= 0x3ff00000) return (hy >= 0) ? y : 0.0;
      return (hy < 0) ? -y : 0.0;
    }
    if (iy == 0x3ff00000) return (hy < 0) ? 1.0 / x : x;
    if (hy == 0x40000000) return x * x;
    if (hy == 0x3fe00000) {
      if (hx >= 0) return sqrt(x);
    }
  }

  ax = fabs(x);
  if (lx == 0) { /* special values of x */
    if (ix == 0x7ff00000 || ix == 0 || ix == 0x3ff00000) {
      z = ax;
      if (hy < 0) z = 1.0 / z;
      if (hx < 0) {
        if (((ix - 0x3ff00000) | yisint) == 0) z = (z - z) / (z - z);
        else if (yisint == 1) z = -z;
      }
      return z;
    }
  }

  n = (hx >> 31) + 1;
  if ((n | yisint) == 0) return (x - x) / (x - x); /* (x<0)**(non-int) */
  s = 1.0;
  if ((n | (yisint - 1)) == 0) s = -1.0;

  if (iy > 0x41e00000) { /* |y| > 2^31 */
    if (iy > 0x43f00000) {
      if (ix <= 0x3fefffff) return (hy < 0) ? huge * huge : tiny * tiny;
      if (ix >= 0x3ff00000) return (hy > 0) ? huge * huge : tiny * tiny;
    }
    if (ix < 0x3fefffff) return (hy < 0) ? s * huge * huge : s * tiny * tiny;
    if (ix > 0x3ff00000) return (hy > 0) ? s * huge * huge : s * tiny * tiny;
    t = ax - 1.0;
    w = (t * t) * (0.5 - t * (0.3333333333333333333333 - t * 0.25));
    u = IVLN2_H * t;
    v = t * IVLN2_L - w * IVLN2;
    t1 = SET_LO(u + v, 0);
    t2 = v - (t1 - u);
  } else {
    double ss, s2, s_h, s_l, t_h, t_l;
    n = 0;
    if (ix < 0x00100000) { ax *= two53; n -= 53; ix = HI(ax); }
    n += ((ix) >> 20) - 0x3ff;
    j = ix & 0x000fffff;
    ix = j | 0x3ff00000;
    if (j <= 0x3988E) k = 0;
    else if (j < 0xBB67A) k = 1;
    else { k = 0; n += 1; ix -= 0x00100000; }
    ax = SET_HI(ax, ix);

    u = ax - bp[k];
    v = 1.0 / (ax + bp[k]);
    ss = u * v;
    s_h = SET_LO(ss, 0);
    t_h = MK((uint32_t)(((ix >> 1) | 0x20000000) + 0x00080000 + (k << 18)), 0);
    t_l = ax - (t_h - bp[k]);
    s_l = v * ((u - s_h * t_h) - s_h * t_l);
    s2 = ss * ss;
    r = s2 * s2 * (PL1 + s2 * (PL2 + s2 * (PL3 + s2 * (PL4 + s2 * (PL5 + s2 * PL6)))));
    r += s_l * (s_h + ss);
    s2 = s_h * s_h;
    t_h = SET_LO(3.0 + s2 + r, 0);
    t_l = r - ((t_h - 3.0) - s2);
    u = s_h * t_h;
    v = s_l * t_h + t_l * ss;
    p_h = SET_LO(u + v, 0);
    p_l = v - (p_h - u);
    z_h = CP_H * p_h;
    z_l = CP_L * p_h + p_l * CP + dp_l[k];
    t = (double)n;
    t1 = SET_LO(((z_h + z_l) + dp_h[k]) + t, 0);
    t2 = z_l - (((t1 - t) - dp_h[k]) - z_h);
  }

  y1 = SET_LO(y, 0);
  p_l = (y - y1) * t1 + y * t2;
  p_h = y1 * t1;
  z = p_l + p_h;
  j = HI(z);
  i = (int32_t)LO(z);
  if (j >= 0x40900000) {
    if (((j - 0x40900000) | i) != 0) return s * huge * huge;
    if (p_l + OVT > z - p_h) return s * huge * huge;
  } else if ((j & 0x7fffffff) >= 0x4090cc00) {
    if (((j - (int32_t)0xc090cc00) | i) != 0) return s * tiny * tiny;
    if (p_l <= z - p_h) return s * tiny * tiny;
  }
  i = j & 0x7fffffff;
  k = (i >> 20) - 0x3ff;
  n = 0;
  if (i > 0x3fe00000) {
    n = j + (0x00100000 >> (k + 1));
    k = ((n & 0x7fffffff) >> 20) - 0x3ff;
    t = MK((uint32_t)(n & ~(0x000fffff >> k)), 0);
    n = ((n & 0x000fffff) | 0x00100000) >> (20 - k);
    if (j < 0) n = -n;
    p_h -= t;
  }
  t = SET_LO(p_l + p_h, 0);
  u = t * LG2_H;
  v = (p_l - (t - p_h)) * LG2 + t * LG2_L;
  z = u + v;
  w = v - (z - u);
  t = z * z;
  t1 = z - t * (PP1 + t * (PP2 + t * (PP3 + t * (PP4 + t * PP5))));
  r = (z * t1) / (t1 - 2.0) - (w + z * w);
  z = 1.0 - (r - z);
  j = HI(z);
  j += (int32_t)((uint32_t)n << 20);
  if ((j >> 20) <= 0) z = scalbn(z, n);
  else z = SET_HI(z, HI(z) + (int32_t)((uint32_t)n << 20));
  return s * z;
}

/* Decimal literals as fdlibm prints them, against the bit patterns above: 0 = all agree. */
int c1o_fd_selfcheck(void) {
  const struct { double bits, dec; } c[] = {
      {S1, -1.66666666666666324348e-01}, {S2, 8.33333333332248946124e-03}, {S3, -1.98412698298579493134e-04},
      {S4, 2.75573137070700676789e-06}, {S5, -2.50507602534068634195e-08}, {S6, 1.58969099521155010221e-10},
      {C1, 4.16666666666666019037e-02}, {C2, -1.38888888888741095749e-03}, {C3, 2.48015872894767294178e-05},
      {C4, -2.75573143513906633035e-07}, {C5, 2.08757232129817482790e-09}, {C6, -1.13596475577881948265e-11},
      {INVPIO2, 6.36619772367581382433e-01}, {PIO2_1, 1.57079632673412561417e+00}, {PIO2_1T, 6.07710050650619224932e-11},
      {PIO2_2, 6.07710050630396597660e-11}, {PIO2_2T, 2.02226624879595063154e-21}, {PIO2_3, 2.02226624871116645580e-21},
      {PIO2_3T, 8.47842766036889956997e-32}, {DP_H1, 5.84962487220764160156e-01}, {DP_L1, 1.35003920212974897128e-08},
      {PL1, 5.99999999999994648725e-01}, {PL2, 4.28571428578550184252e-01}, {PL3, 3.33333329818377432918e-01},
      {PL4, 2.72728123808534006489e-01}, {PL5, 2.30660745775561754067e-01}, {PL6, 2.06975017800338417784e-01},
      {PP1, 1.66666666666666019037e-01}, {PP2, -2.77777777770155933842e-03}, {PP3, 6.61375632143793436117e-05},
      {PP4, -1.65339022054652515390e-06}, {PP5, 4.13813679705723846039e-08}, {LG2, 6.93147180559945286227e-01},
      {LG2_H, 6.93147182464599609375e-01}, {LG2_L, -1.90465429995776804525e-09}, {CP, 9.61796693925975554329e-01},
      {CP_H, 9.61796700954437255859e-01}, {CP_L, -7.02846165095275826516e-09}, {IVLN2, 1.44269504088896338700e+00},
      {IVLN2_H, 1.44269502162933349609e+00}, {IVLN2_L, 1.92596299112661746887e-08},
  };
  int bad = 0;
  for (size_t i = 0; i < sizeof c / sizeof c[0]; i++) bad += memcmp(&c[i].bits, &c[i].dec, 8) != 0;
  /* npio2_hw[n-1] is the high word of n * pi/2 */
  for (int n = 1; n <= 32; n++) bad += HI(n * 1.57079632679489661923) != npio2_hw[n - 1];
  /* pi/2 = pio2_1 + pio2_2 + pio2_3 + ... : the three heads are 33-bit pieces */
  bad += (PIO2_1 + PIO2_1T) != 1.57079632679489661923;
  return bad;
}

static void fd_mdct_table(double *tab, int size, double scale) { /* mdct.js:21-37 with V8's sin / cos */
  const double alpha = (2.0 * M_PI) / (8.0 * size);
  const double omega = (2.0 * M_PI) / size;
  const double scale_root = sqrt(scale / size);
  for (int i = 0; i < size / 4; i++) {
    const double angle = omega * i + alpha;
    tab[2 * i] = scale_root * c1o_fd_cos(angle);
    tab[2 * i + 1] = scale_root * c1o_fd_sin(angle);
  }
}

/* c1o_default_tables with fdlibm's sin / cos / pow in place of the host libm's. */
void c1o_fdlibm_tables(c1o_tables *t) {
  for (int i = 0; i < 32; i++) t->window_short[i] = c1o_fd_sin(((i + 0.5) * M_PI) / 64);
  for (int i = 0; i < 64; i++) t->scale_factors[i] = c1o_fd_pow(2.0, i / 3.0 - 21);
  fd_mdct_table(t->mdct_fwd64, 64, 0.5);
  fd_mdct_table(t->mdct_fwd256, 256, 0.5);
  fd_mdct_table(t->mdct_fwd512, 512, 1.0);
  fd_mdct_table(t->mdct_inv64, 64, 64 * 8);
  fd_mdct_table(t->mdct_inv256, 256, 256 * 8);
  fd_mdct_table(t->mdct_inv512, 512, 512 * 4);
  for (int k = 0; k < 8; k++) {
    const int stride = 2 << k;
    const double angle = (-2 * M_PI) / stride;
    t->fft_w[k][0] = c1o_fd_cos(angle);
    t->fft_w[k][1] = c1o_fd_sin(angle);
  }
}
