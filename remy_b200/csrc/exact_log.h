/*
 * exact_log.h — glibc 2.39 x86-64 log(double), FMA build, restated operation
 * by operation so that host C, the CPU oracle and the sm_100a kernel all return
 * the same bits as the libm the reference links against.
 *
 * Why: the reference samples on/off switch times with
 * std::exponential_distribution (reference src/exponential.hh:22), which is
 * -std::log(1 - u)/lambda in libstdc++ (bits/random.h:4904).  One ULP of
 * difference in a switch time shifts every later event of that sender, so the
 * device log must equal glibc's bit for bit; CUDA's log() is only <= 1 ULP.
 *
 * What is restated: Szabolcs Nagy's table-driven log (glibc
 * sysdeps/ieee754/dbl-64/e_log.c, N = 128) as compiled into the `__log_fma`
 * ifunc variant of libm.so.6 — the variant glibc selects on any x86-64 CPU with
 * FMA+AVX2.  GCC contracts the source expressions of that build into fused
 * multiply-adds; the exact fusion pattern below was read off the machine code
 * of the function (objdump of libm.so.6 at 0x79d50), not off the C source,
 * because the rounding depends on which products are fused.
 *
 * Domain handled: finite x > 0, normal (the simulator only calls it with
 * x = 1 - u, u in [0, 1 - 2^-53], i.e. x in [2^-53, 1]).  x == 1 returns +0.
 * Subnormals, zero, negatives, inf and NaN are out of the simulator's domain
 * and are NOT handled.
 *
 * The constants come from glibc_log_data.h (extracted from the same libm).
 */
#ifndef REMY_EXACT_LOG_H
#define REMY_EXACT_LOG_H

#include <stdint.h>
#include "glibc_log_data.h"

#if defined(__CUDA_ARCH__)
#define REMY_LOG_FN __device__ __forceinline__
#define REMY_FMA(a, b, c) __fma_rn((a), (b), (c))
#define REMY_MUL(a, b) __dmul_rn((a), (b))
#define REMY_ADD(a, b) __dadd_rn((a), (b))
#define REMY_SUB(a, b) __dsub_rn((a), (b))
#define REMY_AS_DOUBLE(u) __longlong_as_double((long long)(u))
#define REMY_AS_U64(d) ((uint64_t)__double_as_longlong(d))
#define REMY_LOG_TABLE_QUAL __device__ const
#else
#include <math.h>
#include <string.h>
#define REMY_LOG_FN static inline
#define REMY_FMA(a, b, c) fma((a), (b), (c))
#define REMY_MUL(a, b) ((a) * (b))
#define REMY_ADD(a, b) ((a) + (b))
#define REMY_SUB(a, b) ((a) - (b))
static inline double remy_as_double_(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t remy_as_u64_(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
#define REMY_AS_DOUBLE(u) remy_as_double_(u)
#define REMY_AS_U64(d) remy_as_u64_(d)
#define REMY_LOG_TABLE_QUAL static const
#endif

REMY_LOG_TABLE_QUAL uint64_t remy_log_A_[5] = GLIBC_LOG_POLY_A;
REMY_LOG_TABLE_QUAL uint64_t remy_log_B_[11] = GLIBC_LOG_POLY_B;
REMY_LOG_TABLE_QUAL uint64_t remy_log_T_[256] = GLIBC_LOG_TAB;

REMY_LOG_FN double remy_exact_log(double x)
{
  const uint64_t ix = REMY_AS_U64(x);
#define A_(i) REMY_AS_DOUBLE(remy_log_A_[i])
#define B_(i) REMY_AS_DOUBLE(remy_log_B_[i])
  /* 1 - 2^-4 <= x < 1 + 0x1.09p-4 : log1p polynomial around 1 */
  if (ix - 0x3fee000000000000ULL < 0x0003090000000000ULL) {
    if (ix == 0x3ff0000000000000ULL) return 0.0;
    const double r = REMY_SUB(x, 1.0);
    double a = REMY_FMA(r, B_(2), B_(1));
    double b = REMY_FMA(r, B_(5), B_(4));
    const double r2 = REMY_MUL(r, r);
    double c = REMY_FMA(r, B_(8), B_(7));
    a = REMY_FMA(r2, B_(3), a);
    b = REMY_FMA(r2, B_(6), b);
    const double r3 = REMY_MUL(r, r2);
    c = REMY_FMA(r2, B_(9), c);
    c = REMY_FMA(r3, B_(10), c);
    double p = REMY_FMA(c, r3, b);
    p = REMY_FMA(p, r3, a);
    /* split r = rhi + rlo so that rhi*rhi*B[0] is exact enough */
    const double two27 = 134217728.0; /* 0x1p27 */
    const double t = REMY_FMA(r, two27, r);          /* r + r*2^27 */
    const double rhi = REMY_FMA(-two27, r, t);       /* t - 2^27*r */
    const double rhi2 = REMY_MUL(rhi, rhi);
    const double rlo = REMY_SUB(r, rhi);
    const double hi = REMY_FMA(rhi2, B_(0), r);
    double lo = REMY_FMA(rhi2, B_(0), REMY_SUB(r, hi));
    lo = REMY_FMA(REMY_MUL(B_(0), rlo), REMY_ADD(r, rhi), lo);
    return REMY_ADD(hi, REMY_FMA(p, r3, lo));
  }
  /* x = 2^k z, z in [OFF, 2 OFF) */
  const uint64_t tmp = ix - 0x3fe6000000000000ULL;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
  const double invc = REMY_AS_DOUBLE(remy_log_T_[2 * i]);
  const double logc = REMY_AS_DOUBLE(remy_log_T_[2 * i + 1]);
  const double z = REMY_AS_DOUBLE(iz);
  const double kd = (double)k;
  const double ln2hi = REMY_AS_DOUBLE(GLIBC_LOG_LN2HI);
  const double ln2lo = REMY_AS_DOUBLE(GLIBC_LOG_LN2LO);
  const double w = REMY_FMA(kd, ln2hi, logc);
  const double r = REMY_FMA(z, invc, -1.0);
  const double p1 = REMY_FMA(r, A_(2), A_(1));
  const double hi = REMY_ADD(r, w);
  const double r2 = REMY_MUL(r, r);
  double lo = REMY_ADD(REMY_SUB(w, hi), r);
  lo = REMY_FMA(kd, ln2lo, lo);
  const double r3 = REMY_MUL(r, r2);
  const double p2 = REMY_FMA(r, A_(4), A_(3));
  const double q = REMY_FMA(r2, A_(0), lo);
  const double p = REMY_FMA(p2, r2, p1);
  return REMY_ADD(REMY_FMA(r3, p, q), hi);
#undef A_
#undef B_
}

#endif /* REMY_EXACT_LOG_H */
