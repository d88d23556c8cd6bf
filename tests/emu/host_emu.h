/*
 * host_emu.h — just enough of the CUDA device vocabulary to compile
 * remy_b200/csrc/sim_kernel.cuh with g++ (TEST INFRASTRUCTURE).
 *
 * Purpose: the kernel restructures the reference's event loop heavily (cached
 * event minima, lazily materialised shuffle, mailbox-in-ring ...).  This build
 * lets the CPU test-suite check that restructuring against the oracle on
 * machines without a GPU.  It is never loaded by the product path
 * (remy_b200/abi.py::Device opens only the nvcc-built library).
 *
 * Built with -ffp-contract=off, so the plain operators below round exactly like
 * the _rn intrinsics they stand in for.
 */
#ifndef REMY_HOST_EMU_H
#define REMY_HOST_EMU_H
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>

#define __device__
#define __global__
#define __forceinline__ inline
#define __align__(n) alignas(n)
#define __launch_bounds__(...)
#define __shared__
#define __syncthreads() ((void)0)

struct emu_dim3 { unsigned x, y, z; };
static const emu_dim3 blockDim = {1, 1, 1}, threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int32_t __double2int_rz(double v)
{ /* cvt.rzi.s32.f64 saturates; NaN -> 0 */
  if (v != v) return 0;
  if (v >= 2147483647.0) return 2147483647;
  if (v <= -2147483648.0) return (int32_t)0x80000000;
  return (int32_t)v;
}
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, sizeof r); return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline long long __double2ll_rn(double v) { return llrint(v); }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
using std::max;
using std::min;
#endif
