// hd.cuh -- host/device portability shims.
//
// All arithmetic of the PBS kernels lives in `B200_HD` functions that take the
// thread index explicitly, so the very same code can be (a) inlined into the
// sm_100a kernels and (b) compiled by g++ into the CPU "CTA emulator"
// (tests/emu) that replays a thread block phase by phase.  The emulator is
// test infrastructure only; the product path always runs the CUDA kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#define B200_D __device__ __forceinline__
#else
#define B200_HD inline
#define B200_D inline
#endif

struct alignas(16) cplx {
  double re, im;
};

B200_HD cplx cmake(double re, double im) {
  cplx r;
  r.re = re;
  r.im = im;
  return r;
}
B200_HD cplx cadd(cplx a, cplx b) { return cmake(a.re + b.re, a.im + b.im); }
B200_HD cplx csub(cplx a, cplx b) { return cmake(a.re - b.re, a.im - b.im); }
// a * b
B200_HD cplx cmul(cplx a, cplx b) {
  return cmake(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
// a * conj(b)
B200_HD cplx cmulc(cplx a, cplx b) {
  return cmake(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
}
// acc + a * b
B200_HD cplx cfma(cplx a, cplx b, cplx acc) {
  return cmake(acc.re + a.re * b.re - a.im * b.im,
               acc.im + a.re * b.im + a.im * b.re);
}
// acc + a * conj(b)
B200_HD cplx cfmac(cplx a, cplx b, cplx acc) {
  return cmake(acc.re + a.re * b.re + a.im * b.im,
               acc.im + a.im * b.re - a.re * b.im);
}
// i * a
B200_HD cplx cmuli(cplx a) { return cmake(-a.im, a.re); }
// 2a - t as ONE fused multiply-add.  Written `2.0 * a - t` the compiler turns
// the product into a + a and emits two DADDs; the fp64 pipe is the kernel's
// scarcest resource, so the FMA is forced.
B200_HD double two_a_minus(double a, double t) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(2.0, a, -t);
#else
  return __builtin_fma(2.0, a, -t);
#endif
}

// ---- scalar conversions ---------------------------------------------------
B200_HD double int_to_double(int32_t x) {
#if defined(__CUDA_ARCH__)
  return __int2double_rn(x);
#else
  return (double)x;
#endif
}
// the same value through the fp64 ADD pipe instead of I2F.F64.S32: plant the
// biased integer in the mantissa of 2^52 + 2^31 and subtract the constant (exact
// for every int32).  I2F.F64 issues once per ~14 cycles per sub-partition on
// B200 (tools/micro/pipes.cu), a DADD once per 2.
B200_HD double int_to_double_splice(int32_t x) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)) -
         4503601774854144.0; // 2^52 + 2^31
#else
  return (double)x;
#endif
}
B200_HD double ll_to_double(int64_t x) {
#if defined(__CUDA_ARCH__)
  return __ll2double_rn(x);
#else
  return (double)x;
#endif
}

// from_torus (tfhe/src/core_crypto/commons/math/torus/mod.rs:75-81):
// frac = x - round(x); round(frac * 2^64) as i64 (saturating) as u64.
// round-to-nearest-even is used for the first rounding (differs from the
// reference's half-away-from-zero only on exact ties, where frac = +-0.5
// either way).
B200_HD uint64_t double_to_torus64(double x) {
  const double magic = 6755399441055744.0; // 1.5 * 2^52, valid for |x| < 2^51
  const double r = (x + magic) - magic;
  const double f = (x - r) * 18446744073709551616.0;
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double2ll_rn(f);
#else
  // mirror the saturating device conversion
  double g = __builtin_rint(f);
  int64_t s;
  if (g >= 9223372036854775808.0)
    s = INT64_MAX;
  else if (g <= -9223372036854775808.0)
    s = INT64_MIN;
  else
    s = (int64_t)g;
  return (uint64_t)s;
#endif
}
