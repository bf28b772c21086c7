// cuda_host_shim.h -- the CUDA intrinsics that 3dscan_b200/csrc/scan3d_math.cuh and scan3d_fused_math.cuh use,
// as plain IEEE host functions, so that the kernels' arithmetic source compiles with g++ for the CPU tests
// (tests/fused_math_host.cpp; built with -ffp-contract=off so that no host operation is fused either).
// TEST INFRASTRUCTURE: nothing in the product includes this file.
//
// Exact equivalents: every *_rn arithmetic intrinsic (IEEE round-to-nearest, as the host's + - * /), fma (libm's
// is correctly rounded), the bit casts, __byte_perm, the saturating conversions.  Two are approximations on the
// device and here: __fdividef (only selects the arctangent table row) and the reciprocal seed rcp.approx.ftz.f64
// (s3d_host_rcp_seed: exact reciprocal truncated to 20 mantissa bits); the results they feed are corrected to
// ~1 ulp in double afterwards, so the float outputs agree except when a value sits within ~2^-50 of a float
// rounding boundary (the GPU tests count those cases on the device itself: none on 2^24 + 8 M inputs).
#pragma once
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fdividef(float a, float b) { return a / b; }

static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
}
static inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(u >> 32); }
static inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }

static inline float __double2float_rn(double d) { return (float)d; }
static inline float __double2float_rz(double d)
{
    float f = (float)d;
    if (f == f && fabs((double)f) > fabs(d)) f = nextafterf(f, 0.0f);
    return f;
}
// cvt.rni.s32: round half to even, saturating, NaN -> 0
static inline int __float2int_rn(float f)
{
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)lrintf(f);
}
static inline int __double2int_rn(double d)
{
    if (!(d == d)) return 0;
    if (d >= 2147483647.5) return 2147483647;
    if (d <= -2147483648.5) return (int)0x80000000;
    return (int)lrint(d);
}
static inline long long __double2ll_rn(double d)
{
    if (!(d == d)) return (long long)0x8000000000000000ull;
    if (d >= 9223372036854775808.0) return 9223372036854775807LL;
    if (d <= -9223372036854775808.0) return (long long)0x8000000000000000ull;
    return llrint(d);
}
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        const uint32_t sel = (s >> (4 * i)) & 0xf;
        uint32_t b = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0x00;   // msb replication mode
        r |= b << (8 * i);
    }
    return r;
}
template <class T> static inline T min(T a, T b) { return b < a ? b : a; }
template <class T> static inline T max(T a, T b) { return a < b ? b : a; }

// rcp.approx.ftz.f64: reciprocal with ~20 good bits, low word zero
double s3d_host_rcp_seed(double x)
{
    double r = 1.0 / x;
    uint64_t u;
    memcpy(&u, &r, 8);
    u &= 0xffffffff00000000ull;
    memcpy(&r, &u, 8);
    return r;
}
