// ef_libm_f32.cuh -- sinf/cosf that reproduce the HOST libm of the CPU reference bit for bit.
//
// The reference's CPU HashSIFT rotates the patch with glibc's cosf/sinf (hash_sift.cpp:119-122).  Those
// are not correctly rounded (about 1e-3 of the arguments in [0, 2pi] differ from round(cos(x)) by one
// ulp), so "bit-exact vs the CPU reference" needs the same function, not a better one.  This is the
// algorithm of glibc 2.39 x86-64 (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h -- the ARM
// optimized-routines sincosf), restated from its published description with the evaluation order of the
// FMA ifunc variant the library selects on every AVX2/FMA host (__sinf_fma/__cosf_fma): double-precision
// range reduction by pi/2 and two degree-limited polynomials.  Coefficients are the published ones
// (they also sit in libm.so.6's .rodata; tests/test_libm_port.py brute-forces this file against the
// host libm over every float in [0, 2*pi] and a sample of large and negative arguments).
//
// Device and host build (the host build exists only for that test).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define EF_HD __host__ __device__ __forceinline__
#else
#define EF_HD static inline
#endif

namespace ef_libm {

// 2/pi bits for the large-argument reduction (glibc __inv_pio4)
#define EF_INV_PIO4_INIT { 0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, \
                           0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, \
                           0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041 }
#if defined(__CUDACC__)
__device__ __constant__ uint32_t ef_inv_pio4_dev[24] = EF_INV_PIO4_INIT;
#endif

EF_HD double fma_(double a, double b, double c) { return fma(a, b, c); }

EF_HD uint32_t f32_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

// cosine / sine polynomials on [-pi/4, pi/4]; `neg` selects the second coefficient table (all cosine
// coefficients negated, sine unchanged), which is an exact sign flip of the cosine result
EF_HD double cos_poly(double x2)
{
    const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
    const double x4 = x2 * x2;
    const double c1_ = fma_(c1, x2, c0);
    const double c2_ = fma_(c4, x2, c3);
    const double x6 = x2 * x4;
    const double c = fma_(x4, c2, c1_);
    return fma_(c2_, x6, c);
}
EF_HD double sin_poly(double x, double x2)
{
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    const double s1_ = fma_(s3, x2, s2);
    const double x3 = x2 * x;
    const double x7 = x2 * x3;
    const double s = fma_(x3, s1, x);
    return fma_(s1_, x7, s);
}
// sinf_poly(x*s, x*x, p, n): n even -> sine polynomial, n odd -> cosine polynomial
EF_HD float poly(double xs, double x2, bool neg, int n)
{
    if ((n & 1) == 0) return (float)sin_poly(xs, x2);
    const double c = cos_poly(x2);
    return (float)(neg ? -c : c);
}

EF_HD double reduce_fast(double x, int* np)
{
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double r = x * hpi_inv;
#if defined(__CUDA_ARCH__)
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
#else
    const int n = ((int32_t)r + 0x800000) >> 24;
#endif
    *np = n;
    return fma_(-(double)n, hpi, x);
}

EF_HD double reduce_large(uint32_t xi, int* np)
{
#if defined(__CUDA_ARCH__)
    const uint32_t* inv_pio4 = ef_inv_pio4_dev;
#else
    static const uint32_t inv_pio4[24] = EF_INV_PIO4_INIT;
#endif
    const double pi63 = 0x1.921FB54442D18p-62;
    const uint32_t* arr = &inv_pio4[(xi >> 26) & 15];
    const int shift = (xi >> 23) & 7;
    uint64_t n, res0, res1, res2;
    xi = (xi & 0xffffff) | 0x800000;
    xi <<= shift;
    res0 = (uint32_t)(xi * arr[0]);
    res1 = (uint64_t)xi * arr[4];
    res2 = (uint64_t)xi * arr[8];
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    n = (res0 + (1ULL << 61)) >> 62;
    res0 -= n << 62;
    const double x = (double)(int64_t)res0;
    *np = (int)n;
    return x * pi63;
}

EF_HD float sign_of(int q) { return (q == 1 || q == 2) ? -1.0f : 1.0f; } // { 1, -1, -1, 1 }

EF_HD float cosf_glibc(float y)
{
    double x = (double)y;
    const uint32_t ix = f32_bits(y);
    const uint32_t top = (ix >> 20) & 0x7ff;
    int n;
    if (top <= 0x3f3) {                  // |y| < pi/4
        const double x2 = x * x;
        if (top <= 0x397) return 1.0f;   // |y| < 2^-12
        return (float)cos_poly(x2);
    } else if (top <= 0x42e) {           // |y| < 120
        x = reduce_fast(x, &n);
        const double s = (double)sign_of(n & 3);
        return poly(x * s, x * x, (n & 2) != 0, n ^ 1);
    } else if (top <= 0x7f7) {
        const int sgn = (int)(ix >> 31);
        x = reduce_large(ix, &n);
        const double s = (double)sign_of((n + sgn) & 3);
        return poly(x * s, x * x, ((n + sgn) & 2) != 0, n ^ 1);
    }
    return y - y; // inf/nan -> nan
}

EF_HD float sinf_glibc(float y)
{
    double x = (double)y;
    const uint32_t ix = f32_bits(y);
    const uint32_t top = (ix >> 20) & 0x7ff;
    int n;
    if (top <= 0x3f3) {
        const double x2 = x * x;
        if (top <= 0x397) return y;
        return (float)sin_poly(x, x2);
    } else if (top <= 0x42e) {
        x = reduce_fast(x, &n);
        const double s = (double)sign_of(n & 3);
        return poly(x * s, x * x, (n & 2) != 0, n);
    } else if (top <= 0x7f7) {
        const int sgn = (int)(ix >> 31);
        x = reduce_large(ix, &n);
        const double s = (double)sign_of((n + sgn) & 3);
        return poly(x * s, x * x, ((n + sgn) & 2) != 0, n);
    }
    return y - y;
}

} // namespace ef_libm
