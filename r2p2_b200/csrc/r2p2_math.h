// r2p2_math.h -- deterministic FP64 math shared by the CUDA kernels (device) and by host code
// that must reproduce the device bits (the port-mode oracle, unit tests).
//
// Why this exists.  The reference does all geometry in CPython floats: `math.sin/cos/atan2`
// (glibc libm), `np.linalg.norm` (OpenBLAS ddot, FMA-contracted tail) and `round()` (half to
// even).  Integer outcomes (which pixel a sonar ray hits, whether the k = limit sample is tested)
// sit on ulp-level knife edges, so the device must produce the *same
// doubles* as libm, not merely accurate ones.
//
//   * r2p2_sin / r2p2_cos : operation-for-operation restatement of glibc 2.39
//     `__sin_fma` / `__cos_fma` (sysdeps/ieee754/dbl-64/s_sin.c as built for x86-64 with FMA:
//     the contraction pattern was read from the disassembly of libm.so.6 and is pinned by
//     tests/test_math_port.py against libm on >1e7 arguments).  Arguments with
//     |x| >= 105414350 (glibc's __branred path) are not supported: *ok is cleared and the
//     caller raises the robot's fault flag.
//   * r2p2_atan2, r2p2_tanh, r2p2_exp : our own <= 2 ulp implementations (constants generated
//     by tools/gen_math_consts.py).  They are NOT bit-identical to libm / numpy; only the
//     1e-5 tolerance applies to what they feed (SURVEY.md H23 iii, a10).
//   * r2p2_norm2 : sqrt(fma(dy,dy, dx*dx)) == np.linalg.norm of a 2-vector on this image.
//
// Every operation is an explicitly rounded IEEE-754 double operation: no contraction is left to
// the compiler (device: __dmul_rn/__dadd_rn/__fma_rn; host: build with -ffp-contract=off).
//
// Provenance and licence of the sin / cos part.  r2p2__do_sin, r2p2__do_cos, r2p2__reduce_sincos, r2p2__taylor_sin, r2p2_sin,
// r2p2_cos and r2p2_sincos follow the algorithm of the GNU C Library's IBM Accurate Mathematical Library routines
// (sysdeps/ieee754/dbl-64/s_sin.c, dosincos / branred helpers; Copyright (C) 2001-2024 Free Software Foundation, Inc.),
// and the polynomial constants and the 440-entry table in r2p2_sincostab.h are the values of glibc 2.39's libm
// (read out of libm.so.6 by tools/gen_sincostab.py).  The GNU C Library is free software; you can redistribute it
// and/or modify it under the terms of the GNU Lesser General Public License as published by the Free Software
// Foundation; either version 2.1 of the License, or (at your option) any later version.  These derived parts are
// distributed under the same terms, WITHOUT ANY WARRANTY; see <https://www.gnu.org/licenses/> for the licence text.
// Everything else in this file (atan2, exp, tanh, logistic, norm) is this project's own work.
#pragma once

#include <stdint.h>
#include <string.h>
#include <math.h>

#include "r2p2_mathconsts.h"
#include "r2p2_sincostab.h"
#include "r2p2_tanhtab.h"

#if defined(__CUDACC__)
#define R2P2_HD __host__ __device__ __forceinline__
#else
#define R2P2_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define R2P2_MUL(a, b) __dmul_rn((a), (b))
#define R2P2_ADD(a, b) __dadd_rn((a), (b))
#define R2P2_SUB(a, b) __dsub_rn((a), (b))
#define R2P2_FMA(a, b, c) __fma_rn((a), (b), (c))
#define R2P2_DIV(a, b) __ddiv_rn((a), (b))
#define R2P2_SQRT(a) __dsqrt_rn((a))
#else
#define R2P2_MUL(a, b) ((a) * (b))
#define R2P2_ADD(a, b) ((a) + (b))
#define R2P2_SUB(a, b) ((a) - (b))
#define R2P2_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define R2P2_DIV(a, b) ((a) / (b))
#define R2P2_SQRT(a) __builtin_sqrt((a))
#endif

// ---------------------------------------------------------------------------------------------
// constant tables: __constant__ copies on the device, plain statics on the host
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
static __device__ __constant__ double r2p2_d_atan_hi[9] = {R2P2_ATAN_HI_VALUES};
static __device__ __constant__ double r2p2_d_atan_lo[9] = {R2P2_ATAN_LO_VALUES};
#endif
static const double r2p2_h_atan_hi[9] = {R2P2_ATAN_HI_VALUES};
static const double r2p2_h_atan_lo[9] = {R2P2_ATAN_LO_VALUES};
static const double r2p2_h_sincostab[R2P2_SINCOSTAB_LEN] = {R2P2_SINCOSTAB_VALUES};
static const double r2p2_h_tanhtab[R2P2_TANH_NINT * R2P2_TANH_ROW] = {R2P2_TANHTAB_VALUES};
#if defined(__CUDACC__)
static __device__ const double r2p2_d_tanhtab[R2P2_TANH_NINT * R2P2_TANH_ROW] = {R2P2_TANHTAB_VALUES};
#endif
#if defined(__CUDA_ARCH__)
#define R2P2_TANHTAB r2p2_d_tanhtab
#else
#define R2P2_TANHTAB r2p2_h_tanhtab
#endif
#if defined(__CUDA_ARCH__)
#define R2P2_ATAN_HI(i) r2p2_d_atan_hi[(i)]
#define R2P2_ATAN_LO(i) r2p2_d_atan_lo[(i)]
#else
#define R2P2_ATAN_HI(i) r2p2_h_atan_hi[(i)]
#define R2P2_ATAN_LO(i) r2p2_h_atan_lo[(i)]
#endif

R2P2_HD uint64_t r2p2_bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
R2P2_HD double r2p2_from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
R2P2_HD int32_t r2p2_lo32(double x) { return (int32_t)(uint32_t)(r2p2_bits(x) & 0xffffffffu); }
R2P2_HD int32_t r2p2_hi32(double x) { return (int32_t)(uint32_t)(r2p2_bits(x) >> 32); }
R2P2_HD double r2p2_fabs(double x) { return r2p2_from_bits(r2p2_bits(x) & 0x7fffffffffffffffULL); }
R2P2_HD double r2p2_copysign(double mag, double sgn) {
    return r2p2_from_bits((r2p2_bits(mag) & 0x7fffffffffffffffULL) | (r2p2_bits(sgn) & 0x8000000000000000ULL));
}
R2P2_HD double r2p2_neg(double x) { return r2p2_from_bits(r2p2_bits(x) ^ 0x8000000000000000ULL); }

// np.linalg.norm((dx, dy)) on this image == sqrt(fma(dy, dy, rn(dx*dx)))   
R2P2_HD double r2p2_norm2sq(double dx, double dy) { return R2P2_FMA(dy, dy, R2P2_MUL(dx, dx)); }
R2P2_HD double r2p2_norm2(double dx, double dy) { return R2P2_SQRT(r2p2_norm2sq(dx, dy)); }

// ---------------------------------------------------------------------------------------------
// glibc 2.39 sin / cos (FMA build), restated
// ---------------------------------------------------------------------------------------------
#define R2P2_S_BIG 0x1.8p45                       /* 1.5 * 2^45: ulp = 1/128 */
#define R2P2_S_SN3 (-0x1.5555555555515p-3)
#define R2P2_S_SN5 0x1.11110e829872fp-7
#define R2P2_S_CS2 0x1.0p-1
#define R2P2_S_CS4 (-0x1.5555555555535p-5)
#define R2P2_S_CS6 0x1.6c16bedd9e239p-10
#define R2P2_S_S1 (-0x1.5555555555555p-3)
#define R2P2_S_S2 0x1.1111111110ecep-7
#define R2P2_S_S3 (-0x1.a01a019db08b8p-13)
#define R2P2_S_S4 0x1.71de27b9a7ed9p-19
#define R2P2_S_S5 (-0x1.addffc2fcdf59p-26)
#define R2P2_S_TOINT 0x1.8p52
#define R2P2_S_HPINV 0x1.45f306dc9c883p-1
#define R2P2_S_MP1 0x1.921fb58000000p+0
#define R2P2_S_MP2 (-0x1.dde973c000000p-27)
#define R2P2_S_PP3 (-0x1.cb3b398000000p-55)
#define R2P2_S_PP4 (-0x1.d747f23e32ed7p-83)
#define R2P2_S_HP0 0x1.921fb54442d18p+0
#define R2P2_S_HP1 0x1.1a62633145c07p-54

R2P2_HD double r2p2__taylor_sin(double xx, double x, double dx) {
    double p = R2P2_FMA(xx, R2P2_S_S5, R2P2_S_S4);
    p = R2P2_FMA(xx, p, R2P2_S_S3);
    p = R2P2_FMA(xx, p, R2P2_S_S2);
    p = R2P2_FMA(xx, p, R2P2_S_S1);
    double t = R2P2_FMA(xx, R2P2_FMA(p, x, r2p2_neg(R2P2_MUL(0.5, dx))), dx);
    return R2P2_ADD(x, t);
}

R2P2_HD double r2p2__do_sin(double x, double dx, const double* tab) {
    const double xold = x;
    double ax = r2p2_fabs(x);
    if (ax < 0.126) return r2p2__taylor_sin(R2P2_MUL(x, x), x, dx);
    if (!(x > 0.0)) dx = r2p2_neg(dx);
    const double u = R2P2_ADD(ax, R2P2_S_BIG);
    x = R2P2_SUB(ax, R2P2_SUB(u, R2P2_S_BIG));
    const double xx = R2P2_MUL(x, x);
    const double s = R2P2_ADD(x, R2P2_FMA(R2P2_MUL(x, xx), R2P2_FMA(xx, R2P2_S_SN5, R2P2_S_SN3), dx));
    const double c = R2P2_FMA(x, dx, R2P2_MUL(xx, R2P2_FMA(xx, R2P2_FMA(xx, R2P2_S_CS6, R2P2_S_CS4), R2P2_S_CS2)));
    const int k = r2p2_lo32(u) * 4;
    const double sn = tab[k], ssn = tab[k + 1], cs = tab[k + 2], ccs = tab[k + 3];
    const double cor = R2P2_FMA(s, cs, R2P2_FMA(r2p2_neg(c), sn, R2P2_FMA(s, ccs, ssn)));
    return r2p2_copysign(R2P2_ADD(sn, cor), xold);
}

R2P2_HD double r2p2__do_cos(double x, double dx, const double* tab) {
    if (x < 0.0) dx = r2p2_neg(dx);
    const double ax = r2p2_fabs(x);
    const double u = R2P2_ADD(ax, R2P2_S_BIG);
    x = R2P2_ADD(R2P2_SUB(ax, R2P2_SUB(u, R2P2_S_BIG)), dx);
    const double xx = R2P2_MUL(x, x);
    const double s = R2P2_FMA(R2P2_MUL(x, xx), R2P2_FMA(xx, R2P2_S_SN5, R2P2_S_SN3), x);
    const double c = R2P2_MUL(xx, R2P2_FMA(xx, R2P2_FMA(xx, R2P2_S_CS6, R2P2_S_CS4), R2P2_S_CS2));
    const int k = r2p2_lo32(u) * 4;
    const double sn = tab[k], ssn = tab[k + 1], cs = tab[k + 2], ccs = tab[k + 3];
    const double cor = R2P2_FMA(r2p2_neg(s), sn, R2P2_FMA(r2p2_neg(c), cs, R2P2_FMA(r2p2_neg(s), ssn, ccs)));
    return R2P2_ADD(cs, cor);
}

// Cody-Waite reduction by pi/2 for 2.426265 < |x| < 105414350; returns the quadrant.
R2P2_HD int r2p2__reduce_sincos(double x, double* a, double* da) {
    const double t = R2P2_FMA(x, R2P2_S_HPINV, R2P2_S_TOINT);
    const double xn = R2P2_SUB(t, R2P2_S_TOINT);
    const double nxn = r2p2_neg(xn);
    const double y = R2P2_FMA(nxn, R2P2_S_MP2, R2P2_FMA(nxn, R2P2_S_MP1, x));
    const int n = r2p2_lo32(t) & 3;
    const double t2 = R2P2_FMA(nxn, R2P2_S_PP3, y);
    double db = R2P2_FMA(nxn, R2P2_S_PP3, R2P2_SUB(y, t2));
    const double b = R2P2_FMA(nxn, R2P2_S_PP4, t2);
    db = R2P2_ADD(db, R2P2_FMA(nxn, R2P2_S_PP4, R2P2_SUB(t2, b)));
    *a = b;
    *da = db;
    return n;
}

R2P2_HD double r2p2__do_sincos(double a, double da, int n, const double* tab) {
    double r = (n & 1) ? r2p2__do_cos(a, da, tab) : r2p2__do_sin(a, da, tab);
    return (n & 2) ? r2p2_neg(r) : r;
}

// ok (may be NULL) is cleared when |x| is outside the supported range (or not finite).
R2P2_HD double r2p2_sin(double x, const double* tab, int* ok) {
    const int32_t k = r2p2_hi32(x) & 0x7fffffff;
    if (k < 0x3e500000) return x;
    if (k < 0x3feb6000) return r2p2__do_sin(x, 0.0, tab);
    if (k < 0x400368fd) {
        const double t = R2P2_SUB(R2P2_S_HP0, r2p2_fabs(x));
        return r2p2_copysign(r2p2__do_cos(t, R2P2_S_HP1, tab), x);
    }
    if (k < 0x419921FB) {
        double a, da;
        const int n = r2p2__reduce_sincos(x, &a, &da);
        return r2p2__do_sincos(a, da, n, tab);
    }
    if (ok) *ok = 0;
    return R2P2_SUB(x, x) / R2P2_SUB(x, x);   // NaN
}

R2P2_HD double r2p2_cos(double x, const double* tab, int* ok) {
    const int32_t k = r2p2_hi32(x) & 0x7fffffff;
    if (k < 0x3e400000) return 1.0;
    if (k < 0x3feb6000) return r2p2__do_cos(x, 0.0, tab);
    if (k < 0x400368fd) {
        const double y = R2P2_SUB(R2P2_S_HP0, r2p2_fabs(x));
        const double a = R2P2_ADD(y, R2P2_S_HP1);
        const double da = R2P2_ADD(R2P2_SUB(y, a), R2P2_S_HP1);
        return r2p2__do_sin(a, da, tab);
    }
    if (k < 0x419921FB) {
        double a, da;
        const int n = r2p2__reduce_sincos(x, &a, &da);
        return r2p2__do_sincos(a, da, n + 1, tab);
    }
    if (ok) *ok = 0;
    return R2P2_SUB(x, x) / R2P2_SUB(x, x);
}

// Both at once, results identical to the two calls above, without data-dependent branches on the common paths.
// In every range of glibc's sin and cos the pair needs exactly ONE do_sin and ONE do_cos evaluation:
//   |x| < 0.855469           sin = do_sin(x, 0)                  cos = do_cos(x, 0)
//   |x| < 2.426265           sin = +-do_cos(hp0 - |x|, hp1)      cos = do_sin(a, da), a + da = (hp0 - |x|) + hp1
//   |x| < 105414350          (a, da, n) = reduce(x): quadrant n for sin, n + 1 for cos: one is do_sin(a, da), the other do_cos(a, da)
// so the three ranges only select the arguments and route the two results.  Lanes of a warp that hold angles of
// different ranges (the rays of a sonar fan; robots of a population) execute one instruction stream instead of up to
// six; inside do_sin the Taylor branch (|x| < 0.126) is evaluated next to the table branch and selected.
// rowlen: doubles between consecutive table rows (4 = glibc's layout; the kernels keep a copy with rows padded to 5 in shared
// memory: with a stride of 8 words the rows of 32 lanes fall on 4 bank groups -- 8-way conflicts on each of the 8 loads
// -- with 10 words on 16).
R2P2_HD void r2p2_sincos_rs(double x, const double* tab, int rowlen, double* s, double* c, int* ok) {
    const int32_t k = r2p2_hi32(x) & 0x7fffffff;
    if (k >= 0x419921FB) {                                  // rare: unsupported range / not finite
        if (ok) *ok = 0;
        *s = R2P2_SUB(x, x) / R2P2_SUB(x, x);
        *c = *s;
        return;
    }
    const int r1 = k < 0x3feb6000, r2 = !r1 && (k < 0x400368fd);
    const double ax = r2p2_fabs(x);
    // range 3: Cody-Waite reduction (always evaluated; harmless for small |x|)
    double a3, da3;
    const int n = r2p2__reduce_sincos(x, &a3, &da3);
    // range 2
    const double t2 = R2P2_SUB(R2P2_S_HP0, ax);
    const double a2 = R2P2_ADD(t2, R2P2_S_HP1);
    const double da2 = R2P2_ADD(R2P2_SUB(t2, a2), R2P2_S_HP1);
    // arguments of the one do_sin and the one do_cos
    const double sx = r1 ? x : (r2 ? a2 : a3);
    double sdx = r1 ? 0.0 : (r2 ? da2 : da3);
    const double cx = r1 ? x : (r2 ? t2 : a3);
    double cdx = r1 ? 0.0 : (r2 ? R2P2_S_HP1 : da3);
    // ---- do_sin(sx, sdx) ----
    double ds;
    {
        const double tay = r2p2__taylor_sin(R2P2_MUL(sx, sx), sx, sdx);
        const double asx = r2p2_fabs(sx);
        if (!(sx > 0.0)) sdx = r2p2_neg(sdx);
        const double u = R2P2_ADD(asx, R2P2_S_BIG);
        const double xr = R2P2_SUB(asx, R2P2_SUB(u, R2P2_S_BIG));
        const double xx = R2P2_MUL(xr, xr);
        const double sp = R2P2_ADD(xr, R2P2_FMA(R2P2_MUL(xr, xx), R2P2_FMA(xx, R2P2_S_SN5, R2P2_S_SN3), sdx));
        const double cp = R2P2_FMA(xr, sdx, R2P2_MUL(xx, R2P2_FMA(xx, R2P2_FMA(xx, R2P2_S_CS6, R2P2_S_CS4), R2P2_S_CS2)));
        int ti = r2p2_lo32(u) * rowlen;
        if (asx < 0.126) ti = 0;                            // the table value is not used; keep the address in range
        const double sn = tab[ti], ssn = tab[ti + 1], cs = tab[ti + 2], ccs = tab[ti + 3];
        const double cor = R2P2_FMA(sp, cs, R2P2_FMA(r2p2_neg(cp), sn, R2P2_FMA(sp, ccs, ssn)));
        ds = (asx < 0.126) ? tay : r2p2_copysign(R2P2_ADD(sn, cor), sx);
    }
    // ---- do_cos(cx, cdx) ----
    double dc;
    {
        if (cx < 0.0) cdx = r2p2_neg(cdx);
        const double acx = r2p2_fabs(cx);
        const double u = R2P2_ADD(acx, R2P2_S_BIG);
        const double xr = R2P2_ADD(R2P2_SUB(acx, R2P2_SUB(u, R2P2_S_BIG)), cdx);
        const double xx = R2P2_MUL(xr, xr);
        const double sp = R2P2_FMA(R2P2_MUL(xr, xx), R2P2_FMA(xx, R2P2_S_SN5, R2P2_S_SN3), xr);
        const double cp = R2P2_MUL(xx, R2P2_FMA(xx, R2P2_FMA(xx, R2P2_S_CS6, R2P2_S_CS4), R2P2_S_CS2));
        const int ti = r2p2_lo32(u) * rowlen;
        const double sn = tab[ti], ssn = tab[ti + 1], cs = tab[ti + 2], ccs = tab[ti + 3];
        const double cor = R2P2_FMA(r2p2_neg(sp), sn, R2P2_FMA(r2p2_neg(cp), cs, R2P2_FMA(r2p2_neg(sp), ssn, ccs)));
        dc = R2P2_ADD(cs, cor);
    }
    // ---- route ----
    double rs, rc;
    if (r1) { rs = ds; rc = dc; }
    else if (r2) { rs = r2p2_copysign(dc, x); rc = ds; }
    else {
        rs = (n & 1) ? dc : ds;
        rc = (n & 1) ? ds : dc;
        if (n & 2) rs = r2p2_neg(rs);
        if ((n + 1) & 2) rc = r2p2_neg(rc);
    }
    if (k < 0x3e500000) rs = x;
    if (k < 0x3e400000) rc = 1.0;
    *s = rs;
    *c = rc;
}

R2P2_HD void r2p2_sincos(double x, const double* tab, double* s, double* c, int* ok) { r2p2_sincos_rs(x, tab, 4, s, c, ok); }

// ---------------------------------------------------------------------------------------------
// own atan2 / exp / tanh (deterministic, <= 2 ulp; not libm-identical)
// ---------------------------------------------------------------------------------------------
R2P2_HD double r2p2__atan_unit(double u, double* lo_out) {
    // atan(u) for 0 <= u <= 1 as hi + lo: nearest table point c = i/8, t = (u-c)/(1+u*c), Taylor in t.
    int i = (int)R2P2_FMA(u, 8.0, 0.5);
    double t;
    if (i == 0) {
        t = u;
    } else {
        const double c = R2P2_MUL((double)i, 0.125);
        t = R2P2_DIV(R2P2_SUB(u, c), R2P2_FMA(u, c, 1.0));
    }
    const double T[9] = {R2P2_ATAN_TAYLOR_VALUES};
    const double tt = R2P2_MUL(t, t);
    double p = T[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 7; j >= 0; --j) p = R2P2_FMA(tt, p, T[j]);
    // atan(t) = t + t*tt*p
    const double corr = R2P2_MUL(R2P2_MUL(t, tt), p);
    *lo_out = R2P2_ADD(R2P2_ATAN_LO(i), R2P2_ADD(corr, t));
    return R2P2_ATAN_HI(i);
}

R2P2_HD double r2p2_atan2(double y, double x) {
    const double ax = r2p2_fabs(x), ay = r2p2_fabs(y);
    if (x != x || y != y) return R2P2_ADD(x, y);
    if (ay == 0.0) {                                   // atan2(+-0, x)
        if ((r2p2_bits(x) >> 63) == 0) return y;       // x >= +0 : +-0
        return r2p2_copysign(R2P2_ADD(R2P2_PI_HI, R2P2_PI_LO), y);
    }
    if (ax == 0.0) return r2p2_copysign(R2P2_PIO2_HI, y);
    const double inf = r2p2_from_bits(0x7ff0000000000000ULL);
    if (ax == inf || ay == inf) {
        double r;
        if (ax == inf && ay == inf) r = (x > 0.0) ? 0x1.921fb54442d18p-1 : 0x1.2d97c7f3321d2p+1;   // pi/4, 3pi/4
        else if (ax == inf) r = (x > 0.0) ? 0.0 : R2P2_PI_HI;
        else r = R2P2_PIO2_HI;
        return r2p2_copysign(r, y);
    }
    double hi, lo, r;
    if (ay <= ax) {
        hi = r2p2__atan_unit(R2P2_DIV(ay, ax), &lo);       // atan(ay/ax) in [0, pi/4]
        if (x > 0.0) r = R2P2_ADD(hi, lo);
        else r = R2P2_ADD(R2P2_SUB(R2P2_PI_HI, hi), R2P2_SUB(R2P2_PI_LO, lo));
    } else {
        hi = r2p2__atan_unit(R2P2_DIV(ax, ay), &lo);       // pi/2 -+ atan(ax/ay)
        if (x > 0.0) r = R2P2_ADD(R2P2_SUB(R2P2_PIO2_HI, hi), R2P2_SUB(R2P2_PIO2_LO, lo));
        else r = R2P2_ADD(R2P2_ADD(R2P2_PIO2_HI, hi), R2P2_ADD(R2P2_PIO2_LO, lo));
    }
    return r2p2_copysign(r, y);
}

// exp(y) - 1 = 2^k (1 + p) - 1 for y <= 0: k = rint(y/ln2), r = y - k ln2, p = expm1(r) by Taylor to r^15.
R2P2_HD double r2p2__expm1_reduced(double y, int* k_out) {
#if defined(__CUDA_ARCH__)
    const double kf = rint(R2P2_MUL(y, R2P2_INV_LN2));
#else
    const double kf = __builtin_rint(R2P2_MUL(y, R2P2_INV_LN2));
#endif
    const double r = R2P2_FMA(r2p2_neg(kf), R2P2_LN2_LO, R2P2_FMA(r2p2_neg(kf), R2P2_LN2_HI, y));
    const double F[14] = {R2P2_INV_FACT_VALUES};
    double p = F[13];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 12; j >= 0; --j) p = R2P2_FMA(r, p, F[j]);
    *k_out = (int)kf;
    return R2P2_FMA(R2P2_MUL(r, r), p, r);
}

R2P2_HD double r2p2__expm1_neg(double y) {   // y <= 0
    if (y < -745.0) return -1.0;
    int k;
    const double p = r2p2__expm1_reduced(y, &k);
    if (k == 0) return p;
    if (k < -1021) return -1.0;
    const double two_k = r2p2_from_bits((uint64_t)(1023 + k) << 52);
    return R2P2_FMA(two_k, p, R2P2_SUB(two_k, 1.0));
}

R2P2_HD double r2p2__exp_neg(double y) {     // y <= 0; results below 2^-1021 flush to 0
    if (y < -745.0) return 0.0;
    int k;
    const double p = r2p2__expm1_reduced(y, &k);
    if (k < -1021) return 0.0;
    const double two_k = r2p2_from_bits((uint64_t)(1023 + k) << 52);
    return R2P2_MUL(two_k, R2P2_ADD(1.0, p));
}

// tanh, exp-based (kept for reference / accuracy cross-check): expm1 + one division, ~446 cycles on B200.
R2P2_HD double r2p2_tanh_exp(double x) {
    const double ax = r2p2_fabs(x);
    if (x != x) return x;
    if (ax >= 22.0) return r2p2_copysign(1.0, x);
    if (ax < 0x1.0p-28) return x;
    const double t = r2p2__expm1_neg(R2P2_MUL(-2.0, ax));   // in (-1, 0)
    const double r = R2P2_DIV(r2p2_neg(t), R2P2_ADD(t, 2.0));
    return r2p2_copysign(r, x);
}

// tanh by table: on [2^-6, 32) a degree-11 Taylor polynomial around the midpoint of the interval selected by the
// exponent and the top four mantissa bits of |x| (tools/gen_tanhtab.py; t = |x| - mid is exact), evaluated with Estrin's
// scheme -- no exp, no division, a 5-deep dependent chain.  Below 2^-6 the odd series at 0; from 20 on, +-1.
R2P2_HD double r2p2_tanh_tab(double x, const double* tanhtab) {
    const double ax = r2p2_fabs(x);
    if (x != x) return x;
    if (ax >= 20.0) return r2p2_copysign(1.0, x);
    if (ax < 0x1.0p-6) {
        if (ax < 0x1.0p-28) return x;
        const double S[4] = {R2P2_TANH_SMALL_VALUES};
        const double xx = R2P2_MUL(x, x);
        double p = R2P2_FMA(xx, S[3], S[2]);
        p = R2P2_FMA(xx, p, S[1]);
        p = R2P2_FMA(xx, p, S[0]);
        return R2P2_FMA(R2P2_MUL(x, xx), p, x);
    }
    const uint64_t top = r2p2_bits(ax) >> 48;                                   // exponent field and 4 mantissa bits
    const int idx = (int)top - ((1023 + R2P2_TANH_E_LO) << 4);
    const double mid = r2p2_from_bits((top << 48) | (1ULL << 47));
    const double t = R2P2_SUB(ax, mid);
    // row: c0 hi, c0 lo, c1..c11.  tanh = c0 + t * (c1 + c2 t + ... + c11 t^10): the bracket by Estrin, the dominant
    // term added last so that its rounding is the only one that matters (|t| <= mid/32)
    const double* c = tanhtab + idx * R2P2_TANH_ROW;
    const double t2 = R2P2_MUL(t, t), t4 = R2P2_MUL(t2, t2), t8 = R2P2_MUL(t4, t4);
    const double a0 = R2P2_FMA(c[3], t, c[2]), a1 = R2P2_FMA(c[5], t, c[4]), a2 = R2P2_FMA(c[7], t, c[6]);
    const double a3 = R2P2_FMA(c[9], t, c[8]), a4 = R2P2_FMA(c[11], t, c[10]);
    const double b0 = R2P2_FMA(a1, t2, a0), b1 = R2P2_FMA(a3, t2, a2), b2 = R2P2_FMA(c[12], t2, a4);
    const double inner = R2P2_FMA(b2, t8, R2P2_FMA(b1, t4, b0));
    const double r = R2P2_ADD(c[0], R2P2_FMA(t, inner, c[1]));
    return r2p2_copysign(r, x);
}

R2P2_HD double r2p2_tanh(double x) { return r2p2_tanh_tab(x, R2P2_TANHTAB); }

// logistic(x) = 1 / (1 + exp(-x))   (sklearn 'logistic' activation = scipy.special.expit)
R2P2_HD double r2p2_logistic(double x) {
    if (x != x) return x;
    const double e = r2p2__exp_neg(r2p2_neg(r2p2_fabs(x)));   // exp(-|x|) in (0, 1]
    const double d = R2P2_ADD(1.0, e);
    return (x >= 0.0) ? R2P2_DIV(1.0, d) : R2P2_DIV(e, d);
}
