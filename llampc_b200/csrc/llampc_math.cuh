// fp32 elementary functions for the Pacejka / bicycle right-hand side, sm_100a.
//
// MUFU has no atan, and MUFU.SIN's absolute error (2^-21.4) is too large for the look-back score, which is a
// squared difference of nearly equal increments (DESIGN.md "Numerics").  Everything here is a short
// branch-light minimax polynomial on the FMA pipe; MUFU.RCP is used only for the range reductions.
// Coefficients: tools/fit_coeffs.py (Chebyshev-node least squares with Lawson reweighting).
#pragma once
#include <cuda_runtime.h>

namespace llampc {

__device__ __forceinline__ float rcp_approx(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

// 1/a to ~0.5 ulp: MUFU.RCP plus one Newton step (no special-case handling: inf/0 propagate).
__device__ __forceinline__ float rcp_newton(float a) {
    float r = rcp_approx(a);
    float e = fmaf(-a, r, 1.0f);
    return fmaf(e, r, r);
}

// atan(q) for |q| <= 1: q + q*s*A(s), s = q^2; polynomial error 1.2e-8.
__device__ __forceinline__ float atan_unit(float q) {
    float s = q * q;
    float p = -2.5096684205e-03f;
    p = fmaf(p, s, 1.4001107524e-02f);
    p = fmaf(p, s, -3.6678340323e-02f);
    p = fmaf(p, s, 6.3189641129e-02f);
    p = fmaf(p, s, -8.6894353736e-02f);
    p = fmaf(p, s, 1.1042151828e-01f);
    p = fmaf(p, s, -1.4279634493e-01f);
    p = fmaf(p, s, 1.9999791356e-01f);
    p = fmaf(p, s, -3.3333332165e-01f);
    return fmaf(q * s, p, q);
}

// atan(t) for |t| <= 0.5 (slip-angle tangents of a car that is not spinning): 5 coefficients, 1.0e-8.
__device__ __forceinline__ float atan_half(float t) {
    float s = t * t;
    float p = -5.6081126704e-02f;
    p = fmaf(p, s, 1.0436029272e-01f);
    p = fmaf(p, s, -1.4229306540e-01f);
    p = fmaf(p, s, 1.9998319291e-01f);
    p = fmaf(p, s, -3.3333325341e-01f);
    return fmaf(t * s, p, t);
}

#define LLAMPC_PIO2_HI 1.57079637050628662109375f
#define LLAMPC_PIO2_LO (-4.37113900018624283e-8f)

// |z| > 1, +-inf, NaN: atan(z) = sign(z)*pi/2 - atan(1/z).  Cold path.
static __device__ __noinline__ float atan_large(float z) {
    float a = atan_unit(1.0f / z);
    float hi = copysignf(LLAMPC_PIO2_HI, z), lo = copysignf(LLAMPC_PIO2_LO, z);
    return (hi - a) + lo;
}

__device__ __forceinline__ float atan_any(float z) {
    if (fabsf(z) <= 1.0f) return atan_unit(z);
    return atan_large(z);
}

// np.arctan2(y, abs(vx)) (llampc/models/dynamic.py:146-147) given avx = |vx| >= 0 and inv = 1/avx.
static __device__ __noinline__ float atan2_pos_slow(float y, float avx) {
    if (y == 0.0f) return (avx != avx) ? avx : y;          // atan2(+-0, x>=0) = +-0 ; NaN propagates
    if (y != y || avx != avx) return y + avx;              // NaN
    float a = atan_unit(avx / y);                          // |avx/y| < 1 here (or 0 when avx = 0)
    float hi = copysignf(LLAMPC_PIO2_HI, y), lo = copysignf(LLAMPC_PIO2_LO, y);
    return (hi - a) + lo;
}

__device__ __forceinline__ float atan2_pos(float y, float avx, float inv) {
    float t = y * inv;
    if (fabsf(t) <= 1.0f) return atan_unit(t);
    return atan2_pos_slow(y, avx);
}

// sin(t) for any finite t: k = rint(t/pi), r = t - k*pi (two-constant Cody-Waite), sin t = (-1)^k sin r.
// Polynomial on |r| <= pi/2: r + r*s*S(s), error 7.6e-10.
__device__ __forceinline__ float sin_any(float t) {
    const float MAGIC = 12582912.0f;                       // 1.5 * 2^23
    float kf = fmaf(t, 0.318309886183790671538f, MAGIC);
    unsigned kbits = __float_as_uint(kf);
    kf -= MAGIC;
    float r = fmaf(kf, -3.1415927410125732421875f, t);
    r = fmaf(kf, 8.74227765734758577e-8f, r);
    float s = r * r;
    float p = -2.4753451528e-08f;
    p = fmaf(p, s, 2.7570330044e-06f);
    p = fmaf(p, s, -1.9841623052e-04f);
    p = fmaf(p, s, 8.3333355270e-03f);
    p = fmaf(p, s, -1.6666666687e-01f);
    float v = fmaf(r * s, p, r);
    return __uint_as_float(__float_as_uint(v) ^ (kbits << 31));
}

// atan(q) for |q| <= 1 with 8 coefficients (7.7e-8): used with the SFU tyre sine, whose own error (2^-21.4) is larger.
__device__ __forceinline__ float atan_unit7(float q) {
    float s = q * q;
    float p = 4.0598551869e-03f;
    p = fmaf(p, s, -2.0706461466e-02f);
    p = fmaf(p, s, 4.9855267454e-02f);
    p = fmaf(p, s, -8.0743718081e-02f);
    p = fmaf(p, s, 1.0888638420e-01f);
    p = fmaf(p, s, -1.4260910336e-01f);
    p = fmaf(p, s, 1.9998927280e-01f);
    p = fmaf(p, s, -3.3333325682e-01f);
    return fmaf(q * s, p, q);
}

// atan(z) for ANY z, branch-free: q = z / max(z^2, 1) is z for |z| <= 1 and 1/z otherwise (one MUFU.RCP, the sign
// rides along), atan z = sign(z) pi/2 - atan(q) in the second case.  +-inf -> +-pi/2, NaN -> NaN.
template <bool SHORT_POLY>
__device__ __forceinline__ float atan_full(float z) {
    const float zz = z * z;
    const float q = z * rcp_approx(fmaxf(zz, 1.0f));
    const float a = SHORT_POLY ? atan_unit7(q) : atan_unit(q);
    const float c = copysignf(LLAMPC_PIO2_HI, z);
    return (zz > 1.0f) ? (c - a) : a;                        // a NaN z gives q = NaN, hence a = NaN
}

// np.arctan2(y, avx) for avx >= 0, ANY magnitudes, branch-free: q = min/max in [0, 1] (one MUFU.RCP),
// octant fix-up, sign of y.  atan2(0, 0) = 0 like NumPy; a NaN input gives NaN.
template <bool RESTORE_NAN = true>
__device__ __forceinline__ float atan2_pos_full(float y, float avx) {
    const float ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ay, avx), 1e-30f), mn = fminf(ay, avx);
    const float q = mn * rcp_approx(mx);
    float a = atan_unit(q);
    a = (ay > avx) ? (LLAMPC_PIO2_HI - a) : a;
    a = copysignf(a, y);
    if (!RESTORE_NAN) return a;                                // caller detects NaN inputs elsewhere (see accel_fast)
    return (y != y || avx != avx) ? (y + avx) : a;             // fmaxf/fminf drop NaN operands: restore them
}

// sin(t) with a 4-coefficient polynomial (2.7e-8) after the same reduction as sin_any: used on the tyre
// curve, where the fp32 evaluation error (1.2e-7) dominates the polynomial error anyway.
__device__ __forceinline__ float sin_tyre(float t) {
    const float MAGIC = 12582912.0f;
    float kf = fmaf(t, 0.318309886183790671538f, MAGIC);
    unsigned kbits = __float_as_uint(kf);
    kf -= MAGIC;
    float r = fmaf(kf, -3.1415927410125732421875f, t);
    r = fmaf(kf, 8.74227765734758577e-8f, r);
    float s = r * r;
    float p = 2.6348915076e-06f;
    p = fmaf(p, s, -1.9822790564e-04f);
    p = fmaf(p, s, 8.3332426307e-03f);
    p = fmaf(p, s, -1.6666665972e-01f);
    float v = fmaf(r * s, p, r);
    return __uint_as_float(__float_as_uint(v) ^ (kbits << 31));
}

// sin/cos of a tiny angle |e| <= 0.125 (RK stage heading offsets of the look-back, O(h^2 * yaw acceleration)):
// truncation errors e^7/5040 < 1e-10 and e^6/720 < 6e-9.  No range check here: the caller tracks max |e|.
__device__ __forceinline__ void sincos_tiny(float e, float& sn, float& cs) {
    float s = e * e;
    float ps = fmaf(s, 8.3333333e-03f, -1.6666667e-01f);
    sn = fmaf(e * s, ps, e);
    float pc = fmaf(s, 4.1666668e-02f, -0.5f);
    cs = fmaf(s, pc, 1.0f);
}

// sincos_small without the range branch (|d| <= 0.5 is the caller's guard)
__device__ __forceinline__ void sincos_half(float d, float& sn, float& cs) {
    float s = d * d;
    float ps = -1.9736421589e-04f;
    ps = fmaf(ps, s, 8.3332314830e-03f);
    ps = fmaf(ps, s, -1.6666666503e-01f);
    sn = fmaf(d * s, ps, d);
    float pc = 2.4801587e-05f;
    pc = fmaf(pc, s, -1.3888889e-03f);
    pc = fmaf(pc, s, 4.1666668e-02f);
    pc = fmaf(pc, s, -0.5f);
    cs = fmaf(s, pc, 1.0f);
}

// MUFU variant (2 instructions); measured against the polynomial in tests/bench, not the default.
__device__ __forceinline__ float sin_mufu(float t) { return __sinf(t); }

static __device__ __noinline__ void sincos_cold(float d, float* sn, float* cs) { sincosf(d, sn, cs); }

// sin/cos of a small heading deviation |d| <= 0.5 (RK stage offsets h*omega); libm beyond that.
__device__ __forceinline__ void sincos_small(float d, float& sn, float& cs) {
    if (fabsf(d) > 0.5f) { sincos_cold(d, &sn, &cs); return; }
    float s = d * d;
    float ps = -1.9736421589e-04f;
    ps = fmaf(ps, s, 8.3332314830e-03f);
    ps = fmaf(ps, s, -1.6666666503e-01f);
    sn = fmaf(d * s, ps, d);
    float pc = 2.4801587e-05f;                              // Taylor: terms through d^8, error < 3e-10 at 0.5
    pc = fmaf(pc, s, -1.3888889e-03f);
    pc = fmaf(pc, s, 4.1666668e-02f);
    pc = fmaf(pc, s, -0.5f);
    cs = fmaf(s, pc, 1.0f);
}

}  // namespace llampc
