// Packed-pair (f32x2) form of the look-back step: TWO candidates per thread, one window row at a time.  sm_100a.
//
// Blackwell issues FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add.rn.f32x2): one instruction, two IEEE fp32 results.  The FMA
// pipe needs two cycles for it, so the FLOP rate is that of scalar FFMA (measured: 0.93 scalar FFMA against 0.47
// FFMA2 per clock per scheduler, tools/ubench/ffma2.cu), but the ISSUE slot is paid once.  K1's step is 80 % fp32
// arithmetic (150 FFMA + 105 FMUL + 24 FADD of 347 instructions) and is bound by issue slots (77 % busy) with the FMA
// pipe at 62 %: packing two candidates into every arithmetic instruction removes 40 % of the issue slots per
// candidate-step and moves the bound to the FMA pipe.  ptxas folds broadcast scalars (R.F32), immediates and
// negations into the packed instruction, so history-row values (shared by both candidates) stay scalar registers
// and model parameters are the packed operands.
//
// Same arithmetic, operation for operation, as lookback_step_fast (llampc_model.cuh), which follows
// Dynamic.calc_forces_batch / _diffequation_batch (llampc/models/dynamic.py:98-154) and odeintRK4_batch
// (llampc/utils/rk6.py:50-68); the only re-association is the sign carried through the rear slip angle (odd functions).
#pragma once
#include "llampc_model.cuh"

namespace llampc {

struct F2 { unsigned long long v; };

__device__ __forceinline__ F2 pk(float a, float b) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(F2 a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v)); }
__device__ __forceinline__ F2 bc(float a) { return pk(a, a); }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 add2(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 sub2(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// sign flip of both components on the ALU pipe (two LOP3) instead of an FMUL2 by -1 on the FMA pipe
__device__ __forceinline__ F2 neg2(F2 a) { F2 r; r.v = a.v ^ 0x8000000080000000ull; return r; }

// component-wise helpers for what has no packed instruction (MUFU, min/max, selects)
__device__ __forceinline__ F2 rcp_abs2(F2 a) {
    float x, y; up(a, x, y);
    return pk(rcp_approx(fabsf(x)), rcp_approx(fabsf(y)));
}

// Two candidates: the packed-bank rows of candidates i0 and i1, with the negated copies the packed code wants
// (a negation of a packed register is not free, a pre-negated parameter is).
struct Cand2 {
    F2 Bf, Cf, Df, Br, Cr, Dr, inv_m, lf, nlr, lf_Iz, lr_Iz, Cm1, Cm2, nCr0, nCr2;
};

__device__ __forceinline__ Cand2 make_cand2(const Cand& a, const Cand& b) {
    Cand2 c;
    c.Bf = pk(a.Bf, b.Bf); c.Cf = pk(a.Cf, b.Cf); c.Df = pk(a.Df, b.Df);
    c.Br = pk(a.Br, b.Br); c.Cr = pk(a.Cr, b.Cr); c.Dr = pk(a.Dr, b.Dr);
    c.inv_m = pk(a.inv_m, b.inv_m); c.lf = pk(a.lf, b.lf); c.nlr = pk(-a.lr, -b.lr);
    c.lf_Iz = pk(a.lf_Iz, b.lf_Iz); c.lr_Iz = pk(a.lr_Iz, b.lr_Iz);
    c.Cm1 = pk(a.Cm1, b.Cm1); c.Cm2 = pk(a.Cm2, b.Cm2); c.nCr0 = pk(-a.Cr0, -b.Cr0); c.nCr2 = pk(-a.Cr2, -b.Cr2);
    return c;
}

__device__ __forceinline__ Cand2 load_cand2(const float4* __restrict__ bank, int Npad, int i0, int i1) {
    return make_cand2(load_cand(bank, Npad, i0), load_cand(bank, Npad, i1));
}

// atan(t), |t| <= 0.5 (atan_half)
__device__ __forceinline__ F2 atan_half2(F2 t) {
    const F2 s = mul2(t, t);
    F2 p = bc(-5.6081126704e-02f);
    p = fma2(p, s, bc(1.0436029272e-01f));
    p = fma2(p, s, bc(-1.4229306540e-01f));
    p = fma2(p, s, bc(1.9998319291e-01f));
    p = fma2(p, s, bc(-3.3333325341e-01f));
    return fma2(mul2(t, s), p, t);
}

// atan(q), |q| <= 1: atan_unit7 (SHORT) or atan_unit
template <bool SHORT_POLY>
__device__ __forceinline__ F2 atan_unit2(F2 q) {
    const F2 s = mul2(q, q);
    F2 p;
    if (SHORT_POLY) {
        p = bc(4.0598551869e-03f);
        p = fma2(p, s, bc(-2.0706461466e-02f));
        p = fma2(p, s, bc(4.9855267454e-02f));
        p = fma2(p, s, bc(-8.0743718081e-02f));
        p = fma2(p, s, bc(1.0888638420e-01f));
        p = fma2(p, s, bc(-1.4260910336e-01f));
        p = fma2(p, s, bc(1.9998927280e-01f));
        p = fma2(p, s, bc(-3.3333325682e-01f));
    } else {
        p = bc(-2.5096684205e-03f);
        p = fma2(p, s, bc(1.4001107524e-02f));
        p = fma2(p, s, bc(-3.6678340323e-02f));
        p = fma2(p, s, bc(6.3189641129e-02f));
        p = fma2(p, s, bc(-8.6894353736e-02f));
        p = fma2(p, s, bc(1.1042151828e-01f));
        p = fma2(p, s, bc(-1.4279634493e-01f));
        p = fma2(p, s, bc(1.9999791356e-01f));
        p = fma2(p, s, bc(-3.3333332165e-01f));
    }
    return fma2(mul2(q, s), p, q);
}

// atan(z) for any z (atan_full): q = z / max(z^2, 1), octant fix-up per component
template <bool SHORT_POLY>
__device__ __forceinline__ F2 atan_full2(F2 z) {
    const F2 zz = mul2(z, z);
    float z0, z1, q0, q1;
    up(zz, q0, q1);
    up(z, z0, z1);
    const F2 q = mul2(z, pk(rcp_approx(fmaxf(q0, 1.0f)), rcp_approx(fmaxf(q1, 1.0f))));
    const F2 a = atan_unit2<SHORT_POLY>(q);
    float a0, a1;
    up(a, a0, a1);
    a0 = (q0 > 1.0f) ? (copysignf(LLAMPC_PIO2_HI, z0) - a0) : a0;
    a1 = (q1 > 1.0f) ? (copysignf(LLAMPC_PIO2_HI, z1) - a1) : a1;
    return pk(a0, a1);
}

// sin_tyre: Cody-Waite reduction by pi + 4-coefficient polynomial (strict mode)
__device__ __forceinline__ F2 sin_tyre2(F2 t) {
    const float MAGIC = 12582912.0f;
    F2 kf = fma2(t, bc(0.318309886183790671538f), bc(MAGIC));
    float k0, k1;
    up(kf, k0, k1);
    const unsigned b0 = __float_as_uint(k0), b1 = __float_as_uint(k1);
    kf = add2(kf, bc(-MAGIC));
    F2 r = fma2(kf, bc(-3.1415927410125732421875f), t);
    r = fma2(kf, bc(8.74227765734758577e-8f), r);
    const F2 s = mul2(r, r);
    F2 p = bc(2.6348915076e-06f);
    p = fma2(p, s, bc(-1.9822790564e-04f));
    p = fma2(p, s, bc(8.3332426307e-03f));
    p = fma2(p, s, bc(-1.6666665972e-01f));
    const F2 v = fma2(mul2(r, s), p, r);
    float v0, v1;
    up(v, v0, v1);
    return pk(__uint_as_float(__float_as_uint(v0) ^ (b0 << 31)), __uint_as_float(__float_as_uint(v1) ^ (b1 << 31)));
}

template <bool MUFU_SIN>
__device__ __forceinline__ F2 pacejka_fast2(F2 B, F2 C, F2 D, F2 alpha) {
    const F2 t = mul2(C, atan_full2<MUFU_SIN>(mul2(B, alpha)));
    if (MUFU_SIN) {
        float t0, t1;
        up(t, t0, t1);
        return mul2(D, pk(sin_mufu(t0), sin_mufu(t1)));
    }
    return mul2(D, sin_tyre2(t));
}

__device__ __forceinline__ void sincos_tiny2(F2 e, F2& sn, F2& cs) {
    const F2 s = mul2(e, e);
    const F2 ps = fma2(s, bc(8.3333333e-03f), bc(-1.6666667e-01f));
    sn = fma2(mul2(e, s), ps, e);
    const F2 pc = fma2(s, bc(4.1666668e-02f), bc(-0.5f));
    cs = fma2(s, pc, bc(1.0f));
}

// sin / cos of a heading offset |e| <= 0.5 (the WIDE form: truncation errors 5e-9 / 3e-10 at 0.5)
__device__ __forceinline__ void sincos_mid2(F2 e, F2& sn, F2& cs) {
    const F2 s = mul2(e, e);
    F2 ps = fma2(s, bc(-1.9841270e-04f), bc(8.3333333e-03f));
    ps = fma2(ps, s, bc(-1.6666667e-01f));
    sn = fma2(mul2(e, s), ps, e);
    F2 pc = fma2(s, bc(2.4801587e-05f), bc(-1.3888889e-03f));
    pc = fma2(pc, s, bc(4.1666668e-02f));
    pc = fma2(pc, s, bc(-0.5f));
    cs = fma2(s, pc, bc(1.0f));
}

struct Deriv2 { F2 vx, nvy, w; };                // nvy = MINUS d(vy)/dt: consumers fold the sign into the (scalar) step size
struct Drive2 { F2 A, nBq; };                     // Frx = A - vx (Cr2 vx + Bq), evaluated as fma(vx, fma(-Cr2, vx, -Bq), A)

__device__ __forceinline__ Drive2 prep_drive2(const Cand2& p, float pwm) {
    Drive2 d;
    d.A = fma2(p.Cm1, bc(pwm), p.nCr0);
    d.nBq = mul2(p.Cm2, bc(-pwm));
    return d;
}
__device__ __forceinline__ F2 drive_force_fast2(const Cand2& p, const Drive2& d, F2 vx) {
    return fma2(vx, fma2(p.nCr2, vx, d.nBq), d.A);
}

// guards: gt = max |slip tangent| (must stay <= 0.5), per component
struct Guard2 { float t0, t1; };

// WIDE: the slip angles come from the full-range atan (any tangent, same 1.2e-8 accuracy as the scalar general step) instead
// of the |t| <= 0.5 polynomial, and the tangent guard is not needed: the form for windows measured at low speed or in a drift,
// where most candidates leave the fast form (a separate kernel instantiation, chosen per launch: LLAMPC_LB_FLAG_WIDE).
template <bool MUFU_SIN, bool WIDE = false>
__device__ __forceinline__ Deriv2 accel_fast2(const Cand2& p, const Ctl& u, const Drive2& drv, F2 vx, F2 vy, F2 w, Guard2& g) {
    const F2 inv = rcp_abs2(vx);
    const F2 tf = mul2(fma2(p.lf, w, vy), inv);
    const F2 ntr = mul2(fma2(p.nlr, w, vy), inv);             // -(lr w - vy) / |vx|: the sign rides through the odd functions
    if (!WIDE) {
        float a, b, c, d;
        up(tf, a, b);
        up(ntr, c, d);
        g.t0 = fmaxf(g.t0, fmaxf(fabsf(a), fabsf(c)));
        g.t1 = fmaxf(g.t1, fmaxf(fabsf(b), fabsf(d)));
    }
    const F2 af = sub2(bc(u.delta), WIDE ? atan_full2<false>(tf) : atan_half2(tf));
    const F2 nar = WIDE ? atan_full2<false>(ntr) : atan_half2(ntr);
    const F2 Frx = drive_force_fast2(p, drv, vx);
    const F2 Ffy = pacejka_fast2<MUFU_SIN>(p.Bf, p.Cf, p.Df, af);
    const F2 nFry = pacejka_fast2<MUFU_SIN>(p.Br, p.Cr, p.Dr, nar);
    const F2 Fc = mul2(Ffy, bc(u.cd));
    Deriv2 r;
    r.vx = fma2(fma2(Ffy, bc(-u.sd), Frx), p.inv_m, mul2(vy, w));
    r.nvy = fma2(sub2(nFry, Fc), p.inv_m, mul2(vx, w));
    r.w = fma2(Fc, p.lf_Iz, mul2(nFry, p.lr_Iz));
    return r;
}

// Two candidates, one history row.  e2 = per-candidate squared increment error; ok0 / ok1 false when a guard tripped
// (the caller redoes that candidate with lookback_step_general).
template <bool GEOM_SHARED, bool MUFU_SIN, bool WIDE = false>
__device__ __forceinline__ F2 lookback_step_fast2(const Cand2& p, const HistRow& r, const StepSize& z, bool& ok0, bool& ok1) {
    const float h = z.h, hh = z.hh;
    const float vx0 = r.q1.z, vy0 = r.q1.w, w0 = r.q2.x;
    Ctl u;
    u.pwm = r.q2.y; u.delta = r.q2.z; u.sd = r.q2.w; u.cd = r.q3.x;
    Guard2 g = {0.0f, 0.0f};
    const Drive2 drv = prep_drive2(p, u.pwm);
    // stage 1
    Deriv2 a1;
    if (GEOM_SHARED) {
        const F2 Frx = drive_force_fast2(p, drv, bc(vx0));
        const F2 Ffy = pacejka_fast2<MUFU_SIN>(p.Bf, p.Cf, p.Df, bc(r.q4.z));
        const F2 nFry = pacejka_fast2<MUFU_SIN>(p.Br, p.Cr, p.Dr, bc(-r.q4.w));
        const F2 Fc = mul2(Ffy, bc(u.cd));
        a1.vx = fma2(fma2(Ffy, bc(-u.sd), Frx), p.inv_m, bc(vy0 * w0));
        a1.nvy = fma2(sub2(nFry, Fc), p.inv_m, bc(vx0 * w0));
        a1.w = fma2(Fc, p.lf_Iz, mul2(nFry, p.lr_Iz));
    } else {
        a1 = accel_fast2<MUFU_SIN, WIDE>(p, u, drv, bc(vx0), bc(vy0), bc(w0), g);
    }
    // stage 2
    const F2 vx2 = fma2(bc(hh), a1.vx, bc(vx0)), vy2 = fma2(bc(-hh), a1.nvy, bc(vy0)), w2 = fma2(bc(hh), a1.w, bc(w0));
    const Deriv2 a2 = accel_fast2<MUFU_SIN, WIDE>(p, u, drv, vx2, vy2, w2, g);
    F2 xs = fma2(vx2, bc(r.q0.w), mul2(vy2, bc(-r.q0.z))), ys = fma2(vx2, bc(r.q0.z), mul2(vy2, bc(r.q0.w)));
    // stage 3
    const F2 vx3 = fma2(bc(hh), a2.vx, bc(vx0)), vy3 = fma2(bc(-hh), a2.nvy, bc(vy0)), w3 = fma2(bc(hh), a2.w, bc(w0));
    const F2 e3 = mul2(bc(hh), mul2(bc(hh), a1.w));
    F2 sd, cd;
    if (WIDE) sincos_mid2(e3, sd, cd); else sincos_tiny2(e3, sd, cd);
    const F2 s3 = fma2(bc(r.q0.z), cd, mul2(bc(r.q0.w), sd)), c3 = fma2(bc(r.q0.w), cd, mul2(bc(-r.q0.z), sd));
    const Deriv2 a3 = accel_fast2<MUFU_SIN, WIDE>(p, u, drv, vx3, vy3, w3, g);
    xs = add2(xs, sub2(mul2(vx3, c3), mul2(vy3, s3)));
    ys = add2(ys, fma2(vx3, s3, mul2(vy3, c3)));
    // stage 4 (front tyre and drivetrain only)
    const F2 vx4 = fma2(bc(h), a3.vx, bc(vx0)), vy4 = fma2(bc(-h), a3.nvy, bc(vy0)), w4 = fma2(bc(h), a3.w, bc(w0));
    const F2 e4 = mul2(bc(h), mul2(bc(hh), a2.w));
    if (WIDE) sincos_mid2(e4, sd, cd); else sincos_tiny2(e4, sd, cd);
    const F2 s4 = fma2(bc(r.q1.x), cd, mul2(bc(r.q1.y), sd)), c4 = fma2(bc(r.q1.y), cd, mul2(bc(-r.q1.x), sd));
    const F2 inv4 = rcp_abs2(vx4);
    const F2 tf4 = mul2(fma2(p.lf, w4, vy4), inv4);
    const F2 Ffy4 = pacejka_fast2<MUFU_SIN>(p.Bf, p.Cf, p.Df, sub2(bc(u.delta), WIDE ? atan_full2<false>(tf4) : atan_half2(tf4)));
    const F2 a4vx = fma2(fma2(Ffy4, bc(-u.sd), drive_force_fast2(p, drv, vx4)), p.inv_m, mul2(vy4, w4));
    const F2 xd4 = sub2(mul2(vx4, c4), mul2(vy4, s4)), yd4 = fma2(vx4, s4, mul2(vy4, c4));
    // increment errors
    const F2 sx = fma2(bc(2.0f), xs, xd4), sy = fma2(bc(2.0f), ys, yd4);
    const F2 sw = add2(add2(a1.w, a2.w), a3.w), sv = fma2(bc(2.0f), add2(a2.vx, a3.vx), add2(a1.vx, a4vx));
    const F2 ex = fma2(bc(z.h6_lo), sx, fma2(bc(z.h6), sx, bc(-r.q3.y)));
    const F2 ey = fma2(bc(z.h6_lo), sy, fma2(bc(z.h6), sy, bc(-r.q3.z)));
    const F2 epsi = fma2(bc(z.hh6_lo), sw, fma2(bc(z.hh6), sw, bc(-r.q3.w)));
    const F2 evx = add2(fma2(bc(z.h6_lo), sv, fma2(bc(z.h6), sv, bc(-r.q4.x))), bc(-r.q4.y));
    const F2 e2 = fma2(ex, ex, fma2(ey, ey, fma2(epsi, epsi, mul2(evx, evx))));
    // guards: slip tangents <= 0.5 (stages 2-4 and, without shared geometry, stage 1), heading offsets <= 0.125
    // (WIDE: no tangent guard, heading offsets <= 0.5 with the longer sin / cos polynomials)
    float t0, t1, o0, o1, p0, p1, q0, q1;
    up(tf4, t0, t1);
    up(e3, o0, o1);
    up(e4, p0, p1);
    up(e2, q0, q1);
    ok0 = (WIDE || fmaxf(g.t0, fabsf(t0)) <= 0.5f) && (fmaxf(fabsf(o0), fabsf(p0)) <= (WIDE ? 0.5f : 0.125f)) && (q0 == q0);
    ok1 = (WIDE || fmaxf(g.t1, fabsf(t1)) <= 0.5f) && (fmaxf(fabsf(o1), fabsf(p1)) <= (WIDE ? 0.5f : 0.125f)) && (q1 == q1);
    return e2;
}

// ---------------------------------------------------------------------------------------------------
// Packed pieces of the look-ahead rollout (lookahead.cu, K2p): two MODELS per thread on one control sequence.
// K2p is bound by FMA-pipe cycles (ncu: math-pipe throttle + fixed-latency wait; 4 CTAs / SM or an Estrin polynomial
// change nothing or lose, profiles/r02_k2p_variants.txt), so everything here is written to need FEWER packed operations.
// ---------------------------------------------------------------------------------------------------
// atan(q), |q| <= 1, 7 coefficients: absolute error 5.0e-7 (tools/fit_coeffs.py, degree 6 in q^2) -- the size of
// MUFU.SIN's own error (2^-21.4 = 3.6e-7), which the tyre force already carries.  One FFMA2 fewer than atan_unit7.
__device__ __forceinline__ F2 atan_unit2_la(F2 q) {
    const F2 s = mul2(q, q);
    F2 p = bc(-6.6536013602e-03f);
    p = fma2(p, s, bc(3.0787179075e-02f));
    p = fma2(p, s, bc(-6.7951120877e-02f));
    p = fma2(p, s, bc(1.0449814999e-01f));
    p = fma2(p, s, bc(-1.4189631985e-01f));
    p = fma2(p, s, bc(1.9994621885e-01f));
    p = fma2(p, s, bc(-3.3333283787e-01f));
    return fma2(mul2(q, s), p, q);
}

// np.arctan2(y, avx) for avx >= 0 and ANY magnitudes (atan2_pos_full), two components.  Candidate models of the
// rollout spin within the horizon (12 % of the warp-steps of config C3 see a slip tangent above 1), so there is no
// guard and no fallback here: q = min / max in [0, 1] (one MUFU.RCP per component), the polynomial in packed
// arithmetic, octant fix-up and sign per component.  A NaN operand is dropped by min / max: a NaN state still reaches
// the increments through the drivetrain force and the vy w / vx w terms.
__device__ __forceinline__ F2 atan2_pos_full2(F2 y, float avx0, float avx1) {
    float y0, y1;
    up(y, y0, y1);
    const float ay0 = fabsf(y0), ay1 = fabsf(y1);
    const float q0 = fminf(ay0, avx0) * rcp_approx(fmaxf(fmaxf(ay0, avx0), 1e-30f));
    const float q1 = fminf(ay1, avx1) * rcp_approx(fmaxf(fmaxf(ay1, avx1), 1e-30f));
    const F2 a = atan_unit2_la(pk(q0, q1));
    float a0, a1;
    up(a, a0, a1);
    a0 = (ay0 > avx0) ? (LLAMPC_PIO2_HI - a0) : a0;
    a1 = (ay1 > avx1) ? (LLAMPC_PIO2_HI - a1) : a1;
    return pk(copysignf(a0, y0), copysignf(a1, y1));
}

// atan(z) for any z (atan_full2) with the 7-coefficient polynomial
__device__ __forceinline__ F2 atan_full2_la(F2 z) {
    const F2 zz = mul2(z, z);
    float z0, z1, q0, q1;
    up(zz, q0, q1);
    up(z, z0, z1);
    const F2 q = mul2(z, pk(rcp_approx(fmaxf(q0, 1.0f)), rcp_approx(fmaxf(q1, 1.0f))));
    const F2 a = atan_unit2_la(q);
    float a0, a1;
    up(a, a0, a1);
    a0 = (q0 > 1.0f) ? (copysignf(LLAMPC_PIO2_HI, z0) - a0) : a0;
    a1 = (q1 > 1.0f) ? (copysignf(LLAMPC_PIO2_HI, z1) - a1) : a1;
    return pk(a0, a1);
}

// D sin(C atan(z)) with z = B alpha already formed
__device__ __forceinline__ F2 pacejka_la2(F2 z, F2 C, F2 D) {
    const F2 t = mul2(C, atan_full2_la(z));
    float t0, t1;
    up(t, t0, t1);
    return mul2(D, pk(sin_mufu(t0), sin_mufu(t1)));
}

// Control input of a packed step.  With one control sequence shared by both models every member is a broadcast pair that
// ptxas folds into the packed instruction as a scalar operand; with one sequence per model the components differ.
struct Ctl2 { F2 pwm, npwm, delta, nsd, cd; };    // npwm = -pwm, nsd = -sin(delta)
__device__ __forceinline__ Ctl2 ctl2_shared(float pwm, float delta, float sd, float cd) {
    Ctl2 u;
    u.pwm = bc(pwm); u.npwm = bc(-pwm); u.delta = bc(delta); u.nsd = bc(-sd); u.cd = bc(cd);
    return u;
}
__device__ __forceinline__ Drive2 prep_drive2(const Cand2& p, const Ctl2& u) {
    Drive2 d;
    d.A = fma2(p.Cm1, u.pwm, p.nCr0);
    d.nBq = mul2(p.Cm2, u.npwm);
    return d;
}

// d/dt of (vx, vy, omega) for two models, slip angles of any size (accel_fast<.., true>)
// Bfd = Bf delta and nBf = -Bf are formed once per step: Bf (delta - a) = nBf a + Bfd is one FFMA2 per stage.
__device__ __forceinline__ Deriv2 accel_full2(const Cand2& p, const Ctl2& u, const Drive2& drv, F2 nBf, F2 Bfd, F2 vx, F2 vy, F2 w) {
    float vx0, vx1;
    up(vx, vx0, vx1);
    const float avx0 = fabsf(vx0), avx1 = fabsf(vx1);
    const F2 zf = fma2(nBf, atan2_pos_full2(fma2(p.lf, w, vy), avx0, avx1), Bfd);      // Bf alpha_f
    const F2 nzr = mul2(p.Br, atan2_pos_full2(fma2(p.nlr, w, vy), avx0, avx1));        // -(Br alpha_r): the sign rides through the odd functions
    const F2 Frx = drive_force_fast2(p, drv, vx);
    const F2 Ffy = pacejka_la2(zf, p.Cf, p.Df);
    const F2 nFry = pacejka_la2(nzr, p.Cr, p.Dr);
    const F2 Fc = mul2(Ffy, u.cd);
    Deriv2 r;
    r.vx = fma2(fma2(Ffy, u.nsd, Frx), p.inv_m, mul2(vy, w));
    r.nvy = fma2(sub2(nFry, Fc), p.inv_m, mul2(vx, w));
    r.w = fma2(Fc, p.lf_Iz, mul2(nFry, p.lr_Iz));
    return r;
}

// sin / cos of the heading psi0 + phi for two models: the rollouts carry the heading RELATIVE to the shared start heading
// (phi stays a small fp32 number), sin / cos phi come from the SFU (MUFU.SIN / MUFU.COS: absolute error 2^-21.4, i.e. a
// position error of v h 4e-7 ~ 1e-8 m per step), and the rotation by the start heading uses its fp64-formed sin / cos as
// broadcast scalars: 4 packed operations + 4 MUFU per stage instead of the 15 packed operations of a polynomial
// sin / cos + rotation, valid for ANY heading change (no tiny-angle guard, no fallback).
struct Head2 { F2 s0, c0, ns0; };                 // sin, cos, -sin of the start heading (broadcast pairs when it is shared)
__device__ __forceinline__ void heading2(F2 phi, const Head2& h0, F2& sn, F2& cs) {
    float a0, a1;
    up(phi, a0, a1);
    const F2 sp = pk(__sinf(a0), __sinf(a1)), cp = pk(__cosf(a0), __cosf(a1));
    sn = fma2(h0.s0, cp, mul2(h0.c0, sp));
    cs = fma2(h0.c0, cp, mul2(h0.ns0, sp));
}

// One RK4 step of two models (rk4_increment_fast, same stages).  State in / out: position relative to the start (X, Y),
// heading relative to the start heading (phi), vx, vy, omega.  The weighted stage sums (k1 + 2 k2 + 2 k3 + k4) are
// accumulated stage by stage and folded into the state with one FFMA2 each.
struct State2 { F2 X, Y, phi, vx, vy, w; };
__device__ __forceinline__ void rk4_step2(const Cand2& p, const Ctl2& u, const Head2& h0, float h, State2& x) {
    const float hh = 0.5f * h, h6 = h * (1.0f / 6.0f);
    const Drive2 drv = prep_drive2(p, u);
    const F2 nBf = neg2(p.Bf), Bfd = mul2(p.Bf, u.delta);
    F2 sn, cs;
    // stage 1
    heading2(x.phi, h0, sn, cs);
    Deriv2 a = accel_full2(p, u, drv, nBf, Bfd, x.vx, x.vy, x.w);
    F2 sx = fma2(x.vx, cs, mul2(neg2(x.vy), sn)), sy = fma2(x.vx, sn, mul2(x.vy, cs));
    F2 sw = x.w, svx = a.vx, snvy = a.nvy, sdw = a.w;
    F2 vx = fma2(bc(hh), a.vx, x.vx), vy = fma2(bc(-hh), a.nvy, x.vy), w = fma2(bc(hh), a.w, x.w);
    heading2(fma2(bc(hh), x.w, x.phi), h0, sn, cs);
    // stage 2
    a = accel_full2(p, u, drv, nBf, Bfd, vx, vy, w);
    sx = fma2(bc(2.0f), fma2(vx, cs, mul2(neg2(vy), sn)), sx);
    sy = fma2(bc(2.0f), fma2(vx, sn, mul2(vy, cs)), sy);
    sw = fma2(bc(2.0f), w, sw); svx = fma2(bc(2.0f), a.vx, svx); snvy = fma2(bc(2.0f), a.nvy, snvy); sdw = fma2(bc(2.0f), a.w, sdw);
    heading2(fma2(bc(hh), w, x.phi), h0, sn, cs);
    vx = fma2(bc(hh), a.vx, x.vx); vy = fma2(bc(-hh), a.nvy, x.vy); w = fma2(bc(hh), a.w, x.w);
    // stage 3
    a = accel_full2(p, u, drv, nBf, Bfd, vx, vy, w);
    sx = fma2(bc(2.0f), fma2(vx, cs, mul2(neg2(vy), sn)), sx);
    sy = fma2(bc(2.0f), fma2(vx, sn, mul2(vy, cs)), sy);
    sw = fma2(bc(2.0f), w, sw); svx = fma2(bc(2.0f), a.vx, svx); snvy = fma2(bc(2.0f), a.nvy, snvy); sdw = fma2(bc(2.0f), a.w, sdw);
    heading2(fma2(bc(h), w, x.phi), h0, sn, cs);
    vx = fma2(bc(h), a.vx, x.vx); vy = fma2(bc(-h), a.nvy, x.vy); w = fma2(bc(h), a.w, x.w);
    // stage 4
    a = accel_full2(p, u, drv, nBf, Bfd, vx, vy, w);
    sx = add2(sx, fma2(vx, cs, mul2(neg2(vy), sn)));
    sy = add2(sy, fma2(vx, sn, mul2(vy, cs)));
    x.X = fma2(bc(h6), sx, x.X);
    x.Y = fma2(bc(h6), sy, x.Y);
    x.phi = fma2(bc(h6), add2(sw, w), x.phi);
    x.vx = fma2(bc(h6), add2(svx, a.vx), x.vx);
    x.vy = fma2(bc(-h6), add2(snvy, a.nvy), x.vy);
    x.w = fma2(bc(h6), add2(sdw, a.w), x.w);
}

}  // namespace llampc
