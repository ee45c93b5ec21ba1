// Dynamic-bicycle / Pacejka right-hand side and the RK4 step in increment form (fp32, registers only).
//
// Follows Dynamic.calc_forces_batch (llampc/models/dynamic.py:117-154, pwm/Pacejka branch :141-149),
// Dynamic._diffequation_batch (dynamic.py:98-115) and odeintRK4_batch (llampc/utils/rk6.py:50-68).
// The state is never advanced as an O(1) fp32 number: every routine returns the INCREMENT of the step
// (sum of weighted stage derivatives) and takes sin/cos of the heading at the start of the step, so the
// caller can compare increments with measured increments (look-back) or add them in higher precision.
#pragma once
#include "llampc_math.cuh"

namespace llampc {

// One candidate model: the packed-bank row (include/llampc_b200.h, llampc_bank_pack_h).
struct Cand {
    float Bf, Cf, Df, Br, Cr, Dr, inv_m, lf, lr, lf_Iz, lr_Iz, Cm1, Cm2, Cr0, Cr2;
};

__device__ __forceinline__ Cand cand_from_groups(const float4& g0, const float4& g1, const float4& g2, const float4& g3) {
    Cand c;
    c.Bf = g0.x; c.Cf = g0.y; c.Df = g0.z; c.Br = g0.w;
    c.Cr = g1.x; c.Dr = g1.y; c.inv_m = g1.z; c.lf = g1.w;
    c.lr = g2.x; c.lf_Iz = g2.y; c.lr_Iz = g2.z; c.Cm1 = g2.w;
    c.Cm2 = g3.x; c.Cr0 = g3.y; c.Cr2 = g3.z;
    return c;
}

__device__ __forceinline__ Cand load_cand(const float4* __restrict__ bank, int Npad, int i) {
    float4 g0 = __ldg(bank + i);
    float4 g1 = __ldg(bank + Npad + i);
    float4 g2 = __ldg(bank + 2 * Npad + i);
    float4 g3 = __ldg(bank + 3 * Npad + i);
    return cand_from_groups(g0, g1, g2, g3);
}

struct Ctl { float pwm, delta, sd, cd; };      // input and sin/cos of the steering angle

template <bool MUFU_SIN>
__device__ __forceinline__ float psin(float t) { return MUFU_SIN ? sin_mufu(t) : sin_any(t); }

// Pacejka lateral force D*sin(C*atan(B*alpha))   (dynamic.py:148-149)
template <bool MUFU_SIN>
__device__ __forceinline__ float pacejka(float B, float C, float D, float alpha) {
    return D * psin<MUFU_SIN>(C * atan_any(B * alpha));
}

// Frx = (Cm1 - Cm2 vx) pwm - Cr0 - Cr2 vx^2   (dynamic.py:145)
__device__ __forceinline__ float drive_force(const Cand& p, float pwm, float vx) {
    float a = fmaf(-p.Cm2, vx, p.Cm1);
    float b = fmaf(-p.Cr2 * vx, vx, -p.Cr0);
    return fmaf(a, pwm, b);
}

// slip angles (dynamic.py:146-147): alphaf = delta - atan2(lf w + vy, |vx|), alphar = atan2(lr w - vy, |vx|)
__device__ __forceinline__ void slip_angles(const Cand& p, float delta, float vx, float vy, float w,
                                            float& af, float& ar) {
    float avx = fabsf(vx);
    float inv = rcp_newton(avx);
    af = delta - atan2_pos(fmaf(p.lf, w, vy), avx, inv);
    ar = atan2_pos(fmaf(p.lr, w, -vy), avx, inv);
}

struct Deriv { float vx, vy, w; };             // d/dt of (vx, vy, omega)   (dynamic.py:108-113)

template <bool MUFU_SIN>
__device__ __forceinline__ Deriv accel_from_slip(const Cand& p, const Ctl& u, float vx, float vy, float w,
                                                 float af, float ar) {
    float Frx = drive_force(p, u.pwm, vx);
    float Ffy = pacejka<MUFU_SIN>(p.Bf, p.Cf, p.Df, af);
    float Fry = pacejka<MUFU_SIN>(p.Br, p.Cr, p.Dr, ar);
    float Fc = Ffy * u.cd;
    Deriv d;
    d.vx = fmaf(fmaf(-Ffy, u.sd, Frx), p.inv_m, vy * w);
    d.vy = fmaf(Fry + Fc, p.inv_m, -vx * w);
    d.w = fmaf(Fc, p.lf_Iz, -Fry * p.lr_Iz);
    return d;
}

template <bool MUFU_SIN>
__device__ __forceinline__ Deriv accel(const Cand& p, const Ctl& u, float vx, float vy, float w) {
    float af, ar;
    slip_angles(p, u.delta, vx, vy, w, af, ar);
    return accel_from_slip<MUFU_SIN>(p, u, vx, vy, w, af, ar);
}

// longitudinal acceleration only (front tyre + drivetrain): what the last RK stage of the look-back needs
template <bool MUFU_SIN>
__device__ __forceinline__ float accel_vx_only(const Cand& p, const Ctl& u, float vx, float vy, float w) {
    float avx = fabsf(vx);
    float inv = rcp_newton(avx);
    float af = u.delta - atan2_pos(fmaf(p.lf, w, vy), avx, inv);
    float Frx = drive_force(p, u.pwm, vx);
    float Ffy = pacejka<MUFU_SIN>(p.Bf, p.Cf, p.Df, af);
    return fmaf(fmaf(-Ffy, u.sd, Frx), p.inv_m, vy * w);
}

// ---------------------------------------------------------------------------------------------------
// Straight-line ("fast") variants for the look-back kernel: no branches inside the step.  The tyre atan is
// branch-free for any argument; the slip-angle atan and the tiny heading rotations are only valid on
// |t| <= 0.5 and |e| <= 0.125, and `guard` accumulates max(2|t|, 8|e|) so that the caller can redo the whole
// step with the general routines in the rare case the guard exceeds 1 (or is NaN).
// ---------------------------------------------------------------------------------------------------
// Drivetrain force with the per-step constants hoisted: Frx = (Cm1 pwm - Cr0) - vx (Cm2 pwm + Cr2 vx)
struct Drive { float A, Bq; };
__device__ __forceinline__ Drive prep_drive(const Cand& p, float pwm) {
    Drive d;
    d.A = fmaf(p.Cm1, pwm, -p.Cr0);
    d.Bq = p.Cm2 * pwm;
    return d;
}
__device__ __forceinline__ float drive_force_fast(const Cand& p, const Drive& d, float vx) {
    return fmaf(-vx, fmaf(p.Cr2, vx, d.Bq), d.A);
}

template <bool MUFU_SIN>
__device__ __forceinline__ float pacejka_fast(float B, float C, float D, float alpha) {
    const float t = C * atan_full<MUFU_SIN>(B * alpha);
    return D * (MUFU_SIN ? sin_mufu(t) : sin_tyre(t));
}

// FULL_SLIP = false: slip tangents up to 0.5 (5-coefficient atan, guarded) -- the look-back, where every step
// starts from a measured, non-spinning state; true: branch-free atan2 for any slip angle -- the look-ahead rollouts,
// where unstable candidate models spin within the horizon (30 % of the warp-steps of config C3 see |t| > 0.5).
template <bool MUFU_SIN, bool FULL_SLIP = false>
__device__ __forceinline__ Deriv accel_fast(const Cand& p, const Ctl& u, const Drive& drv, float vx, float vy, float w,
                                            float& guard) {
    float af, ar;
    if (FULL_SLIP) {
        // no NaN restoration inside the slip angles: a NaN vx, vy or w also reaches the result through the drivetrain
        // force and the vy * w / vx * w terms below, and the rollout kernel re-does any step whose increments are NaN
        const float avx = fabsf(vx);
        af = u.delta - atan2_pos_full<false>(fmaf(p.lf, w, vy), avx);
        ar = atan2_pos_full<false>(fmaf(p.lr, w, -vy), avx);
    } else {
        const float inv = rcp_approx(fabsf(vx));        // 1 ulp: the tangents are small, |error| <= 1.2e-7 |t|
        const float tf = fmaf(p.lf, w, vy) * inv, tr = fmaf(p.lr, w, -vy) * inv;
        guard = fmaxf(guard, 2.0f * fmaxf(fabsf(tf), fabsf(tr)));
        af = u.delta - atan_half(tf);
        ar = atan_half(tr);
    }
    const float Frx = drive_force_fast(p, drv, vx);
    const float Ffy = pacejka_fast<MUFU_SIN>(p.Bf, p.Cf, p.Df, af);
    const float Fry = pacejka_fast<MUFU_SIN>(p.Br, p.Cr, p.Dr, ar);
    const float Fc = Ffy * u.cd;
    Deriv d;
    d.vx = fmaf(fmaf(-Ffy, u.sd, Frx), p.inv_m, vy * w);
    d.vy = fmaf(Fry + Fc, p.inv_m, -vx * w);
    d.w = fmaf(Fc, p.lf_Iz, -Fry * p.lr_Iz);
    return d;
}

// ---------------------------------------------------------------------------------------------------
// Generic RK4 step in increment form.  (s0, c0) = sin/cos of the heading at the start of the step.
// inc[] = y(t+h) - y(t) for (x, y, psi, vx, vy, omega).
// ---------------------------------------------------------------------------------------------------
template <bool MUFU_SIN>
__device__ __forceinline__ void rk4_increment(const Cand& p, const Ctl& u, float s0, float c0,
                                              float vx0, float vy0, float w0, float h, float inc[6]) {
    const float hh = 0.5f * h, h6 = h * (1.0f / 6.0f);
    // stage 1
    Deriv a1 = accel<MUFU_SIN>(p, u, vx0, vy0, w0);
    float xd1 = fmaf(vx0, c0, -vy0 * s0), yd1 = fmaf(vx0, s0, vy0 * c0);
    // stage 2: y0 + k1/2
    float vx2 = fmaf(hh, a1.vx, vx0), vy2 = fmaf(hh, a1.vy, vy0), w2 = fmaf(hh, a1.w, w0);
    float sd, cd;
    sincos_small(hh * w0, sd, cd);
    float s2 = fmaf(s0, cd, c0 * sd), c2 = fmaf(c0, cd, -s0 * sd);
    Deriv a2 = accel<MUFU_SIN>(p, u, vx2, vy2, w2);
    float xd2 = fmaf(vx2, c2, -vy2 * s2), yd2 = fmaf(vx2, s2, vy2 * c2);
    // stage 3: y0 + k2/2
    float vx3 = fmaf(hh, a2.vx, vx0), vy3 = fmaf(hh, a2.vy, vy0), w3 = fmaf(hh, a2.w, w0);
    sincos_small(hh * w2, sd, cd);
    float s3 = fmaf(s0, cd, c0 * sd), c3 = fmaf(c0, cd, -s0 * sd);
    Deriv a3 = accel<MUFU_SIN>(p, u, vx3, vy3, w3);
    float xd3 = fmaf(vx3, c3, -vy3 * s3), yd3 = fmaf(vx3, s3, vy3 * c3);
    // stage 4: y0 + k3
    float vx4 = fmaf(h, a3.vx, vx0), vy4 = fmaf(h, a3.vy, vy0), w4 = fmaf(h, a3.w, w0);
    sincos_small(h * w3, sd, cd);
    float s4 = fmaf(s0, cd, c0 * sd), c4 = fmaf(c0, cd, -s0 * sd);
    Deriv a4 = accel<MUFU_SIN>(p, u, vx4, vy4, w4);
    float xd4 = fmaf(vx4, c4, -vy4 * s4), yd4 = fmaf(vx4, s4, vy4 * c4);
    // y += (k1 + 2 k2 + 2 k3 + k4) / 6      (rk6.py:64-66)
    inc[0] = h6 * ((xd1 + xd4) + 2.0f * (xd2 + xd3));
    inc[1] = h6 * ((yd1 + yd4) + 2.0f * (yd2 + yd3));
    inc[2] = h6 * ((w0 + w4) + 2.0f * (w2 + w3));
    inc[3] = h6 * ((a1.vx + a4.vx) + 2.0f * (a2.vx + a3.vx));
    inc[4] = h6 * ((a1.vy + a4.vy) + 2.0f * (a2.vy + a3.vy));
    inc[5] = h6 * ((a1.w + a4.w) + 2.0f * (a2.w + a3.w));
}

// Straight-line generic RK4 step (look-ahead rollouts): same arithmetic as rk4_increment, no branches.  `guard`
// accumulates max(2|slip tangent|, 2|heading offset|); the step is only valid while guard <= 1.
template <bool MUFU_SIN>
__device__ __forceinline__ void rk4_increment_fast(const Cand& p, const Ctl& u, float s0, float c0, float vx0, float vy0,
                                                   float w0, float h, float inc[6], float& guard) {
    const float hh = 0.5f * h, h6 = h * (1.0f / 6.0f);
    // weighted sums (k1 + 2 k2 + 2 k3 + k4) are accumulated stage by stage to keep few values live
    const Drive drv = prep_drive(p, u.pwm);
    Deriv a = accel_fast<MUFU_SIN, true>(p, u, drv, vx0, vy0, w0, guard);
    float sx = fmaf(vx0, c0, -vy0 * s0), sy = fmaf(vx0, s0, vy0 * c0);
    float sw = w0, svx = a.vx, svy = a.vy, sdw = a.w;
    float vx = fmaf(hh, a.vx, vx0), vy = fmaf(hh, a.vy, vy0), w = fmaf(hh, a.w, w0);
    float sd, cd, d = hh * w0, dmax = fabsf(d);
    sincos_half(d, sd, cd);
    float sn = fmaf(s0, cd, c0 * sd), cs = fmaf(c0, cd, -s0 * sd);
    // stage 2
    a = accel_fast<MUFU_SIN, true>(p, u, drv, vx, vy, w, guard);
    sx = fmaf(2.0f, fmaf(vx, cs, -vy * sn), sx);
    sy = fmaf(2.0f, fmaf(vx, sn, vy * cs), sy);
    sw = fmaf(2.0f, w, sw); svx = fmaf(2.0f, a.vx, svx); svy = fmaf(2.0f, a.vy, svy); sdw = fmaf(2.0f, a.w, sdw);
    d = hh * w;
    dmax = fmaxf(dmax, fabsf(d));
    vx = fmaf(hh, a.vx, vx0); vy = fmaf(hh, a.vy, vy0); w = fmaf(hh, a.w, w0);
    sincos_half(d, sd, cd);
    sn = fmaf(s0, cd, c0 * sd); cs = fmaf(c0, cd, -s0 * sd);
    // stage 3
    a = accel_fast<MUFU_SIN, true>(p, u, drv, vx, vy, w, guard);
    sx = fmaf(2.0f, fmaf(vx, cs, -vy * sn), sx);
    sy = fmaf(2.0f, fmaf(vx, sn, vy * cs), sy);
    sw = fmaf(2.0f, w, sw); svx = fmaf(2.0f, a.vx, svx); svy = fmaf(2.0f, a.vy, svy); sdw = fmaf(2.0f, a.w, sdw);
    d = h * w;
    dmax = fmaxf(dmax, fabsf(d));
    vx = fmaf(h, a.vx, vx0); vy = fmaf(h, a.vy, vy0); w = fmaf(h, a.w, w0);
    sincos_half(d, sd, cd);
    sn = fmaf(s0, cd, c0 * sd); cs = fmaf(c0, cd, -s0 * sd);
    // stage 4
    a = accel_fast<MUFU_SIN, true>(p, u, drv, vx, vy, w, guard);
    sx += fmaf(vx, cs, -vy * sn);
    sy += fmaf(vx, sn, vy * cs);
    guard = fmaxf(guard, 2.0f * dmax);
    inc[0] = h6 * sx;
    inc[1] = h6 * sy;
    inc[2] = h6 * (sw + w);
    inc[3] = h6 * (svx + a.vx);
    inc[4] = h6 * (svy + a.vy);
    inc[5] = h6 * (sdw + a.w);
}

// out-of-line general step for the rare fallback of the rollout kernel
struct Inc6 { float v[6]; };
static __device__ __noinline__ Inc6 rk4_increment_general(const float4* __restrict__ bank, int Npad, int cand, float pwm,
                                                          float delta, float s0, float c0, float vx0, float vy0,
                                                          float w0, float h) {
    const Cand p = load_cand(bank, Npad, cand);
    Ctl u;
    u.pwm = pwm; u.delta = delta;
    sincos_small(delta, u.sd, u.cd);
    Inc6 r;
    rk4_increment<false>(p, u, s0, c0, vx0, vy0, w0, h, r.v);
    return r;
}

// ---------------------------------------------------------------------------------------------------
// Look-back step.  All candidates start the step from the SAME measured state, so everything that does
// not depend on the candidate was computed once on the host in fp64 and sits in the history row:
//   q0 = sin,cos(psi0)            sin,cos(psi0 + h w0/2)          (stage 1 / stages 2,3 base heading)
//   q1 = sin,cos(psi0 + h w0)     vx0 vy0                         (stage 4 base heading)
//   q2 = w0 pwm delta sin(delta)
//   q3 = cos(delta)  mdx mdy mdpsi     measured increments MINUS their candidate-invariant part:
//        mdx = dx_meas - h/6 xdot(x0), mdy likewise, mdpsi = dpsi_meas - h w0
//   q4 = dvx_meas (hi, lo split)  alphaf1 alphar1 (stage-1 slip angles, valid when lf, lr are bank-wide)
// Returns sum over (x, y, psi, vx) of the squared increment error = 4 * the errors of rt.py:349.
// ---------------------------------------------------------------------------------------------------
struct HistRow { float4 q0, q1, q2, q3, q4; };

// Step sizes: the stage states use float(h); the final increments use h/6 and h^2/6 split into hi + lo floats
// (float(0.02) is off by 2e-8 relative, which would bias every increment by ~1e-9, a visible fraction of the
// 1e-4-sized differences being squared).
struct StepSize { float h, hh, h6, h6_lo, hh6, hh6_lo; };

static inline StepSize make_step(double Ts) {
    StepSize z;
    z.h = (float)Ts;
    z.hh = (float)(0.5 * Ts);
    z.h6 = (float)(Ts / 6.0);
    z.h6_lo = (float)(Ts / 6.0 - (double)z.h6);
    z.hh6 = (float)(Ts * Ts / 6.0);
    z.hh6_lo = (float)(Ts * Ts / 6.0 - (double)z.hh6);
    return z;
}

template <bool GEOM_SHARED, bool MUFU_SIN>
__device__ __forceinline__ float lookback_step(const Cand& p, const HistRow& r, const StepSize& z) {
    const float h = z.h, hh = z.hh;
    const float vx0 = r.q1.z, vy0 = r.q1.w, w0 = r.q2.x;
    Ctl u;
    u.pwm = r.q2.y; u.delta = r.q2.z; u.sd = r.q2.w; u.cd = r.q3.x;
    // stage 1 (heading and kinematics are candidate-invariant)
    Deriv a1;
    if (GEOM_SHARED) a1 = accel_from_slip<MUFU_SIN>(p, u, vx0, vy0, w0, r.q4.z, r.q4.w);
    else             a1 = accel<MUFU_SIN>(p, u, vx0, vy0, w0);
    // stage 2: heading psi0 + h w0/2 is candidate-invariant
    float vx2 = fmaf(hh, a1.vx, vx0), vy2 = fmaf(hh, a1.vy, vy0), w2 = fmaf(hh, a1.w, w0);
    Deriv a2 = accel<MUFU_SIN>(p, u, vx2, vy2, w2);
    float xs = fmaf(vx2, r.q0.w, -vy2 * r.q0.z), ys = fmaf(vx2, r.q0.z, vy2 * r.q0.w);
    // stage 3: heading = (psi0 + h w0/2) + (h/2)(w2 - w0)
    float vx3 = fmaf(hh, a2.vx, vx0), vy3 = fmaf(hh, a2.vy, vy0), w3 = fmaf(hh, a2.w, w0);
    float sd, cd;
    sincos_small(hh * (hh * a1.w), sd, cd);
    float s3 = fmaf(r.q0.z, cd, r.q0.w * sd), c3 = fmaf(r.q0.w, cd, -r.q0.z * sd);
    Deriv a3 = accel<MUFU_SIN>(p, u, vx3, vy3, w3);
    xs += fmaf(vx3, c3, -vy3 * s3);
    ys += fmaf(vx3, s3, vy3 * c3);
    // stage 4: heading = (psi0 + h w0) + h (w3 - w0); only xdot, ydot, vxdot are needed for the score
    float vx4 = fmaf(h, a3.vx, vx0), vy4 = fmaf(h, a3.vy, vy0), w4 = fmaf(h, a3.w, w0);
    sincos_small(h * (hh * a2.w), sd, cd);
    float s4 = fmaf(r.q1.x, cd, r.q1.y * sd), c4 = fmaf(r.q1.y, cd, -r.q1.x * sd);
    float a4vx = accel_vx_only<MUFU_SIN>(p, u, vx4, vy4, w4);
    float xd4 = fmaf(vx4, c4, -vy4 * s4), yd4 = fmaf(vx4, s4, vy4 * c4);
    // increment errors
    const float sx = fmaf(2.0f, xs, xd4), sy = fmaf(2.0f, ys, yd4);
    const float sw = (a1.w + a2.w) + a3.w, sv = (a1.vx + a4vx) + 2.0f * (a2.vx + a3.vx);
    float ex = fmaf(z.h6_lo, sx, fmaf(z.h6, sx, -r.q3.y));
    float ey = fmaf(z.h6_lo, sy, fmaf(z.h6, sy, -r.q3.z));
    float epsi = fmaf(z.hh6_lo, sw, fmaf(z.hh6, sw, -r.q3.w));
    float evx = fmaf(z.h6_lo, sv, fmaf(z.h6, sv, -r.q4.x)) - r.q4.y;
    return fmaf(ex, ex, fmaf(ey, ey, fmaf(epsi, epsi, evx * evx)));
}

// Fast look-back step (same arithmetic, straight-line code).  Returns the squared increment error and sets
// `ok` to false when a guard tripped (caller falls back to lookback_step).
template <bool GEOM_SHARED, bool MUFU_SIN>
__device__ __forceinline__ float lookback_step_fast(const Cand& p, const HistRow& r, const StepSize& z, bool& ok) {
    const float h = z.h, hh = z.hh;
    const float vx0 = r.q1.z, vy0 = r.q1.w, w0 = r.q2.x;
    Ctl u;
    u.pwm = r.q2.y; u.delta = r.q2.z; u.sd = r.q2.w; u.cd = r.q3.x;
    float guard = 0.0f;
    const Drive drv = prep_drive(p, u.pwm);
    // stage 1
    Deriv a1;
    if (GEOM_SHARED) {
        const float Frx = drive_force_fast(p, drv, vx0);
        const float Ffy = pacejka_fast<MUFU_SIN>(p.Bf, p.Cf, p.Df, r.q4.z);
        const float Fry = pacejka_fast<MUFU_SIN>(p.Br, p.Cr, p.Dr, r.q4.w);
        const float Fc = Ffy * u.cd;
        a1.vx = fmaf(fmaf(-Ffy, u.sd, Frx), p.inv_m, vy0 * w0);
        a1.vy = fmaf(Fry + Fc, p.inv_m, -vx0 * w0);
        a1.w = fmaf(Fc, p.lf_Iz, -Fry * p.lr_Iz);
    } else {
        a1 = accel_fast<MUFU_SIN>(p, u, drv, vx0, vy0, w0, guard);
    }
    // stage 2
    const float vx2 = fmaf(hh, a1.vx, vx0), vy2 = fmaf(hh, a1.vy, vy0), w2 = fmaf(hh, a1.w, w0);
    const Deriv a2 = accel_fast<MUFU_SIN>(p, u, drv, vx2, vy2, w2, guard);
    float xs = fmaf(vx2, r.q0.w, -vy2 * r.q0.z), ys = fmaf(vx2, r.q0.z, vy2 * r.q0.w);
    // stage 3
    const float vx3 = fmaf(hh, a2.vx, vx0), vy3 = fmaf(hh, a2.vy, vy0), w3 = fmaf(hh, a2.w, w0);
    const float e3 = hh * (hh * a1.w);
    float sd, cd;
    sincos_tiny(e3, sd, cd);
    const float s3 = fmaf(r.q0.z, cd, r.q0.w * sd), c3 = fmaf(r.q0.w, cd, -r.q0.z * sd);
    const Deriv a3 = accel_fast<MUFU_SIN>(p, u, drv, vx3, vy3, w3, guard);
    xs += fmaf(vx3, c3, -vy3 * s3);
    ys += fmaf(vx3, s3, vy3 * c3);
    // stage 4 (front tyre and drivetrain only)
    const float vx4 = fmaf(h, a3.vx, vx0), vy4 = fmaf(h, a3.vy, vy0), w4 = fmaf(h, a3.w, w0);
    const float e4 = h * (hh * a2.w);
    sincos_tiny(e4, sd, cd);
    const float s4 = fmaf(r.q1.x, cd, r.q1.y * sd), c4 = fmaf(r.q1.y, cd, -r.q1.x * sd);
    guard = fmaxf(guard, 8.0f * fmaxf(fabsf(e3), fabsf(e4)));
    const float inv4 = rcp_approx(fabsf(vx4));
    const float tf4 = fmaf(p.lf, w4, vy4) * inv4;
    guard = fmaxf(guard, 2.0f * fabsf(tf4));
    const float Ffy4 = pacejka_fast<MUFU_SIN>(p.Bf, p.Cf, p.Df, u.delta - atan_half(tf4));
    const float a4vx = fmaf(fmaf(-Ffy4, u.sd, drive_force_fast(p, drv, vx4)), p.inv_m, vy4 * w4);
    const float xd4 = fmaf(vx4, c4, -vy4 * s4), yd4 = fmaf(vx4, s4, vy4 * c4);
    // increment errors
    const float sx = fmaf(2.0f, xs, xd4), sy = fmaf(2.0f, ys, yd4);
    const float sw = (a1.w + a2.w) + a3.w, sv = (a1.vx + a4vx) + 2.0f * (a2.vx + a3.vx);
    const float ex = fmaf(z.h6_lo, sx, fmaf(z.h6, sx, -r.q3.y));
    const float ey = fmaf(z.h6_lo, sy, fmaf(z.h6, sy, -r.q3.z));
    const float epsi = fmaf(z.hh6_lo, sw, fmaf(z.hh6, sw, -r.q3.w));
    const float evx = fmaf(z.h6_lo, sv, fmaf(z.h6, sv, -r.q4.x)) - r.q4.y;
    const float e2 = fmaf(ex, ex, fmaf(ey, ey, fmaf(epsi, epsi, evx * evx)));
    ok = (guard <= 1.0f) && (e2 == e2);       // fmaxf drops NaN operands, so a NaN result is checked explicitly
    return e2;
}

// out-of-line general step for the rare fallback: reloads the candidate and the history row itself so that
// the hot loop does not have to keep them addressable (no local-memory traffic on the fast path)
template <bool GEOM_SHARED, bool MUFU_SIN>
__device__ __noinline__ float lookback_step_general(const float4* __restrict__ bank, int Npad, int cand,
                                                    const float4* srow_w, StepSize z) {
    const Cand p = load_cand(bank, Npad, cand);
    HistRow r;
    r.q0 = srow_w[0]; r.q1 = srow_w[1]; r.q2 = srow_w[2]; r.q3 = srow_w[3]; r.q4 = srow_w[4];
    return lookback_step<GEOM_SHARED, MUFU_SIN>(p, r, z);
}

}  // namespace llampc
