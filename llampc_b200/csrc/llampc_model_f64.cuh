// fp64 device restatement of the model, used (a) to re-score the look-back finalists exactly and (b) for the
// plant step of the Monte-Carlo configuration.  Same operation order as the NumPy reference:
// Dynamic.calc_forces / _diffequation (llampc/models/dynamic.py:76-96,156-193; batch twins :98-154),
// odeintRK4_batch (llampc/utils/rk6.py:50-68), odeintRK6 (rk6.py:13-28).
#pragma once
#include <cuda_runtime.h>

namespace llampc {

struct Params64 { double lf, lr, mass, Iz, Bf, Br, Cf, Cr, Df, Dr, Cm1, Cm2, Cr0, Cr2; };

__device__ __forceinline__ void rhs64(const Params64& p, const double y[6], double pwm, double steer, double f[6]) {
    double psi = y[2], vx = y[3], vy = y[4], om = y[5];
    double Frx = (p.Cm1 - p.Cm2 * vx) * pwm - p.Cr0 - p.Cr2 * (vx * vx);
    double alphaf = steer - atan2(p.lf * om + vy, fabs(vx));
    double alphar = atan2(p.lr * om - vy, fabs(vx));
    double Ffy = p.Df * sin(p.Cf * atan(p.Bf * alphaf));
    double Fry = p.Dr * sin(p.Cr * atan(p.Br * alphar));
    double sp, cp, sdl, cdl;
    sincos(psi, &sp, &cp);
    sincos(steer, &sdl, &cdl);
    f[0] = vx * cp - vy * sp;
    f[1] = vx * sp + vy * cp;
    f[2] = om;
    f[3] = 1 / p.mass * (Frx - Ffy * sdl) + vy * om;
    f[4] = 1 / p.mass * (Fry + Ffy * cdl) - vx * om;
    f[5] = 1 / p.Iz * (Ffy * p.lf * cdl - Fry * p.lr);
}

__device__ __forceinline__ void rk4_step64(const Params64& p, const double y0[6], double pwm, double steer,
                                           double h, double out[6]) {
    double k1[6], k2[6], k3[6], k4[6], t[6];
    rhs64(p, y0, pwm, steer, k1);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k1[i] *= h; t[i] = y0[i] + k1[i] / 2; }
    rhs64(p, t, pwm, steer, k2);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k2[i] *= h; t[i] = y0[i] + k2[i] / 2; }
    rhs64(p, t, pwm, steer, k3);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k3[i] *= h; t[i] = y0[i] + k3[i]; }
    rhs64(p, t, pwm, steer, k4);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k4[i] *= h; out[i] = y0[i] + (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6; }
}

// The same RK4 step evaluated by a PAIR of adjacent lanes: the even lane evaluates the front tyre, the odd lane the
// rear tyre (the two Pacejka chains are the bulk of the fp64 latency), the forces are swapped with one shuffle per
// stage; every other quantity is computed redundantly by both lanes with the same operations, so both lanes hold
// bit-identical results, equal to rk4_step64's.  Must be called by both lanes of the pair (full-warp shuffles).
__device__ __forceinline__ void rhs64_pair(const Params64& p, const double y[6], double pwm, double steer, double sdl,
                                           double cdl, bool rear, double f[6]) {
    const double psi = y[2], vx = y[3], vy = y[4], om = y[5];
    const double Frx = (p.Cm1 - p.Cm2 * vx) * pwm - p.Cr0 - p.Cr2 * (vx * vx);
    double F;
    if (rear) {
        const double alphar = atan2(p.lr * om - vy, fabs(vx));
        F = p.Dr * sin(p.Cr * atan(p.Br * alphar));
    } else {
        const double alphaf = steer - atan2(p.lf * om + vy, fabs(vx));
        F = p.Df * sin(p.Cf * atan(p.Bf * alphaf));
    }
    const double other = __shfl_xor_sync(0xffffffffu, F, 1);
    const double Ffy = rear ? other : F, Fry = rear ? F : other;
    double sp, cp;
    sincos(psi, &sp, &cp);
    f[0] = vx * cp - vy * sp;
    f[1] = vx * sp + vy * cp;
    f[2] = om;
    f[3] = 1 / p.mass * (Frx - Ffy * sdl) + vy * om;
    f[4] = 1 / p.mass * (Fry + Ffy * cdl) - vx * om;
    f[5] = 1 / p.Iz * (Ffy * p.lf * cdl - Fry * p.lr);
}

__device__ __forceinline__ void rk4_step64_pair(const Params64& p, const double y0[6], double pwm, double steer,
                                                double h, bool rear, double out[6]) {
    double sdl, cdl;
    sincos(steer, &sdl, &cdl);                     // the input is constant over the step
    double k1[6], k2[6], k3[6], k4[6], t[6];
    rhs64_pair(p, y0, pwm, steer, sdl, cdl, rear, k1);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k1[i] *= h; t[i] = y0[i] + k1[i] / 2; }
    rhs64_pair(p, t, pwm, steer, sdl, cdl, rear, k2);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k2[i] *= h; t[i] = y0[i] + k2[i] / 2; }
    rhs64_pair(p, t, pwm, steer, sdl, cdl, rear, k3);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k3[i] *= h; t[i] = y0[i] + k3[i]; }
    rhs64_pair(p, t, pwm, steer, sdl, cdl, rear, k4);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k4[i] *= h; out[i] = y0[i] + (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]) / 6; }
}

// 6-stage Runge-Kutta-Fehlberg step with the 5th-order weights (rk6.py:14,19-27)
__device__ __forceinline__ void rk6_step64(const Params64& p, const double y0[6], double pwm, double steer,
                                           double h, double out[6]) {
    double k1[6], k2[6], k3[6], k4[6], k5[6], k6[6], t[6];
    rhs64(p, y0, pwm, steer, k1);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k1[i] *= h; t[i] = y0[i] + k1[i] / 4; }
    rhs64(p, t, pwm, steer, k2);
#pragma unroll
    for (int i = 0; i < 6; ++i) { k2[i] *= h; t[i] = y0[i] + 3.0 / 32 * k1[i] + 9.0 / 32 * k2[i]; }
    rhs64(p, t, pwm, steer, k3);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        k3[i] *= h;
        t[i] = y0[i] + 1932.0 / 2197 * k1[i] - 7200.0 / 2197 * k2[i] + 7296.0 / 2197 * k3[i];
    }
    rhs64(p, t, pwm, steer, k4);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        k4[i] *= h;
        t[i] = y0[i] + 439.0 / 216 * k1[i] - 8 * k2[i] + 3680.0 / 513 * k3[i] - 845.0 / 4104 * k4[i];
    }
    rhs64(p, t, pwm, steer, k5);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        k5[i] *= h;
        t[i] = y0[i] - 8.0 / 27 * k1[i] + 2 * k2[i] - 3544.0 / 2565 * k3[i] + 1859.0 / 4104 * k4[i] - 11.0 / 40 * k5[i];
    }
    rhs64(p, t, pwm, steer, k6);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        k6[i] *= h;
        // gamma . K in index order, like np.matmul(gamma, K) (gamma[1] = 0)
        double acc = 16.0 / 135 * k1[i];
        acc += 0.0 * k2[i];
        acc += 6656.0 / 12825 * k3[i];
        acc += 28561.0 / 56430 * k4[i];
        acc += -9.0 / 50 * k5[i];
        acc += 2.0 / 55 * k6[i];
        out[i] = y0[i] + acc;
    }
}

}  // namespace llampc
