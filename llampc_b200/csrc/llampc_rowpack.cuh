// History-row packing shared by the host entry point (llampc_hist_row_pack_h) and the device kernel that packs
// the rows of many vehicles (Monte-Carlo layout).  fp64 throughout; see llampc_model.cuh (HistRow) for the layout.
#pragma once
#include <math.h>
#include "../../include/llampc_b200.h"

#ifdef __CUDACC__
#define LLAMPC_HD __host__ __device__
#else
#define LLAMPC_HD
#endif

namespace llampc {

LLAMPC_HD inline void pack_hist_row(const double* x_k, const double* u_k, const double* x_k1, double h, double lf_shared,
                                    double lr_shared, float* r, double* row64) {
    const double psi = x_k[2], vx = x_k[3], vy = x_k[4], w = x_k[5];
    const double pwm = u_k[0], delta = u_k[1];
    // q0: heading at stage 1 and at stages 2/3 (psi + h w/2);  q1: heading base of stage 4 (psi + h w), vx, vy
    r[0] = (float)sin(psi);               r[1] = (float)cos(psi);
    r[2] = (float)sin(psi + 0.5 * h * w); r[3] = (float)cos(psi + 0.5 * h * w);
    r[4] = (float)sin(psi + h * w);       r[5] = (float)cos(psi + h * w);
    r[6] = (float)vx;                     r[7] = (float)vy;
    // q2
    r[8] = (float)w; r[9] = (float)pwm; r[10] = (float)delta; r[11] = (float)sin(delta);
    // q3: cos(delta) and the measured increments minus their candidate-invariant parts
    const double xd0 = vx * cos(psi) - vy * sin(psi), yd0 = vx * sin(psi) + vy * cos(psi);
    r[12] = (float)cos(delta);
    r[13] = (float)((x_k1[0] - x_k[0]) - h / 6.0 * xd0);
    r[14] = (float)((x_k1[1] - x_k[1]) - h / 6.0 * yd0);
    r[15] = (float)((x_k1[2] - x_k[2]) - h * w);
    // q4: measured vx increment as hi + lo, stage-1 slip angles for bank-wide geometry
    const double dvx = x_k1[3] - x_k[3];
    const float hi = (float)dvx;
    r[16] = hi;
    r[17] = (float)(dvx - (double)hi);
    if (lf_shared == lf_shared && lr_shared == lr_shared) {
        r[18] = (float)(delta - atan2(lf_shared * w + vy, fabs(vx)));
        r[19] = (float)atan2(lr_shared * w - vy, fabs(vx));
    } else {
        r[18] = 0.0f; r[19] = 0.0f;
    }
    if (row64) {
        for (int i = 0; i < 6; ++i) row64[i] = x_k[i];
        row64[6] = pwm; row64[7] = delta;
        for (int i = 0; i < 4; ++i) row64[8 + i] = x_k1[i];
    }
}

// One candidate of the packed bank (layout in include/llampc_b200.h, llampc_bank_pack_h) from its 14 fp64
// parameters in LLAMPC_NPARAM order: lf lr mass Iz Bf Br Cf Cr Df Dr Cm1 Cm2 Cr0 Cr2.
LLAMPC_HD inline void pack_candidate(const double* v, float* packed, int Npad, int i) {
    float* g0 = packed + ((size_t)0 * Npad + i) * 4;
    float* g1 = packed + ((size_t)1 * Npad + i) * 4;
    float* g2 = packed + ((size_t)2 * Npad + i) * 4;
    float* g3 = packed + ((size_t)3 * Npad + i) * 4;
    g0[0] = (float)v[4]; g0[1] = (float)v[6]; g0[2] = (float)v[8]; g0[3] = (float)v[5];          // Bf Cf Df Br
    g1[0] = (float)v[7]; g1[1] = (float)v[9]; g1[2] = (float)(1.0 / v[2]); g1[3] = (float)v[0];  // Cr Dr 1/m lf
    g2[0] = (float)v[1]; g2[1] = (float)(v[0] / v[3]); g2[2] = (float)(v[1] / v[3]); g2[3] = (float)v[10];   // lr lf/Iz lr/Iz Cm1
    g3[0] = (float)v[11]; g3[1] = (float)v[12]; g3[2] = (float)v[13]; g3[3] = 0.0f;              // Cm2 Cr0 Cr2
}

}  // namespace llampc
