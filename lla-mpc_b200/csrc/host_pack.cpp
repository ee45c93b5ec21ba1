// Host-side packing of reference-layout fp64 data into the device layouts (no CUDA calls here).
#include <math.h>
#include <stddef.h>

#include "../../include/llampc_b200.h"
#include "llampc_rowpack.cuh"

extern "C" int llampc_abi_version(void) { return LLAMPC_ABI_VERSION; }

extern "C" const char* llampc_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case LLAMPC_E_ARG: return "llampc: null pointer or non-positive / unsupported size";
        case LLAMPC_E_ALIGN: return "llampc: pointer or stride is not 16-byte aligned";
        case LLAMPC_E_RANGE: return "llampc: W, K or H outside the compiled limits";
        default: return code > 0 ? "llampc: CUDA runtime error (value is the cudaError_t)" : "llampc: unknown error";
    }
}

// order of LLAMPC_NPARAM: lf lr mass Iz Bf Br Cf Cr Df Dr Cm1 Cm2 Cr0 Cr2
enum { P_LF, P_LR, P_MASS, P_IZ, P_BF, P_BR, P_CF, P_CR, P_DF, P_DR, P_CM1, P_CM2, P_CR0, P_CR2 };

extern "C" int llampc_bank_pack_h(const double* const* params_h, const int* is_array, int N, int Npad, float* packed_h) {
    if (!params_h || !is_array || !packed_h || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    for (int j = 0; j < LLAMPC_NPARAM; ++j)
        if (!params_h[j]) return LLAMPC_E_ARG;
    for (int i = 0; i < Npad; ++i) {
        const int s = i < N ? i : N - 1;
        double v[LLAMPC_NPARAM];
        for (int j = 0; j < LLAMPC_NPARAM; ++j) v[j] = params_h[j][is_array[j] ? s : 0];
        float* g0 = packed_h + ((size_t)0 * Npad + i) * 4;
        float* g1 = packed_h + ((size_t)1 * Npad + i) * 4;
        float* g2 = packed_h + ((size_t)2 * Npad + i) * 4;
        float* g3 = packed_h + ((size_t)3 * Npad + i) * 4;
        g0[0] = (float)v[P_BF]; g0[1] = (float)v[P_CF]; g0[2] = (float)v[P_DF]; g0[3] = (float)v[P_BR];
        g1[0] = (float)v[P_CR]; g1[1] = (float)v[P_DR]; g1[2] = (float)(1.0 / v[P_MASS]); g1[3] = (float)v[P_LF];
        g2[0] = (float)v[P_LR]; g2[1] = (float)(v[P_LF] / v[P_IZ]); g2[2] = (float)(v[P_LR] / v[P_IZ]); g2[3] = (float)v[P_CM1];
        g3[0] = (float)v[P_CM2]; g3[1] = (float)v[P_CR0]; g3[2] = (float)v[P_CR2]; g3[3] = 0.0f;
    }
    return 0;
}

extern "C" int llampc_hist_row_pack_h(const double* x_k, const double* u_k, const double* x_k1, double Ts,
                                      double lf_shared, double lr_shared, float* r, double* row64_h) {
    if (!x_k || !u_k || !x_k1 || !r || !(Ts > 0.0)) return LLAMPC_E_ARG;
    llampc::pack_hist_row(x_k, u_k, x_k1, Ts, lf_shared, lr_shared, r, row64_h);
    return 0;
}
