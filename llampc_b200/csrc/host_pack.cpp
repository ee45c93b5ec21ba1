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
        case LLAMPC_E_PEER: return "llampc: multi-GPU exchange timed out (a peer rank did not deliver within ~1 s): no decision for this tick";
        default: return code > 0 ? "llampc: CUDA runtime error (value is the cudaError_t)" : "llampc: unknown error";
    }
}

extern "C" int llampc_bank_pack_h(const double* const* params_h, const int* is_array, int N, int Npad, float* packed_h,
                                  float* sin_arg_max_h) {
    if (!params_h || !is_array || !packed_h || N <= 0 || Npad < N) return LLAMPC_E_ARG;
    for (int j = 0; j < LLAMPC_NPARAM; ++j)
        if (!params_h[j]) return LLAMPC_E_ARG;
    double cmax = 0.0;                             // max(|Cf|, |Cr|): bounds the tyre-sine argument C atan(B alpha) by cmax pi/2
    for (int i = 0; i < Npad; ++i) {
        const int s = i < N ? i : N - 1;
        double v[LLAMPC_NPARAM];
        for (int j = 0; j < LLAMPC_NPARAM; ++j) v[j] = params_h[j][is_array[j] ? s : 0];
        llampc::pack_candidate(v, packed_h, Npad, i);
        const double c = fmax(fabs(v[6]), fabs(v[7]));
        if (!(c <= cmax)) cmax = c == c ? c : INFINITY;      // a NaN parameter forces the range-independent sine
    }
    if (sin_arg_max_h) *sin_arg_max_h = (float)(cmax * 1.5707963267948966);
    return 0;
}

extern "C" int llampc_hist_row_pack_h(const double* x_k, const double* u_k, const double* x_k1, double Ts,
                                      double lf_shared, double lr_shared, float* r, double* row64_h) {
    if (!x_k || !u_k || !x_k1 || !r || !(Ts > 0.0)) return LLAMPC_E_ARG;
    llampc::pack_hist_row(x_k, u_k, x_k1, Ts, lf_shared, lr_shared, r, row64_h);
    return 0;
}
