// Device-side look-ahead reference generator: the reference's ConstantSpeed planner (llampc/mpc/planner.py:12-67)
// for V vehicles at once, fp64, one thread per vehicle.  Pieces restated: Track.project_fast
// (llampc/tracks/track.py:147-160) with Projection (llampc/utils/projection.py:11-38) over a 10-point window of
// the raceline, arc-length march with the mu-interpolated speed profile, Spline / Spline2D evaluation
// (llampc/utils/pycubicspline.py:47-65,155-162; bisect index, a + b dx + c dx^2 + d dx^3).
// The spline coefficient tables are produced once on the host (llampc_b200/tracks.py).
// Compiled with --fmad=false semantics where it matters (explicit __dmul_rn/__dadd_rn) so the arithmetic follows
// the reference's operation order.
#include "llampc_common.cuh"

namespace llampc {

__device__ __forceinline__ double dnorm2(double a, double b) { return sqrt(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b))); }

// projection.py:11-38 -- projection of (px, py) on the segment (x1, y1)-(x2, y2); returns the distance
__device__ double project_segment(double px, double py, double x1, double y1, double x2, double y2) {
    double d1x = x2 - x1, d1y = y2 - y1;
    const double n1 = dnorm2(d1x, d1y);
    d1x /= n1; d1y /= n1;
    const double t = __dadd_rn(__dmul_rn(px - x1, d1x), __dmul_rn(py - y1, d1y));
    double qx = __dadd_rn(x1, __dmul_rn(d1x, t)), qy = __dadd_rn(y1, __dmul_rn(d1y, t));
    double d2x = qx - x1, d2y = qy - y1, d3x = qx - x2, d3y = qy - y2;
    const double n2 = dnorm2(d2x, d2y), n3 = dnorm2(d3x, d3y);
    if (n2 > 0 && n3 > 0) {
        d2x /= n2; d2y /= n2; d3x /= n3; d3y /= n3;
        const bool on_line = dnorm2(d2x - d3x, d2y - d3y) > 1e-10;
        if (!on_line) {
            if (dnorm2(x1 - qx, y1 - qy) < dnorm2(x2 - qx, y2 - qy)) { qx = x1; qy = y1; }
            else { qx = x2; qy = y2; }
        }
    }
    return dnorm2(px - qx, py - qy);
}

// bisect.bisect(s, t) - 1
__device__ __forceinline__ int seg_index(const double* __restrict__ s, int n, double t) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t < s[mid]) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

__device__ __forceinline__ double cubic(const double* __restrict__ c4, double dx) {
    const double dx2 = __dmul_rn(dx, dx), dx3 = __dmul_rn(dx2, dx);
    return __dadd_rn(__dadd_rn(__dadd_rn(c4[0], __dmul_rn(c4[1], dx)), __dmul_rn(c4[2], dx2)), __dmul_rn(c4[3], dx3));
}

__global__ void __launch_bounds__(128)
planner_kernel(const double* __restrict__ s, const double* __restrict__ xy, const double* __restrict__ coef,
               const double* __restrict__ mus, int n, int n_mu, const double* __restrict__ states, int V,
               const int* __restrict__ projidx_in, const double* __restrict__ curr_mu, int mu_shared, int N, double Ts,
               double scale, float* __restrict__ xref32, double* __restrict__ xref64, int* __restrict__ projidx_out,
               double* __restrict__ vr_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const double px = states[(size_t)v * 6], py = states[(size_t)v * 6 + 1], v0 = states[(size_t)v * 6 + 3];
    const double mu = curr_mu[mu_shared ? 0 : v];
    int pid = projidx_in[v];
    // project onto raceline[:, pid:pid+10]  (planner.py:26, numpy slicing truncates at the end of the table)
    const int n_way = max(0, min(10, n - pid));
    int best = 0;
    double best_d = 0.0;
    for (int i = 0; i + 1 < n_way; ++i) {
        const double* a = xy + (size_t)(pid + i) * 2;
        const double d = project_segment(px, py, a[0], a[1], a[2], a[3]);
        if (i == 0 || d < best_d) { best_d = d; best = i; }
    }
    pid += best;
    // arc length of raceline[:, :pid+2]  (planner.py:30,36)
    double dist = s[min(pid + 1, n - 1)];
    const double s_last = s[n - 1];
    const int stride = 4 * (2 + n_mu);
    // speed table selection (planner.py:49-62)
    int i_hi, i_lo;
    double w_lo = 0.0, w_hi = 1.0;
    if (mu < mus[0]) { i_hi = 0; i_lo = -1; }
    else if (mu > mus[n_mu - 1]) { i_hi = n_mu - 1; i_lo = -1; }
    else {
        i_hi = 0;
        for (int i = 0; i < n_mu; ++i) { i_hi = i; if (mus[i] >= mu) break; }
        i_lo = (i_hi - 1 + n_mu) % n_mu;                          // Python's spline_v[i-1] wraps for i = 0
        const double den = mus[i_hi] - mus[i_lo];
        w_lo = (mus[i_hi] - mu) / den;
        w_hi = (mu - mus[i_lo]) / den;
    }
    if (xref32) { xref32[(size_t)v * (N + 1) * 2] = (float)px; xref32[(size_t)v * (N + 1) * 2 + 1] = (float)py; }
    if (xref64) { xref64[(size_t)v * (N + 1) * 2] = px; xref64[(size_t)v * (N + 1) * 2 + 1] = py; }
    double vel = fmax(v0, .01), vr = 0.0;
    // segment index = bisect(s, dist) - 1 (pycubicspline.py:104).  dist starts at s[pid+1] and advances by a few
    // centimetres per step, so the index is tracked incrementally (same result as the bisection, a couple of loads
    // instead of ~10 dependent ones); a wrap past the end of the table restarts the scan with a bisection.
    int seg = min(pid + 1, n - 1);
    for (int idh = 1; idh <= N; ++idh) {
        dist = __dadd_rn(dist, __dmul_rn(__dmul_rn(scale, vel), Ts));
        dist = fmod(dist, s_last);
        if (dist < s[seg]) seg = seg_index(s, n, dist);
        else if (seg + 8 < n && s[seg + 8] <= dist) seg = seg_index(s, n, dist);          // large jump (huge speed)
        else while (seg + 1 < n && s[seg + 1] <= dist) ++seg;
        seg = min(max(seg, 0), n - 2);
        const double dx = dist - s[seg];
        const double* c = coef + (size_t)seg * stride;
        const double rx = cubic(c, dx), ry = cubic(c + 4, dx);
        if (xref32) { xref32[((size_t)v * (N + 1) + idh) * 2] = (float)rx; xref32[((size_t)v * (N + 1) + idh) * 2 + 1] = (float)ry; }
        if (xref64) { xref64[((size_t)v * (N + 1) + idh) * 2] = rx; xref64[((size_t)v * (N + 1) + idh) * 2 + 1] = ry; }
        if (i_lo < 0) vel = cubic(c + 8 + 4 * i_hi, dx);
        else {
            const double vb = cubic(c + 8 + 4 * i_lo, dx), va = cubic(c + 8 + 4 * i_hi, dx);
            // vb*(mus[i]-mu)/(mus[i]-mus[i-1]) + va*(mu-mus[i-1])/(mus[i]-mus[i-1])  evaluated left to right
            const double den = mus[i_hi] - mus[i_lo];
            vel = __dadd_rn(__dmul_rn(vb, mus[i_hi] - mu) / den, __dmul_rn(va, mu - mus[i_lo]) / den);
            (void)w_lo; (void)w_hi;
        }
        if (idh == 1) vr = __dmul_rn(vel, scale);
    }
    if (projidx_out) projidx_out[v] = pid;
    if (vr_out) vr_out[v] = vr;
}

}  // namespace llampc

using namespace llampc;

extern "C" int llampc_planner_constant_speed_f64(const double* s, const double* xy, const double* coef, const double* mus,
                                                 int n, int n_mu, const double* states, int V, const int* projidx_in,
                                                 const double* curr_mu, int mu_shared, int N, double Ts, double scale,
                                                 float* xref32, double* xref64, int* projidx_out, double* vr_out,
                                                 llampc_stream_t stream) {
    if (!s || !xy || !coef || !mus || !states || !projidx_in || !curr_mu || n < 3 || n_mu < 1 || V <= 0 || N <= 0)
        return LLAMPC_E_ARG;
    if (N > LLAMPC_MAX_H) return LLAMPC_E_RANGE;
    planner_kernel<<<(V + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        s, xy, coef, mus, n, n_mu, states, V, projidx_in, curr_mu, mu_shared, N, Ts, scale, xref32, xref64, projidx_out, vr_out);
    return (int)cudaGetLastError();
}
