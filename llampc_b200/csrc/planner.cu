// Device-side look-ahead reference generator: the reference's ConstantSpeed planner (llampc/mpc/planner.py:12-67)
// for V vehicles at once, fp64, eight lanes per vehicle, table windows staged in shared memory by TMA.  Pieces restated: Track.project_fast
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

// ---------------------------------------------------------------------------------------------------
// planner_kernel: SIXTEEN LANES PER VEHICLE, the vehicle's table window staged in shared memory by TMA.
//   * the nine segment projections of the 10-point window run on nine lanes at once, every lane then replays the
//     reference's sequential arg-min scan over the nine distances (same first-minimum / NaN behaviour);
//   * the projection fixes the first segment of the arc-length march, so the window of the tables the march can reach
//     -- PL_WSEG segments: arc lengths, the x(s) / y(s) cubics, and the TWO speed cubics that the friction level selects --
//     is copied global -> shared by three bulk copies per vehicle (cp.async.bulk + mbarrier, SASS UBLKCP): one L2 round
//     trip instead of two dependent L2 loads in each of the N steps.  A march that leaves the window re-stages it at the
//     new position (one more round trip); a step that wraps past the end of the table reads global memory;
//   * in the march lanes 0..3 evaluate the four cubics x, y, v_lo, v_hi of a step concurrently; the speed blend and the
//     arc-length update are computed redundantly on every lane from the shuffled values (identical fp64 operations, so
//     no divergence between lanes and bit-identical to one thread doing it all).
// Same fp64 operations in the same order as the one-thread-per-vehicle kernel it replaces (golden tests bit-equal).
// Tables (see include/llampc_b200.h): s is readable LLAMPC_PLAN_SPAD entries past any even index <= n - 1, coef_xy /
// coef_vp PL_WSEG rows past any row <= n - 2 (the host pads them), so every window copy has one fixed size.
// ---------------------------------------------------------------------------------------------------
constexpr int PL_LANES = 16;                       // the nine segment projections in ONE round
constexpr int PL_VEH = 8;                          // vehicles per CTA
constexpr int PL_THREADS = PL_LANES * PL_VEH;      // 128
constexpr int PL_WSEG = LLAMPC_PLAN_WSEG;          // 48 segments: 0.9 x 2.5 m/s x 0.02 s x 20 steps / 0.023 m per segment = 39

constexpr int PL_SWIN = PL_WSEG + PL_LANES + 2;    // arc-length marks staged: the probes of the last staged segment stay inside
struct __align__(16) PlanWin {
    double s[PL_SWIN];                             // arc length at points base .. base + PL_SWIN - 1
    double cxy[PL_WSEG][8];                        // a b c d of x(s), then of y(s), per segment
    double cv[PL_WSEG][8];                         // a b c d of v_lo(s), then of v_hi(s), per segment
};

__global__ void __launch_bounds__(PL_THREADS)
planner_kernel(const double* __restrict__ s, const double* __restrict__ xy, const double* __restrict__ coef_xy,
               const double* __restrict__ coef_vp, const double* __restrict__ mus, int n, int n_mu,
               const double* __restrict__ states, int V, const int* __restrict__ projidx_in,
               const double* __restrict__ curr_mu, int mu_shared, int N, double Ts, double scale,
               float* __restrict__ xref32, double* __restrict__ xref64, int* __restrict__ projidx_out,
               double* __restrict__ vr_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar[PL_VEH];
    const int tid = threadIdx.x, g = tid / PL_LANES, l = tid % PL_LANES;
    const int v_raw = blockIdx.x * PL_VEH + g;
    const bool live = v_raw < V;                                  // a padding group repeats the last vehicle and stores nothing:
    const int v = live ? v_raw : V - 1;                           // every warp stays whole, so full-warp syncs are safe
    if (l == 0) mbar_init(&mbar[g], 1);
    __syncthreads();
    const unsigned gmask = ((1u << PL_LANES) - 1u) << (((tid & 31) / PL_LANES) * PL_LANES);
    const int lane0 = ((tid & 31) / PL_LANES) * PL_LANES;        // first lane of this vehicle's group inside the warp
    PlanWin* win = reinterpret_cast<PlanWin*>(smem_raw) + g;
    int pid = projidx_in[v];
    // ---- stage the window the march can reach BEFORE the projection (the copies fly while the nine segments are
    // projected): the projection moves the index forward by at most 9, so the window starts at the first segment the march
    // can start from, base = (pid + 1) rounded down to even (16-byte aligned rows of s)
    int base = min(max(pid + 1, 0), n - 2) & ~1;
    // coef_vp[j] = the (v_{j-1 mod n_mu}, v_j) cubics of every segment, contiguous: pair j is picked from the friction level
    // (planner.py:49-62).  The levels sit in registers (lane l holds mus[l] and mus[16 + l]); the search is one ballot.
    const double mu = curr_mu[mu_shared ? 0 : v];
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    const double m0 = l < n_mu ? mus[l] : INF, m1 = PL_LANES + l < n_mu ? mus[PL_LANES + l] : INF;
    const double px = states[(size_t)v * 6], py = states[(size_t)v * 6 + 1], v0 = states[(size_t)v * 6 + 3];
    int i_hi, i_lo;
    double mu_hi, mu_lo;
    if (n_mu <= 2 * PL_LANES) {
        const unsigned full = 0xffffffffu;
        const unsigned b0 = (__ballot_sync(full, m0 >= mu) >> lane0) & 0xffffu, b1 = (__ballot_sync(full, m1 >= mu) >> lane0) & 0xffffu;
        const unsigned ge = b0 | (b1 << PL_LANES);                // bit i: mus[i] >= mu (padding levels are +inf)
        const double first = __shfl_sync(full, m0, lane0);
        const int il = n_mu - 1;
        const double last = il < PL_LANES ? __shfl_sync(full, m0, lane0 + il) : __shfl_sync(full, m1, lane0 + il - PL_LANES);
        if (mu < first) { i_hi = 0; i_lo = -1; }
        else if (mu > last) { i_hi = n_mu - 1; i_lo = -1; }
        else {
            i_hi = ge ? __ffs((int)ge) - 1 : n_mu - 1;             // first level >= mu (a NaN mu falls through to the last)
            i_hi = min(i_hi, n_mu - 1);
            i_lo = (i_hi - 1 + n_mu) % n_mu;                      // Python's spline_v[i-1] wraps for i = 0
        }
        const int jl = i_lo < 0 ? i_hi : i_lo;
        const double h0 = __shfl_sync(full, m0, lane0 + (i_hi & (PL_LANES - 1))), h1 = __shfl_sync(full, m1, lane0 + (i_hi & (PL_LANES - 1)));
        const double l0 = __shfl_sync(full, m0, lane0 + (jl & (PL_LANES - 1))), l1 = __shfl_sync(full, m1, lane0 + (jl & (PL_LANES - 1)));
        mu_hi = i_hi < PL_LANES ? h0 : h1;
        mu_lo = jl < PL_LANES ? l0 : l1;
    } else {                                                       // more than 32 friction levels: the plain scan
        if (mu < mus[0]) { i_hi = 0; i_lo = -1; }
        else if (mu > mus[n_mu - 1]) { i_hi = n_mu - 1; i_lo = -1; }
        else {
            i_hi = 0;
            for (int i = 0; i < n_mu; ++i) { i_hi = i; if (mus[i] >= mu) break; }
            i_lo = (i_hi - 1 + n_mu) % n_mu;
        }
        mu_hi = mus[i_hi];
        mu_lo = mus[i_lo < 0 ? i_hi : i_lo];
    }
    const double* vp = coef_vp + (size_t)i_hi * (size_t)(n - 1 + PL_WSEG) * 8;
    if (l == 0) {
        mbar_expect_tx(&mbar[g], (unsigned)sizeof(PlanWin));
        tma_bulk_g2s(win->s, s + base, (unsigned)sizeof(win->s), &mbar[g]);
        tma_bulk_g2s(win->cxy, coef_xy + (size_t)base * 8, (unsigned)sizeof(win->cxy), &mbar[g]);
        tma_bulk_g2s(win->cv, vp + (size_t)base * 8, (unsigned)sizeof(win->cv), &mbar[g]);
    }
    const double s_last = s[n - 1];
    // ---- project onto raceline[:, pid:pid+10]  (planner.py:26, numpy slicing truncates at the end of the table)
    const int n_way = max(0, min(10, n - pid));
    double d_mine = 0.0;
    if (l + 1 < n_way) {                                          // segment l of the (at most) nine
        const double* a = xy + (size_t)(pid + l) * 2;
        d_mine = project_segment(px, py, a[0], a[1], a[2], a[3]);
    }
    __syncwarp();
    int best = 0;
    double best_d = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {                                 // the reference's sequential scan, on every lane
        const double d = __shfl_sync(0xffffffffu, d_mine, lane0 + i);
        if (i + 1 < n_way && (i == 0 || d < best_d)) { best_d = d; best = i; }
    }
    pid += best;
    int seg = min(pid + 1, n - 1);
    if (l < 2 && live) {
        const double p0 = l == 0 ? px : py;
        if (xref32) xref32[(size_t)v * (N + 1) * 2 + l] = (float)p0;
        if (xref64) xref64[(size_t)v * (N + 1) * 2 + l] = p0;
    }
    double vel = fmax(v0, .01), vr = 0.0;
    // the two vehicles of a warp wait on their own barriers: re-converge afterwards, or the halves of the warp would run
    // the whole march one after the other (measured: 2x the per-step latency)
    mbar_wait_parity(&mbar[g], 0);
    __syncwarp();
    // arc length of raceline[:, :pid+2]  (planner.py:30,36): s[pid + 1], which the window holds (base <= pid + 1 <= base + 11)
    double dist = seg - base < PL_SWIN ? win->s[seg - base] : s[seg];
    // segment index = bisect(s, dist) - 1 (pycubicspline.py:104).  dist starts at s[pid+1] and advances by a few
    // centimetres per step, so the index is tracked incrementally (same result as the bisection).  The loop body is
    // straight-line code for the common case -- no wrap, the segment inside the staged window, an advance of fewer than
    // sixteen segments -- so that the four vehicles sharing a warp do not serialise each other: the eight lanes probe the
    // next sixteen arc-length marks at once (one ballot instead of a dependent scan), ALL lanes evaluate a cubic (lanes
    // 4..15 repeat 0..3) and their share of the friction blend, and the results meet through shuffles.
    const int q4 = l & 3;                                          // 0: x, 1: y, 2: v_lo, 3: v_hi
    const double blend_w = (q4 & 1) ? (mu - mu_lo) : (mu_hi - mu); // lane 2: v_lo weight, lane 3: v_hi weight
    const double den = i_lo < 0 ? 1.0 : mu_hi - mu_lo;
    // x / den for the same den in every step: r = RN(1 / den), q0 = RN(x r), q = RN(q0 + RN(x - q0 den) r).  With the
    // exact remainder from the FMA this final correction yields the correctly rounded quotient (Markstein), i.e. the
    // very bits of the IEEE division the reference performs, in 3 dependent operations instead of a 440-cycle divide
    // per step (tools/ubench/dfma.cu).
    const double rden = 1.0 / den;
    // this lane's cubic inside the window: row k of cxy (lanes 0, 1) or cv (lanes 2, 3), first or second half of the row
    const double* lane_win = (q4 < 2 ? &win->cxy[0][0] : &win->cv[0][0]) + 4 * (q4 & 1);
    const double* lane_glb = (q4 < 2 ? coef_xy : vp) + 4 * (q4 & 1);
    float* o32 = (xref32 && l < 2 && live) ? xref32 + (size_t)v * (N + 1) * 2 + l : nullptr;
    double* o64 = (xref64 && l < 2 && live) ? xref64 + (size_t)v * (N + 1) * 2 + l : nullptr;
    const int sh = lane0;                                          // this vehicle's 16 bits of a full-warp ballot
    int k = seg - base;                                            // segment inside the window (0 or 1 at the start)
    unsigned phase = 0;                                            // parity of the last completed window copy
    bool restaged = false;
    for (int idh = 1; idh <= N; ++idh) {
        dist = __dadd_rn(dist, __dmul_rn(__dmul_rn(scale, vel), Ts));
        // the sixteen lanes probe the next sixteen arc-length marks at once (index clamped into the window so that the
        // ballot can be unconditional and warp-wide: no divergent collective)
        const int kc = min(max(k, 0), PL_WSEG - 1);
        const unsigned bits = (__ballot_sync(0xffffffffu, win->s[kc + 1 + l] <= dist) >> sh) & ((1u << PL_LANES) - 1u);
        const int adv = __ffs((int)~bits) - 1;                    // marks passed (the marks are increasing): 0 .. PL_LANES
        double c0, c1, c2, c3, s_seg;                              // the lane's cubic and the arc length at the segment start
        if (k == kc && k + adv < PL_WSEG && adv < PL_LANES && dist < s_last && dist >= win->s[k]) {
            k += adv;
            seg = base + k;
            const double2* cp = reinterpret_cast<const double2*>(lane_win + k * 8);
            const double2 lo = cp[0], hi = cp[1];
            c0 = lo.x; c1 = lo.y; c2 = hi.x; c3 = hi.y;
            s_seg = win->s[k];
        } else {                                                   // wrap, backward move, large jump or end of the window
            if (!(dist > -s_last && dist < s_last)) dist = fmod(dist, s_last);   // fmod is the identity inside (-s_last, s_last)
            if (dist < s[seg]) seg = seg_index(s, n, dist);
            else if (seg + 8 < n && s[seg + 8] <= dist) seg = seg_index(s, n, dist);      // large jump (huge speed)
            else while (seg + 1 < n && s[seg + 1] <= dist) ++seg;
            seg = min(max(seg, 0), n - 2);
            const double* cg = lane_glb + (size_t)seg * 8;         // this step reads global memory ...
            c0 = cg[0]; c1 = cg[1]; c2 = cg[2]; c3 = cg[3];
            s_seg = s[seg];
            // ... and the window is re-staged at the new position for the following steps (every lane of the vehicle
            // is here, so nobody still reads the old window; the copies land while this step's cubic is evaluated)
            base = seg & ~1;
            k = seg - base;
            __syncwarp(gmask);
            if (l == 0) {
                mbar_expect_tx(&mbar[g], (unsigned)sizeof(PlanWin));
                tma_bulk_g2s(win->s, s + base, (unsigned)sizeof(win->s), &mbar[g]);
                tma_bulk_g2s(win->cxy, coef_xy + (size_t)base * 8, (unsigned)sizeof(win->cxy), &mbar[g]);
                tma_bulk_g2s(win->cv, vp + (size_t)base * 8, (unsigned)sizeof(win->cv), &mbar[g]);
            }
            restaged = true;
        }
        const double dx = dist - s_seg;
        const double dx2 = __dmul_rn(dx, dx), dx3 = __dmul_rn(dx2, dx);
        const double q = __dadd_rn(__dadd_rn(__dadd_rn(c0, __dmul_rn(c1, dx)), __dmul_rn(c2, dx2)), __dmul_rn(c3, dx3));
        if (o32) o32[idh * 2] = (float)q;
        if (o64) o64[idh * 2] = q;
        // vb*(mus[i]-mu)/(mus[i]-mus[i-1]) + va*(mu-mus[i-1])/(mus[i]-mus[i-1])  evaluated left to right (planner.py:62):
        // each product / quotient on the lane that holds the cubic, the sum on every lane
        const double xw = __dmul_rn(q, blend_w);
        const double t0 = __dmul_rn(xw, rden);
        const double t = __fma_rn(__fma_rn(-t0, den, xw), rden, t0);
        if (restaged) {                                            // the new window must have landed before the next step
            phase ^= 1u;
            mbar_wait_parity(&mbar[g], phase);
            restaged = false;
        }
        __syncwarp();                                              // both vehicles of the warp are here: warp-wide shuffles
        const double tb = __shfl_sync(0xffffffffu, t, lane0 + 2), ta = __shfl_sync(0xffffffffu, t, lane0 + 3);
        const double va = __shfl_sync(0xffffffffu, q, lane0 + 3);
        vel = i_lo < 0 ? va : __dadd_rn(tb, ta);
        if (idh == 1) vr = __dmul_rn(vel, scale);
    }
    if (l == 0 && live) {
        if (projidx_out) projidx_out[v] = pid;
        if (vr_out) vr_out[v] = vr;
    }
}

}  // namespace llampc

using namespace llampc;

extern "C" int llampc_planner_constant_speed_f64(const double* s, const double* xy, const double* coef_xy,
                                                 const double* coef_vp, const double* mus, int n, int n_mu,
                                                 const double* states, int V, const int* projidx_in,
                                                 const double* curr_mu, int mu_shared, int N, double Ts, double scale,
                                                 float* xref32, double* xref64, int* projidx_out, double* vr_out,
                                                 llampc_stream_t stream) {
    if (!s || !xy || !coef_xy || !coef_vp || !mus || !states || !projidx_in || !curr_mu || n < 3 || n_mu < 1 || V <= 0 || N <= 0)
        return LLAMPC_E_ARG;
    if (N > LLAMPC_MAX_H) return LLAMPC_E_RANGE;
    if ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(coef_xy) | reinterpret_cast<uintptr_t>(coef_vp)) & 15u)
        return LLAMPC_E_ALIGN;
    const size_t smem = sizeof(PlanWin) * PL_VEH;
    LLAMPC_CUDA_TRY(cudaFuncSetAttribute(planner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    planner_kernel<<<(V + PL_VEH - 1) / PL_VEH, PL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        s, xy, coef_xy, coef_vp, mus, n, n_mu, states, V, projidx_in, curr_mu, mu_shared, N, Ts, scale, xref32, xref64,
        projidx_out, vr_out);
    return (int)cudaGetLastError();
}
