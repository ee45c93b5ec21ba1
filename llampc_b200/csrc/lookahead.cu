// K2 look-ahead rollout (M models x K control sequences x H RK4 steps + NMPC cost) and K3 plant step.
//
// Integrator: Model._integrate_batch chained H times (llampc/models/model.py:32-40, llampc/utils/rk6.py:50-68);
// cost: llampc/mpc/nmpc.py:48,66-71,111 with the weights of run_nmpc_orca_llampc_rt.py:60-62;
// plant: Model._integrate -> odeintRK6 (model.py:18-30, rk6.py:13-28).
#include "llampc_common.cuh"
#include "llampc_model.cuh"
#include "llampc_model_f64.cuh"
#include "llampc_packed.cuh"
#include "llampc_launch.cuh"

namespace llampc {

#ifndef LLAMPC_LA_MIN_BLOCKS
#define LLAMPC_LA_MIN_BLOCKS 4
#endif
constexpr int LA_THREADS = 128;
constexpr int LA_WARPS = LA_THREADS / 32;

// One warp per model, lanes over control sequences (K = 32 -> one sequence per lane, the per-model arg-min
// is a warp shuffle).  Shared control table [K][H][2] and the reference path [H+1][2] are staged once per
// CTA with TMA bulk copies.  Positions are integrated relative to the start position so the fp32 state
// never carries the O(1) track coordinate.
__global__ void __launch_bounds__(LA_THREADS, LLAMPC_LA_MIN_BLOCKS)
lookahead_kernel(const float4* __restrict__ bank, int Mpad, const int* __restrict__ model_idx, int M,
                 const double* __restrict__ x0, int n_x0, const float* __restrict__ U, int K, int H,
                 const float* __restrict__ xref, const float* __restrict__ uprev, int per_model,
                 float q0, float q1, float r0, float r1, float p0, float p1, float h,
                 float* __restrict__ J, int* __restrict__ best_k, double* __restrict__ x_final,
                 double* __restrict__ x_traj) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    const bool u_pm = per_model & 1, xref_pm = per_model & 2, uprev_pm = per_model & 4, force_general = per_model & 8;
    const unsigned u_bytes = u_pm ? 0u : (unsigned)(K * H * 8);
    const unsigned u_bytes_pad = (u_bytes + 15u) & ~15u;
    const unsigned xr_bytes = xref_pm ? 0u : (unsigned)(((H + 1) * 8 + 15) & ~15);
    float2* sU = reinterpret_cast<float2*>(smem_raw);
    float2* sXr = reinterpret_cast<float2*>(smem_raw + u_bytes_pad);
    // per-warp reference path relative to the warp's start position (formed in fp64 once, used as fp32)
    float2* sRel = reinterpret_cast<float2*>(smem_raw + u_bytes_pad + xr_bytes) + (threadIdx.x >> 5) * (H + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&mbar, u_bytes_pad + xr_bytes);
        if (u_bytes_pad) tma_bulk_g2s(sU, U, u_bytes_pad, &mbar);       // caller pads the table to 16 B
        if (xr_bytes) tma_bulk_g2s(sXr, xref, xr_bytes, &mbar);
    }
    const int m = blockIdx.x * LA_WARPS + warp;
    const bool m_ok = m < M;
    const int mm = m_ok ? m : M - 1;
    const int bi = model_idx ? model_idx[mm] : mm;
    const Cand p = load_cand(bank, Mpad, bi);
    const double* xs = x0 + (n_x0 > 1 ? (size_t)mm * 6 : 0);
    double s0d, c0d;
    sincos(xs[2], &s0d, &c0d);
    const float s_start = (float)s0d, c_start = (float)c0d;
    const float2 up = uprev_pm ? reinterpret_cast<const float2*>(uprev)[mm] : reinterpret_cast<const float2*>(uprev)[0];
    const float2* Ug = reinterpret_cast<const float2*>(U) + (u_pm ? (size_t)mm * K * H : 0);
    const float2* Xg = reinterpret_cast<const float2*>(xref) + (xref_pm ? (size_t)mm * (H + 1) : 0);
    mbar_wait(&mbar, 0);
    if (!m_ok) return;
    {
        const double X0 = xs[0], Y0 = xs[1];
        for (int j = lane; j <= H; j += 32) {
            const float2 xr = xref_pm ? __ldg(Xg + j) : sXr[j];
            sRel[j] = make_float2((float)((double)xr.x - X0), (float)((double)xr.y - Y0));
        }
        __syncwarp();
    }

    u64 best = ~0ull;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const bool k_ok = k < K;
        const int kk = k_ok ? k : K - 1;
        float X = 0.0f, Y = 0.0f, dpsi = 0.0f;
        float s = s_start, c = c_start;
        float vx = (float)xs[3], vy = (float)xs[4], w = (float)xs[5];
        float2 uq = up;
        float Jt = 0.0f, Ja = 0.0f, ex = 0.0f, ey = 0.0f;
        double* tr = (x_traj && k_ok) ? x_traj + ((size_t)m * K + k) * (size_t)(H + 1) * 6 : nullptr;
        if (tr) {
#pragma unroll
            for (int i = 0; i < 6; ++i) tr[i] = xs[i];
        }
        for (int hh = 0; hh < H; ++hh) {
            const float2 u = u_pm ? __ldg(Ug + (size_t)kk * H + hh) : sU[kk * H + hh];
            const float du0 = u.x - uq.x, du1 = u.y - uq.y;       // nmpc.py:66-69 (du_0 = u_0 - uprev)
            Ja = fmaf(r0 * du0, du0, fmaf(r1 * du1, du1, Ja));
            uq = u;
            Ctl ctl;
            ctl.pwm = u.x; ctl.delta = u.y;
            sincos_half(u.y, ctl.sd, ctl.cd);
            float inc[6];
            float guard = 2.0f * fabsf(u.y);
            rk4_increment_fast<true>(p, ctl, s, c, vx, vy, w, h, inc, guard);   // SFU tyre sine: far below the cost tolerance
            guard = fmaxf(guard, 2.0f * fabsf(inc[2]));
            if (force_general) guard = 2.0f;
            if (!(guard <= 1.0f) || !(inc[3] + inc[5] == inc[3] + inc[5])) {   // rare: spinning, huge steering / yaw rate, NaN
                guard = 2.0f;
                const Inc6 g = rk4_increment_general(bank, Mpad, bi, u.x, u.y, s, c, vx, vy, w, h);
#pragma unroll
                for (int i = 0; i < 6; ++i) inc[i] = g.v[i];
            }
            X += inc[0]; Y += inc[1]; dpsi += inc[2];
            vx += inc[3]; vy += inc[4]; w += inc[5];
            float sr, cr;
            if (guard <= 1.0f) sincos_half(inc[2], sr, cr);
            else sincos_small(inc[2], sr, cr);
            const float sn = fmaf(s, cr, c * sr), cn = fmaf(c, cr, -s * sr);
            s = sn; c = cn;
            if (tr) {                                             // state after step hh (trajectory output, small K only)
                double* o = tr + (size_t)(hh + 1) * 6;
                o[0] = xs[0] + (double)X; o[1] = xs[1] + (double)Y; o[2] = xs[2] + (double)dpsi;
                o[3] = (double)vx; o[4] = (double)vy; o[5] = (double)w;
            }
            const float2 xr = sRel[hh + 1];
            ex = X - xr.x;
            ey = Y - xr.y;
            Jt = fmaf(q0 * ex, ex, fmaf(q1 * ey, ey, Jt));        // nmpc.py:70-71
        }
        Jt = fmaf(p0 * ex, ex, fmaf(p1 * ey, ey, Jt));            // terminal cost nmpc.py:48
        const float Jk = Jt + Ja;
        if (k_ok) {
            J[(size_t)m * K + k] = Jk;
            best = u64_min(best, pack_key(Jk, (unsigned)k));
            if (x_final) {
                double* o = x_final + ((size_t)m * K + k) * 6;
                o[0] = xs[0] + (double)X; o[1] = xs[1] + (double)Y; o[2] = xs[2] + (double)dpsi;
                o[3] = (double)vx; o[4] = (double)vy; o[5] = (double)w;
            }
        }
    }
    best = warp_min_u64(best);
    if (lane == 0) best_k[m] = (int)(best & 0xffffffffull);
}

// ---------------------------------------------------------------------------------------------------
// K2p.  The rollout with TWO MODELS per thread in packed f32x2 arithmetic (llampc_packed.cuh), for the layout of config
// C3: every model rolls the SAME K control sequences from the SAME start state along the same reference path.  One warp =
// one model pair, lanes over control sequences; the model parameters are the packed operands (as in K1p), everything that
// depends only on the control sequence is a scalar register that ptxas folds into the packed instruction as a broadcast.
// What does not depend on the model is computed ONCE PER CTA into shared memory after the TMA staging of the tables:
//   sCtl [K][H]  (pwm, delta, sin delta, cos delta) of every sampled input,
//   sJa  [K]     the control-effort cost sum_h du_h' R du_h of every sequence (nmpc.py:66-69),
//   sRel [H+1]   the reference path relative to the start position (formed in fp64).
// The step is straight-line code with no guard and no fallback (llampc_packed.cuh: rk4_step2): slip angles by the
// branch-free full-range atan2 (candidate models spin inside the horizon: 12 % of the warp-steps of C3 see a slip tangent
// above 1, so a guarded small-angle form would send every eighth warp-step through a fallback), headings carried
// relative to the start heading with sin / cos from the SFU, valid for any heading change.
// ---------------------------------------------------------------------------------------------------
#ifndef LLAMPC_LA2_MIN_BLOCKS
#define LLAMPC_LA2_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(LA_THREADS, LLAMPC_LA2_MIN_BLOCKS)
lookahead2_kernel(const float4* __restrict__ bank, int Mpad, int M, const double* __restrict__ x0,
                  const float* __restrict__ U, int K, int H, const float* __restrict__ xref,
                  const float* __restrict__ uprev, float q0, float q1, float r0, float r1, float p0, float p1, float h,
                  float* __restrict__ J, int* __restrict__ best_k, double* __restrict__ x_final) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ float s_x0[8];                      // sin psi0, cos psi0, vx0, vy0, w0
    const unsigned u_bytes_pad = ((unsigned)(K * H * 8) + 15u) & ~15u;
    const unsigned xr_bytes = (unsigned)(((H + 1) * 8 + 15) & ~15);
    float2* sU = reinterpret_cast<float2*>(smem_raw);
    float2* sXr = reinterpret_cast<float2*>(smem_raw + u_bytes_pad);
    float2* sRel = reinterpret_cast<float2*>(smem_raw + u_bytes_pad + xr_bytes);
    float* sJa = reinterpret_cast<float*>(sRel + (H + 1) + ((H + 1) & 1));
    float4* sCtl = reinterpret_cast<float4*>(smem_raw + ((u_bytes_pad + xr_bytes + (unsigned)(H + 2) * 8 + (unsigned)K * 4 + 15u) & ~15u));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&mbar, u_bytes_pad + xr_bytes);
        tma_bulk_g2s(sU, U, u_bytes_pad, &mbar);                   // caller pads the tables to 16 B
        tma_bulk_g2s(sXr, xref, xr_bytes, &mbar);
        double s0d, c0d;
        sincos(x0[2], &s0d, &c0d);
        s_x0[0] = (float)s0d; s_x0[1] = (float)c0d; s_x0[2] = (float)x0[3]; s_x0[3] = (float)x0[4]; s_x0[4] = (float)x0[5];
    }
    const int m0 = (blockIdx.x * LA_WARPS + warp) * 2, m1 = m0 + 1;
    const bool ok0 = m0 < M, ok1 = m1 < M;
    const int i0 = ok0 ? m0 : M - 1, i1 = ok1 ? m1 : M - 1;
    const Cand2 p = load_cand2(bank, Mpad, i0, i1);               // overlaps the bulk copies
    mbar_wait(&mbar, 0);
    // ---- model-invariant tables, once per CTA
    const float2 up0 = reinterpret_cast<const float2*>(uprev)[0];
    for (int i = tid; i < K * H; i += LA_THREADS) {
        const float2 u = sU[i];
        float sd, cd;
        sincosf(u.y, &sd, &cd);                                    // any steering angle (once per CTA)
        sCtl[i] = make_float4(u.x, u.y, sd, cd);
    }
    for (int k = tid; k < K; k += LA_THREADS) {
        float2 uq = up0;
        float ja = 0.0f;
        for (int hh = 0; hh < H; ++hh) {
            const float2 u = sU[k * H + hh];
            const float du0 = u.x - uq.x, du1 = u.y - uq.y;        // nmpc.py:66-69 (du_0 = u_0 - uprev)
            ja = fmaf(r0 * du0, du0, fmaf(r1 * du1, du1, ja));
            uq = u;
        }
        sJa[k] = ja;
    }
    {
        const double X0 = x0[0], Y0 = x0[1];
        for (int j = tid; j <= H; j += LA_THREADS) {
            const float2 xr = sXr[j];
            sRel[j] = make_float2((float)((double)xr.x - X0), (float)((double)xr.y - Y0));
        }
    }
    __syncthreads();
    if (!ok0) return;                                              // whole warp

    Head2 h0;
    h0.s0 = bc(s_x0[0]); h0.c0 = bc(s_x0[1]); h0.ns0 = bc(-s_x0[0]);
    u64 best0 = ~0ull, best1 = ~0ull;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const bool k_ok = k < K;
        const int kk = k_ok ? k : K - 1;
        State2 x;
        x.X = bc(0.0f); x.Y = bc(0.0f); x.phi = bc(0.0f);
        x.vx = bc(s_x0[2]); x.vy = bc(s_x0[3]); x.w = bc(s_x0[4]);
        F2 Jt = bc(0.0f), ex = bc(0.0f), ey = bc(0.0f);
        const float4* ctlk = sCtl + kk * H;
#pragma unroll 1
        for (int hh = 0; hh < H; ++hh) {
            const float4 cq = ctlk[hh];
            rk4_step2(p, ctl2_shared(cq.x, cq.y, cq.z, cq.w), h0, h, x);      // straight-line code: no guard, no fallback
            const float2 xr = sRel[hh + 1];
            ex = add2(x.X, bc(-xr.x));
            ey = add2(x.Y, bc(-xr.y));
            Jt = fma2(mul2(bc(q0), ex), ex, fma2(mul2(bc(q1), ey), ey, Jt));      // nmpc.py:70-71
        }
        Jt = fma2(mul2(bc(p0), ex), ex, fma2(mul2(bc(p1), ey), ey, Jt));          // terminal cost nmpc.py:48
        const F2 Jk = add2(Jt, bc(sJa[kk]));
        float j0, j1;
        up(Jk, j0, j1);
        // programmatic dependent launch (per_model_flags & 32; no-ops otherwise): after the last rollout of this warp the
        // next launch on the stream may start its rollouts; nothing is written before the previous launch has completed
        if (k0 + 32 >= K) pdl_launch_dependents();
        pdl_wait();
        if (k_ok) {
            J[(size_t)m0 * K + k] = j0;
            best0 = u64_min(best0, pack_key(j0, (unsigned)k));
            if (ok1) {
                J[(size_t)m1 * K + k] = j1;
                best1 = u64_min(best1, pack_key(j1, (unsigned)k));
            }
            if (x_final) {
                float a[6], b[6];
                up(x.X, a[0], b[0]); up(x.Y, a[1], b[1]); up(x.phi, a[2], b[2]); up(x.vx, a[3], b[3]); up(x.vy, a[4], b[4]);
                up(x.w, a[5], b[5]);
                double* o = x_final + ((size_t)m0 * K + k) * 6;
                o[0] = x0[0] + (double)a[0]; o[1] = x0[1] + (double)a[1]; o[2] = x0[2] + (double)a[2];
                o[3] = (double)a[3]; o[4] = (double)a[4]; o[5] = (double)a[5];
                if (ok1) {
                    o = x_final + ((size_t)m1 * K + k) * 6;
                    o[0] = x0[0] + (double)b[0]; o[1] = x0[1] + (double)b[1]; o[2] = x0[2] + (double)b[2];
                    o[3] = (double)b[3]; o[4] = (double)b[4]; o[5] = (double)b[5];
                }
            }
        }
    }
    best0 = warp_min_u64(best0);
    best1 = warp_min_u64(best1);
    if (lane == 0) {
        best_k[m0] = (int)(best0 & 0xffffffffull);
        if (ok1) best_k[m1] = (int)(best1 & 0xffffffffull);
    }
}

// ---------------------------------------------------------------------------------------------------
// K2q.  The packed rollout for the per-model layouts (Monte-Carlo closed loop: one model, start state, control table,
// reference path and previous input PER VEHICLE): two models per thread as in K2p, but the control input and the start
// heading now differ between the two components, so they are packed operands too (Ctl2 / Head2) and nothing is hoisted
// per CTA.  Any of x0 / U / xref / uprev may still be shared (stride 0).  The inputs of the next step are loaded while
// the current one is integrated; sin / cos of the steering angle come from the SFU.
// ---------------------------------------------------------------------------------------------------
// CTA = 2 warps = 4 models, registers capped for 7 CTAs per SM (14 warps): the 4,096-vehicle closed loop is 2,048
// pair-warps = 13.8 per SM, so every warp is resident in ONE wave (at 3 CTAs of 4 warps the last 1.8 warps per SM ran a
// second, latency-bound round: 70 us against 58 us for the scalar kernel).
#ifndef LLAMPC_LQ_MIN_BLOCKS
#define LLAMPC_LQ_MIN_BLOCKS 7
#endif
constexpr int LQ_THREADS = 64;
constexpr int LQ_WARPS = LQ_THREADS / 32;
__global__ void __launch_bounds__(LQ_THREADS, LLAMPC_LQ_MIN_BLOCKS)
lookahead2_permodel_kernel(const float4* __restrict__ bank, int Mpad, const int* __restrict__ model_idx, int M,
                           const double* __restrict__ x0, int n_x0, const float* __restrict__ U, int K, int H,
                           const float* __restrict__ xref, const float* __restrict__ uprev, int per_model,
                           float q0, float q1, float r0, float r1, float p0, float p1, float h,
                           float* __restrict__ J, int* __restrict__ best_k, double* __restrict__ x_final) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const bool u_pm = per_model & 1, xref_pm = per_model & 2, uprev_pm = per_model & 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* sRel = reinterpret_cast<float2*>(smem_raw) + (size_t)warp * 2 * (H + 1);      // [2][H+1] per warp
    const int m0 = (blockIdx.x * LQ_WARPS + warp) * 2, m1 = m0 + 1;
    if (m0 >= M) return;                                           // whole warp
    const bool ok1 = m1 < M;
    const int mm1 = ok1 ? m1 : m0;
    const int i0 = model_idx ? model_idx[m0] : m0, i1 = model_idx ? model_idx[mm1] : mm1;
    const Cand2 p = load_cand2(bank, Mpad, i0, i1);
    const double* xa = x0 + (n_x0 > 1 ? (size_t)m0 * 6 : 0);
    const double* xb = x0 + (n_x0 > 1 ? (size_t)mm1 * 6 : 0);
    double sa, ca, sb, cb;
    sincos(xa[2], &sa, &ca);
    sincos(xb[2], &sb, &cb);
    Head2 h0;
    h0.s0 = pk((float)sa, (float)sb); h0.c0 = pk((float)ca, (float)cb); h0.ns0 = pk(-(float)sa, -(float)sb);
    const float2* Ua = reinterpret_cast<const float2*>(U) + (u_pm ? (size_t)m0 * K * H : 0);
    const float2* Ub = reinterpret_cast<const float2*>(U) + (u_pm ? (size_t)mm1 * K * H : 0);
    const float2* Xa = reinterpret_cast<const float2*>(xref) + (xref_pm ? (size_t)m0 * (H + 1) : 0);
    const float2* Xb = reinterpret_cast<const float2*>(xref) + (xref_pm ? (size_t)mm1 * (H + 1) : 0);
    const float2 upa = reinterpret_cast<const float2*>(uprev)[uprev_pm ? m0 : 0];
    const float2 upb = reinterpret_cast<const float2*>(uprev)[uprev_pm ? mm1 : 0];
    for (int j = lane; j <= H; j += 32) {                          // reference paths relative to the start positions (fp64)
        const float2 ra = __ldg(Xa + j), rb = __ldg(Xb + j);
        sRel[j] = make_float2((float)((double)ra.x - xa[0]), (float)((double)ra.y - xa[1]));
        sRel[(H + 1) + j] = make_float2((float)((double)rb.x - xb[0]), (float)((double)rb.y - xb[1]));
    }
    __syncwarp();
    u64 best0 = ~0ull, best1 = ~0ull;
    for (int k0 = 0; k0 < K; k0 += 32) {
        const int k = k0 + lane;
        const bool k_ok = k < K;
        const int kk = k_ok ? k : K - 1;
        State2 x;
        x.X = bc(0.0f); x.Y = bc(0.0f); x.phi = bc(0.0f);
        x.vx = pk((float)xa[3], (float)xb[3]); x.vy = pk((float)xa[4], (float)xb[4]); x.w = pk((float)xa[5], (float)xb[5]);
        F2 Jt = bc(0.0f), Ja = bc(0.0f), ex = bc(0.0f), ey = bc(0.0f);
        F2 uq0 = pk(upa.x, upb.x), uq1 = pk(upa.y, upb.y);        // previous input (pwm pair, steer pair)
        const float2* ua = Ua + (size_t)kk * H;
        const float2* ub = Ub + (size_t)kk * H;
        float2 na = __ldg(ua), nb = __ldg(ub);
#pragma unroll 1
        for (int hh = 0; hh < H; ++hh) {
            const float2 a = na, b = nb;
            if (hh + 1 < H) { na = __ldg(ua + hh + 1); nb = __ldg(ub + hh + 1); }          // next step's inputs in flight
            Ctl2 u;
            u.pwm = pk(a.x, b.x); u.npwm = pk(-a.x, -b.x); u.delta = pk(a.y, b.y);
            u.nsd = pk(-__sinf(a.y), -__sinf(b.y)); u.cd = pk(__cosf(a.y), __cosf(b.y));
            const F2 du0 = sub2(u.pwm, uq0), du1 = sub2(u.delta, uq1);                     // nmpc.py:66-69 (du_0 = u_0 - uprev)
            Ja = fma2(mul2(bc(r0), du0), du0, fma2(mul2(bc(r1), du1), du1, Ja));
            uq0 = u.pwm; uq1 = u.delta;
            rk4_step2(p, u, h0, h, x);
            const float2 ra = sRel[hh + 1], rb = sRel[(H + 1) + hh + 1];
            ex = sub2(x.X, pk(ra.x, rb.x));
            ey = sub2(x.Y, pk(ra.y, rb.y));
            Jt = fma2(mul2(bc(q0), ex), ex, fma2(mul2(bc(q1), ey), ey, Jt));               // nmpc.py:70-71
        }
        Jt = fma2(mul2(bc(p0), ex), ex, fma2(mul2(bc(p1), ey), ey, Jt));                   // terminal cost nmpc.py:48
        float j0, j1;
        up(add2(Jt, Ja), j0, j1);
        if (k_ok) {
            J[(size_t)m0 * K + k] = j0;
            best0 = u64_min(best0, pack_key(j0, (unsigned)k));
            if (ok1) {
                J[(size_t)m1 * K + k] = j1;
                best1 = u64_min(best1, pack_key(j1, (unsigned)k));
            }
            if (x_final) {
                float a[6], b[6];
                up(x.X, a[0], b[0]); up(x.Y, a[1], b[1]); up(x.phi, a[2], b[2]); up(x.vx, a[3], b[3]); up(x.vy, a[4], b[4]);
                up(x.w, a[5], b[5]);
                double* o = x_final + ((size_t)m0 * K + k) * 6;
                o[0] = xa[0] + (double)a[0]; o[1] = xa[1] + (double)a[1]; o[2] = xa[2] + (double)a[2];
                o[3] = (double)a[3]; o[4] = (double)a[4]; o[5] = (double)a[5];
                if (ok1) {
                    o = x_final + ((size_t)m1 * K + k) * 6;
                    o[0] = xb[0] + (double)b[0]; o[1] = xb[1] + (double)b[1]; o[2] = xb[2] + (double)b[2];
                    o[3] = (double)b[3]; o[4] = (double)b[4]; o[5] = (double)b[5];
                }
            }
        }
    }
    best0 = warp_min_u64(best0);
    best1 = warp_min_u64(best1);
    if (lane == 0) {
        best_k[m0] = (int)(best0 & 0xffffffffull);
        if (ok1) best_k[m1] = (int)(best1 & 0xffffffffull);
    }
}

__global__ void __launch_bounds__(128)
plant_rk6_kernel(const double* __restrict__ params, int V, const double* __restrict__ x, const double* __restrict__ u,
                 double h, double* __restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    Params64 p;
    double* pp = reinterpret_cast<double*>(&p);
#pragma unroll
    for (int j = 0; j < LLAMPC_NPARAM; ++j) pp[j] = params[(size_t)v * LLAMPC_NPARAM + j];
    double y0[6], y1[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y0[i] = x[(size_t)v * 6 + i];
    rk6_step64(p, y0, u[(size_t)v * 2], u[(size_t)v * 2 + 1], h, y1);
#pragma unroll
    for (int i = 0; i < 6; ++i) out[(size_t)v * 6 + i] = y1[i];
}

}  // namespace llampc

using namespace llampc;

extern "C" int llampc_lookahead_rollout_f32(const float* bank, int Mpad, const int* model_idx, int M,
                                            const double* x0, int n_x0, const float* U, int K, int H,
                                            const float* xref, const float* uprev, int per_model_flags,
                                            const float* qrp_h, double Ts, float* J, int* best_k, double* x_final,
                                            double* x_traj, llampc_stream_t stream) {
    if (!bank || !x0 || !U || !xref || !uprev || !qrp_h || !J || !best_k || M <= 0 || K <= 0 || H <= 0) return LLAMPC_E_ARG;
    if (n_x0 != 1 && n_x0 != M) return LLAMPC_E_ARG;
    if (H > LLAMPC_MAX_H) return LLAMPC_E_RANGE;
    if ((reinterpret_cast<uintptr_t>(bank) | reinterpret_cast<uintptr_t>(U) | reinterpret_cast<uintptr_t>(xref)) & 15u)
        return LLAMPC_E_ALIGN;
    const bool u_pm = per_model_flags & 1, xref_pm = per_model_flags & 2;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // shared start state, control table, reference path and previous input (config C3): two models per thread (K2p).
    // per_model_flags & 8 (diagnostics: force the general step) keeps the scalar kernel for this layout too.
    if ((per_model_flags & 15) == 0 && !(per_model_flags & 16) && n_x0 == 1 && !model_idx && !x_traj && M >= 2) {
        const size_t tab = (((size_t)K * H * 8 + 15) & ~(size_t)15) + ((((size_t)H + 1) * 8 + 15) & ~(size_t)15);
        const size_t smem2 = ((tab + ((size_t)H + 2) * 8 + (size_t)K * 4 + 15) & ~(size_t)15) + (size_t)K * H * 16;
        if (smem2 <= 160 * 1024) {
            LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookahead2_kernel, smem2));
            const dim3 grid2((M + 2 * LA_WARPS - 1) / (2 * LA_WARPS));
            if (per_model_flags & 32)              // programmatic dependent launch: back-to-back independent rollouts
                return issue_pdl(lookahead2_kernel, grid2, dim3(LA_THREADS), smem2, st, reinterpret_cast<const float4*>(bank),
                                 Mpad, M, x0, U, K, H, xref, uprev, qrp_h[0], qrp_h[1], qrp_h[2], qrp_h[3], qrp_h[4], qrp_h[5],
                                 (float)Ts, J, best_k, x_final);
            lookahead2_kernel<<<grid2, LA_THREADS, smem2, st>>>(
                reinterpret_cast<const float4*>(bank), Mpad, M, x0, U, K, H, xref, uprev, qrp_h[0], qrp_h[1], qrp_h[2],
                qrp_h[3], qrp_h[4], qrp_h[5], (float)Ts, J, best_k, x_final);
            return (int)cudaGetLastError();
        }
    }
    // any per-model layout (Monte-Carlo closed loop: model_idx, per-vehicle x0 / U / xref / uprev): the packed kernel K2q
    if (!(per_model_flags & (8 | 16)) && !x_traj) {
        const size_t smemq = (size_t)LQ_WARPS * 2 * (H + 1) * 8;
        LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookahead2_permodel_kernel, smemq));
        lookahead2_permodel_kernel<<<(M + 2 * LQ_WARPS - 1) / (2 * LQ_WARPS), LQ_THREADS, smemq, st>>>(
            reinterpret_cast<const float4*>(bank), Mpad, model_idx, M, x0, n_x0, U, K, H, xref, uprev, per_model_flags & 7,
            qrp_h[0], qrp_h[1], qrp_h[2], qrp_h[3], qrp_h[4], qrp_h[5], (float)Ts, J, best_k, x_final);
        return (int)cudaGetLastError();
    }
    // the scalar kernel K2 (one model per warp): trajectory output and the diagnostic general-step mode
    size_t smem = (u_pm ? 0 : (((size_t)K * H * 8 + 15) & ~(size_t)15)) + (xref_pm ? 0 : ((((size_t)H + 1) * 8 + 15) & ~(size_t)15)) +
                  (size_t)LA_WARPS * (H + 1) * 8;
    if (smem > 96 * 1024) return LLAMPC_E_RANGE;
    LLAMPC_CUDA_TRY((cudaError_t)raise_dynamic_smem(lookahead_kernel, smem));
    lookahead_kernel<<<(M + LA_WARPS - 1) / LA_WARPS, LA_THREADS, smem, st>>>(
        reinterpret_cast<const float4*>(bank), Mpad, model_idx, M, x0, n_x0, U, K, H, xref, uprev, per_model_flags & 15,
        qrp_h[0], qrp_h[1], qrp_h[2], qrp_h[3], qrp_h[4], qrp_h[5], (float)Ts, J, best_k, x_final, x_traj);
    return (int)cudaGetLastError();
}

extern "C" int llampc_plant_rk6_f64(const double* params64, int V, const double* x64, const double* u64v, double Ts,
                                    double* out64, llampc_stream_t stream) {
    if (!params64 || !x64 || !u64v || !out64 || V <= 0) return LLAMPC_E_ARG;
    plant_rk6_kernel<<<(V + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(params64, V, x64, u64v, Ts, out64);
    return (int)cudaGetLastError();
}
