"""Monte-Carlo closed loop on the device: V independent vehicles / friction scenarios, each running the LLA-MPC
tick -- planner, look-ahead over sampled control sequences, plant step, look-back adaptation, friction estimate --
with every array resident in HBM and no host round trip inside a tick (BASELINE config 4).

Per tick ``idt`` and vehicle, in the reference's own order (loop body run_nmpc_orca_llampc_rt.py:269-389):
  1. xref   = ConstantSpeed(x[:2], vx, track, H, Ts, projidx[, curr_mu=MU_pred, scale=v_factor])   planner.py:12-67
              -- the friction estimate and the speed scale only from tick W + 2 on, the planner's defaults
              (curr_mu = 1, scale = 1) before (rt.py:278-282)
  2. U      = clip(nominal + eps)                K sampled control sequences around the previous best one
  3. J, k*  = look-ahead rollout of the vehicle's CURRENT best model over U, NMPC cost  (replaces the IPOPT solve rt.py:305:
              there is no NLP solver on this path; the controller is best-of-K)
  4. u      = U[k*][0]; nominal = shift(U[k*])
  5. plant: x+ = RK6(x, u; true parameters with the vehicle's friction schedule)         rt.py:274,311
  6. friction lists: the W + 1 warm-up seeds while idt <= W (written once at construction), afterwards the mean Dr, Df
     of the top-K models of the PREVIOUS tick's look-back; MU_pred = raw `smoothing_mu`-tick moving average (feeds the
     planner), smoothed x 0.95 = the logged value                                          rt.py:326-344
  7. from tick 1 on (the reference skips the transition of tick 0, rt.py:346): push (x, u, x+) into the vehicle's
     history ring; once W transitions are in (tick W): score the whole bank over the window, arg-min + top-K per
     vehicle -> the model of the next tick's look-ahead                                    rt.py:347-366

Because the friction estimate of a tick only needs the previous tick's top-K (that is the reference's order: estimate at
rt.py:326-344, look-back at :347-366), the look-back of tick t and steps 1-2 of tick t + 1 are independent: `tick()` runs
them concurrently on two streams (fork after the estimate, join before returning), so after `tick()` the plan (`xref`,
`U`, `projidx`) is already the one of the NEXT tick; the constructor plans for tick 0.
"""

import numpy as np

from .. import _lib
from ..bank import ModelBank, PARAM_NAMES
from ..tracks import RacelineTable
from .lookback import LookbackLaunch


class MonteCarlo:
    def __init__(self, bank_params, table, x_init, projidx_init, plant_params, drop_start, V=None, W=20, K_models=10,
                 K_seq=32, H=20, Ts=0.02, scale=0.9, mu_init=1.0, seed=4, sigma_pwm=0.1, sigma_steer=0.05,
                 drop_rate=1.0 / 22.0, drop_len=0.2, initial_model=0, smoothing_mu=20, mu_alpha=0.08, limits=None,
                 lookback_mode="rolling", use_graphs=False, fast_sin=None, lookback_kernel=None, overlap="lookback_first"):
        torch = _lib.require_cuda()
        self.torch, self.L = torch, _lib.lib()
        self.bank = bank_params if isinstance(bank_params, ModelBank) else ModelBank(bank_params)
        if not isinstance(table, RacelineTable):
            table = RacelineTable.from_track(table)
        self.table = table
        dev = self.bank.device
        self.dev = dev
        x_init = np.atleast_2d(np.asarray(x_init, dtype=np.float64))
        self.V = V = x_init.shape[0] if V is None else V
        x_init = np.ascontiguousarray(np.broadcast_to(x_init, (V, 6)))
        self.W, self.Km, self.Ks, self.H, self.Ts, self.scale = W, K_models, K_seq, H, float(Ts), float(scale)
        if K_models > _lib.LIST_LEN:
            raise ValueError("K_models <= %d" % _lib.LIST_LEN)
        N = self.bank.N
        f64, f32, i32, i64 = torch.float64, torch.float32, torch.int32, torch.int64
        dv = lambda a, dt: torch.from_numpy(np.array(a, copy=True, order="C")).to(dev).to(dt)
        self.x = dv(x_init, f64)
        self.x_next = torch.empty_like(self.x)
        pp = np.stack([np.broadcast_to(np.asarray(plant_params[k], dtype=np.float64), (V,)) for k in PARAM_NAMES], axis=1)
        self.plant = dv(pp, f64)                                   # [V][14] true parameters (friction evolves)
        self.drop_start = dv(np.broadcast_to(np.asarray(drop_start, dtype=np.float64), (V,)), f64)
        self.drop_rate, self.drop_len = float(drop_rate), float(drop_len)
        self.projidx = dv(np.broadcast_to(np.asarray(projidx_init), (V,)).astype(np.int32), i32)
        # friction estimate (rt.py:326-344): mu_pred = MU_pred, the raw moving average the planner receives from tick
        # W + 2 on (rt.py:278-280); mu_display = smoothed x 0.95, the value the reference logs (MU_preds); mu_default = the
        # planner's default curr_mu = 1 used before (planner.py:12)
        self.mu_pred = torch.full((V,), float("nan"), dtype=f64, device=dev)
        self.mu_display = torch.full((V,), float(mu_init), dtype=f64, device=dev)
        self.mu_default = torch.ones((V,), dtype=f64, device=dev)
        self.smoothing = int(smoothing_mu)
        self.mu_alpha = float(mu_alpha)
        self.mu_state = torch.zeros((V, 2 * self.smoothing + 3), dtype=f64, device=dev)
        m0, lf0, lr0 = (float(np.ravel(self.bank.param(k, 0))[0]) for k in ("mass", "lf", "lr"))
        with torch.cuda.device(dev):                               # the W + 1 warm-up entries of rt.py:326-330 (g = 9.8)
            _lib.check(self.L.llampc_mu_seed_f64(self.mu_state.data_ptr(), V, self.smoothing, W + 1,
                                                 float(mu_init) * m0 * 9.8 * lr0 / (lf0 + lr0),
                                                 float(mu_init) * m0 * 9.8 * lf0 / (lf0 + lr0),
                                                 torch.cuda.current_stream().cuda_stream), "mu_seed")
        lim = limits or {"min_pwm": -0.1, "max_pwm": 1.0, "min_steer": -0.35, "max_steer": 0.35}
        self.box = np.array([lim["min_pwm"], lim["max_pwm"], lim["min_steer"], lim["max_steer"]], dtype=np.float32)
        rng = np.random.RandomState(seed)
        eps = np.stack([sigma_pwm * rng.randn(K_seq, H), sigma_steer * rng.randn(K_seq, H)], axis=-1)
        eps[0] = 0.0                                               # sequence 0 = the unperturbed nominal
        self.eps = dv(eps, f32)
        self.nominal = torch.zeros((V, H, 2), dtype=f32, device=dev)
        self.nominal[:, :, 0] = 0.5
        self.uprev = torch.zeros((V, 2), dtype=f32, device=dev)
        self.uprev[:, 0] = 0.5
        pad = 4                                                    # bulk-copy granularity of the shared-table path
        self.U = torch.zeros(V * K_seq * H * 2 + pad, dtype=f32, device=dev)
        self.xref32 = torch.zeros(V * (H + 1) * 2 + pad, dtype=f32, device=dev)
        self.J = torch.empty((V, K_seq), dtype=f32, device=dev)
        self.best_k = torch.zeros(V, dtype=i32, device=dev)
        self.u_applied = torch.zeros((V, 2), dtype=f64, device=dev)
        self.model_idx = torch.full((V,), int(initial_model), dtype=i32, device=dev)
        self.hist = torch.zeros((V, W, _lib.HIST_ROW), dtype=f32, device=dev)
        if lookback_mode not in ("rolling", "recompute"):
            raise ValueError("lookback_mode must be 'rolling' or 'recompute'")
        # "rolling" = the reference's own bookkeeping (rt.py:352-354): per-vehicle (W, N) error ring, one new column per
        # tick; "recompute" re-integrates every vehicle's whole window every tick (W times the look-back work)
        self.rolling = lookback_mode == "rolling"
        self.err_ring = torch.zeros((V, _lib.ring_rows(W), self.bank.Npad), dtype=f32, device=dev) if self.rolling else None
        self.topk = torch.zeros((V, _lib.LIST_LEN + 1), dtype=i64, device=dev)
        # the look-back launch of every tick (llampc_lookback_launch): rolling -> K1v (one CTA per vehicle), recompute ->
        # K1p over (candidate tile, vehicle) with the last-CTA merge; tyre sine None = automatic (SFU while the bank's
        # tyre-sine argument stays within [-pi, pi]), True / False force a mode
        self.lb = LookbackLaunch(self.bank, self.hist, W, self.Ts, K=K_models, n_vehicles=V, hist_stride_rows=W,
                                 mode=lookback_mode, err_ring=self.err_ring, out=self.topk, fast_sin=fast_sin,
                                 kernel=lookback_kernel)
        self.qrp = np.array([1.0, 1.0, 5e-3, 1.0, 0.0, 0.0], dtype=np.float32)
        self.tick_count = 0
        self.lookback_steps = 0
        self.lookahead_steps = 0
        # the time of the friction schedule lives on the device so that a whole tick can be replayed as a CUDA graph
        self.t_dev = torch.zeros((), dtype=f64, device=dev)
        self.use_graphs = bool(use_graphs)
        self._graphs = {}                                          # ring slot -> captured tick (steady state only)
        # planner + control sampling of the next tick run beside the look-back of this one, on a side stream.
        #   overlap = "lookback_first" (default): the look-back is enqueued first on the main stream, the planner and the
        #       control sampling follow on a side stream of the same priority and take the SM slots the look-back's last
        #       wave leaves free (measured, graph replay: 210 us per tick; with a high-priority side stream the planner's
        #       CTAs displace look-back CTAs and stretch it from 98 to 131 us: 228 us);
        #   "plan_first": the round-1 order (planner on a high-priority stream ahead of the look-back: right for the old
        #       32-CTA planner, but the 512-CTA planner then takes every SM's shared memory first and delays the look-back);
        #   "none": everything on one stream.
        if overlap not in ("lookback_first", "plan_first", "none"):
            raise ValueError("overlap must be 'lookback_first', 'plan_first' or 'none'")
        self.overlap = overlap
        self._side = torch.cuda.Stream(device=dev, priority=-1 if overlap == "plan_first" else 0)
        with torch.cuda.device(self.dev):
            self._plan(torch.cuda.current_stream().cuda_stream, 0)   # the plan of tick 0

    # ------------------------------------------------------------------ one tick, asynchronous on the current stream
    def tick(self):
        """One closed-loop tick for all vehicles.  Once the windows are full the tick for each of the W ring slots is
        captured once as a CUDA graph (all arguments are then static) and replayed: one graph launch per tick."""
        if self.use_graphs and self.tick_count >= self.W + 2:
            torch = self.torch
            slot = self.tick_count % self.W
            g = self._graphs.get(slot)
            if g is None:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._tick_body()
                self._graphs[slot] = g                             # capture does not execute: replay runs the tick
            g.replay()
            self._after_tick()
            return
        self._tick_body()
        self._after_tick()

    def _after_tick(self):
        V, bank = self.V, self.bank
        if self.rolling:
            self.lookback_steps += V * bank.N if self.tick_count >= 1 else 0
        elif self.tick_count >= self.W:
            self.lookback_steps += V * bank.N * self.W
        self.lookahead_steps += V * self.Ks * self.H
        self.tick_count += 1

    def _plan(self, st, idt):
        """Steps 1-2 for tick `idt`: reference path + sampled control sequences (reads x, projidx, mu_pred, nominal).  The
        planner gets MU_pred and scale = v_factor only when idt > W + 1, its defaults (1, 1) before (rt.py:278-282)."""
        L, V, chk = self.L, self.V, _lib.check
        dev, s, xy, cxy, cvp, mus = self.table.device_tables()
        fed = idt > self.W + 1
        chk(L.llampc_planner_constant_speed_f64(s.data_ptr(), xy.data_ptr(), cxy.data_ptr(), cvp.data_ptr(), mus.data_ptr(), self.table.n,
                                                self.table.n_mu, self.x.data_ptr(), V, self.projidx.data_ptr(),
                                                (self.mu_pred if fed else self.mu_default).data_ptr(), 0, self.H, self.Ts,
                                                self.scale if fed else 1.0,
                                                self.xref32.data_ptr(), None, self.projidx.data_ptr(), None, st), "planner")
        chk(L.llampc_sample_controls_f32(self.nominal.data_ptr(), self.eps.data_ptr(), V, self.Ks, self.H,
                                         self.box.ctypes.data, self.U.data_ptr(), st), "sample_controls")

    def _tick_body(self):
        torch, L, V = self.torch, self.L, self.V
        main = torch.cuda.current_stream()
        st = main.cuda_stream
        bank = self.bank
        chk = _lib.check
        with torch.cuda.device(self.dev):
            chk(L.llampc_lookahead_rollout_f32(bank.packed.data_ptr(), bank.Npad, self.model_idx.data_ptr(), V,
                                               self.x.data_ptr(), V, self.U.data_ptr(), self.Ks, self.H,
                                               self.xref32.data_ptr(), self.uprev.data_ptr(), 1 | 2 | 4, self.qrp.ctypes.data,
                                               self.Ts, self.J.data_ptr(), self.best_k.data_ptr(), None, None, st), "lookahead")
            chk(L.llampc_apply_best_f32(self.U.data_ptr(), self.best_k.data_ptr(), V, self.Ks, self.H, self.nominal.data_ptr(),
                                        self.uprev.data_ptr(), self.u_applied.data_ptr(), st), "apply_best")
            # friction schedule ('sudden' style of run_nmpc_orca_llampc_nrt_avg_runs.py:163-166): Df, Dr decay while the
            # vehicle's drop interval is active
            chk(L.llampc_mc_friction_schedule_f64(self.plant.data_ptr(), V, 8, 2, self.drop_start.data_ptr(), self.drop_len,
                                                  self.drop_rate, self.t_dev.data_ptr(), st), "friction schedule")
            chk(L.llampc_plant_rk6_f64(self.plant.data_ptr(), V, self.x.data_ptr(), self.u_applied.data_ptr(), self.Ts,
                                       self.x_next.data_ptr(), st), "plant")
            idt = self.tick_count
            pushing = idt > 0                                      # rt.py:346: the transition of tick 0 is not scored
            slot = (idt - 1) % self.W
            if pushing:
                chk(L.llampc_pack_rows_f64(self.x.data_ptr(), self.u_applied.data_ptr(), self.x_next.data_ptr(), V, self.Ts,
                                           bank.lf_shared, bank.lr_shared, slot, self.W, self.hist.data_ptr(), None, st), "pack_rows")
            if idt > self.W:                                       # rt.py:331-344: the previous tick's look-back left a top-K
                chk(L.llampc_mu_estimate_f64(self.topk.data_ptr(), _lib.LIST_LEN + 1, self.Km, 0, bank.bank64.data_ptr(),
                                             bank.N, V, self.smoothing, self.mu_alpha, 0.95, 9.81, self.mu_state.data_ptr(),
                                             self.mu_pred.data_ptr(), self.mu_display.data_ptr(), st), "mu_estimate")
            # fork: x <- x_next, t += Ts and the plan of the next tick on the side stream, beside the look-back of this tick
            # on the main stream (the look-back reads the history rings, not x)
            side = main if self.overlap == "none" else self._side
            full = idt >= self.W

            def lookback():
                if pushing and self.rolling:
                    self.lb.launch(st, slot=slot, emit=int(full))
                elif pushing and full:
                    self.lb.launch(st)

            def plan_next():
                if side is not main:
                    side.wait_stream(main)
                ss = side.cuda_stream
                chk(L.llampc_mc_advance_tick_f64(None, 0, None, self.x.data_ptr(), self.x_next.data_ptr(), V,
                                                 self.t_dev.data_ptr(), self.Ts, ss), "advance tick")
                self._plan(ss, idt + 1)

            if self.overlap == "plan_first":
                plan_next()
                lookback()
            else:
                fork = None
                if side is not main:                               # the side work depends on the tick so far, not on the look-back
                    fork = torch.cuda.Event()
                    fork.record(main)
                lookback()
                if fork is not None:
                    side.wait_event(fork)
                    ss = side.cuda_stream
                    chk(L.llampc_mc_advance_tick_f64(None, 0, None, self.x.data_ptr(), self.x_next.data_ptr(), V,
                                                     self.t_dev.data_ptr(), self.Ts, ss), "advance tick")
                    self._plan(ss, idt + 1)
                else:
                    plan_next()
            if full:                                               # the model of the next tick's look-ahead
                chk(L.llampc_mc_advance_tick_f64(self.topk.data_ptr(), _lib.LIST_LEN + 1, self.model_idx.data_ptr(), None,
                                                 None, V, None, self.Ts, st), "model index")
            if side is not main:
                main.wait_stream(side)                             # join

    def run(self, n):
        for _ in range(n):
            self.tick()

    # ------------------------------------------------------------------ host views (tests, logging)
    def host(self):
        V, H, Ks = self.V, self.H, self.Ks
        c = lambda t: t.cpu().numpy()
        keys = c(self.topk).view(np.uint64)
        return {
            "x": c(self.x), "projidx": c(self.projidx), "mu_pred": c(self.mu_pred), "mu_display": c(self.mu_display),
            "model_idx": c(self.model_idx),
            "xref": np.swapaxes(c(self.xref32)[:V * (H + 1) * 2].reshape(V, H + 1, 2), 1, 2),
            "U": c(self.U)[:V * Ks * H * 2].reshape(V, Ks, H, 2), "J": c(self.J), "best_k": c(self.best_k),
            "u_applied": c(self.u_applied), "nominal": c(self.nominal), "uprev": c(self.uprev), "plant": c(self.plant),
            "hist": c(self.hist), "topk_idx": (keys[:, 1:] & np.uint64(0xFFFFFFFF)).astype(np.int64),
            "best_idx": (keys[:, 0] & np.uint64(0xFFFFFFFF)).astype(np.int64),
        }
