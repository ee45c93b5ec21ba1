#!/usr/bin/env python
"""Benchmark of the B200-native LLA-MPC look-back hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): candidate-model RK4 steps/s.  One bench "step" = one MPC tick of the look-back step =
one pass of the hot path over the whole bank and window:
    N = 1 : config C2  - 65,536 candidates (6 Pacejka + mass varied) x 50-step window, per-tick arg-min + top-10
    N > 1 : config C5  - 1,048,576 candidates x 50-step window sharded by contiguous index range over N GPUs,
            local reduce + ONE min-loc exchange (packed 64-bit key) per tick, fused into the launch over NVLink peer
            memory (LLAMPC_BENCH_NCCL=1: NCCL MIN all-reduce); the line carries a `parity` block (all ranks agree, equal
            to the NCCL variant and to the float64 oracle) and the C5-on-1-GPU rate the scaling is to be read against
`value` is measured with the inputs resident in HBM (CUDA events on the launching stream, L2 flushed between
ticks); `e2e` goes through the public Python API (LookBack.push) with host NumPy inputs: host packing,
pinned H2D of the history row, kernels, D2H of the selected indices, every tick.
`--impl reference` times the reference's own CPU path on all host cores: the unmodified evaluate_models_vectorized
(imported from the reference package when it resolves -- baseline/_ref on the GPU box -- kind "reference"; else the
NumPy oracle port, kind "port") + scoring, float64, on the full 65,536-candidate bank per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RF_READS_PER_WARP_STEP = 1231          # register source operands of the K1p loop body (tools/sass_reg_reads.py)
RF_READS_PER_CLK = 1.86                # measured per scheduler (tools/ubench/ffma2_operands.cu)
F_ALG = 652.0            # FP32 flop per candidate-RK4-step (SURVEY.md 8(d) convention)
S_ALG = 40.0             # SFU-class ops per step (same convention)
NPARAM_PACKED = 16       # floats per candidate in the packed bank
W_C2, N_C2 = 50, 65536
N_C5 = 1 << 20
TS = 0.02
C2_VARIATION = (("Br", 0.2), ("Cr", 0.1), ("Dr", 0.5), ("Bf", 0.2), ("Cf", 0.1), ("Df", 0.5), ("mass", 0.15))
NOMINAL = {"lf": 0.029, "lr": 0.033, "mass": 0.041, "Iz": 27.8e-6, "Bf": 2.579, "Br": 3.3852, "Cf": 1.2, "Cr": 1.2691,
           "Df": 0.192, "Dr": 0.1737, "Cm1": 0.287, "Cm2": 0.0545, "Cr0": 0.0518, "Cr2": 0.00035}


def make_bank(n, seed, lo=0, hi=None):
    """Synthetic bank: nominal x (1 + sigma randn), drawn like run_nmpc_orca_llampc_rt.py:145-179 (per model,
    per varied parameter) from a seeded RandomState; [lo, hi) selects a shard."""
    hi = n if hi is None else hi
    z = np.random.RandomState(seed).randn(n, len(C2_VARIATION))[lo:hi]
    bank = dict(NOMINAL)
    for j, (name, sigma) in enumerate(C2_VARIATION):
        bank[name] = NOMINAL[name] * (1 + sigma * z[:, j])
    return bank


def make_bank_rt(n, seed):
    """C1-style bank: the six Pacejka parameters varied (run_nmpc_orca_llampc_rt.py:153-158), shared mass."""
    z = np.random.RandomState(seed).randn(n, 6)
    bank = dict(NOMINAL)
    for j, (name, sigma) in enumerate(C2_VARIATION[:6]):
        bank[name] = NOMINAL[name] * (1 + sigma * z[:, j])
    return bank


def synthetic_history(n_ticks, plant_step):
    """History B of SURVEY.md 8(d): ORCA plant (RK6) from the ETHZMobil start, sinusoidal inputs, sudden
    friction drop; `plant_step(params, x, u)` is the integrator used (GPU plant kernel or the CPU port)."""
    x = np.array([1.2, 0.9, 0.0, 1.0, 0.0, 0.0])
    p = dict(NOMINAL)
    S, U = [x.copy()], []
    for k in range(n_ticks):
        t = k * TS
        if 0.6 < t < 0.8:                                    # 'sudden' friction drop (avg_runs.py:163-166 shape)
            p["Df"] *= (1 - 1 / 22.0)
            p["Dr"] *= (1 - 1 / 22.0)
        u = np.array([0.45 + 0.35 * np.sin(0.7 * t), 0.25 * np.sin(1.3 * t)])
        x = plant_step(p, x, u)
        S.append(x.copy())
        U.append(u)
    return np.array(S).T, np.array(U).T


class ClockSampler:
    """SM clock and throttle reasons sampled with NVML (every ~2 ms) during the timed region."""

    def __init__(self, index=0):
        self.index, self.sm, self.mx, self.reasons, self.stop = index, [], None, set(), False
        self.th = None

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:                                   # no NVML: report an empty record
            self.err = repr(e)
        return self

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80}
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.002)

    def __exit__(self, *a):
        self.stop = True
        if self.th is not None:
            self.th.join(timeout=2)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0, "error": getattr(self, "err", None)}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU reference arm: the reference's own evaluate_models_vectorized + scoring (float64 NumPy) when the reference package
# resolves (oracle/reference_adapter.py: LLAMPC_REFERENCE_ROOT, /root/reference, or the offline install under
# baseline/_ref), else the oracle port of the same lines; sharded over host processes
# ----------------------------------------------------------------------------------------------------------
_REF = {"ns": None, "m0": None}


def reference_kind():
    try:
        from oracle import reference_adapter as ra
        if ra.available():
            ra.load()
            return "reference"
    except Exception as e:                                      # broken tree: fall back to the port and say so
        sys.stderr.write("reference package not usable (%r): timing the oracle port\n" % (e,))
    return "port"


def _cpu_worker(args):
    bank, S, U, W, kind = args
    t0 = time.perf_counter()
    if kind == "reference":
        # the reference's own per-tick lines (rt.py:349 + :357-360), once per window row: the model list only supplies
        # len() and models[0] (evaluate_models_vectorized.py:6,16-19), so one Dynamic repeated stands for the bank
        from oracle import reference_adapter as ra
        ref = ra.load()
        if _REF["m0"] is None:
            _REF["m0"] = ref.Dynamic(**ref.ORCA(control='pwm'))
        n = len(bank["Bf"])
        models = [_REF["m0"]] * n
        pp = tuple(bank[k] for k in ("Bf", "Cf", "Df", "Br", "Cr", "Dr"))
        ew = np.empty((n, W))
        for j in range(W):
            pred = ref.evaluate_models_vectorized(models, n, S[:, j], U[:, j], TS, pp)
            ew[:, j] = np.mean((pred - S[0:4, j + 1]) ** 2, axis=1)
        avg = np.mean(ew, axis=1)
        best, topk = int(np.argmin(avg)), avg.argsort()[:10]
    else:
        from oracle import llampc_oracle as orc
        ew = orc.window_errors(bank, S, U, W - 1, W, TS)
        avg = ew.mean(axis=1)
        best, topk = orc.select(avg, 10)
    return time.perf_counter() - t0, best


class CpuReference:
    """The reference algorithm (float64 NumPy) on `procs` host processes over contiguous candidate shards."""

    def __init__(self, n_cand, W, procs, S, U, kind):
        import multiprocessing as mp
        bank = make_bank(n_cand, seed=1)
        self.n_cand, self.W, self.shards, self.kind = n_cand, W, [], kind
        per = (n_cand + procs - 1) // procs
        for r in range(procs):
            lo, hi = r * per, min(n_cand, (r + 1) * per)
            if lo < hi:
                self.shards.append(({k: (v[lo:hi] if np.ndim(v) else v) for k, v in bank.items()}, S[:, :W + 1], U[:, :W], W, kind))
        self.pool = mp.get_context("fork").Pool(len(self.shards)) if len(self.shards) > 1 else None

    def step(self):
        """One pass over the sample; returns seconds."""
        t0 = time.perf_counter()
        if self.pool is None:
            _cpu_worker(self.shards[0])
        else:
            self.pool.map(_cpu_worker, self.shards)
        return time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.terminate()


def cpu_reference_rate(n_cand, W, procs, S, U, kind, repeats=1):
    ref = CpuReference(n_cand, W, procs, S, U, kind)
    ref.step()                                                 # warm-up (imports, page-in)
    best_t = min(ref.step() for _ in range(repeats))
    ref.close()
    return n_cand * W / best_t, best_t


def cpu_history():
    from oracle import llampc_oracle as orc
    return synthetic_history(W_C2 + 2, lambda p, x, u: orc.rk6_step(p, x, u, 0, TS))


REF_WHAT = {"reference": "the reference's own evaluate_models_vectorized (llampc/mpc/evaluate_models_vectorized.py, imported "
                         "unmodified) called once per window row + errors / mean / argmin / argsort()[:10] (rt.py:349-360), "
                         "float64 NumPy; mass from models[0] as that signature dictates",
            "port": "float64 NumPy port of evaluate_models_vectorized + scoring (oracle/llampc_oracle.py; the reference "
                    "package did not resolve on this box)"}


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = reference_kind()
    S, U = cpu_history()
    n_total = N_C2 if args.gpus == 1 else N_C5
    n_sample = N_C2                                             # the whole C2 bank per step (C5: a 65,536-candidate sample)
    ref = CpuReference(n_sample, W_C2, cores, S, U, kind)
    t_first = ref.step()                                        # also the first warm-up step
    budget = 150.0                                              # seconds for warm-up + timed steps
    if t_first * (args.steps + max(args.warmup, 1)) > budget and n_sample > 16384:
        ref.close()                                             # a slow box: bound the sample so the run ends in minutes
        n_sample = int(max(16384, n_sample * budget / (t_first * (args.steps + max(args.warmup, 1)))) // 1024 * 1024)
        ref = CpuReference(n_sample, W_C2, cores, S, U, kind)
        ref.step()
    for _ in range(max(args.warmup, 1) - 1):
        ref.step()
    total = sum(ref.step() for _ in range(args.steps))
    ref.close()
    value = n_sample * W_C2 * args.steps / total
    sample = "%d of %d candidates x %d-step window per step; %s; %d processes over contiguous candidate shards" % (
        n_sample, n_total, W_C2, REF_WHAT[kind], cores)
    line = {"impl": "reference", "metric": "candidate-model RK4 steps/s (look-back window)", "value": value,
            "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


L2_BYTES = 126e6                       # B200 L2


def bank_copies(n_local, ticks):
    """Copies of the rank's packed bank (64 B per candidate) for a run of `ticks` launches: one per launch, so that no tick
    ever re-reads a buffer (capped at 16 GB; never fewer than what exceeds the L2 by a third)."""
    floor = max(2, int(np.ceil(1.35 * L2_BYTES / (64.0 * n_local))))
    return max(floor, min(int(ticks), int(16e9 // (64 * n_local))))


def l2_note(gpus):
    return ("inputs larger than L2: every tick of the run (warm-up included) reads its OWN copy of the %s bank -- distinct HBM "
            "buffers, written once and followed by a 256 MiB read-modify-write L2 flush before the run, never read before their tick (beyond 16 GB "
            "of copies they are reused in rotation) -- so the K timed ticks are launched back to back between ONE pair of CUDA "
            "events with no flush inside the timed region; the per-tick figure with a 256 MiB read-modify-write L2 flush before every tick is "
            "reported as l2_flushed" % ("rank-local" if gpus > 1 else "whole"))


def workload_config(gpus):
    if gpus == 1:
        return {"workload": "C2: look-back, 65,536 candidates (6 Pacejka + mass varied) x 50-step window, arg-min + top-10 per tick",
                "candidates": N_C2, "window": W_C2, "Ts": TS, "l2": l2_note(gpus)}
    return {"workload": "C5: look-back sweep, 1,048,576 candidates x 50-step window sharded over %d GPUs, one min-loc exchange per tick "
                        "(fused into the kernels over NVLink peer memory; LLAMPC_BENCH_NCCL=1 = NCCL MIN all-reduce)" % gpus,
            "candidates": N_C5, "window": W_C2, "Ts": TS, "l2": l2_note(gpus)}


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def cpu_baseline_leg():
    """Reference algorithm on the host cores (bounded sample), run BEFORE CUDA is initialised (fork safety)."""
    cores = os.cpu_count() or 1
    kind = reference_kind()
    Sc, Uc = cpu_history()
    n_all = N_C2 if cores >= 8 else int(max(16384, 2048 * cores))      # ~10-30 s of CPU work in total
    r1, _ = cpu_reference_rate(16384, W_C2, 1, Sc, Uc, kind, repeats=1)
    rall, _ = cpu_reference_rate(n_all, W_C2, cores, Sc, Uc, kind, repeats=4)
    return {"value": rall, "unit": "steps/s", "cores": cores, "kind": kind,
            "sample": "%d of the 65,536 C2 candidates x 50-step window (best of 4 passes); %s; sharded over %d processes; "
                      "1-core figure: 16,384 candidates" % (n_all, REF_WHAT[kind], cores),
            "value_1core_as_reference_runs_it": r1}


def run_b200(args, emit):
    cpu_base = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_cpu:
        cpu_base = cpu_baseline_leg()
    import torch
    import torch.distributed as td
    from llampc_b200 import _lib
    from llampc_b200.bank import ModelBank
    from llampc_b200.models import Dynamic
    from llampc_b200.mpc import LookBack
    from llampc_b200.mpc.lookback import LookbackLaunch, decode_keys

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    # ---- synthetic inputs (plant = the product's own fp64 RK6 kernel)
    plant = Dynamic(**NOMINAL)

    def plant_step(p, x, u):
        plant.Df, plant.Dr = p["Df"], p["Dr"]
        return plant.plant_step(x[None], u[None], TS)[0]

    S, U = synthetic_history(W_C2 + 2 + args.steps + args.warmup + 8, plant_step)
    n_total = N_C2 if world == 1 else N_C5
    per = (n_total + world - 1) // world
    lo, hi = rank * per, min(n_total, (rank + 1) * per)
    bank = ModelBank(make_bank(n_total, seed=1 if world == 1 else 5, lo=lo, hi=hi))
    n_local = hi - lo
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def window_rows(b, t0=0):
        rows = np.zeros((W_C2, _lib.HIST_ROW), dtype=np.float32)
        for j in range(W_C2):
            xk, uk, xk1 = (np.ascontiguousarray(a) for a in (S[:, t0 + j], U[:, t0 + j], S[:, t0 + j + 1]))
            L.llampc_hist_row_pack_h(xk.ctypes.data, uk.ctypes.data, xk1.ctypes.data, TS, b.lf_shared, b.lr_shared,
                                     rows[j].ctypes.data, None)
        return torch.from_numpy(rows).to(dev)

    hist = window_rows(bank)
    avg_err = torch.empty(n_local, dtype=torch.float32, device=dev)
    # N > 1: the min-loc across GPUs is fused into the same launch over NVLink peer memory (symmetric buffers);
    # LLAMPC_BENCH_NCCL=1 times the NCCL MIN all-reduce variant instead
    peer = None
    if world > 1 and os.environ.get("LLAMPC_BENCH_NCCL", "0") != "1":
        from llampc_b200.dist import PeerExchange
        try:
            peer = PeerExchange(device=dev)
        except Exception as e:                                  # symmetric memory unavailable: NCCL min-loc instead
            peer = None
            sys.stderr.write("peer exchange unavailable (%r), using NCCL\n" % (e,))
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        td.all_reduce(ok, op=td.ReduceOp.MIN)                   # every rank must take the same path
        if int(ok.item()) == 0:
            peer = None
    # the tick: ONE launch (scores + per-CTA lists + in-kernel merge tree -> arg-min and top-10; for N > 1 the root of the
    # tree also runs the NVLink min-loc); sine mode and kernel are the library's automatic choices, reported below
    # (programmatic dependent launch: in the back-to-back timed region the bank loads and RK4 rows of tick t + 1 start beside
    # the selection / merge-tree tail of tick t; LLAMPC_BENCH_PDL=0 turns it off)
    use_pdl = os.environ.get("LLAMPC_BENCH_PDL", "1") == "1"
    tick = LookbackLaunch(bank, hist, W_C2, TS, K=10, idx_offset=lo, avg_err=avg_err, peer=peer, pdl=use_pdl)
    scores_only = LookbackLaunch(bank, hist, W_C2, TS, K=0, idx_offset=lo, avg_err=avg_err)
    assert tick.plan.launches == 1

    def tick_device():
        tick.launch()
        if world > 1 and peer is None:
            td.all_reduce(tick.out[0, :1], op=td.ReduceOp.MIN)

    def tick_kernel_only():
        """the same launch without the cross-GPU exchange: the kernel the roofline is quoted on"""
        tick.launch(use_peer=False)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in evs:
            flush.add_(1)                                     # evict L2 between timed iterations (not timed)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        ms = np.array([a.elapsed_time(b) for a, b in evs])
        tot = torch.tensor([ms.sum()], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(tot, op=td.ReduceOp.MAX)
        return float(tot.item()), ms

    # ---- the timed region: inputs larger than L2.  Every tick reads a different copy of the bank (distinct HBM buffers whose
    # total exceeds the L2), so the K ticks can be launched back to back between one pair of events: what is measured is the
    # sustained tick rate, without the ~5 us that a pair of CUDA events around a single launch adds (tools/gpu_fixed_cost.py:
    # 4.8 - 6.2 us around a trivial kernel).  The flushed per-tick protocol of round 1 is kept as a second figure.
    warm = max(args.warmup, 3)
    n_rot = bank_copies(n_local, warm + args.steps)
    rot = [bank.packed.clone() for _ in range(n_rot)]
    rot_ptrs = [t.data_ptr() for t in rot]

    def tick_rot(i):
        tick.desc.bank = rot_ptrs[i % n_rot]
        tick_device()

    def tick_rot_kernel_only(i):
        tick.desc.bank = rot_ptrs[i % n_rot]
        tick.launch(use_peer=False)

    def timed_rot(fn, steps, warmup):
        flush.add_(1)                                         # the copies were written (or read by the previous run): evict them
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps):
            fn(warmup + i)
        b.record()
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        tot = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(tot, op=td.ReduceOp.MAX)
        return float(tot.item())

    with ClockSampler(local) as clk:
        total_ms = timed_rot(tick_rot, args.steps, warm)
        k1_total_ms = timed_rot(tick_rot_kernel_only, args.steps, warm)
        tick.desc.bank = bank.packed.data_ptr()
        fl_total_ms, per_ms = timed(tick_device, min(args.steps, 50), 3)
        _, k1_ms = timed(tick_kernel_only, min(args.steps, 50), 3)
        _, k1_bare_ms = timed(scores_only.launch, min(args.steps, 50), 3)
    clocks = clk.summary()
    del rot
    steps_per_tick = n_total * W_C2
    value = steps_per_tick * args.steps / (total_ms * 1e-3)
    k1_avg_s = k1_total_ms * 1e-3 / args.steps
    k1_rate = n_local * W_C2 / k1_avg_s
    l2_flushed = {"value": steps_per_tick * len(per_ms) / (fl_total_ms * 1e-3), "ms_per_step": fl_total_ms / len(per_ms),
                  "kernel_us": float(np.mean(k1_ms)) * 1e3, "ticks": int(len(per_ms)),
                  "how": "round-1 protocol: 256 MiB L2 flush before every tick (outside the timed events), one pair of CUDA "
                         "events around each tick launch; includes the ~5 us a pair of events adds to any single launch"}

    # ---- parity of the timed path: the key the timed launch leaves in out[0] must be the same on every rank, equal the
    # NCCL variant (MIN all-reduce of the rank-local arg-min keys) and the arg-min of the float64 oracle on a sample
    parity = None
    tick_device()
    torch.cuda.synchronize()
    key_dev = tick.out[0, 0:1].clone()
    local_key = tick.out[0, 1:2].clone()                        # rank-local arg-min = head of the local top-10
    if world > 1:
        gathered = torch.zeros(world, dtype=torch.int64, device=dev)
        td.all_gather_into_tensor(gathered, key_dev)
        nccl_key = local_key.clone()
        td.all_reduce(nccl_key, op=td.ReduceOp.MIN)
        allk = gathered.cpu().numpy()
        ranks_agree = bool((allk == allk[0]).all())
        vs_nccl = bool(int(key_dev.item()) == int(nccl_key.item()))
    else:
        ranks_agree, vs_nccl = True, None
    gkey = np.array([int(key_dev.item())], dtype=np.int64).view(np.uint64)
    g_err, g_idx = decode_keys(gkey)
    # float64 oracle (checker) on the winner + a random sample of this rank's shard: the winner must beat the sample
    from oracle import llampc_oracle as orc
    bank_h = make_bank(n_total, seed=1 if world == 1 else 5)
    rng = np.random.RandomState(123)
    samp = np.unique(np.concatenate([rng.randint(0, n_total, 2048), g_idx]))
    sub = {k: (v[samp] if np.ndim(v) else v) for k, v in bank_h.items()}
    ref = np.mean(orc.window_errors(sub, S, U, W_C2 - 1, W_C2, TS), axis=1)
    w = int(np.flatnonzero(samp == g_idx[0])[0])
    oracle_ok = bool(int(np.argmin(ref)) == w and abs(float(g_err[0]) - ref[w]) <= 1e-4 * ref[w])
    parity = {"ranks_agree": ranks_agree, "vs_nccl": vs_nccl, "key": "0x%016x" % int(gkey[0]), "index": int(g_idx[0]),
              "vs_oracle_f64": oracle_ok,
              "how": "out[0] of the timed launch all-gathered over the ranks; NCCL MIN all-reduce of the rank-local arg-min "
                     "keys; float64 oracle on the winner + 2,048 random candidates of the whole bank"}
    if not (ranks_agree and oracle_ok and vs_nccl in (True, None)):
        raise SystemExit("bench: parity of the timed path FAILED: %r" % (parity,))

    # ---- roofline of the dominant kernel (K1): FP32 pipe, plus the SFU and HBM readings for context
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12                       # TFLOP/s at the max SM clock
    achieved = k1_rate * F_ALG / 1e12
    hbm_bytes = n_local * (NPARAM_PACKED * 4 + 4) + W_C2 * 80
    kname = {"K1": "lookback_window_kernel", "K1p": "lookback_window2_kernel", "K1b": "lookback_balanced_kernel"}.get(tick.kernel_name, tick.kernel_name)
    roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one tick launch at C2: NOT measured in this run, taken
                # from the committed ncu --set full capture named in traffic_source
                "traffic": 4264448 if world == 1 else None,
                "traffic_source": "ncu --set full capture profiles/r02_k1p_c2_ncu.md (4,264,448 B read, 0 B written back within the "
                                  "launch: the scores stay in L2) and, per timed launch of this very command under ncu "
                                  "--cache-control none, profiles/r02_bench_dram_per_launch.csv (4.26 MB read by every timed "
                                  "tick); not captured for the sharded C5 launches",
                "kernel": "%s (%s: scores + selection + in-kernel tree merge = the whole tick), window split %d, tyre sine %s"
                          % (kname, tick.kernel_name, tick.plan.split, tick.sine_name),
                "kernel_us": k1_avg_s * 1e6, "kernel_us_l2_flushed": float(np.mean(k1_ms)) * 1e3,
                "kernel_us_scores_only_l2_flushed": float(np.mean(k1_bare_ms)) * 1e3,
                "peak_source": "148 SM x 128 FP32 lanes x 2 x %.0f MHz (sm_max_mhz of MEASURED_PEAKS.json; tensor/HBM peaks do not bound this elementwise ODE kernel)" % sm_max,
                "flop_per_step": F_ALG, "steps_per_launch": n_local * W_C2,
                "sfu": {"achieved_Tops": k1_rate * S_ALG / 1e12, "peak_Tops": 148 * 16 * sm_max * 1e6 / 1e12},
                "hbm": {"algorithmic_bytes": hbm_bytes, "achieved_GBs": hbm_bytes / k1_avg_s / 1e9,
                        "peak_GBs": peaks.get("hbm_gbs", 6650.0), "peak_kind": "measured" if peaks else "fallback"}}
    if clocks.get("sm_mhz"):
        roofline["frac_at_observed_clock"] = achieved / (148 * 128 * 2 * clocks["sm_mhz"] * 1e6 / 1e12)
    if tick.kernel_name == "K1p":
        # What actually binds the packed step (profiles/r02_register_bandwidth.md): a B200 scheduler delivers ~1.86 32-bit
        # register source operands per clock (tools/ubench/ffma2_operands.cu: FFMA2 with three distinct register pairs
        # issues every 3.2 clocks, not 2), and the loop body of K1p reads 1,231 of them per warp-step of 64 candidate-steps
        # (tools/sass_reg_reads.py on the shipped library) against 559 FMA-pipe clocks.
        reads = n_local * W_C2 / 64.0 * RF_READS_PER_WARP_STEP
        cap = 148 * 4 * RF_READS_PER_CLK * sm_max * 1e6 * k1_avg_s
        roofline["register_file"] = {"reads_per_warp_step": RF_READS_PER_WARP_STEP, "peak_reads_per_clk_per_scheduler": RF_READS_PER_CLK,
                                     "frac_whole_launch": reads / cap,
                                     "clocks_per_warp_step_bound": RF_READS_PER_WARP_STEP / RF_READS_PER_CLK,
                                     "fma_pipe_clocks_per_warp_step": 559,
                                     "source": "tools/sass_reg_reads.py (SASS of the shipped K1p loop) and "
                                               "tools/ubench/ffma2_operands.cu (profiles/r02_register_bandwidth.md)"}
    # SM clock sustained under an FMA-bound load (clock64 against globaltimer, ~300 us of work on every SM)
    try:
        probe = torch.zeros(2, dtype=torch.int64, device=dev)
        sink = torch.zeros(1, dtype=torch.float32, device=dev)
        for _ in range(3):
            _lib.check(L.llampc_clock_probe(40000, probe.data_ptr(), sink.data_ptr(), st), "clock_probe")
        torch.cuda.synchronize()
        cyc, ns = (int(v) for v in probe.cpu().numpy())
        mhz = cyc / ns * 1e3
        roofline["sm_clock_under_fma_load_mhz"] = mhz
        roofline["frac_at_measured_clock"] = achieved / (148 * 128 * 2 * mhz * 1e6 / 1e12)
    except Exception as e:
        roofline["sm_clock_under_fma_load_mhz"] = repr(e)

    # ---- end to end through the public API (rank-local bank; host inputs every tick).  The wall-clock loops run with the cyclic
    # garbage collector off: a generation-2 collection over the CPU leg's objects stalled ONE push of a run for 42 ms.
    import gc
    gc.collect()
    gc.disable()
    lbe = LookBack(bank, W=W_C2, Ts=TS, K=10, refine=16, idx_offset=lo, group=(td.group.WORLD if world > 1 else None))
    for t in range(W_C2 + max(args.warmup, 3)):
        lbe.push(S[:, t], U[:, t], S[:, t + 1])
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    lats = []
    t_base = W_C2 + max(args.warmup, 3)
    t0 = time.perf_counter()
    e2e_best = None
    push_outs = []
    for i in range(args.steps):
        t = t_base + i
        a = time.perf_counter()
        e2e_best, topk, err = lbe.push(S[:, t], U[:, t], S[:, t + 1])
        lats.append(time.perf_counter() - a)
        push_outs.append((e2e_best, list(topk), err))
    wall = time.perf_counter() - t0
    if world > 1:
        wt = torch.tensor([wall], dtype=torch.float64, device=dev)
        td.all_reduce(wt, op=td.ReduceOp.MAX)
        wall = float(wt.item())
        eb = torch.tensor([e2e_best], dtype=torch.int64, device=dev)
        ebs = torch.zeros(world, dtype=torch.int64, device=dev)
        td.all_gather_into_tensor(ebs, eb)
        parity["e2e_ranks_agree"] = bool((ebs == ebs[0]).all().item())
    h2d = _lib.HIST_ROW * 4 + _lib.HIST64_ROW * 8          # the row travels as kernel parameters
    d2h = (1 + 2 * lbe.Kt * (world if world > 1 else 1)) * 8
    # the same K ticks through LookBack.replay: a recorded run handed over as host arrays, up to 4 ticks in flight (every tick
    # still carries its row in the launch parameters and returns its decision through its own mapped pinned slot)
    # (the SAME K transitions as the push loop above, on a second object primed with the same window; decisions must agree)
    lbp = LookBack(bank, W=W_C2, Ts=TS, K=10, refine=16, idx_offset=lo, group=(td.group.WORLD if world > 1 else None))
    tp = np.arange(0, t_base)
    lbp.replay(S[:, tp].T, U[:, tp].T, S[:, tp + 1].T)                              # window primed, slots allocated, paths warm
    ts = np.arange(t_base, t_base + args.steps)
    torch.cuda.synchronize()
    if world > 1:
        td.barrier()
    t0 = time.perf_counter()
    outs = lbp.replay(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    wall_r = time.perf_counter() - t0
    assert len(outs) == args.steps and all(o[0] is not None for o in outs)
    replay_equals_push = all(o[0] == q[0] and list(o[1]) == q[1] and o[2] == q[2] for o, q in zip(outs, push_outs))
    parity["e2e_replay_equals_push"] = bool(replay_equals_push)
    if not replay_equals_push:
        raise SystemExit("bench: LookBack.replay and LookBack.push disagree")
    del lbp
    if world > 1:
        wt = torch.tensor([wall_r], dtype=torch.float64, device=dev)
        td.all_reduce(wt, op=td.ReduceOp.MAX)
        wall_r = float(wt.item())
        eb = torch.tensor([outs[-1][0]], dtype=torch.int64, device=dev)
        ebs = torch.zeros(world, dtype=torch.int64, device=dev)
        td.all_gather_into_tensor(ebs, eb)
        parity["e2e_replay_ranks_agree"] = bool((ebs == ebs[0]).all().item())
    e2e = {"value": steps_per_tick * args.steps / wall_r, "unit": "steps/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h,
           "api": "LookBack.replay (K recorded transitions as host NumPy arrays in, one decision per tick out: arg-min + top-10 "
                  "indices + float64 best error; up to 4 ticks in flight; every tick carries its row to the device in the launch "
                  "parameters, is re-scored in fp64 (16 finalists; for N > 1 the finalist all-gather over NVLink included) and "
                  "returns through its own mapped pinned slot)",
           "sync_push_value": steps_per_tick * args.steps / wall,
           "sync_push_api": "LookBack.push, one tick at a time (the host waits for every decision): the latency of tick_latency"}
    lat = {"p50_us": float(np.percentile(lats, 50) * 1e6), "p95_us": float(np.percentile(lats, 95) * 1e6),
           "max_us": float(np.max(lats) * 1e6), "n_over_1ms": int(np.sum(np.array(lats) > 1e-3)),
           "mode": "recompute (the whole 50-row window re-integrated every tick)"}
    if world == 1:
        # the reference's own rolling bookkeeping (one new error column per tick): same decisions, 1/W of the work
        lbr = LookBack(bank, W=W_C2, Ts=TS, K=10, refine=16, mode="rolling")
        for t in range(W_C2 + 5):
            lbr.push(S[:, t], U[:, t], S[:, t + 1])
        lr = []
        for i in range(args.steps):
            t = W_C2 + 5 + i
            a = time.perf_counter()
            lbr.push(S[:, t], U[:, t], S[:, t + 1])
            lr.append(time.perf_counter() - a)
        lat["rolling_mode_p50_us"] = float(np.percentile(lr, 50) * 1e6)
        lat["rolling_mode_p95_us"] = float(np.percentile(lr, 95) * 1e6)
        del lbr
    del lbe
    gc.enable()

    extras = {}
    scaling_base = None
    if world == 1 and not args.no_extras:
        extras = secondary_configs(torch, L, _lib, S, U, st, flush)
    elif world > 1 and rank == 0:
        # the same workload (C5, 1,048,576 x 50) on ONE GPU, measured in this run on rank 0: the base the strong-scaling
        # efficiency of this line is to be read against (N = 1 of the driver's sweep runs C2, a different workload)
        scaling_base = c5_single_gpu(torch, _lib, S, U, flush)
    if world > 1:
        td.barrier()

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    line = {"metric": "candidate-model RK4 steps/s (look-back window)", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "tyre_sine": "%s (chosen by the library from the bank: max |C| pi/2 = %.2f rad <= pi)" % (tick.sine_name, bank.sin_arg_max),
            "gpu_launches": args.steps * tick.plan.launches, "clocks": clocks, "roofline": roofline, "parity": parity,
            "launch_overlap": ("programmatic dependent launch: the bank loads and RK4 rows of tick t + 1 run beside the selection / "
                               "merge-tree tail of tick t (LLAMPC_LB_FLAG_PDL)") if use_pdl else "none (stream order)",
            "l2_flushed": l2_flushed}
    if world > 1:
        line["exchange"] = "nvlink-peer-memory min-loc inside the kernel" if peer is not None else "nccl all_reduce(MIN) of the packed key"
        if scaling_base:
            line["scaling_base_value"] = scaling_base["steps_per_s"]
            line["scaling_base"] = scaling_base
            line["strong_scaling_efficiency_vs_c5_on_1_gpu"] = value / (world * scaling_base["steps_per_s"])
    line["e2e"] = e2e
    line["tick_latency"] = lat
    if extras:
        line["other_configs"] = extras
        line["other_configs_protocol"] = ("per-launch CUDA events with a 256 MiB L2 flush before every launch (the l2_flushed protocol), "
                                          "except C5_1gpu ms_per_tick and C3 ms_per_call, which use the protocol of the headline (one bank copy per launch, back to back)")
    if cpu_base:
        line["cpu_baseline"] = cpu_base
    emit(line)
    if world > 1:
        td.destroy_process_group()


def _time_it(torch, flush, fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        flush.add_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in evs])) * 1e-3


def c5_single_gpu(torch, _lib, S, U, flush, with_push=False):
    """C5 on one GPU: 1,048,576 candidates x 50 (the sharded sweep's single-GPU reference point), device-timed tick and
    optionally the end-to-end push."""
    from llampc_b200.bank import ModelBank
    from llampc_b200.mpc import LookBack
    from llampc_b200.mpc.lookback import LookbackLaunch
    bank = ModelBank(make_bank(N_C5, seed=5))
    lb = LookBack(bank, W=W_C2, Ts=TS, K=10, refine=16)
    ts = np.arange(0, W_C2)
    lb.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    ll = LookbackLaunch(bank, lb.hist, W_C2, TS, K=10, avg_err=lb.avg_err, pdl=os.environ.get("LLAMPC_BENCH_PDL", "1") == "1")
    dt_flushed = _time_it(torch, flush, ll.launch, 20)
    # the headline protocol: copies of the bank in rotation (larger than L2), ticks back to back between one pair of events
    reps = 20
    n_rot = bank_copies(N_C5, reps + 3)
    rot = [bank.packed.clone() for _ in range(n_rot)]
    ptrs = [t.data_ptr() for t in rot]
    flush.add_(1)
    for i in range(3):
        ll.desc.bank = ptrs[i % n_rot]
        ll.launch()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        ll.desc.bank = ptrs[(3 + i) % n_rot]
        ll.launch()
    b.record()
    torch.cuda.synchronize()
    ll.desc.bank = bank.packed.data_ptr()
    dt = a.elapsed_time(b) * 1e-3 / reps
    del rot
    out = {"steps_per_s": N_C5 * W_C2 / dt, "ms_per_tick": dt * 1e3, "ms_per_tick_l2_flushed": dt_flushed * 1e3,
           "kernel": ll.kernel_name, "split": ll.plan.split,
           "tyre_sine": ll.sine_name, "roofline_frac": N_C5 * W_C2 / dt * F_ALG / (148 * 128 * 2 * 1965e6)}
    if with_push:
        lat = []
        for t in range(W_C2, W_C2 + 23):
            a0 = time.perf_counter()
            lb.push(S[:, t], U[:, t], S[:, t + 1])
            lat.append(time.perf_counter() - a0)
        p50 = float(np.percentile(lat[3:], 50))
        out["e2e_push_p50_us"] = p50 * 1e6
        out["e2e_steps_per_s"] = N_C5 * W_C2 / p50
    del lb, ll
    return out


def secondary_configs(torch, L, _lib, S, U, st, flush):
    """Device-timed throughput of the other BASELINE configs on one GPU (same units: RK4 steps/s), each with an end-to-end
    figure through the public API with host arrays in and out."""
    from llampc_b200.bank import ModelBank
    from llampc_b200.mpc import LookBack, LookAhead
    from llampc_b200.mpc.lookback import LookbackLaunch
    out = {}
    time_it = lambda fn, reps: _time_it(torch, flush, fn, reps)
    peak_rate = 148 * 128 * 2 * 1965e6 / F_ALG

    # strict mode of the headline config: FMA-pipe polynomial tyre sine instead of MUFU.SIN
    bank2 = ModelBank(make_bank(N_C2, seed=1))
    lbs = LookBack(bank2, W=W_C2, Ts=TS, K=10, refine=0)
    ts = np.arange(0, W_C2)
    lbs.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    ll = LookbackLaunch(bank2, lbs.hist, W_C2, TS, K=10, avg_err=lbs.avg_err, fast_sin=False)
    dt = time_it(ll.launch, 50)
    out["C2_strict_polynomial_sin"] = {"steps_per_s": N_C2 * W_C2 / dt, "us_per_tick": dt * 1e6, "roofline_frac": N_C2 * W_C2 / dt / peak_rate}
    del lbs, ll, bank2

    out["C5_1gpu_lookback_1048576x50"] = c5_single_gpu(torch, _lib, S, U, flush, with_push=True)

    # C1: 1,024 candidates x 20 (the reference's own CPU-runnable case): latency-bound
    bank1 = ModelBank(make_bank_rt(1024, seed=0))
    lb1 = LookBack(bank1, W=20, Ts=TS, K=10, refine=0)
    ts = np.arange(0, 20)
    lb1.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    ll = LookbackLaunch(bank1, lb1.hist, 20, TS, K=10, avg_err=lb1.avg_err)
    dt = time_it(ll.launch, 50)
    out["C1_lookback_1024x20"] = {"steps_per_s": 1024 * 20 / dt, "us_per_tick": dt * 1e6, "kernel": ll.kernel_name, "split": ll.plan.split}
    del lb1, ll
    for mode in ("recompute", "rolling"):                        # the same case end to end (LookBack.push, host in / out)
        lbp = LookBack(bank1, W=20, Ts=TS, K=10, refine=16, mode=mode)
        for t in range(30):
            lbp.push(S[:, t], U[:, t], S[:, t + 1])
        lat = []
        for t in range(30, 60):
            a0 = time.perf_counter()
            lbp.push(S[:, t % 60], U[:, t % 60], S[:, t % 60 + 1])
            lat.append(time.perf_counter() - a0)
        out["C1_lookback_1024x20"]["push_p50_us_" + mode] = float(np.percentile(lat, 50) * 1e6)
        del lbp

    # C3 look-ahead: 16,384 models x 32 control sequences x 20-step horizon with the raceline-tracking cost
    M, K, H = 16384, 32, 20
    rng = np.random.RandomState(3)
    t0 = W_C2
    u_nom = U[:, t0:t0 + H].T
    Useq = u_nom[None] + np.stack([0.1 * rng.randn(K, H), 0.05 * rng.randn(K, H)], axis=-1)
    Useq[..., 0] = np.clip(Useq[..., 0], -0.1, 1.0)
    Useq[..., 1] = np.clip(Useq[..., 1], -0.35, 0.35)
    xref = S[:2, t0:t0 + H + 1]
    big = make_bank(N_C2, seed=2)
    lbm = LookBack(big, W=W_C2, Ts=TS, K=10, refine=0)
    ts = np.arange(0, W_C2)
    lbm.load_window(S[:, ts].T, U[:, ts].T, S[:, ts + 1].T)
    lbm.evaluate()
    keep = np.argsort(lbm.avg_errors(), kind="stable")[:M]          # the 16,384 best-adapted models
    la = LookAhead({k: (v[keep] if np.ndim(v) else v) for k, v in big.items()}, Ts=TS)
    del lbm
    plan = la.plan(S[:, t0], Useq, xref, U[:, t0 - 1])
    dt_flushed = time_it(plan.run, 20)
    # sustained, the headline's protocol: every call reads its own copy of the 1 MiB model bank (never read before, 135 MiB in
    # all against 126 MB of L2), calls back to back between one pair of events, programmatic dependent launch
    reps = int(np.ceil(1.35 * L2_BYTES / (64.0 * la.bank.Npad)))
    copies = [la.bank.packed.clone() for _ in range(reps + 3)]
    flush.add_(1)
    use_pdl = os.environ.get("LLAMPC_BENCH_PDL", "1") == "1"
    for c in copies[:3]:
        plan.run(bank_ptr=c.data_ptr(), pdl=use_pdl)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for c in copies[3:]:
        plan.run(bank_ptr=c.data_ptr(), pdl=use_pdl)
    eb.record()
    torch.cuda.synchronize()
    dt = ea.elapsed_time(eb) * 1e-3 / reps
    del copies
    out["C3_lookahead_16384x32x20"] = {"steps_per_s": M * K * H / dt, "ms_per_call": dt * 1e3, "roofline_frac": M * K * H / dt / peak_rate,
                                       "protocol": "one bank copy per call, %d calls back to back, %s" % (reps, "programmatic dependent launch" if use_pdl else "stream order"),
                                       "ms_per_call_l2_flushed": dt_flushed * 1e3, "roofline_frac_l2_flushed": M * K * H / dt_flushed / peak_rate}
    # end to end: LookAhead.rollout with host arrays in (x0, U, xref, uprev: 5.3 KB) and J (M, K) + best_k (M) back on the host
    lat = []
    for _ in range(8):
        a0 = time.perf_counter()
        J, bk = la.rollout(S[:, t0], Useq, xref, U[:, t0 - 1])
        lat.append(time.perf_counter() - a0)
    p50 = float(np.percentile(lat[2:], 50))
    out["C3_lookahead_16384x32x20"].update({"e2e_rollout_p50_ms": p50 * 1e3, "e2e_steps_per_s": M * K * H / p50,
                                            "e2e_h2d_bytes": int(Useq.size * 4 + xref.size * 4 + 6 * 8 + 2 * 4),
                                            "e2e_d2h_bytes": int(M * K * 4 + M * 4)})
    del la, plan

    # C4 Monte-Carlo closed loop: 4,096 vehicles x (look-back 1,024 candidates x 20 window + look-ahead 32 x 20 + planner,
    # plant, friction estimate) per tick, everything device-resident
    try:
        from llampc_b200.mpc.montecarlo import MonteCarlo
        from llampc_b200.tracks import RacelineTable
        rl = np.load(os.path.join(ROOT, "tests", "golden", "raceline_ethzmobil.npz"))
        tab = RacelineTable(rl["x"], rl["y"], rl["speeds"], rl["mus"])
        Vn = 4096
        r4 = np.random.RandomState(4)
        start = r4.randint(0, 400, Vn)
        x_init = np.zeros((Vn, 6))
        x_init[:, 0] = 0.6 * rl["x"][start + 1] + 0.4 * rl["x"][start + 2]
        x_init[:, 1] = 0.6 * rl["y"][start + 1] + 0.4 * rl["y"][start + 2]
        x_init[:, 2] = np.arctan2(rl["y"][start + 2] - rl["y"][start + 1], rl["x"][start + 2] - rl["x"][start + 1])
        x_init[:, 3] = 1.0
        for mode in ("rolling", "recompute"):
            mc = MonteCarlo(make_bank_rt(1024, seed=0), tab, x_init, start, NOMINAL, r4.uniform(3.0, 15.0, Vn), W=20,
                            K_models=10, K_seq=32, H=20, Ts=TS, seed=4, lookback_mode=mode, use_graphs=True)
            mc.run(48)                                           # fill the windows, capture the per-slot tick graphs
            torch.cuda.synchronize()
            n_t = 20
            s0_lb, s0_la = mc.lookback_steps, mc.lookahead_steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            mc.run(n_t)
            b.record()
            torch.cuda.synchronize()
            dt = a.elapsed_time(b) * 1e-3 / n_t
            steps = (mc.lookback_steps - s0_lb + mc.lookahead_steps - s0_la) / n_t
            # the look-back launch of the tick alone (L2 flushed): the kernel the C4 roofline fraction is quoted on
            lb_dt = time_it((lambda: mc.lb.launch(slot=3, emit=1)) if mode == "rolling" else mc.lb.launch, 10)
            lb_steps = Vn * 1024 * (1 if mode == "rolling" else 20)
            out["C4_montecarlo_4096veh_" + mode] = {"steps_per_s": steps / dt, "ms_per_tick": dt * 1e3,
                                                     "rk4_steps_per_tick": steps, "vehicle_ticks_per_s": Vn / dt,
                                                     "roofline_frac_tick": steps / dt / peak_rate,
                                                     "lookback_kernel": mc.lb.kernel_name, "lookback_sine": mc.lb.sine_name,
                                                     "lookback_launch_us": lb_dt * 1e6,
                                                     "lookback_launch_steps_per_s": lb_steps / lb_dt,
                                                     "lookback_launch_roofline_frac": lb_steps / lb_dt / peak_rate}
            del mc
    except Exception as e:                                       # the secondary configs never block the main line
        out["C4_montecarlo_4096veh"] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (C1, C3, C5 on one GPU)")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): anything a library prints there meanwhile (the NCCL version banner of the
    # first collective, for one) is sent to stderr at the file-descriptor level until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, emit)
    else:
        run_b200(args, emit)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)


if __name__ == "__main__":
    main()
