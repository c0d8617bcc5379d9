#!/usr/bin/env python
"""bench.py -- benchmarks of the LUDVM vortex-velocity hot path on B200.

    python bench.py [--gpus G] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload selfconv|flowfield|sweep] [--nvortices N] [--no-extra-legs]

Workloads (BASELINE.json `configs`):
  selfconv   configs[2], the headline: synthetic all-pairs self-convection of N = 2^20 Vatistas vortices, seed 20260101,
             x~U(-20,0), z~U(-4,4), Gamma~N(0,1)*1e-2, v_core = 0.065, dt = 0.05; one "step" = one forward-Euler
             self-convection step = N^2 pair interactions (LUDVM.py:549-570 inside LUDVM.py:1095-1127).  With --gpus G the
             target rows are sharded over G ranks (strong scaling: the problem stays N) and the updated positions are
             exchanged each step (peer stores fused into the kernel epilogue, or NCCL all-gather).
  flowfield  configs[4]: velocity + vorticity on a 4096 x 4096 grid induced by 200 000 vortices (LUDVM.py:1186-1298),
             grid x-rows sharded over the ranks with one halo row per side, no collective.
  sweep      configs[3]: 4096 independent README-size LUDVM cases (LESPcrit 0.1-0.4 x reduced frequency 0.1-1.0), one
             CTA per case, cases split over the ranks, no collective.
The default (selfconv) line also carries one-step `flowfield` and `sweep` legs unless --no-extra-legs.

Prints ONE JSON line (rank 0).  `value` = the metric with inputs resident in HBM; `e2e` = the same metric through the
public host-buffer API (H2D/D2H inside the timed region); `parity` = a check of the timed code against the CPU oracle
made in the same run.  `--impl reference` times the CPU oracle port (oracle/ludvm_oracle.c, all host threads, thread
count passed explicitly) on a bounded sample of the same workload -- the reference itself is a Python/numpy file that
does not travel to the GPU box.
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
sys.path.insert(0, ROOT)

UNIT = "pair-interactions/s"
METRICS = {"selfconv": ("biot_savart_pair_interactions_per_s_fp64", UNIT),
           "flowfield": ("flowfield_pair_interactions_per_s_fp64", UNIT),
           "sweep": ("ludvm_sweep_case_timesteps_per_s", "case-timesteps/s")}
SEED, VCORE, DT = 20260101, 0.065, 0.05
FF_SEED, FF_NSRC, FF_DR = 20260102, 200000, 0.005
SLOTS_PER_PAIR = 13          # 7 DFMA + 4 DMUL + 2 DADD FP64-pipe issue slots per pair (SASS-counted)
FLOP_PER_PAIR = 20           # N-body convention (FMA = 2)
EXACT_OPS_PER_PAIR = 31      # exact mode: 7 (r^4 + vc^4) + 8 (sqrt) + 1 (2 pi) + 5 (reciprocal) + 6 (two quotients) + 2 (x Gamma) + 2 (sums)
NOMINAL_DFMA_PER_S = 148 * 64 * 1.965e9      # 148 SMs x 64 FP64 lanes x 1.965 GHz (clocks.max.sm)
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused kernel at N = 2^20 on one GPU, from the
# `ncu --set full` capture under profiles/ (see profiles/README.md).  Algorithmic: 32 B read + 16 B written per vortex =
# 50.3 MB; the capture sees the 25 MB of sources read once and almost none of the 16.8 MB of results, which are still
# dirty in the 126 MB L2 when the kernel ends.
NCU_DRAM_BYTES_PER_LAUNCH = {"fast_fused": 25.52e6,       # profiles/r02g_k_fast_fused_raw.csv: 25.34 MB read + 0.18 MB written
                             "fast_tiled_tma": None}      # (round-1 partial-sum path: 41.7 MB with 8 chunks, not re-captured with 16)
README = dict(t0=0, tf=20, dt=5e-2, chord=1, rho=1.225, Uinf=1, Npoints=81, Ncoeffs=30, LESPcrit=0.2, Naca="0012")
README_PAIRS_PER_CASE = 8.262e7   # pair evaluations of one README run (SURVEY.md / BASELINE.md, cProfile of the reference)

try:
    HOST_THREADS = len(os.sched_getaffinity(0))
except AttributeError:
    HOST_THREADS = os.cpu_count() or 1


def make_cloud(n, seed=SEED):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-20, 0, n)
    z = rng.uniform(-4, 4, n)
    g = rng.standard_normal(n) * 1e-2
    return g, x, z


def ff_axes():
    x1, z1 = np.arange(-20.48, 0, FF_DR), np.arange(-10.24, 10.24, FF_DR)
    assert len(x1) == 4096 and len(z1) == 4096, (len(x1), len(z1))
    return x1, z1


def sweep_cases():
    from ludvm_b200 import sweep
    return sweep.lespcrit_k_grid(np.linspace(0.1, 0.4, 64), np.linspace(0.1, 1.0, 64), **README)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle as the timed CPU baseline -- the one place besides tests/ and smoke() that may execute oracle/)
# ---------------------------------------------------------------------------------------------------------------
def numpy_reference_rate(g, x, z, seconds=3.0):
    """The reference's own formulation (numpy temporaries, one thread; LUDVM.py:555-569) on 32-row chunks."""
    vc4, n, rows, done, t0 = VCORE ** 4, len(x), 32, 0, time.perf_counter()
    i = 0
    while time.perf_counter() - t0 < seconds:
        xd = x[i:i + rows, None] - x[None, :]
        zd = z[i:i + rows, None] - z[None, :]
        den = 2 * np.pi * np.sqrt((xd ** 2 + zd ** 2) ** 2 + vc4)
        u = np.sum(g * (zd / den), axis=1)
        w = np.sum(-(g * (xd / den)), axis=1)
        done += rows * n
        i = (i + rows) % (n - rows)
    return done / (time.perf_counter() - t0), float(u[0] + w[0])


def calibrate_rows(oracle, g, x, z, xp, zp, target_s, threads):
    """Targets of the N-source problem the CPU port evaluates in ~target_s seconds (thread pool warmed first)."""
    n, rows = len(xp), 256
    for _ in range(2):   # the first parallel regions run far below steady state (thread pool / scheduler warm-up)
        oracle.induced_velocity(g, x, z, xp[:min(n, 2048)], zp[:min(n, 2048)], VCORE, nthreads=threads)
    while True:
        t = time.perf_counter()
        oracle.induced_velocity(g, x, z, xp[:rows], zp[:rows], VCORE, nthreads=threads)
        dt = max(time.perf_counter() - t, 1e-4)
        if dt > 0.3 or rows >= n:
            break
        rows = min(n, rows * 4)
    return int(min(n, max(rows, rows * target_s / dt)))


def cpu_oracle_rate(g, x, z, target_s, threads=HOST_THREADS):
    """oracle/ludvm_oracle.c with `threads` host threads on a bounded row sample; returns (pairs/s, rows, threads)."""
    from oracle import ludvm_oracle as oracle
    n = len(x)
    rows = calibrate_rows(oracle, g, x, z, x, z, target_s, threads)
    t = time.perf_counter()
    oracle.induced_velocity(g, x, z, x[:rows], z[:rows], VCORE, nthreads=threads)
    dt = time.perf_counter() - t
    return rows * n / dt, rows, threads


def oracle_cases_parallel(cases, threads):
    """Run OracleLUDVM (single-threaded C time loop; ctypes releases the GIL) for every case on a thread pool."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ludvm_oracle as oracle
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda kw: oracle.OracleLUDVM(**kw), cases))


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ludvm_oracle as oracle
    oracle.build()
    metric, unit = METRICS[args.workload]
    threads = HOST_THREADS
    if args.workload == "sweep":
        cases = sweep_cases()
        batch = threads                                    # a step = one case per host thread (400 time steps each)
        pick = lambda k: [cases[(k * batch + j) * 37 % len(cases)] for j in range(batch)]  # noqa: E731
        for k in range(args.warmup):
            oracle_cases_parallel(pick(k), threads)
        t0 = time.perf_counter()
        for k in range(args.steps):
            oracle_cases_parallel(pick(args.warmup + k), threads)
        el = time.perf_counter() - t0
        value = batch * 400.0 * args.steps / el
        sample = "%d of the 4096 cases per step (one per host thread), 400 time steps each, scalar C port" % batch
        workload = "configs[3]: 4096-case LESPcrit x k sweep of the README case"
    else:
        if args.workload == "flowfield":
            g, x, z = make_cloud(FF_NSRC, FF_SEED)
            x1, z1 = ff_axes()
            pts = np.random.default_rng(1).choice(len(x1) * len(z1), 1 << 18, replace=False)
            xp, zp = x1[pts // len(z1)], z1[pts % len(z1)]
            workload = "configs[4]: flow field of %d vortices on a 4096x4096 grid" % FF_NSRC
        else:
            g, x, z = make_cloud(args.n)
            xp, zp = x, z
            workload = "configs[2]: synthetic all-pairs self-convection, N=%d vortices" % args.n
        n = len(x)
        rows = calibrate_rows(oracle, g, x, z, xp, zp, args.ref_step_s, threads)   # a step = a sample worth ~ref_step_s s
        for _ in range(args.warmup):
            oracle.induced_velocity(g, x, z, xp[:rows], zp[:rows], VCORE, nthreads=threads)
        t0 = time.perf_counter()
        for k in range(args.steps):
            r0 = (k * rows) % max(1, len(xp) - rows)
            oracle.induced_velocity(g, x, z, xp[r0:r0 + rows], zp[r0:r0 + rows], VCORE, nthreads=threads)
        el = time.perf_counter() - t0
        value = rows * n * args.steps / el
        sample = "%d targets x %d sources per step (%.3g pairs), scalar C port with OpenMP, %d threads" % (
            rows, n, rows * n, threads)
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "note": "CPU baseline on a bounded sample of the same workload"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port", "sample": sample,
                         "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------------------
# GPU legs
# ---------------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide set-up shared by the legs: device, context on torch's current stream, rank plumbing."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from ludvm_b200 import _lib
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; ludvm_b200 has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = _lib.Context(self.local, torch.cuda.current_stream().cuda_stream)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.oracle_threads = max(1, HOST_THREADS // self.world)     # every rank checks its own rows at the same time

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v, op="max"):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN,
                                        "sum": self.dist.ReduceOp.SUM}[op])
        return float(t.item())

    def time_events(self, fn, steps, warmup):
        """`warmup` untimed + `steps` timed calls of fn with an L2 flush before each; returns the summed device
        milliseconds, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        tot = 0.0
        for _ in range(steps):
            self.flush.fill_(1)                              # evict L2 between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            tot += e0.elapsed_time(e1)
        self.barrier()
        return self.reduce(tot, "max")

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def rel_err(a, ref):
    return float(np.max(np.abs(a - ref)) / np.max(np.abs(ref)))


def selfconv_parity(env, g_h, x_h, z_h, g, x, z, n, transport):
    """The timed code against the oracle: one step of this rank's shard from the initial cloud, 512 sampled rows of the
    shard (u and w, 1e-12 of max |ref|), the Euler update of every row bitwise; with G > 1 also a small sharded run
    against the same run unsharded on this rank, bitwise."""
    from ludvm_b200 import ops
    from ludvm_b200.sharded import ShardedSelfConvection
    from oracle import ludvm_oracle as oracle
    torch = env.torch
    shard = n // env.world
    row0 = env.rank * shard
    xo, zo = torch.empty_like(x), torch.empty_like(z)
    u, w = (torch.empty(shard, dtype=torch.float64, device=env.dev) for _ in range(2))
    ops.selfconv_step(env.ctx, "fast", g, x, z, VCORE ** 4, DT, xo, zo, row0=row0, nrows=shard, u_out=u, w_out=w)
    torch.cuda.synchronize()
    plan = env.ctx.last_plan()
    rows = np.sort(np.random.default_rng(7 + env.rank).choice(shard, min(512, shard), replace=False)) + row0
    uo, wo = oracle.induced_velocity(g_h, x_h, z_h, x_h[rows], z_h[rows], VCORE, nthreads=env.oracle_threads)
    u_h, w_h = u.cpu().numpy(), w.cpu().numpy()
    eu, ew = rel_err(u_h[rows - row0], uo), rel_err(w_h[rows - row0], wo)
    sl = slice(row0, row0 + shard)
    euler = bool(np.array_equal(xo[sl].cpu().numpy(), x_h[sl] + DT * u_h) and
                 np.array_equal(zo[sl].cpu().numpy(), z_h[sl] + DT * w_h))
    sharded_ok = None
    if env.world > 1:
        m = 1 << 17
        sc = ShardedSelfConvection(g[:m].clone(), x[:m].clone(), z[:m].clone(), VCORE, DT, mode="fast", ctx=env.ctx,
                                   transport=transport)
        for _ in range(2):
            xs, zs = sc.step()
        xa, za = x[:m].clone(), z[:m].clone()
        xb, zb = torch.empty_like(xa), torch.empty_like(za)
        gm = g[:m].clone()
        for _ in range(2):
            ops.selfconv_step(env.ctx, "fast", gm, xa, za, VCORE ** 4, DT, xb, zb)
            xa, xb, za, zb = xb, xa, zb, za
        torch.cuda.synchronize()
        sharded_ok = bool(torch.equal(xs, xa) and torch.equal(zs, za))
        sharded_ok = env.reduce(1.0 if sharded_ok else 0.0, "min") == 1.0
    eu, ew = env.reduce(eu), env.reduce(ew)
    euler = env.reduce(1.0 if euler else 0.0, "min") == 1.0
    ok = eu <= 1e-12 and ew <= 1e-12 and euler and sharded_ok is not False
    return {"ok": bool(ok), "checked": "ludvm_selfconv_step, fast mode, this run's launch configuration",
            "rows_checked": int(len(rows)) * env.world, "max_err_u_over_max_ref": eu, "max_err_w_over_max_ref": ew,
            "tolerance": 1e-12, "euler_update_bitwise": euler,
            "sharded_2steps_n131072_equals_unsharded_bitwise": sharded_ok, "plan": plan}


def selfconv_leg(env, args):
    from ludvm_b200 import ops
    from ludvm_b200.sharded import ShardedSelfConvection
    torch, world, rank, ctx = env.torch, env.world, env.rank, env.ctx
    n = args.n
    assert n % world == 0, "N must divide evenly over the ranks"
    shard, row0 = n // world, rank * (n // world)
    g_h, x_h, z_h = make_cloud(n)
    g, x, z = (torch.tensor(a, device=env.dev) for a in (g_h, x_h, z_h))
    info = {}

    def timed_steps(mode, steps, warmup):
        sc = ShardedSelfConvection(g, x.clone(), z.clone(), VCORE, DT, mode=mode, ctx=ctx, transport=args.transport)
        info["transport"] = sc.transport
        l0 = [0]

        def warm_done():
            l0[0] = ctx.launch_count()
        for _ in range(warmup):
            sc.step()
        warm_done()
        ms = env.time_events(sc.step, steps, 0)              # kernel + fused peer-store all-gather (or NCCL) if G > 1
        info["plan_" + mode] = ctx.last_plan()
        return ms, ctx.launch_count() - l0[0]

    sampler = ClockSampler(env.local)
    sampler.start()
    total_ms, launches = timed_steps("fast", args.steps, args.warmup)
    clocks = sampler.stop()
    ms_per_step = total_ms / args.steps
    value = float(n) * n / (ms_per_step * 1e-3)
    f32_steps = max(1, min(args.steps, 2))
    f32_ms, _ = timed_steps("fp32", f32_steps, 1)            # fp32-fast, reported separately
    f32_value = float(n) * n / (f32_ms / f32_steps * 1e-3)
    f12_ms, _ = timed_steps("fast12", f32_steps, 1)          # opt-in 12-slot pair (4.3e-13 per pair), reported separately
    f12_value = float(n) * n / (f12_ms / f32_steps * 1e-3)
    ex_ms, _ = timed_steps("exact", 1, 1)                    # exact mode, reported separately
    ex_value = float(n) * n / (ex_ms * 1e-3)

    # e2e through the public host-buffer API: host arrays in, host arrays out (pinned and pageable)
    xs_p = {}
    e2e = {}
    for kind in ("pinned", "pageable"):
        conv = (lambda a: torch.tensor(a).pin_memory().numpy()) if kind == "pinned" else np.array   # noqa: E731
        gp_, xp_, zp_ = conv(g_h), conv(x_h), conv(z_h)
        xs, zs = xp_[row0:row0 + shard], zp_[row0:row0 + shard]
        ops.induced_velocity(gp_, xp_, zp_, xs, zs, VCORE, mode="fast", ctx=ctx)   # one untimed call: staging buffers
        env.barrier()
        e2e_steps = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            u_h, w_h = ops.induced_velocity(gp_, xp_, zp_, xs, zs, VCORE, mode="fast", ctx=ctx)
            xs_p[kind] = xs + DT * u_h                       # the host reads the step's result
        e2e[kind] = float(n) * n * e2e_steps / env.reduce(time.perf_counter() - t0)

    parity = selfconv_parity(env, g_h, x_h, z_h, g, x, z, n, args.transport)
    dfma = ctx.fp64_fma_rate(300.0)
    ffma = ctx.fp32_fma_rate(200.0)
    per_gpu = value / world
    plan = info["plan_fast"]
    tr = info["transport"]
    out = {
        "ms_per_step": ms_per_step, "value": value, "clocks": clocks, "gpu_launches": int(launches),
        "config": {"workload": "configs[2]: synthetic all-pairs self-convection of N=%d Vatistas vortices, target rows "
                               "sharded over %d GPU(s), positions exchanged each step (%s)"
                               % (n, world, {"p2p": "fused into the kernel epilogue as NVLink peer stores + "
                                                    "symmetric-memory barrier", "nccl": "NCCL all_gather_into_tensor",
                                             "none": "single GPU: no exchange"}[tr]),
                   "n_vortices": n, "pairs_per_step": float(n) * n, "mode": "fast_f64 (FMA + MUFU.RSQ64H rsqrt)",
                   "kernel": plan, "l2": "256 MiB buffer written between timed iterations (inputs are 24 MiB < L2)",
                   "parallelism": "row-shard x%d" % world, "transport": tr},
        "e2e": {"value": e2e["pinned"], "unit": UNIT, "h2d_bytes_per_step": 8 * (3 * n + 2 * shard),
                "d2h_bytes_per_step": 8 * 2 * shard, "pageable_host_value": e2e["pageable"],
                "api": "ludvm_b200.ops.induced_velocity(host numpy buffers) = C ABI ludvm_induced_velocity(PTR_HOST); "
                       "value: torch-pinned host arrays, pageable_host_value: plain numpy arrays"},
        "parity": parity,
        "roofline": {
            "bound": "fp64_fma_pipe", "unit": "TFLOP/s",
            "achieved": per_gpu * SLOTS_PER_PAIR * 2 / 1e12, "peak": dfma * 2 / 1e12,
            "frac": per_gpu * SLOTS_PER_PAIR / dfma,
            "peak_nominal": NOMINAL_DFMA_PER_S * 2 / 1e12, "frac_of_nominal": per_gpu * SLOTS_PER_PAIR / NOMINAL_DFMA_PER_S,
            "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(plan["kernel"]) if (world == 1 and n == (1 << 20)) else None,
            "algorithmic_bytes_per_launch": 48.0 * n / world,
            "note": "per GPU; achieved = pairs/s x 13 FP64-pipe issue slots x 2 flop; peak = DFMA issue rate measured "
                    "live by ludvm_measure_fp64_fma_rate (MEASURED_PEAKS.json has no FP64 entry), peak_nominal = 148 SM "
                    "x 64 lanes x 1.965 GHz; one launch per step (no partial sums in global memory, no combine "
                    "kernel); traffic = dram bytes read+written per launch from the ncu --set full capture of this "
                    "launch under profiles/ (N=2^20, 1 GPU)",
            "flop20_tflops": per_gpu * FLOP_PER_PAIR / 1e12,
            "hbm_algorithmic_gbs": 48.0 * n / world / (ms_per_step * 1e-3) / 1e9, "mufu_per_s": per_gpu},
        "fp32_fast": {"value": f32_value, "unit": UNIT, "ffma_per_s_measured": ffma, "kernel": info["plan_fp32"],
                      "note": "fp32 pair arithmetic, fp64 accumulation across tiles; accuracy ~1e-5 relative"},
        "fast12_f64": {"value": f12_value, "unit": UNIT, "fp64_slots_per_pair": 12, "kernel": info["plan_fast12"],
                       "frac_of_dfma_rate": f12_value / world * 12 / dfma,
                       "note": "opt-in mode LUDVM_FAST12_F64: one second-order refinement of the centred MUFU.RSQ64H seed, "
                               "|error| <= 4.3e-13 per pair (inside the 1e-12 statement) instead of 2.7e-16; not the headline"},
        "exact_f64": {"value": ex_value, "unit": UNIT, "ms_per_step": ex_ms, "fp64_ops_per_pair": EXACT_OPS_PER_PAIR,
                      "frac_of_dfma_rate": ex_value / world * EXACT_OPS_PER_PAIR / dfma,
                      "frac_of_nominal": ex_value / world * EXACT_OPS_PER_PAIR / NOMINAL_DFMA_PER_S,
                      "kernel": info["plan_exact"],
                      "note": "bitwise equal to the reference's numpy result (pairwise summation tree, correctly rounded "
                              "division and square root); 31 FP64-pipe operations + 2 MUFU per pair, SASS-counted"},
    }
    if world == 1 and rank == 0:
        cpu_v, rows, cores = cpu_oracle_rate(g_h, x_h, z_h, args.cpu_seconds)
        np_v, _ = numpy_reference_rate(g_h, x_h, z_h, 3.0)
        out["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": "%d target rows x %d sources (%.3g pairs) through oracle/ludvm_oracle.c, "
                                         "OpenMP over rows, %d threads" % (rows, n, float(rows) * n, cores),
                               "numpy_1core_value": np_v,
                               "numpy_1core_note": "the reference's own numpy formulation (LUDVM.py:555-569) on 32-row "
                                                   "chunks, 1 thread, 3 s sample"}
        out["timesteps_per_s"] = timesteps_leg(env)
    return out


def tree_leg(env, n, order, steps, warmup, rows_checked, all_pairs_ms=None):
    """Self-convection step through the O(N log N) treecode (csrc/tree.cu): target rows sharded over the
    ranks like the all-pairs step, every rank builds the same tree, NCCL all-gather of the positions.  The reference has
    no treecode: parity is the error of sampled rows against the oracle's ALL-PAIRS sum, relative to sum_j |term_ij|
    (SURVEY.md 8(d)-3), and the Euler update bitwise."""
    from ludvm_b200 import ops
    from ludvm_b200.sharded import ShardedSelfConvection
    from oracle import ludvm_oracle as oracle
    torch, world, rank, ctx = env.torch, env.world, env.rank, env.ctx
    from ludvm_b200.sharded import morton_order
    shard, row0 = n // world, rank * (n // world)
    g_h, x_h, z_h = make_cloud(n)
    perm = morton_order(x_h, z_h)          # the same cloud relabelled along the Z-order curve: row shards compact in space
    g_h, x_h, z_h = (np.ascontiguousarray(a[perm]) for a in (g_h, x_h, z_h))
    g, x, z = (torch.tensor(a, device=env.dev) for a in (g_h, x_h, z_h))
    sc = ShardedSelfConvection(g, x.clone(), z.clone(), VCORE, DT, mode="tree", ctx=ctx, transport="nccl", order=order)
    for _ in range(warmup):
        sc.step()
    l0 = ctx.launch_count()
    ms = env.time_events(sc.step, steps, 0) / steps
    launches = ctx.launch_count() - l0
    # parity of one step from the initial cloud: this rank's shard, sampled rows against the oracle
    xo, zo = torch.empty_like(x), torch.empty_like(z)
    u, w = (torch.empty(shard, dtype=torch.float64, device=env.dev) for _ in range(2))
    st = ops.selfconv_step_tree(ctx, g, x, z, VCORE ** 4, DT, xo, zo, row0=row0, nrows=shard, u_out=u, w_out=w, order=order,
                                return_stats=True)
    torch.cuda.synchronize()
    rows = np.sort(np.random.default_rng(11 + rank).choice(shard, max(1, rows_checked // world), replace=False)) + row0
    uo, wo = oracle.induced_velocity(g_h, x_h, z_h, x_h[rows], z_h[rows], VCORE, nthreads=env.oracle_threads)
    den = np.empty(len(rows))
    ga = g.abs()
    for k0 in range(0, len(rows), 8):                              # sum_j |term_ij| of the checked rows (on the device)
        rr = torch.as_tensor(rows[k0:k0 + 8], device=env.dev)
        dx, dz = x[rr, None] - x[None, :], z[rr, None] - z[None, :]
        r2 = dx * dx + dz * dz
        den[k0:k0 + 8] = ((ga[None, :] * r2.sqrt() / (r2 * r2 + VCORE ** 4).sqrt()).sum(1) / (2 * np.pi)).cpu().numpy()
        del dx, dz, r2
    u_h, w_h = u.cpu().numpy(), w.cpu().numpy()
    err = float(np.max(np.hypot(u_h[rows - row0] - uo, w_h[rows - row0] - wo) / den))
    sl = slice(row0, row0 + shard)
    euler = bool(np.array_equal(xo[sl].cpu().numpy(), x_h[sl] + DT * u_h) and np.array_equal(zo[sl].cpu().numpy(), z_h[sl] + DT * w_h))
    err = env.reduce(err)
    euler = env.reduce(1.0 if euler else 0.0, "min") == 1.0
    evals = env.reduce(st["pair_evaluations"], "sum")
    out = {"metric": "fmm_selfconv_ms_per_step", "n_vortices": n, "order": order, "ms_per_step": ms, "steps": steps,
           "gpu_launches": int(launches), "leaf_level": st["leaf_level"], "proxies_per_cell": st["proxies_per_cell"],
           "pair_evaluations_per_step": evals, "pairs_left_frac": evals / (float(n) * n),
           "eval_pairs_per_s": evals / (ms * 1e-3), "all_pairs_equivalent_pairs_per_s": float(n) * n / (ms * 1e-3),
           "frac_of_dfma_rate_whole_step": evals / (ms * 1e-3) / world * SLOTS_PER_PAIR / ctx.fp64_fma_rate(100.0),
           "build_ms_this_rank": st["build_ms"], "eval_ms_this_rank": st["eval_ms"], "arena_bytes": st["arena_bytes"],
           "parity": {"ok": bool(err <= 1e-12 and euler), "rows_checked": int(len(rows)) * world,
                      "max_err_over_sum_abs_terms": err, "tolerance": 1e-12, "euler_update_bitwise": euler,
                      "checker": "oracle all-pairs sum (the reference has no treecode)"},
           "note": "approximates the same all-pairs sum; not the headline (value stays the all-pairs fp64 kernel)"}
    if all_pairs_ms:
        out["all_pairs_ms_per_step"] = all_pairs_ms
        out["speedup_vs_all_pairs"] = all_pairs_ms / ms
    del sc, g, x, z, xo, zo
    torch.cuda.empty_cache()
    return out


def timesteps_leg(env):
    """LUDVM timesteps/s, README case (BASELINE.json configs[0]), end to end (table upload + result download)."""
    from ludvm_b200 import LUDVM
    from oracle import ludvm_oracle as oracle
    ts = {}
    for mode in ("exact", "fast"):
        s = LUDVM(**README, verbose=False, run=False, mode=mode, ctx=env.ctx, steps_per_graph=400)
        s.time_loop(); s.compute_coefficients()          # warm-up (graph capture)
        best = 1e9
        for _ in range(3):
            t = time.perf_counter(); s.time_loop(); s.compute_coefficients(); best = min(best, time.perf_counter() - t)
        ts[mode] = 400.0 / best
        if mode == "exact":
            cl = s.Cl.copy()
        s.close()
    t = time.perf_counter(); o = oracle.OracleLUDVM(**README, run=False); o.time_loop(); o.compute_coefficients()
    t_or = time.perf_counter() - t
    return {"workload": "configs[0]: README case, 400 steps, time_loop+compute_coefficients incl. table upload and "
                        "result download",
            "exact": ts["exact"], "fast": ts["fast"], "cpu_oracle_1core": 400.0 / t_or,
            "parity_exact_Cl_bitwise_vs_oracle": bool(np.array_equal(cl.view(np.uint64), o.Cl.view(np.uint64))),
            "reference_numpy_NOT_measured_here": 60.0,
            "reference_numpy_note": "59-64 steps/s is the unmodified reference timed on the development container's Xeon "
                                    "(BASELINE.md section 2), quoted for orientation only; the reference (a Python file) "
                                    "does not travel to the GPU box, the figure timed on this box is cpu_oracle_1core"}


def flowfield_leg(env, steps, warmup, with_e2e):
    """configs[4]: u, w and the vorticity stencil on the 4096 x 4096 grid from 200 000 vortices; this rank evaluates its
    x-row slab plus one halo row per interior side (sharded.grid_slab), no collective."""
    from ludvm_b200 import ops
    from ludvm_b200._lib import check, load, ptr, PTR_DEVICE
    from ludvm_b200.sharded import grid_slab
    from oracle import ludvm_oracle as oracle
    torch, ctx = env.torch, env.ctx
    g_h, xw_h, zw_h = make_cloud(FF_NSRC, FF_SEED)
    x1_h, z1_h = ff_axes()
    nx, nz = len(x1_h), len(z1_h)
    r0, r1, h0, h1 = grid_slab(nx, env.world, env.rank)
    g, xw, zw, x1, z1 = (torch.tensor(a, device=env.dev) for a in (g_h, xw_h, zw_h, x1_h, z1_h))
    ne = h1 - h0
    u, w, ome = (torch.empty((ne, nz), dtype=torch.float64, device=env.dev) for _ in range(3))
    vc4 = VCORE ** 4
    x1s = x1[h0:h1].contiguous()

    def step():
        ops.flowfield_velocity_device(ctx, "fast", g, xw, zw, vc4, x1, z1, h0, ne, u, w)
        check(load().ludvm_flowfield_vorticity(ctx.handle, ptr(x1s), ne, ptr(z1), nz, ptr(u), ptr(w), 1, ptr(ome),
                                               PTR_DEVICE))
    l0 = ctx.launch_count()
    ms = env.time_events(step, steps, warmup) / steps
    launches = (ctx.launch_count() - l0) // (steps + warmup) * steps
    plan = ctx.last_plan()
    pairs = float(nx) * nz * FF_NSRC
    value = pairs / (ms * 1e-3)
    # parity: sampled points of the owned slab against the oracle; the stencil on the slab's first 64 x 64 block
    rng = np.random.default_rng(11 + env.rank)
    ii, jj = rng.integers(r0, r1, 256), rng.integers(0, nz, 256)
    uo, wo = oracle.induced_velocity(g_h, xw_h, zw_h, x1_h[ii], z1_h[jj], VCORE, nthreads=env.oracle_threads)
    u_h, w_h = u.cpu().numpy(), w.cpu().numpy()
    eu, ew = rel_err(u_h[ii - h0, jj], uo), rel_err(w_h[ii - h0, jj], wo)
    blk = slice(0, min(64, ne))
    X, Z = np.meshgrid(x1_h[h0:h1][blk], z1_h[:64], indexing="ij")
    om_o = oracle.vorticity(X, Z, u_h[None, blk, :64], w_h[None, blk, :64])[0]
    om_h = ome.cpu().numpy()
    # interior of the block only: its last row/column are one-sided in the oracle's small block but centred in ours
    vort_ok = bool(np.array_equal(om_h[blk, :64][:-1, :-1][(1 if h0 else 0):], om_o[:-1, :-1][(1 if h0 else 0):]))
    eu, ew = env.reduce(eu), env.reduce(ew)
    vort_ok = env.reduce(1.0 if vort_ok else 0.0, "min") == 1.0
    out = {"metric": METRICS["flowfield"][0], "value": value, "unit": UNIT, "ms_per_step": ms, "steps": steps,
           "pairs_per_step": pairs, "gpu_launches": int(launches), "kernel": plan, "scaling": "strong",
           "config": {"workload": "configs[4]: velocity + vorticity on a %dx%d grid (dr=%g) from %d vortices, x-rows "
                                  "sharded over %d GPU(s) with one halo row per interior side, no collective"
                                  % (nx, nz, FF_DR, FF_NSRC, env.world), "mode": "fast_f64",
                      "rows_evaluated_this_rank": int(ne), "rows_owned_this_rank": int(r1 - r0)},
           "parity": {"ok": bool(eu <= 1e-12 and ew <= 1e-12 and vort_ok), "points_checked": 256 * env.world,
                      "max_err_u_over_max_ref": eu, "max_err_w_over_max_ref": ew, "tolerance": 1e-12,
                      "vorticity_block_bitwise_vs_oracle_stencil": vort_ok}}
    # the same slab through the hierarchical far field (csrc/tree.cu), reported separately: not pairs/s, an approximation
    # of the same sums whose error against the oracle is part of the figure
    ut, wt = torch.empty_like(u), torch.empty_like(w)
    dens = 1.0 / (FF_DR * FF_DR)

    def tree_step():
        ops.flowfield_velocity_tree_device(ctx, g, xw, zw, vc4, x1, z1, h0, ne, ut, wt, dens, order=18)
        check(load().ludvm_flowfield_vorticity(ctx.handle, ptr(x1s), ne, ptr(z1), nz, ptr(ut), ptr(wt), 1, ptr(ome), PTR_DEVICE))
    tms = env.time_events(tree_step, 2, 1) / 2
    st = ops.flowfield_velocity_tree_device(ctx, g, xw, zw, vc4, x1, z1, h0, ne, ut, wt, dens, order=18, return_stats=True)
    torch.cuda.synchronize()
    ut_h, wt_h = ut.cpu().numpy(), wt.cpu().numpy()
    den = np.empty(len(ii))
    ga = g.abs()
    for k0 in range(0, len(ii), 32):
        xs_, zs_ = x1[torch.as_tensor(ii[k0:k0 + 32], device=env.dev)], z1[torch.as_tensor(jj[k0:k0 + 32], device=env.dev)]
        dx, dz = xs_[:, None] - xw[None, :], zs_[:, None] - zw[None, :]
        r2 = dx * dx + dz * dz
        den[k0:k0 + 32] = ((ga[None, :] * r2.sqrt() / (r2 * r2 + vc4).sqrt()).sum(1) / (2 * np.pi)).cpu().numpy()
    terr = env.reduce(float(np.max(np.hypot(ut_h[ii - h0, jj] - uo, wt_h[ii - h0, jj] - wo) / den)))
    evals = env.reduce(st["pair_evaluations"], "sum")
    out["far_field"] = {"metric": "fmm_flowfield_ms_per_step", "order": 18, "ms_per_step": tms, "all_pairs_ms_per_step": ms,
                        "speedup_vs_all_pairs": ms / tms, "pair_evaluations_per_step": evals,
                        "pairs_left_frac": evals / pairs, "leaf_level": st["leaf_level"],
                        "parity": {"ok": bool(terr <= 1e-12), "points_checked": 256 * env.world,
                                   "max_err_over_sum_abs_terms": terr, "tolerance": 1e-12,
                                   "checker": "oracle all-pairs sum (the reference has no far-field method)"}}
    if with_e2e:   # public host API: host sources/axes in, host u/w out, then the stencil on host fields
        t0 = time.perf_counter()
        uh, wh, omh = ops.flowfield(g_h, xw_h, zw_h, None, None, None, vc4, x1_h, z1_h, row0=h0, nrows=ne, mode="fast",
                                    ctx=ctx)
        el = env.reduce(time.perf_counter() - t0)
        out["e2e"] = {"value": pairs / el, "unit": UNIT, "h2d_bytes_per_step": 8 * (3 * FF_NSRC + nx + nz),
                      "d2h_bytes_per_step": 8 * 3 * ne * nz, "checksum": float(omh[1, 1]),
                      "api": "ops.flowfield = C ABI ludvm_flowfield(PTR_HOST): host sources and axes in, host u, w, ome out"}
    return out


def sweep_leg(env, steps, warmup, modes=("fast", "exact")):
    """configs[3]: 4096 independent README-size cases, one CTA per case; this rank runs its case slice."""
    from ludvm_b200 import sweep
    from ludvm_b200.sharded import case_slice
    from oracle import ludvm_oracle as oracle
    cases = sweep_cases()
    sl = case_slice(len(cases), env.world, env.rank)
    res, out = {}, {}
    for mode in modes:
        for _ in range(warmup):
            sweep.run_sweep(cases, mode=mode, ctx=env.ctx, case_slice=sl)
        env.barrier()
        l0 = env.ctx.launch_count()
        t_dev = t_all = 0.0
        for _ in range(steps):
            t0 = time.perf_counter()
            r = sweep.run_sweep(cases, mode=mode, ctx=env.ctx, case_slice=sl)
            t_all += time.perf_counter() - t0
            t_dev += r["timing"]["ludvm_sweep_run_s"]
        res[mode] = r
        launches = env.ctx.launch_count() - l0
        t_dev, t_all = env.reduce(t_dev) / steps, env.reduce(t_all) / steps
        out[mode] = {"value": len(cases) * 400.0 / t_dev, "e2e_value": len(cases) * 400.0 / t_all,
                     "s_per_sweep": t_dev, "s_per_sweep_with_host_tables": t_all, "gpu_launches": int(launches),
                     "equiv_pairs_per_s_estimate": len(cases) * README_PAIRS_PER_CASE / t_dev}
    # parity: two cases of this rank's slice against the oracle, exact mode, bit for bit
    mine = list(range(len(cases)))[sl]
    pick = [mine[0], mine[len(mine) // 2 + 7 * (env.rank + 1) % max(1, len(mine) // 2)]]
    ok = True
    if "exact" in res:
        for c in pick:
            o = oracle.OracleLUDVM(**cases[c])
            a = c - mine[0]
            for k in ("Cl", "Cd", "Cm", "LESP", "LEV_shed"):
                ok = ok and bool(np.array_equal(res["exact"][k][a].view(np.uint64), getattr(o, k).view(np.uint64)))
    ok = env.reduce(1.0 if ok else 0.0, "min") == 1.0
    head = out[modes[0]]
    return {"metric": METRICS["sweep"][0], "value": head["value"], "unit": METRICS["sweep"][1], "mode": modes[0],
            "ms_per_step": 1e3 * head["s_per_sweep"], "steps": steps, "gpu_launches": head["gpu_launches"],
            "scaling": "strong", "modes": out,
            "config": {"workload": "configs[3]: 4096 LUDVM cases (LESPcrit 0.1-0.4 x k 0.1-1.0, README otherwise), 400 "
                                   "time steps each, one CTA per case, cases split over %d GPU(s), no collective"
                                   % env.world, "cases_this_rank": len(mine),
                       "timed": "ludvm_sweep_run: parameter/table upload, one persistent launch, history download"},
            "e2e": {"value": head["e2e_value"], "unit": METRICS["sweep"][1],
                    "h2d_bytes_per_step": None, "d2h_bytes_per_step": 8 * 12 * 401 * len(mine),
                    "api": "ludvm_b200.sweep.run_sweep (host kwargs in, numpy histories out)"},
            "parity": {"ok": ok, "cases_checked": 2 * env.world if "exact" in res else 0,
                       "what": "Cl, Cd, Cm, LESP, LEV_shed of sampled cases bit-equal to the oracle (exact mode)"},
            "roofline": {"bound": "latency / instruction issue (one CTA per case, wakes <= 604 vortices)",
                         "equiv_pairs_per_s_estimate": head["equiv_pairs_per_s_estimate"],
                         "note": "estimate = cases x 8.262e7 pair evaluations of the README run / time"}}


def run_ours(args):
    env = Env()
    world, rank = env.world, env.rank
    metric, unit = METRICS[args.workload]
    base = {"metric": metric, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    if args.workload == "selfconv":
        leg = selfconv_leg(env, args)
        out = dict(base, **leg)
        if not args.no_extra_legs:
            out["flowfield"] = flowfield_leg(env, 1, 1, with_e2e=False)
            out["sweep"] = sweep_leg(env, 1, 1)
            out["tree"] = tree_leg(env, args.n, 18, 3, 2, 512, all_pairs_ms=leg["ms_per_step"])
            out["tree"]["order14"] = tree_leg(env, args.n, 14, 3, 2, 512, all_pairs_ms=leg["ms_per_step"])
            out["tree"]["n_2p24"] = tree_leg(env, 1 << 24, 18, 2, 1, 128)
    elif args.workload == "flowfield":
        sampler = ClockSampler(env.local)
        sampler.start()
        leg = flowfield_leg(env, args.steps, args.warmup, with_e2e=True)
        clocks = sampler.stop()
        dfma = env.ctx.fp64_fma_rate(300.0)
        per_gpu = leg["value"] / world
        out = dict(base, **leg)
        out["clocks"] = clocks
        out["roofline"] = {"bound": "fp64_fma_pipe", "unit": "TFLOP/s", "achieved": per_gpu * SLOTS_PER_PAIR * 2 / 1e12,
                           "peak": dfma * 2 / 1e12, "frac": per_gpu * SLOTS_PER_PAIR / dfma,
                           "frac_of_nominal": per_gpu * SLOTS_PER_PAIR / NOMINAL_DFMA_PER_S, "traffic": None,
                           "algorithmic_bytes_per_launch": 24.0 * 4096 * 4096 / world + 24.0 * FF_NSRC,
                           "note": "halo rows are evaluated but not counted as work"}
        if world == 1:
            g_h, xw_h, zw_h = make_cloud(FF_NSRC, FF_SEED)
            cpu_v, rows, cores = cpu_oracle_rate(g_h, xw_h, zw_h, min(args.cpu_seconds, 6.0))
            out["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": "%d targets x %d sources through oracle/ludvm_oracle.c" % (rows, FF_NSRC)}
    else:
        sampler = ClockSampler(env.local)
        sampler.start()
        leg = sweep_leg(env, args.steps, args.warmup)
        out = dict(base, **leg)
        out["clocks"] = sampler.stop()
        if world == 1:
            cases = sweep_cases()
            t0 = time.perf_counter()
            oracle_cases_parallel([cases[j * 37 % len(cases)] for j in range(HOST_THREADS)], HOST_THREADS)
            el = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": HOST_THREADS * 400.0 / el, "unit": unit, "cores": HOST_THREADS, "kind": "port",
                                   "sample": "%d cases (one per host thread), 400 steps each, scalar C port" % HOST_THREADS}
    if rank == 0:
        print(json.dumps(out))
    env.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="selfconv", choices=["selfconv", "flowfield", "sweep"])
    ap.add_argument("--n", "--nvortices", dest="n", type=int, default=1 << 20)   # (torchrun's own parser trips over a bare --n)
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-s", type=float, default=4.0)
    ap.add_argument("--no-extra-legs", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
