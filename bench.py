#!/usr/bin/env python
"""bench.py — trust-region steps/sec (fp64) of LevenbergMarquardt(QR()) on the dense synthetic problem of
BASELINE.json configs[1] (J 100 000 x 1 000), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--m M --n N] [--no-other-configs]

A "step" is one pass of the `while` body of levenberg_marquardt.jl:72-140 (colsumabs2!, damping, the damped QR
solve, J'f, x -= δ, f!, ||Jδ - f||², ρ / Δ update, and g! after an accepted step), i.e. one trust-region step.
`value` runs real LM iterations with everything resident in HBM (restarting from x0 on convergence);
`e2e` is the same hot-path body driven from HOST buffers: J (pinned) and f are copied H2D every step and δ plus
the step scalars are read back, user f!/g! evaluation excluded.  N > 1 shards the rows of J over ranks
(TSQR: local QR, NCCL all-gather of the n x (n+1) R factors, replicated QR of the stack) = strong scaling.

The same JSON line carries `other_configs` (bench_other.py): BASELINE.json configs[2] (sparse LM(LSMR)) and
configs[4] (bounded Dogleg(QR) at 200k x 10k) at N = 1, and configs[3] (row-sharded LM(Cholesky), 250 000 x 4 000 rows
per GPU = 2M x 4k at N = 8, ONE NCCL all-reduce of the packed [J'J | J'f] per step) at every N.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

_NCPU = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
if "reference" in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core
    # (BASELINE.md §4), so the BLAS pool is sized before numpy / scipy load their OpenBLAS and again at run time.
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_NCPU)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240607 + 2
C_MODEL = 0.1
NOISE = 1e-3


def workload_name(m, n):
    """One string for both arms (the driver compares them)."""
    return f"dense synthetic J {m}x{n} fp64, LevenbergMarquardt(QR())"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=100000)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-resolve", action="store_true", help="skip the rejected-step re-solve measurement (ncu launch-list passes)")
    ap.add_argument("--only", default="", help="comma list of other configs to run (c3,c4,c5); default: all that apply")
    return ap.parse_args()


def blas_pool(threads=None):
    """Size every OpenBLAS pool in the process to `threads` (default: all host cores) and report what is in force:
    {"cores", "blas_threads", "openblas_version"} — printed beside every CPU number."""
    info = {"cores": _NCPU, "blas_threads": None, "openblas_version": None}
    try:
        import scipy.linalg  # noqa: F401  (loads scipy's OpenBLAS so that it is sized too)
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=threads or _NCPU, user_api="blas")
        pools = [p for p in threadpoolctl.threadpool_info() if p.get("user_api") == "blas"]
        if pools:
            info["blas_threads"] = min(int(p["num_threads"]) for p in pools)
            info["openblas_version"] = "/".join(sorted({str(p.get("version")) for p in pools}))
    except Exception as e:     # pragma: no cover
        info["blas_note"] = f"threadpoolctl unavailable: {e}"
    return info


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        busy = sm[len(sm) // 2:] or [0.0]          # upper half = samples under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's LM(QR) (LAPACK dgeqp3-based dgelsy through OpenBLAS) on host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_lm_steps(m, n, nsteps, nwarm, A=None, seed=SEED):
    from oracle import reference_port as O
    from oracle import synth_ref as S
    model = S.DenseModel(m, n, seed, c=C_MODEL, noise=NOISE, A=A)
    J = np.zeros((m, n), order="F")
    # run nwarm + nsteps iterations once, timing the last nsteps (tolerances off so the loop never exits early)

    def f_(out, x):
        model.f(out, x)

    t_iter = []
    orig_assess = O.assess_convergence

    def assess(*a, **k):       # called exactly once per iteration, at its end: use it as the iteration clock
        t_iter.append(time.perf_counter())
        return orig_assess(*a, **k)

    O.assess_convergence = assess
    try:
        t0 = time.perf_counter()
        r = O.levenberg_marquardt(f_, model.g, model.x0, J, m, solver="qr", x_tol=-1, f_tol=-1, g_tol=-1,
                                  iterations=nwarm + nsteps)
    finally:
        O.assess_convergence = orig_assess
    stamps = [t0] + t_iter
    dt = stamps[nwarm + nsteps] - stamps[nwarm]
    return nsteps / dt, dt / nsteps, r


def run_reference(args, rank, world):
    if rank != 0:
        return
    pool = blas_pool()
    K = max(1, min(args.steps, 3))
    W = max(0, min(args.warmup, 1))
    sps, sec, r = cpu_lm_steps(args.m, args.n, K, W)
    line = {
        "impl": "reference", "metric": "trust-region steps/sec (fp64)", "value": sps, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.m, args.n), "baseline_config": "BASELINE.json configs[1]",
                   "note": "oracle restatement of LeastSquaresOptim.jl's LM(QR) on the host cores (Julia is not in this "
                           "image): LAPACK dgelsy (dgeqp3 + dormqr + dtrtrs) and dgemv through OpenBLAS; "
                           f"steps capped at {K} (requested {args.steps}), warm-up {W} (a step is seconds of CPU work)",
                   "omp_num_threads_env_inherited": os.environ.get("OMP_NUM_THREADS")},
        "cpu_baseline": {"value": sps, "unit": "steps/s", "kind": "port", **pool,
                         "sample": f"{K} full LM(QR) steps at {args.m}x{args.n} after {W} warm-up"},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
class DeviceProblem:
    """Synthetic dense model resident in HBM; f! / g! are CUDA kernels (lso_synth_*), rows [row0, row0 + m_loc)."""

    def __init__(self, L, ctx, m_loc, n, row0, seed):
        from lsob200._lib import check, lib
        self.L, self.ctx, self.m, self.n = L, ctx, m_loc, n
        self.lib, self.check = lib(), check
        self.A = L.DenseMatrix(ctx, m_loc, n)
        check(self.lib.lso_synth_dense_matrix(ctx.handle, m_loc, n, row0, seed, self.A.ptr, self.A.ld), ctx.handle)
        self.xstar, self.x0, self.b = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), L.DeviceVector(ctx, m_loc)
        self.t = L.DeviceVector(ctx, m_loc)
        noise, pert = L.DeviceVector(ctx, m_loc), L.DeviceVector(ctx, n)
        check(self.lib.lso_synth_vector(ctx.handle, n, 0, seed + 11, 1.0, self.xstar.ptr), ctx.handle)
        check(self.lib.lso_synth_vector(ctx.handle, m_loc, row0, seed + 12, 1.0, noise.ptr), ctx.handle)
        check(self.lib.lso_synth_vector(ctx.handle, n, 0, seed + 13, 1.0, pert.ptr), ctx.handle)
        zero = L.DeviceVector(ctx, m_loc)
        check(self.lib.lso_synth_residual(ctx.handle, m_loc, n, self.A.ptr, self.A.ld, self.xstar.ptr, zero.ptr, C_MODEL,
                                          self.t.ptr, self.b.ptr), ctx.handle)        # b = t + c t^2 at x*
        self.b.axpy(NOISE, noise)
        self.x0.copyto(self.xstar).axpy(0.1, pert)

    def f_(self, out, x):
        self.check(self.lib.lso_synth_residual(self.ctx.handle, self.m, self.n, self.A.ptr, self.A.ld, x.ptr, self.b.ptr,
                                               C_MODEL, self.t.ptr, out.ptr), self.ctx.handle)

    def g_(self, J, x):
        self.check(self.lib.lso_dense_gemv_n(self.ctx.handle, self.m, self.n, 1.0, self.A.ptr, self.A.ld, x.ptr, 0.0,
                                             self.t.ptr), self.ctx.handle)
        self.check(self.lib.lso_synth_jacobian(self.ctx.handle, self.m, self.n, self.A.ptr, self.A.ld, self.t.ptr,
                                               C_MODEL, J.ptr, J.ld), self.ctx.handle)


class Env:
    """What every measurement needs: the library, the context, its stream as a torch stream, the ranks."""

    def __init__(self, L, ctx, torch, dist, rank, world, local_rank):
        self.L, self.ctx, self.torch, self.dist = L, ctx, torch, dist
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))

    def sync_all(self):
        self.ctx.sync()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, reps):
        """CUDA-event time (ms) of `reps` calls of fn on the context stream, barrier + synchronize on both sides,
        max over ranks."""
        torch = self.torch
        self.sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        with torch.cuda.stream(self.stream):
            ev0.record(self.stream)
            for _ in range(reps):
                fn()
            ev1.record(self.stream)
        self.sync_all()
        wall = (time.perf_counter() - t0) * 1e3
        return self.max_over_ranks(ev0.elapsed_time(ev1)), wall

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return float(v)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import lsob200 as L

    torch.cuda.set_device(local_rank)
    ctx = L.Context(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
        uid = [L.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])
    env = Env(L, ctx, torch, dist if world > 1 else None, rank, world, local_rank)

    m, n = args.m, args.n
    rows = [(m * r) // world for r in range(world + 1)]
    row0, m_loc = rows[rank], rows[rank + 1] - rows[rank]
    prob = DeviceProblem(L, ctx, m_loc, n, row0, SEED)
    x = L.DeviceVector(ctx, n).copyto(prob.x0)
    y = L.DeviceVector(ctx, m_loc)
    J = L.DenseMatrix(ctx, m_loc, n)
    nls = L.LeastSquaresProblem(x=x, y=y, f_=prob.f_, g_=prob.g_, J=J, device_callbacks=True, ctx=ctx)
    anls = L.allocate(nls, L.LevenbergMarquardt(L.QR()), sharded=(world > 1))

    state = {"run": None, "restarts": 0, "accepted": 0, "rejected": 0}

    def one_step():
        run = state["run"]
        if run is None or run.converged:
            x.copyto(prob.x0)
            state["run"] = run = L.LMRun(anls)
            state["restarts"] += 1
        if run.iterate():
            state["accepted"] += 1
        else:
            state["rejected"] += 1

    for _ in range(args.warmup):
        one_step()
    env.sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ctx.set_option("profile", 1)
    ctx.launch_count(reset=True)
    ctx.stat("qr_update_flops", reset=True)
    state["accepted"] = state["rejected"] = 0
    ms, wall_ms = env.timed(one_step, args.steps)
    launches = ctx.launch_count(reset=True)
    kern_ms, kern_launches = ctx.profile_read()
    coll_ms, coll_calls = ctx.profile_read_collective()
    update_flops = ctx.stat("qr_update_flops", reset=True)
    ctx.set_option("profile", 0)
    clocks = sampler.stop() if rank == 0 else None
    value = args.steps / (ms * 1e-3)
    last_ssr = state["run"].ssr if state["run"] else None

    # ---- (f3) what a REJECTED step's solve costs: same J and f, larger damping (levenberg_marquardt.jl:77-87,135) ----
    resolve = None
    if world == 1 and not args.no_resolve:
        resolve = measure_resolve(env, prob, anls, n)

    # ---- e2e: hot-path body from HOST buffers (J + f uploaded each step, δ + scalars downloaded) ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(env, prob, anls, m_loc, n, args)

    # ---- roofline of the dominant kernel (QR trailing update on the fp64 tensor pipe) ----
    roofline = cpu_baseline = None
    dmma_peak = None
    if rank == 0:
        import ctypes as C
        from lsob200._lib import check, lib
        out = C.c_double()
        check(lib().lso_bench_fp64_mma_peak(ctx.handle, 20000, C.byref(out)), ctx.handle)
        dmma_peak = out.value
        # numerator: the flops the update kernel itself performs usefully — sum over panels of 4*32*(active rows)*(trailing
        # columns), counted by the library as it launches (lso_ctx_stat); the panel factorisation's own flops are NOT
        # credited to this kernel
        achieved = update_flops / (kern_ms * 1e-3) / 1e12 if kern_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "apply_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roofline = {
            "kernel": "qr_apply_pp_kernel_t (CAQR trailing update: mma.sync m8n8k4 f64 = DMMA, two ping-pong consumer groups, "
                      "all global traffic as cp.async.bulk loads / stores issued by a producer warp)",
            "bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
            "frac": achieved / dmma_peak if dmma_peak else None, "traffic": traffic,
            "peak_source": "fp64 DMMA issue-bound micro-benchmark measured in this run (lso_bench_fp64_mma_peak); "
                           "MEASURED_PEAKS.json holds only HBM and bf16 peaks, tcgen05 has no f64 kind",
            "launches": kern_launches, "kernel_ms_per_step": kern_ms / args.steps,
            "kernel_share_of_step": (kern_ms / args.steps) / (ms / args.steps),
            "algorithmic_flops_per_step": update_flops / args.steps,
            "numerator": "trailing-update flops only (sum over panels of 4*32*rows*trailing columns, lso_ctx_stat "
                         "\"qr_update_flops\"), per rank",
        }
        if world == 1 and not args.no_cpu_baseline:
            pool = blas_pool()
            A_host = prob.A.download()
            sps, sec, _ = cpu_lm_steps(m, n, 1, 1, A=A_host)
            del A_host
            cpu_baseline = {"value": sps, "unit": "steps/s", "kind": "port", **pool,
                            "sample": f"1 full LM(QR) step at {m}x{n} (after 1 warm-up step) of the oracle restatement: "
                                      f"LAPACK dgelsy/dgeqp3 + dgemv via OpenBLAS, {sec:.2f} s/step"}

    # ---- the other BASELINE.json configs (their own roofline / cpu_baseline / parity each) ----
    other = None
    if not args.no_other_configs:
        del prob, anls, nls, J, y, x
        state["run"] = None
        import gc
        gc.collect()
        import bench_other
        only = [s for s in args.only.split(",") if s]
        other = bench_other.run_all(env, only, dmma_peak)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = {
        "metric": "trust-region steps/sec (fp64)", "value": value, "unit": "steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(m, n), "baseline_config": "BASELINE.json configs[1]",
                   "rows_per_gpu": m_loc, "parallelism": "single GPU" if world == 1 else f"row-sharded TSQR x{world}",
                   "l2_policy": f"inputs larger than L2: J is {8 * m_loc * n / 1e6:.0f} MB per GPU vs 126 MB L2",
                   "lm_restarts_in_run": state["restarts"], "last_ssr": last_ssr, "steps_accepted": state["accepted"],
                   "steps_rejected": state["rejected"],
                   "wall_ms_per_step": wall_ms / args.steps, "rejected_step_solve": resolve},
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "collective": ({"what": "ncclAllGather of the n x (n+1) [R | Q'f] factors, per step", "ms_per_step": coll_ms / args.steps,
                        "calls": coll_calls} if world > 1 else None),
        "measured_peaks": {"hbm_gbs": peaks.get("hbm_gbs"), "fp64_dmma_tflops": dmma_peak},
        "other_configs": other,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_resolve(env, prob, anls, n):
    """One damped solve on a fresh J (the QR of [J; sqrt(D)]) against the re-solve a rejected step needs (same J and f,
    Δ halved: lso_qr_solve_redamp factors the 2n x n stack [R; sqrt(D_new - D_last)])."""
    L, ctx = env.L, env.ctx
    x = anls.x
    x.copyto(prob.x0)
    prob.f_(anls.fcur, x)
    prob.g_(anls.J, x)
    dtd, d2, dx = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), L.DeviceVector(ctx, n)
    anls.J.colsumabs2(dtd)
    from lsob200.api import _lm_damping
    _lm_damping(ctx, dtd, 0.1)
    state = {"delta": 1.0}

    def fresh():
        d2.copyto(dtd)
        anls.solver.ldiv(dx, anls.J, anls.fcur, d2, same_J=False)

    def rejected():
        state["delta"] *= 2.0
        d2.copyto(dtd).rmul(state["delta"])
        anls.solver.ldiv(dx, anls.J, anls.fcur, d2, same_J=True)

    fresh()
    ms_f, _ = env.timed(fresh, 3)
    fresh()
    rejected()                      # first use creates (and tunes) the 2n x n stack workspace
    before = anls.solver.solves_redamped
    ms_r, _ = env.timed(rejected, 5)
    return {"fresh_J_solve_ms": ms_f / 3, "rejected_step_solve_ms": ms_r / 5,
            "redamped_solves": anls.solver.solves_redamped - before,
            "note": "a rejected LM step re-solves with the same J, f and a larger damping: the triangular factor of the "
                    "previous solve is re-damped (QR of [R; sqrt(D_new - D_last)], 2n x n banded) instead of refactoring "
                    "[J; sqrt(D)]; parity tests/test_gpu_solvers.py::test_qr_redamp_after_rejected_steps"}


def run_e2e(env, prob, anls, m_loc, n, args):
    """Same metric end to end through the package API with HOST buffers: every step uploads J and f, runs the
    hot-path body of one LM iteration on the device and reads δ and the step scalars back.  Measured twice: from
    PINNED host memory (the headline `value`) and from ordinary PAGEABLE memory (what a Julia `Matrix` is)."""
    import ctypes as C
    from lsob200._lib import check, lib
    L, ctx = env.L, env.ctx
    x = anls.x
    x.copyto(prob.x0)
    prob.f_(anls.fcur, x)
    prob.g_(anls.J, x)
    ctx.sync()
    hJ, hf = C.c_void_p(), C.c_void_p()
    check(lib().lso_host_alloc_pinned(ctx.handle, m_loc * n * 8, C.byref(hJ)), ctx.handle)
    check(lib().lso_host_alloc_pinned(ctx.handle, m_loc * 8, C.byref(hf)), ctx.handle)
    check(lib().lso_download(ctx.handle, hJ, anls.J.ptr, m_loc * n * 8), ctx.handle)
    check(lib().lso_download(ctx.handle, hf, anls.fcur.ptr, m_loc * 8), ctx.handle)
    hstep = L.HostStep(anls)
    dx_host = np.zeros(n)
    K = max(3, min(args.steps, 10))
    res = {}

    def measure(pJ, pf):
        for _ in range(2):
            hstep.run(pJ, pf, 10.0, dx_host)
        out = {}

        def one():
            out["scal"] = hstep.run(pJ, pf, 10.0, dx_host)
        ms, wall = env.timed(one, K)
        return ms, wall, out["scal"]

    ms, wall, scal = measure(hJ.value, hf.value)
    # pageable: plain numpy arrays (malloc'd, not page-locked), the memory a Julia Array lives in
    Jp = np.empty(m_loc * n)
    fp = np.empty(m_loc)
    C.memmove(Jp.ctypes.data, hJ.value, m_loc * n * 8)
    C.memmove(fp.ctypes.data, hf.value, m_loc * 8)
    ms_p, wall_p, _ = measure(Jp.ctypes.data, fp.ctypes.data)
    # the same arrays page-locked in place once (lso_host_register): what the glue does with a Julia J at allocation
    check(lib().lso_host_register(ctx.handle, Jp.ctypes.data, m_loc * n * 8), ctx.handle)
    check(lib().lso_host_register(ctx.handle, fp.ctypes.data, m_loc * 8), ctx.handle)
    ms_r, wall_r, _ = measure(Jp.ctypes.data, fp.ctypes.data)
    lib().lso_host_unregister(ctx.handle, Jp.ctypes.data)
    lib().lso_host_unregister(ctx.handle, fp.ctypes.data)
    lib().lso_host_free_pinned(ctx.handle, hJ)
    lib().lso_host_free_pinned(ctx.handle, hf)
    res = {"value": K / (ms * 1e-3), "unit": "steps/s", "h2d_bytes_per_step": (m_loc * n + m_loc) * 8,
           "d2h_bytes_per_step": (n + 4) * 8, "steps": K, "ms_per_step": ms / K, "wall_ms_per_step": wall / K,
           "note": "per step: H2D of J and f from pinned host memory (in row chunks, each chunk factorised while the next is in "
                   "flight when row_chunks > 1), colsumabs2 + damping + QR solve + J'f + predicted ssr on the device, D2H of δ "
                   "and 4 scalars; user f!/g! evaluation excluded",
           "pageable": {"value": K / (max(ms_p, wall_p) * 1e-3), "ms_per_step": max(ms_p, wall_p) / K,
                        "note": "same step with J and f in ordinary pageable host memory (numpy / Julia Array); "
                                "the larger of device and wall time is reported because the driver stages pageable "
                                "copies synchronously"},
           "registered": {"value": K / (max(ms_r, wall_r) * 1e-3), "ms_per_step": max(ms_r, wall_r) / K,
                          "note": "the same pageable arrays after ONE lso_host_register (cudaHostRegister in place): J, x, y "
                                  "persist over the whole optimize! run, so the glue registers them at allocation"},
           "row_chunks": hstep.chunks,
           "scalars": scal}
    return res


if __name__ == "__main__":
    main()
