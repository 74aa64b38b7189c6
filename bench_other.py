"""bench_other.py — the BASELINE.json configs other than the headline one, measured by bench.py into `other_configs`:

  c3  configs[2]  sparse CSC J 5 000 000 x 500 000, 200 entries per column (~20 per row, nnz 1e8), LevenbergMarquardt(LSMR())
                  on one GPU: ms per LM step, ms per LSMR iteration, HBM fraction of the two sparse products
  c4  configs[3]  dense J row-sharded, 250 000 x 4 000 rows per GPU (2M x 4k at N = 8), LevenbergMarquardt(Cholesky()):
                  ONE ncclAllReduce of the packed [upper(J'J) | J'f] per solve (timed by itself), DMMA fraction of the syrk,
                  δ compared with the row-sharded TSQR solve of the same damped system (independent algorithm)
  c5  configs[4]  bounded fit n = 10 000, m = 200 000, Dogleg(QR()) on one GPU: s per step, DMMA fraction of the update

Each entry has its own `roofline` and (N = 1, rank 0) `cpu_baseline`.  Parity at these sizes against the oracle is in
tests/test_gpu_named_sizes.py; here only cheap cross-checks run.
"""
import ctypes as C
import json
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
C_MODEL, NOISE = 0.1, 1e-3


def _hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p.get("hbm_gbs")), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6545.6, "fallback: B200_PROFILING.md measured copy bandwidth"


def run_all(env, only, dmma_peak):
    out = {}
    want = (lambda k: (not only) or k in only)
    if want("c3"):
        out["c3"] = _guard(run_c3, env)
    if want("c4"):
        out["c4"] = _guard(run_c4, env, dmma_peak)
    if want("c5"):
        out["c5"] = _guard(run_c5, env, dmma_peak)
    return out


def _all_ok(env, ok: bool) -> bool:
    """True only when EVERY rank reports ok (one all-reduce on the torch process group): a rank that failed to allocate
    must not leave the others waiting inside a library collective."""
    if env.world == 1:
        return ok
    t = env.torch.tensor([1 if ok else 0], device="cuda", dtype=env.torch.int32)
    env.dist.all_reduce(t, op=env.dist.ReduceOp.MIN)
    return bool(int(t.item()))


def _guard(fn, *a):
    import gc
    try:
        r = fn(*a)
    except Exception as e:      # one config failing must not take the headline line down; the failure is reported
        r = {"error": f"{type(e).__name__}: {e}"}
    gc.collect()
    return r


# ---------------------------------------------------------------------------------------------------------------
# c3: sparse LM(LSMR)
# ---------------------------------------------------------------------------------------------------------------
def run_c3(env, m=5_000_000, n=500_000, k=200, steps=10, warmup=2):
    """N = 1: the config as BASELINE.json names it.  N > 1 (beyond the north star's single-GPU LSMR): the
    ROWS of the same pattern are sharded over the ranks (strong scaling) and LSMR exchanges [J'u | ||u||²] once per
    iteration (lso_lsmr_solve_sharded)."""
    from lsob200._lib import check, lib
    L, ctx = env.L, env.ctx
    world, rank = env.world, env.rank
    h = ctx.handle
    nnz = n * k
    m_glob, nnz_glob = m, nnz
    err = None
    try:        # everything that can fail on ONE rank (host or device memory) happens before the ranks agree to go on
        colptr = np.zeros(n + 1, dtype=np.int64)
        rowval = np.zeros(nnz, dtype=np.int64)
        check(lib().lso_synth_csc_pattern(m, n, k, 20240609, colptr.ctypes.data, rowval.ctypes.data))
        colptr -= 1
        rowval -= 1
        if world > 1:
            row0, row1 = (m * rank) // world, (m * (rank + 1)) // world
            mask = (rowval >= row0) & (rowval < row1)
            cs = np.zeros(nnz + 1, dtype=np.int64)
            np.cumsum(mask, out=cs[1:])
            colptr = cs[colptr]
            rowval = rowval[mask] - row0
            del mask, cs
            m, nnz = row1 - row0, int(rowval.size)
        A0 = L.CSCMatrix(ctx, m, n, colptr, rowval)          # A  (f! needs t = A x)
        Jac = L.CSCMatrix(ctx, m, n, colptr, rowval)         # J = diag(1 + 2 c t) A
        aval, aval_r = L.DeviceVector(ctx, nnz), L.DeviceVector(ctx, nnz)
        xs, x, pert = (L.DeviceVector(ctx, n) for _ in range(3))
        t, b, noise = (L.DeviceVector(ctx, m) for _ in range(3))
        zero = L.DeviceVector(ctx, m)
    except Exception as e:
        err = f"{type(e).__name__}: {e}"
    if not _all_ok(env, err is None):
        return {"error": err or "another rank failed to allocate"}
    check(lib().lso_synth_vector(h, nnz, 0, 99 + rank, 1.0, aval.ptr), h)
    check(lib().lso_csc_set_values_dev(A0.handle, aval.ptr), h)
    A0.gather_csr(aval, aval_r)
    check(lib().lso_synth_vector(h, n, 0, 7, 1.0, xs.ptr), h)
    check(lib().lso_synth_vector(h, m, 0, 12 + 1000 * rank, NOISE, noise.ptr), h)
    check(lib().lso_synth_vector(h, n, 0, 13, 0.1, pert.ptr), h)
    A0.mul(t, xs, 1.0, 0.0)
    check(lib().lso_synth_residual_from_t(h, m, t.ptr, zero.ptr, C_MODEL, b.ptr), h)      # b = t + c t^2 at x*
    b.axpy(1.0, noise)
    x0 = L.DeviceVector(ctx, n).copyto(xs).axpy(1.0, pert)
    x.copyto(x0)
    del zero, noise, pert

    def f_(out, xx):
        A0.mul(t, xx, 1.0, 0.0)
        check(lib().lso_synth_residual_from_t(h, m, t.ptr, b.ptr, C_MODEL, out.ptr), h)

    def g_(JJ, xx):            # device g!: writes the CSC and the CSR image (no mirror gather)
        A0.mul(t, xx, 1.0, 0.0)
        check(lib().lso_synth_csc_jacobian_both(JJ.handle, aval.ptr, aval_r.ptr, t.ptr, C_MODEL), h)

    nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m), f_=f_, g_=g_, J=Jac, device_callbacks=True, ctx=ctx)
    anls = L.allocate(nls, L.LevenbergMarquardt(L.LSMR()), sharded=(world > 1))
    st = {"run": None, "iters": [], "acc": 0}

    def one_step():
        if st["run"] is None or st["run"].converged:
            x.copyto(x0)
            st["run"] = L.LMRun(anls)
        st["acc"] += int(st["run"].iterate())
        st["iters"].append(anls.solver.last_iters)

    for _ in range(warmup):
        one_step()
    st["iters"], st["acc"] = [], 0
    ctx.launch_count(reset=True)
    ms, wall = env.timed(one_step, steps)
    launches = ctx.launch_count(reset=True)

    # LSMR iteration cost by itself: a fixed number of iterations (all stopping rules off)
    ws = anls.solver
    NIT = 20
    dtd, dx, fcur = L.DeviceVector(ctx, n), L.DeviceVector(ctx, n), anls.fcur
    damp_h = None

    def fixed_iters():
        Jac.colsumabs2(dtd)
        if world > 1:
            ctx.allreduce(dtd)
        check(lib().lso_lm_damping(h, n, dtd.ptr, 1e-6, 1e32, 0.1), h)
        it, istop = C.c_int64(), C.c_int()
        if world > 1:
            check(lib().lso_lsmr_solve_sharded(ws._h, Jac.handle, fcur.ptr, dtd.ptr, dx.ptr, 0.0, 0.0, 0.0, NIT, m_glob, None,
                                               C.byref(it), C.byref(istop)), h)
        else:
            check(lib().lso_lsmr_solve(ws._h, Jac.handle, None, 0, fcur.ptr, dtd.ptr, dx.ptr, 0.0, 0.0, 0.0, NIT,
                                       C.byref(it), C.byref(istop)), h)
        assert it.value == NIT, (it.value, istop.value)

    fixed_iters()
    ctx.set_option("profile", 1)
    ctx.profile_read()
    ctx.stat("spmv_bytes", reset=True)
    ms_fix, _ = env.timed(fixed_iters, 3)
    sp_ms, sp_launches = ctx.profile_read()
    sp_bytes = ctx.stat("spmv_bytes", reset=True)
    ctx.set_option("profile", 0)
    lsmr_launches, lsmr_syncs = ws.stats()
    peak, peak_src = _hbm_peak()
    achieved = sp_bytes / (sp_ms * 1e-3) / 1e9 if sp_ms > 0 else 0.0
    ms_per_it = ms_fix / 3 / NIT
    if rank != 0:
        return None
    res = {
        "workload": f"sparse CSC J {m_glob}x{n}, {k} entries per column (~{nnz_glob // m_glob} per row, nnz {nnz_glob:.0e}) fp64, "
                    "LevenbergMarquardt(LSMR())",
        "baseline_config": "BASELINE.json configs[2]" + ("" if world == 1 else " with the rows sharded over the ranks (SURVEY §8 f4; "
                                                                              "the north star keeps LSMR on one GPU)"),
        "n_gpus": world, "scaling": "strong", "rows_per_gpu": m,
        "collective": ({"what": "ONE ncclAllReduce of [J'u | ||u||²] per LSMR iteration (+ colsumabs2 once per solve)",
                        "doubles": n + 1, "bytes": 8 * (n + 1)} if world > 1 else None),
        "metric": "trust-region steps/sec (fp64)", "value": steps / (ms * 1e-3), "unit": "steps/s",
        "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps, "steps": steps, "warmup": warmup,
        "steps_accepted": st["acc"], "lsmr_iterations_per_step": st["iters"], "gpu_launches": launches,
        "ms_per_lsmr_iteration": ms_per_it,
        "lsmr": {"fixed_iterations": NIT, "launches_per_solve": lsmr_launches, "host_syncs_per_solve": lsmr_syncs},
        "roofline": {"kernel": "spmv_warp_kernel (CSR-mirror J v and CSC J'u with the LSMR vector algebra and norms fused "
                               "in; paired 128-bit value / 64-bit index loads, G lanes per row / column, shuffle reductions)",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "launches": sp_launches, "kernel_ms_per_launch": sp_ms / max(sp_launches, 1),
                     "algorithmic_bytes_per_launch": sp_bytes / max(sp_launches, 1),
                     "numerator": "12 B per stored entry (8 B value + 4 B index) + 8 B per element of both vectors, per product",
                     "whole_iteration_frac": (2 * (12.0 * nnz + 8.0 * (m + n)) + 8.0 * 8 * n) / (ms_per_it * 1e-3) / 1e9 / peak,
                     "binding_resource": "L2 sector bandwidth / L1TEX wavefronts of the random gathers: one 32-byte sector per "
                                         "gathered double (ncu: 1.75e8 sectors = 5.6 GB through L2 per product, ~11.7 TB/s vs the "
                                         "~12.4 TB/s LTS cap), not HBM — profiles/r2_ncu_spmv_warp.txt",
                     "traffic": 1.38e9 if world == 1 else None},
        "last_ssr": st["run"].ssr,
    }
    # cpu baseline: the reference's serial CSC products (SparseArrays mul! is single-threaded), two LSMR iterations' worth
    if world > 1:
        res["cpu_baseline"] = {"note": "timed at N = 1 (same pattern): see that line"}
        return res
    try:
        import scipy.sparse as sp
        vals = Jac_values_host(L, ctx, Jac, nnz)
        A = sp.csc_matrix((vals, rowval.astype(np.int32), colptr.astype(np.int32)), shape=(m, n))
        v, u = np.ones(n), np.ones(m)
        A @ v
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            u2 = A @ v
            v2 = A.T @ u
            v = v + 1e-3 * v2 / max(np.abs(v2).max(), 1.0)         # the vector work of an iteration is a few axpys
            u = u + 1e-3 * u2 / max(np.abs(u2).max(), 1.0)
        sec = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"value": sec * 1e3, "unit": "ms per LSMR iteration", "cores": 1, "kind": "port",
                               "sample": f"{reps} iterations' worth of serial CSC J*v and J'*u (scipy, the reference's "
                                         "SparseArrays mul! is single-threaded) at the full pattern",
                               "gpu_over_cpu": sec * 1e3 / ms_per_it}
    except Exception as e:
        res["cpu_baseline"] = {"error": str(e)}
    return res


def Jac_values_host(L, ctx, Jac, nnz):
    from lsob200._lib import check, lib
    out = np.empty(nnz)
    check(lib().lso_download(ctx.handle, out.ctypes.data, Jac.values_ptr(), nnz * 8), ctx.handle)
    return out


# ---------------------------------------------------------------------------------------------------------------
# c4: row-sharded LM(Cholesky), weak scaling (250 000 rows per GPU)
# ---------------------------------------------------------------------------------------------------------------
def run_c4(env, dmma_peak, m_loc=250_000, n=4_000, steps=3, warmup=1):
    import bench
    from lsob200._lib import check, lib
    L, ctx = env.L, env.ctx
    h = ctx.handle
    world, rank = env.world, env.rank
    err = None
    try:
        prob = bench.DeviceProblem(L, ctx, m_loc, n, rank * m_loc, 20240607 + 4)
        x = L.DeviceVector(ctx, n).copyto(prob.x0)
        J = L.DenseMatrix(ctx, m_loc, n)
        nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m_loc), f_=prob.f_, g_=prob.g_, J=J, device_callbacks=True, ctx=ctx)
        anls = L.allocate(nls, L.LevenbergMarquardt(L.Cholesky()), sharded=(world > 1))
        qr = L.DenseQRAllocatedSolver(ctx, m_loc, n, damped=(world == 1), sharded=(world > 1))      # for the parity cross-check
        if world > 1:
            check(lib().lso_qr_prepare_sharded(qr._h), h)
    except Exception as e:
        err = f"{type(e).__name__}: {e}"
    if not _all_ok(env, err is None):
        return {"error": err or "another rank failed to allocate"} if rank == 0 else None
    st = {"run": None, "acc": 0}

    def one_step():
        if st["run"] is None or st["run"].converged:
            x.copyto(prob.x0)
            st["run"] = L.LMRun(anls)
        st["acc"] += int(st["run"].iterate())

    for _ in range(warmup):
        one_step()
    st["acc"] = 0
    ctx.set_option("profile", 1)
    ctx.profile_read(); ctx.profile_read_collective()
    ctx.stat("syrk_flops", reset=True)
    ctx.stat("syrk_i8_macs", reset=True)
    ctx.launch_count(reset=True)
    ms, wall = env.timed(one_step, steps)
    launches = ctx.launch_count(reset=True)
    syrk_ms, syrk_launches = ctx.profile_read()
    coll_ms, coll_calls = ctx.profile_read_collective()
    syrk_flops = ctx.stat("syrk_flops", reset=True)
    i8_macs = ctx.stat("syrk_i8_macs", reset=True)
    ctx.set_option("profile", 0)
    coll_ms = env.max_over_ranks(coll_ms)
    last_ssr = st["run"].ssr

    # parity cross-check at size: δ of the first damped solve, Cholesky + all-reduce against TSQR (QR of the shards,
    # all-gather of the R factors, QR of the stack) — two independent algorithms on the same sharded system
    x.copyto(prob.x0)
    fcur = anls.fcur
    prob.f_(fcur, x)
    prob.g_(J, x)
    dtd, grad, d_ch, d_qr = (L.DeviceVector(ctx, n) for _ in range(4))
    J.colsumabs2_and_grad(dtd, grad, fcur)
    if world > 1:
        ctx.allreduce(dtd)
        ctx.allreduce(grad)
    check(lib().lso_lm_damping(h, n, dtd.ptr, 1e-6, 1e32, 0.1), h)
    damp2 = L.DeviceVector(ctx, n).copyto(dtd)
    anls.solver.ldiv(d_ch, J, fcur, dtd)
    parity = {}
    try:
        qr.ldiv(d_qr, J, fcur, damp2)
        a, b2 = d_ch.download(), d_qr.download()
        parity = {"vs": "row-sharded TSQR solve of the same damped system (independent algorithm, same shards)",
                  "rel_diff": float(np.linalg.norm(a - b2) / np.linalg.norm(b2)),
                  "note": "Cholesky squares the condition number; agreement is bounded by cond(J'J + D) * eps"}
        del qr
    except Exception as e:
        parity = {"error": f"{type(e).__name__}: {e}"}
    if rank != 0:
        return None
    achieved = syrk_flops / (syrk_ms * 1e-3) / 1e12 if syrk_ms > 0 else 0.0
    pk = n * (n + 1) // 2 + n
    res = {
        "workload": f"dense synthetic J {m_loc * world}x{n} fp64 row-sharded over {world} GPU(s) ({m_loc} rows each), "
                    "LevenbergMarquardt(Cholesky())",
        "baseline_config": "BASELINE.json configs[3] (2M x 4k on 8 GPUs = this shard size x 8)", "n_gpus": world,
        "scaling": "weak", "metric": "trust-region steps/sec (fp64)", "value": steps / (ms * 1e-3), "unit": "steps/s",
        "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps, "steps": steps, "warmup": warmup,
        "steps_accepted": st["acc"], "gpu_launches": launches, "last_ssr": last_ssr,
        "collective": {"what": "ONE ncclAllReduce of the packed [upper(J'J) by columns | J'f] per solve",
                       "doubles": pk, "bytes": 8 * pk, "calls": coll_calls,
                       "ms_per_call_max_over_ranks": coll_ms / max(coll_calls, 1),
                       "ms_per_step": coll_ms / steps} if world > 1 else None,
        "roofline": _c4_roofline(achieved, dmma_peak, syrk_launches, syrk_ms, steps, ms, i8_macs),
        "parity": parity,
    }
    if world == 1:
        res["cpu_baseline"] = _c4_cpu(m_loc, n)
    return res


def _c4_roofline(fp64_tflops, dmma_peak, launches, syrk_ms, steps, ms, i8_macs):
    """J'J runs on tcgen05 (int8 digit matrices, Ozaki scheme) when the library chose that path (i8_macs > 0): the pipe it
    occupies is the int8 tensor pipe, so `achieved` / `peak` are int8 TOP/s (peak = 2 x the measured dense bf16 peak of
    MEASURED_PEAKS.json: kind::i8 issues at twice the kind::f16 rate); the fp64-equivalent rate is reported beside the DMMA
    peak it replaces.  Otherwise the DMMA syrk's roofline."""
    base = {"bound": "tensor", "launches": launches, "kernel_ms_per_step": syrk_ms / steps, "kernel_share_of_step": syrk_ms / ms,
            "fp64_equivalent_tflops": fp64_tflops, "fp64_dmma_peak_tflops": dmma_peak,
            "fp64_equivalent_over_dmma_peak": fp64_tflops / dmma_peak if dmma_peak else None, "traffic": None}
    if i8_macs > 0:      # ncu (profiles/r2_ncu_ozaki_syrk_tcgen05.txt): dram bytes of the tile kernel + split + exponents per J'J
        base["traffic"] = 89.9e9 + 0.2e9 + 16.0e9 + 8.0e9
    if i8_macs > 0:
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            bf16 = float(pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops"))
            src = "2 x MEASURED_PEAKS.json dense bf16 (sustained figure: the kernel runs tens of ms inside a step)"
        except Exception:
            bf16, src = 2250.0, "2 x nominal dense bf16 (2.25 PFLOP/s)"
        tops = 2.0 * i8_macs / (syrk_ms * 1e-3) / 1e12 if syrk_ms > 0 else 0.0
        base.update({"kernel": "oz_syrk_kernel (tcgen05.mma.kind::i8 128x128x32 into TMEM, TMA-fed 3-stage ring; incl. the digit "
                               "split and the split-K reduce)",
                     "achieved": tops, "peak": 2.0 * bf16, "unit": "TOP/s (int8)", "frac": tops / (2.0 * bf16), "peak_source": src,
                     "numerator": "2 x int8 MACs issued (S(S+1)/2 digit products x 128x128 tiles of the upper triangle x rows), "
                                  "lso_ctx_stat \"syrk_i8_macs\"; time = the whole J'J (split + tiles + reduce)"})
    else:
        base.update({"kernel": "syrk_mma_kernel (J'J upper tiles on the fp64 tensor pipe, DMMA m8n8k4)", "achieved": fp64_tflops,
                     "peak": dmma_peak, "unit": "TFLOP/s", "frac": fp64_tflops / dmma_peak if dmma_peak else None,
                     "numerator": "m n (n + 1) flops per J'J (lso_ctx_stat \"syrk_flops\"), per rank"})
    return base


def _c4_cpu(m, n, block=16_000):
    """dense_cholesky.jl:43-59 on host cores: dsyrk streamed over row blocks (the full J does not fit in host RAM at
    2M rows), dpotrf, two dtrsv; linear in the rows, so one block is timed and scaled."""
    import bench
    pool = bench.blas_pool()
    from scipy.linalg import blas, lapack
    rng = np.random.default_rng(4)
    Jb = np.asfortranarray(rng.standard_normal((block, n)))
    f = rng.standard_normal(block)
    blas.dsyrk(1.0, Jb[:2000], trans=1)
    t0 = time.perf_counter()
    Cm = blas.dsyrk(1.0, Jb, trans=1)
    g = Jb.T @ f
    t_blk = time.perf_counter() - t0
    Cm[np.diag_indices(n)] += 1.0 + np.einsum("ij,ij->j", Jb, Jb)
    t0 = time.perf_counter()
    c, info = lapack.dpotrf(Cm, lower=0)
    lapack.dpotrs(c, g, lower=0)
    t_fac = time.perf_counter() - t0
    sec = t_blk * (m / block) + t_fac
    return {"value": 1.0 / sec, "unit": "solves/s", "kind": "port", **pool,
            "sample": f"dsyrk + J'f on one {block} x {n} row block ({t_blk:.2f} s) scaled to {m} rows, plus dpotrf + dpotrs "
                      f"at n = {n} ({t_fac:.2f} s): {sec:.1f} s per damped Cholesky solve (solve only, EXTRAPOLATED linearly in rows)"}


# ---------------------------------------------------------------------------------------------------------------
# c5: bounded Dogleg(QR) at 200 000 x 10 000
# ---------------------------------------------------------------------------------------------------------------
def run_c5(env, dmma_peak, m=200_000, n=10_000, steps=2, warmup=1):
    import bench
    L, ctx = env.L, env.ctx
    world, rank = env.world, env.rank
    rows = [(m * r) // world for r in range(world + 1)]
    row0, m_loc = rows[rank], rows[rank + 1] - rows[rank]          # strong scaling: the rows of the fixed problem are sharded
    from lsob200._lib import check, lib
    err = None
    try:
        prob = bench.DeviceProblem(L, ctx, m_loc, n, row0, 20240607 + 5)
        x = L.DeviceVector(ctx, n).copyto(prob.x0)
        nls = L.LeastSquaresProblem(x=x, y=L.DeviceVector(ctx, m_loc), f_=prob.f_, g_=prob.g_, J=L.DenseMatrix(ctx, m_loc, n),
                                    device_callbacks=True, ctx=ctx)
        anls = L.allocate(nls, L.Dogleg(L.QR()), sharded=(world > 1))
        if world > 1:
            check(lib().lso_qr_prepare_sharded(anls.solver._h), ctx.handle)
    except Exception as e:
        err = f"{type(e).__name__}: {e}"
    if not _all_ok(env, err is None):
        return {"error": err or "another rank failed to allocate"} if rank == 0 else None
    xs, x0 = prob.xstar.download(), prob.x0.download()
    lo, hi = np.full(n, -np.inf), np.full(n, np.inf)
    idx = np.arange(n) % 5 == 0                       # 20 % of the coordinates are boxed around x*
    lo[idx] = np.minimum(xs[idx] - 0.05, x0[idx])
    hi[idx] = np.maximum(xs[idx] + 0.05, x0[idx])
    run = L.DoglegRun(anls, lower=lo, upper=hi)
    st = {"acc": 0}

    def one_step():
        st["acc"] += int(run.iterate())

    for _ in range(warmup):
        one_step()
    st["acc"] = 0
    ctx.set_option("profile", 1)
    ctx.profile_read()
    ctx.stat("qr_update_flops", reset=True)
    ctx.stat("qr_flops", reset=True)
    ctx.launch_count(reset=True)
    ms, wall = env.timed(one_step, steps)
    launches = ctx.launch_count(reset=True)
    k_ms, k_launches = ctx.profile_read()
    coll_ms, coll_calls = ctx.profile_read_collective()
    uflops = ctx.stat("qr_update_flops", reset=True)
    qflops = ctx.stat("qr_flops", reset=True)
    ctx.set_option("profile", 0)
    coll_ms = env.max_over_ranks(coll_ms)
    if rank != 0:
        return None
    achieved = uflops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
    res = {
        "workload": f"bounded dense synthetic fit m={m} n={n} fp64 (20 % of the coordinates boxed), Dogleg(QR())",
        "baseline_config": "BASELINE.json configs[4] (1 -> 8 GPU sweep: the rows of the fixed problem are sharded, TSQR)",
        "n_gpus": world, "scaling": "strong", "rows_per_gpu": m_loc,
        "collective": ({"what": "ncclAllGather of the n x (n+1) [R | Q'f] factors, per factorisation",
                        "bytes_per_rank": 8 * n * (n + 1), "calls": coll_calls,
                        "ms_per_call_max_over_ranks": coll_ms / max(coll_calls, 1)} if world > 1 else None),
        "metric": "trust-region steps/sec (fp64)", "value": steps / (ms * 1e-3), "unit": "steps/s",
        "s_per_step": ms / steps * 1e-3, "wall_s_per_step": wall / steps * 1e-3, "steps": steps, "warmup": warmup,
        "steps_accepted": st["acc"],
        "gpu_launches": launches, "last_ssr": run.ssr,
        "roofline": {"kernel": "qr_apply_pp_kernel_t (CAQR trailing update, DMMA)", "bound": "tensor", "achieved": achieved,
                     "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak if dmma_peak else None,
                     "launches": k_launches, "kernel_s_per_step": k_ms / steps * 1e-3, "kernel_share_of_step": k_ms / ms,
                     "numerator": "trailing-update flops only (lso_ctx_stat \"qr_update_flops\")", "traffic": None},
    }
    del run, anls, nls, prob
    if world == 1:
        res["cpu_baseline"] = _c5_cpu(m, n)
    return res


def _c5_cpu(m, n, mr=20_000, nr=1_000):
    """dgelsy (dgeqp3 + dormqr + dtrtrs, dense_qr.jl:30-42 via the stdlib) does not finish in bench time at 200k x 10k
    (~3.3e13 flop, about an hour on a dozen cores, see oracle/make_golden_c5.py): one solve is timed at a reduced shape
    and scaled by the flop model 2 m n^2 - 2/3 n^3 — flagged EXTRAPOLATED."""
    import bench
    from oracle import reference_port as O
    pool = bench.blas_pool()
    rng = np.random.default_rng(5)
    Jr = np.asfortranarray(rng.standard_normal((mr, nr)))
    fr = rng.standard_normal(mr)
    O.qr_ldiv(Jr[:4000, :200].copy(order="F"), fr[:4000])
    t0 = time.perf_counter()
    O.qr_ldiv(Jr, fr)
    t = time.perf_counter() - t0
    flop = lambda a, b: 2.0 * a * b * b - 2.0 * b ** 3 / 3.0
    sec = t * flop(m, n) / flop(mr, nr)
    golden = os.path.join(ROOT, "tests", "golden", "c5_first_solve.npz")
    meas = None
    if os.path.exists(golden):
        g = np.load(golden)
        meas = {"dgelsy_seconds_full_size": float(g["dgelsy_seconds"]), "cores": int(g["cores"]),
                "where": "oracle/make_golden_c5.py on the build container (not this box)"}
    return {"value": 1.0 / sec, "unit": "solves/s", "kind": "port", "extrapolated": True, **pool,
            "sample": f"one undamped QR solve (dgelsy) at {mr} x {nr}: {t:.2f} s, scaled by the flop model to {m} x {n}: "
                      f"{sec:.0f} s per solve (EXTRAPOLATED)",
            "measured_elsewhere": meas}
