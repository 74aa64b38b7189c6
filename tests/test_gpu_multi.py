"""Multi-GPU (one process per GPU, NCCL through the C ABI): row-sharded QR (TSQR), Cholesky (one all-reduce of
[J'J | J'y]) and LSMR (one all-reduce of [J'u | ||u||²] per iteration) solves, and sharded LM / Dogleg runs, against the
single-process oracle.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import ctypes as C
    sys.path.insert(0, ROOT)
    from lsob200._lib import lib
    c = C.c_int()
    lib().lso_device_count(C.byref(c))
    return c.value


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import lsob200 as L
    from lsob200.sharding import init_comm, row_partition
    from oracle import reference_port as O
    from oracle import synth_ref as S
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = init_comm(L.Context(rank), dist)
    res = {}
    # ---- one damped solve, both solvers ----
    rng = np.random.default_rng(4)
    m, n = 9001, 130
    Jh = rng.standard_normal((m, n)) * np.exp2(rng.integers(-3, 4, n))
    yh = rng.standard_normal(m)
    damp = np.einsum("ij,ij->j", Jh, Jh) / 10
    row0, rows = row_partition(m, world)[rank]
    Jk, yk = L.DenseMatrix(ctx, rows, n, Jh[row0:row0 + rows]), L.DeviceVector(ctx, rows, yh[row0:row0 + rows])
    d, x = L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    xr, _ = O.qr_ldiv(Jh, yh, damp.copy())
    L.DenseQRAllocatedSolver(ctx, rows, n, damped=False, sharded=True).ldiv(x, Jk, yk, d)
    res["qr"] = float(np.linalg.norm(x.download() - xr) / np.linalg.norm(xr))
    L.DenseCholeskyAllocatedSolver(ctx, rows, n, damped=True, sharded=True).ldiv(x, Jk, yk, d)
    res["chol"] = float(np.linalg.norm(x.download() - xr) / np.linalg.norm(xr))
    # a REPLICATED (non-sharded) problem on a context that has a communicator stays local: no all-reduce, same answer
    Jf, yf = L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh)
    L.DenseCholeskyAllocatedSolver(ctx, m, n, damped=True).ldiv(x, Jf, yf, d)
    res["chol_local"] = float(np.linalg.norm(x.download() - xr) / np.linalg.norm(xr))
    # ---- sharded LM(QR) and LM(Cholesky) on the synthetic model vs the oracle on the whole problem ----
    m, n, seed = 8000, 64, 31337
    model = S.DenseModel(m, n, seed, c=0.1, noise=1e-3)
    ro = O.levenberg_marquardt(model.f, model.g, model.x0, np.zeros((m, n), order="F"), m, solver="qr")
    row0, rows = row_partition(m, world)[rank]
    for name, sol in (("lm_qr", L.QR), ("lm_chol", L.Cholesky)):
        prob = bench.DeviceProblem(L, ctx, rows, n, row0, seed)
        xx = L.DeviceVector(ctx, n).copyto(prob.x0)
        nls = L.LeastSquaresProblem(x=xx, y=L.DeviceVector(ctx, rows), f_=prob.f_, g_=prob.g_,
                                    J=L.DenseMatrix(ctx, rows, n), device_callbacks=True, ctx=ctx)
        r = L.optimize_(L.allocate(nls, L.LevenbergMarquardt(sol()), sharded=True))
        res[name] = (r.iterations, ro.iterations, float(np.linalg.norm(r.minimizer.download() - ro.minimizer) /
                                                        np.linalg.norm(ro.minimizer)), r.converged)
    # ---- sharded Dogleg(QR) (undamped TSQR; BASELINE.json configs[4]'s scaling form) with bounds ----
    rod = O.dogleg(model.f, model.g, model.x0, np.zeros((m, n), order="F"), m, solver="qr")
    prob = bench.DeviceProblem(L, ctx, rows, n, row0, seed)
    xx = L.DeviceVector(ctx, n).copyto(prob.x0)
    nls = L.LeastSquaresProblem(x=xx, y=L.DeviceVector(ctx, rows), f_=prob.f_, g_=prob.g_,
                                J=L.DenseMatrix(ctx, rows, n), device_callbacks=True, ctx=ctx)
    r = L.optimize_(L.allocate(nls, L.Dogleg(L.QR()), sharded=True))
    res["dogleg_qr"] = (r.iterations, rod.iterations, float(np.linalg.norm(r.minimizer.download() - rod.minimizer) /
                                                            np.linalg.norm(rod.minimizer)), r.converged)
    # ---- row-sharded LSMR on a sparse J: one all-reduce of [J'u | ||u||²] per iteration ----
    import scipy.sparse as sp
    m, n = 30011, 900
    Asp = sp.random(m, n, density=0.01, random_state=17, format="csr")
    Asp.data = Asp.data * 2 - 1
    yh = np.random.default_rng(8).standard_normal(m)
    row0, rows = row_partition(m, world)[rank]
    Ak = Asp[row0:row0 + rows].tocsc()
    Ak.sort_indices()
    Jk, yk = L.CSCMatrix.from_scipy(ctx, Ak), L.DeviceVector(ctx, rows, yh[row0:row0 + rows])
    for damped in (True, False):
        damp = np.asarray(Asp.multiply(Asp).sum(axis=0)).ravel() / 10 + 1e-3
        xr, nmul_r, it_r, istop_r = O.lsmr_ldiv(Asp.tocsc(), yh, damp.copy() if damped else None)
        ws = (L.LSMRDampenedAllocatedSolver if damped else L.LSMRAllocatedSolver)(ctx, rows, n, sharded=True, m_total=m)
        x = L.DeviceVector(ctx, n)
        if damped:
            _, nmul = ws.ldiv(x, Jk, yk, L.DeviceVector(ctx, n, damp))
        else:
            _, nmul = ws.ldiv(x, Jk, yk)
        res["lsmr_damped" if damped else "lsmr"] = (ws.last_iters, it_r, ws.last_istop, istop_r, nmul, nmul_r,
                                                    float(np.linalg.norm(x.download() - xr) / np.linalg.norm(xr)), ws.stats())
    # ---- sharded LM(LSMR) from host callbacks on this rank's rows: r_i(x) = (A x)_i + c (A x)_i² - b_i ----
    xs = np.random.default_rng(9).standard_normal(n)
    t_star = Asp @ xs
    bh = t_star + 0.05 * t_star ** 2
    Ak_csr = Asp[row0:row0 + rows].tocsr()
    bk = bh[row0:row0 + rows]

    def f_loc(out_, xx):
        t = Ak_csr @ xx
        out_[:] = t + 0.05 * t * t - bk

    def g_loc(JJ, xx):
        t = Ak_csr @ xx
        JJ.data[:] = sp.diags(1 + 0.1 * t).dot(Ak_csr).tocsc().data

    def f_all(out_, xx):
        t = Asp @ xx
        out_[:] = t + 0.05 * t * t - bh

    def g_all(JJ, xx):
        t = Asp @ xx
        JJ.data[:] = sp.diags(1 + 0.1 * t).dot(Asp).tocsc().data

    x0 = xs + 0.1 * np.random.default_rng(10).standard_normal(n)
    Jall = Asp.tocsc().copy(); Jall.sort_indices()
    ro = O.levenberg_marquardt(f_all, g_all, x0.copy(), Jall, m, solver="lsmr")
    Jloc = Ak.copy()
    nls = L.LeastSquaresProblem(x=x0.copy(), y=np.zeros(rows), f_=f_loc, g_=g_loc, J=Jloc, ctx=ctx)
    r = L.optimize_(L.allocate(nls, L.LevenbergMarquardt(L.LSMR()), sharded=True))
    xm = r.minimizer.download() if hasattr(r.minimizer, "download") else np.asarray(r.minimizer)
    res["lm_lsmr"] = (r.iterations, ro.iterations, float(np.linalg.norm(xm - ro.minimizer) / np.linalg.norm(ro.minimizer)),
                      r.converged)
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_solves_and_lm(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, 29700 + os.getpid() % 1000, out), nprocs=world, join=True)
    for rank in range(world):
        res = out[rank]
        assert res["qr"] <= 1e-10 and res["chol"] <= 1e-10 and res["chol_local"] <= 1e-10, res
        for k in ("lm_qr", "lm_chol", "dogleg_qr"):
            it, it_ref, err, conv = res[k]
            assert conv and it == it_ref and err <= 1e-9, (k, res[k])
        for k in ("lsmr_damped", "lsmr"):
            it, it_r, istop, istop_r, nmul, nmul_r, err, (launches, syncs) = res[k]
            assert (it, istop, nmul) == (it_r, istop_r, nmul_r) and err <= 2e-5, (k, res[k])
            assert syncs <= max(it, 1) + 1
        it, it_ref, err, conv = res["lm_lsmr"]
        assert conv and it == it_ref and err <= 1e-6, res["lm_lsmr"]
