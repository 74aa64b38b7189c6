"""Multi-GPU (one process per GPU, NCCL through the C ABI): row-sharded QR (TSQR) and Cholesky (one all-reduce of
[J'J | J'y]) solves, and a sharded LM run, against the single-process oracle.  Needs >= 2 GPUs (gpurun --gpus 2)."""
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
