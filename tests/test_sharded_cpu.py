"""N > 1 host logic on CPU: world_size-2 gloo process group.  Checks the row partition, that the TSQR stack layout
used by lso_qr_solve_sharded reproduces the single-process solve (R_k from each rank's rows, all-gathered, stacked with
the sqrt(damp) rows), and that the packed [J'J | J'y] all-reduce reproduces the Cholesky solve."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import lsob200  # noqa: F401
    from lsob200.sharding import packed_upper_layout, row_partition, stack_layout
    from oracle import reference_port as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(42)          # same data on every rank, each takes its rows
    m, n = 1003, 17
    J = rng.standard_normal((m, n)) * np.exp2(rng.integers(-4, 5, n))
    y = rng.standard_normal(m)
    damp = np.einsum("ij,ij->j", J, J) / 10
    row0, rows = row_partition(m, world)[rank]
    Jk, yk = J[row0:row0 + rows], y[row0:row0 + rows]
    # ---- TSQR: local [R_k | c_k], all-gather, stack, QR ----
    Q, R = np.linalg.qr(np.hstack([Jk, yk[:, None]]), mode="reduced")
    Rk = np.triu(R[:n, :])                                   # n x (n+1): [R_k | Q_k' y_k]
    gathered = [torch.zeros(n, n + 1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(Rk.copy()))
    lay = stack_layout(n, world)                             # the INTERLEAVED layout the device kernel builds
    S = np.zeros((lay["rows"], lay["cols"]))
    for k in range(world):
        for r in range(n):
            S[lay["row_of"](k, r)] = gathered[k].numpy()[r]
    for r in range(n):
        S[lay["row_of"](lay["damp_triangle"], r), r] = np.sqrt(damp[r])
    # the band property the replicated QR relies on: stack row rho has no entry left of column rho // Q
    for rho in range(lay["rows"]):
        assert not S[rho, :min(rho // lay["Q"], n)].any()
    x_tsqr = np.linalg.lstsq(S[:, :n], S[:, n], rcond=None)[0]
    x_ref, _ = O.qr_ldiv(J, y, damp.copy())
    # ---- Cholesky: one all-reduce of the packed [J'J | J'y] ----
    pk = packed_upper_layout(n)                              # [upper(J'J) by columns | J'y]: n(n+1)/2 + n doubles
    G, g = Jk.T @ Jk, Jk.T @ yk
    buf = np.zeros(pk["len"])
    for j in range(n):
        for i in range(j + 1):
            buf[pk["index"](i, j)] = G[i, j]
    buf[pk["rhs0"]:] = g
    packed = torch.from_numpy(buf)
    dist.all_reduce(packed)
    tot = packed.numpy()
    C = np.zeros((n, n))
    for j in range(n):
        for i in range(j + 1):
            C[i, j] = C[j, i] = tot[pk["index"](i, j)]
    C += np.diag(damp)
    x_chol = np.linalg.solve(C, tot[pk["rhs0"]:])
    out[rank] = (float(np.linalg.norm(x_tsqr - x_ref) / np.linalg.norm(x_ref)),
                 float(np.linalg.norm(x_chol - x_ref) / np.linalg.norm(x_ref)), rows)
    dist.barrier()
    dist.destroy_process_group()


def test_row_partition():
    sys.path.insert(0, ROOT)
    from lsob200.sharding import lsmr_packed_layout, row_partition
    assert lsmr_packed_layout(500000) == {"len": 500001, "adjoint0": 0, "sumsq": 500000}
    for m, w in [(100000, 8), (1003, 2), (7, 7), (2_000_000, 8)]:
        parts = row_partition(m, w)
        assert parts[0][0] == 0 and sum(r for _, r in parts) == m
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert max(r for _, r in parts) - min(r for _, r in parts) <= 1
    with pytest.raises(ValueError):
        row_partition(3, 4)


def test_world2_gloo_tsqr_and_allreduce():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        e_tsqr, e_chol, rows = out[rank]
        assert e_tsqr <= 1e-12 and e_chol <= 1e-11
    assert sum(out[r][2] for r in range(world)) == 1003


def _lsmr_worker(rank, world, port, out):
    """Row-sharded LSMR (lso_lsmr_solve_sharded's algebra) on the oracle's own recurrences: u is sharded like the rows, the
    n-vectors are replicated, J'u and ||u||² are summed over the ranks; the damping part of u is replicated and counted once."""
    sys.path.insert(0, ROOT)
    import math
    import torch
    import scipy.sparse as sp
    from lsob200.sharding import row_partition
    from oracle import reference_port as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, n = 2003, 150
    J = sp.random(m, n, density=0.03, random_state=11, format="csr")
    rng = np.random.default_rng(3)
    y = rng.standard_normal(m)
    res = {}
    for damped in (True, False):
        damp = np.asarray(J.multiply(J).sum(axis=0)).ravel() / 10 + 1e-3
        x_ref, _, it_ref, istop_ref = O.lsmr_ldiv(J.tocsc(), y, damp.copy() if damped else None)
        row0, rows = row_partition(m, world)[rank]
        Jk, yk = J[row0:row0 + rows], y[row0:row0 + rows]

        def allsum(a):
            t = torch.from_numpy(np.atleast_1d(np.asarray(a, dtype=np.float64)).copy())
            dist.all_reduce(t)
            return t.numpy()

        class Shard(O._Precond):                       # vectors: [y_k (this rank's rows) ; x (replicated)]
            def mul_t(self, b, a, alpha, beta):
                tmp = allsum(self.J.T @ a[:self.m])     # the ONE exchange of an iteration (n doubles) ...
                if self.diag is not None:
                    tmp = tmp + 1.0 * a[self.m:] * self.diag
                tmp2 = tmp * self.P
                if beta != 1:
                    b *= beta
                b += alpha * tmp2
                return b

            def norm(self, b):                          # ... with ||u_k||² riding along
                sy = float(allsum(np.dot(b[:self.m], b[:self.m]))[0])
                sx = float(np.dot(b[self.m:], b[self.m:])) if self.diag is not None else 0.0
                return math.sqrt(sy + sx)

        P = allsum(np.asarray(Jk.multiply(Jk).sum(axis=0)).ravel())
        if damped:
            P = P + damp
        P = np.where(P > 0, 1.0 / np.sqrt(np.where(P > 0, P, 1.0)), 0.0)
        A = Shard(Jk, P, np.sqrt(damp) if damped else None)
        b = np.concatenate([yk, np.zeros(n)]) if damped else yk.copy()
        x, it, istop = O.lsmr(np.zeros(n), A, b, btol=0.5 if damped else 1e-6, maxiter=max(m + (n if damped else 0), n))
        x = x * P
        res[damped] = (it, it_ref, istop, istop_ref, float(np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref)))
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_row_sharded_lsmr():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_lsmr_worker, args=(world, port, out), nprocs=world, join=True)
    for rank in range(world):
        for damped in (True, False):
            it, it_ref, istop, istop_ref, err = out[rank][damped]
            assert (it, istop) == (it_ref, istop_ref) and err <= 1e-10, (rank, damped, out[rank][damped])
