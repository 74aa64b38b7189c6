"""Row-sharding of a tall dense Jacobian over ranks: host-side layout logic shared by bench.py,
the multi-GPU tests and the CPU (gloo) tests.

  * `row_partition(m, world)`            contiguous row blocks, sizes differ by at most one
  * `init_comm(ctx)`                     NCCL communicator for the C-ABI context; the 128-byte id travels over
                                         torch.distributed (plumbing only)
  * `stack_layout(n, world)`             the row-interleaved stack of the R_k factors and the sqrt(D) triangle that
                                         `lso_qr_solve_sharded` factorises on every rank
  * `packed_upper_layout(n)`             the packed [upper(J'J) | J'y] buffer of the sharded Cholesky path's all-reduce
  * `lsmr_packed_layout(n)`              the [J'u | ||u||²] buffer the row-sharded LSMR all-reduces once per iteration
"""
from __future__ import annotations


def row_partition(m: int, world: int):
    """[(row0, rows)] for each rank; every row appears exactly once, in order."""
    if world < 1 or m < world:
        raise ValueError("need at least one row per rank")
    edges = [(m * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1] - edges[r]) for r in range(world)]


def stack_layout(n: int, world: int, damped: bool = True):
    """TSQR stack as `stack_assemble_kernel` (csrc/dense_solve.cu) builds it: the P = world upper-triangular R_k and, when
    damped, diag(sqrt(damp)) as one more triangle, with their rows INTERLEAVED — stack row Q*r + i is row r of triangle i
    (Q = P + 1 with damping, else P).  Stack row rho then has no entry left of column rho // Q, which is what lets panel j
    of the replicated QR stop at row Q*32*(j+1) (`QRPlan::band`).  The right-hand side column is interleaved the same way
    ([Q_k'y_k] for the R triangles, 0 for the damping triangle).
    Returns {"Q", "rows", "cols", "row_of": f(triangle, r) -> stack row, "damp_triangle": index or None}."""
    Q = world + 1 if damped else world
    return {"Q": Q, "rows": Q * n, "cols": n + 1, "row_of": (lambda tri, r: Q * r + tri),
            "damp_triangle": world if damped else None}


def packed_upper_layout(n: int):
    """The buffer of the ONE all-reduce of the row-sharded Cholesky path (`chol_pack_kernel`, csrc/chol.cu):
    [upper(J'J) by columns | J'y]: entry (i, j), i <= j, at j*(j+1)//2 + i; J'y at n*(n+1)//2 + i; n(n+1)/2 + n doubles."""
    return {"len": n * (n + 1) // 2 + n, "index": (lambda i, j: j * (j + 1) // 2 + i), "rhs0": n * (n + 1) // 2}


def lsmr_packed_layout(n: int):
    """The buffer of the ONE all-reduce per iteration of `lso_lsmr_solve_sharded` (csrc/lsmr.cu): this rank's J_k'u_k (u not yet
    normalised) in [0, n) and its ||u_k||² at n; the damping part of u is replicated and is added AFTER the sum, once."""
    return {"len": n + 1, "adjoint0": 0, "sumsq": n}


def init_comm(ctx, dist=None):
    """Create the NCCL communicator of `ctx` from the torch.distributed process group (rank 0 makes the id)."""
    if dist is None:
        import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ctx
    from .device import Context
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    return ctx
