"""Row-sharding of a tall dense Jacobian over ranks: host-side layout logic shared by bench.py,
the multi-GPU tests and the CPU (gloo) tests.

  * `row_partition(m, world)`            contiguous row blocks, sizes differ by at most one
  * `init_comm(ctx)`                     NCCL communicator for the C-ABI context; the 128-byte id travels over
                                         torch.distributed (plumbing only)
  * `stack_layout(n, world)`             row offsets of the R_k factors / sqrt(D) block in the TSQR stack that
                                         `lso_qr_solve_sharded` factorises on every rank
"""
from __future__ import annotations


def row_partition(m: int, world: int):
    """[(row0, rows)] for each rank; every row appears exactly once, in order."""
    if world < 1 or m < world:
        raise ValueError("need at least one row per rank")
    edges = [(m * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1] - edges[r]) for r in range(world)]


def stack_layout(n: int, world: int):
    """TSQR stack: rows [k*n, (k+1)*n) hold R_k (upper triangular), rows [world*n, world*n + n) hold diag(sqrt(damp)),
    the right-hand side column holds [Q_0'y_0; ...; Q_{P-1}'y_{P-1}; 0]."""
    return {"R_rows": [(k * n, (k + 1) * n) for k in range(world)], "damp_rows": (world * n, world * n + n),
            "rows": world * n + n, "cols": n + 1}


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
