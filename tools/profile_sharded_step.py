"""torchrun --nproc-per-node N tools/profile_sharded_step.py : wall time of each operation of the sharded LM(QR) step
on rank 0 (each bracketed by stream syncs, so the sum is an upper bound of the un-instrumented step)."""
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import lsob200 as L
import bench
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
ctx = L.Context(lr)
if world > 1:
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", lr))
    uid = [L.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
m, n = 100000, 1000
rows = [(m * r) // world for r in range(world + 1)]
row0, m_loc = rows[rank], rows[rank + 1] - rows[rank]
prob = bench.DeviceProblem(L, ctx, m_loc, n, row0, bench.SEED)
x = L.DeviceVector(ctx, n).copyto(prob.x0); y = L.DeviceVector(ctx, m_loc); J = L.DenseMatrix(ctx, m_loc, n)
nls = L.LeastSquaresProblem(x=x, y=y, f_=prob.f_, g_=prob.g_, J=J, device_callbacks=True, ctx=ctx)
anls = L.allocate(nls, L.LevenbergMarquardt(L.QR()), sharded=(world > 1))
run = L.LMRun(anls)
for _ in range(3): run.iterate()
acc = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        ctx.sync(); t0 = time.perf_counter(); r = f(*a, **k); ctx.sync(); acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
wrap(anls, "g", "g!"); wrap(anls, "f", "f!"); wrap(J, "colsumabs2_and_grad", "colsumabs2+J'f"); wrap(ctx, "allreduce", "allreduce (n-vectors, scalars)")
wrap(anls.solver, "ldiv", "ldiv (local QR + allgather + stack QR)")
import lsob200.api as _api
wrap(_api, "_step_tail", "step tail (ssr, predicted ssr, maxabs; 1 allreduce of 2 scalars; 1 sync)")
wrap(_api, "_box_project", "box projection"); wrap(_api, "_lm_damping", "damping")
K = 6
ctx.sync(); t0 = time.perf_counter()
for _ in range(K): run.iterate()
ctx.sync(); tot = (time.perf_counter() - t0) / K * 1e3
if rank == 0:
    print(f"world {world}: instrumented step {tot:.2f} ms")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]): print(f"  {k:45s} {v / K * 1e3:7.3f} ms")
    print(f"  {'(rest: vector kernels, host logic)':45s} {tot - sum(acc.values()) / K * 1e3 + acc.get('allreduce (n-vectors, scalars)', 0) / K * 1e3 * 0:7.3f} ms")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
