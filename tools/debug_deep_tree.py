import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import lsob200 as L
from oracle import reference_port as O
ctx = L.Context.default(0)
def run(m, n, kern):
    ctx.set_option("qr_apply", kern)
    rng = np.random.default_rng(m * 7 + n)
    Jh = np.asfortranarray(rng.standard_normal((m, n))); yh = rng.standard_normal(m)
    dtd = np.einsum("ij,ij->j", Jh, Jh); damp = dtd / 10.0
    ws = L.DenseQRAllocatedSolver(ctx, m, n, damped=True)
    J, y, d, x = L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    xs = []
    for rep in range(3):
        ws.ldiv(x, J, y, d); xs.append(x.download())
    xr, _ = O.qr_ldiv(Jh, yh, damp)
    errs = [np.linalg.norm(v - xr) / np.linalg.norm(xr) for v in xs]
    print(m, n, "kernel", kern, "rel err per rep", ["%.1e" % e for e in errs], "identical", all(np.array_equal(xs[0], v) for v in xs), flush=True)
for m in (70000, 100000, 160000, 320000, 640000, 1100000):
    for kern in (2, 1):
        run(m, 32, kern)
run(300000, 64, 2)
