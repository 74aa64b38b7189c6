"""Bit-reproducibility sweep: many random shapes, 4 solves each (damped + undamped QR), also vs. the plain-FMA
cross-check update kernel for shapes it can handle.  Prints failures only."""
import sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
ctx = L.Context.default(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
nshape = int(sys.argv[2]) if len(sys.argv) > 2 else 60
bad = 0
for it in range(nshape):
    n = int(rng.choice([16, 32, 40, 64, 96, 128, 200, 300, 500]))
    m = int(rng.integers(max(n, 256), 120000 if n <= 128 else 60000))
    Jh = np.asfortranarray(rng.standard_normal((m, n))); yh = rng.standard_normal(m)
    damp = np.einsum("ij,ij->j", Jh, Jh) / 10.0
    J, y, d, x = L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    for damped in (True, False):
        ws = L.DenseQRAllocatedSolver(ctx, m, n, damped=damped)
        xs = []
        for _ in range(4):
            ws.ldiv(x, J, y, d if damped else None); xs.append(x.download())
        same = all(np.array_equal(xs[0], v) for v in xs)
        # residual check of the (damped) normal equations instead of a CPU reference
        r = Jh.T @ (Jh @ xs[0] - yh) + (damp * xs[0] if damped else 0.0)
        ok = np.linalg.norm(r) <= 1e-8 * np.linalg.norm(Jh.T @ yh)
        if not (same and ok):
            bad += 1
            print("FAIL", m, n, "damped" if damped else "undamped", "identical", same, "normal-eq residual ok", ok, flush=True)
        del ws
    del J, y, d, x
print("shapes", nshape, "failures", bad)
