"""Randomised shape sweep of the damped / undamped QR solve and the Cholesky solve against the oracle (determinism + parity)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import lsob200 as L
from oracle import reference_port as O
ctx = L.Context.default(0)
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
if len(sys.argv) > 3: ctx.set_option("qr_lookahead", int(sys.argv[3]))
if len(sys.argv) > 4: ctx.set_option("qr_apply", int(sys.argv[4]))
shapes = []
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    n = int(rng.choice([1, 2, 7, 31, 32, 33, 63, 64, 65, 96, 100, 127, 129, 200, 257, 300]))
    m = int(n + rng.integers(0, 5)) if rng.random() < 0.2 else int(rng.integers(n, 60000 if n > 100 else 400000))
    shapes.append((m, n))
shapes += [(255, 32), (256, 32), (257, 32), (287, 32), (288, 32), (289, 32), (8192, 32), (8193, 33), (65536, 64), (2047, 1), (33, 33), (64, 64), (296 * 256 - 32, 32), (296 * 256 + 1, 32)]
for (m, n) in shapes:
    Jh = np.asfortranarray(rng.standard_normal((m, n)) * np.exp2(rng.integers(-6, 7, n))); yh = rng.standard_normal(m)
    dtd = np.einsum("ij,ij->j", Jh, Jh)
    damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) / 10.0
    J, y, d, x = L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    for name, ws, dd, ref in (("qr damped", L.DenseQRAllocatedSolver(ctx, m, n, True), d, lambda: O.qr_ldiv(Jh, yh, damp)[0]),
                              ("qr undamped", L.DenseQRAllocatedSolver(ctx, m, n, False), None, lambda: O.qr_ldiv(Jh, yh, None)[0]),
                              ("chol damped", L.DenseCholeskyAllocatedSolver(ctx, m, n, True), d, lambda: O.qr_ldiv(Jh, yh, damp)[0])):
        xs = []
        for rep in range(2):
            ws.ldiv(x, J, y, dd); xs.append(x.download())
        xr = ref()
        err = np.linalg.norm(xs[0] - xr) / max(np.linalg.norm(xr), 1e-300)
        tol = 1e-10 if name.startswith("qr") else 1e-6
        ok = np.array_equal(xs[0], xs[1]) and err <= tol
        if not ok:
            bad += 1
            print("FAIL", name, m, n, "err %.2e" % err, "identical", np.array_equal(xs[0], xs[1]), flush=True)
    del J, y, d, x
print("shapes", len(shapes), "failures", bad)
