import sys
import numpy as np
sys.path.insert(0, ".")
import lsob200 as L
ctx = L.Context.default(0)
def run(m, n, damped, reps=8, seed=5):
    rng = np.random.default_rng(seed)
    Jh = np.asfortranarray(rng.standard_normal((m, n))); yh = rng.standard_normal(m)
    dtd = np.einsum("ij,ij->j", Jh, Jh); damp = dtd / 10.0
    ws = L.DenseQRAllocatedSolver(ctx, m, n, damped=damped)
    J, y, d, x = L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n, damp), L.DeviceVector(ctx, n)
    xs = []
    Rs = []
    for rep in range(reps):
        ws.ldiv(x, J, y, d if damped else None); xs.append(x.download()); Rs.append(ws.factor())
    nd = len({v.tobytes() for v in xs}); ndr = len({v.tobytes() for v in Rs})
    diff = max(np.abs(v - xs[0]).max() for v in xs)
    print(m, n, "damped" if damped else "undamped", "distinct x:", nd, "distinct R:", ndr, "max |dx|: %.2e" % diff, flush=True)
for (m, n) in ((49059, 300), (49059, 300), (49059, 300), (30011, 300), (70003, 200)):
    run(m, n, False); run(m, n, True)
