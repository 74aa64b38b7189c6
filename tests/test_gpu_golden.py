"""GPU plugin against the committed golden fixtures (no oracle import needed at run time): per-solve δ and the
LM / Dogleg runs of the reference's test problems."""
import json
import os

import numpy as np
import pytest

import problems as P

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_solves_match_golden(ctx):
    import lsob200 as L
    sol = json.load(open(os.path.join(G, "solves.json")))
    for key, c in sol.items():
        if key.startswith("_"):
            continue
        Jh, yh = np.asfortranarray(c["J"]), np.array(c["y"])
        m, n = Jh.shape
        J, y, x = L.DenseMatrix(ctx, m, n, Jh), L.DeviceVector(ctx, m, yh), L.DeviceVector(ctx, n)
        if key.startswith("rank_deficient"):
            ws = L.DenseQRAllocatedSolver(ctx, m, n, damped=False)
            ws.ldiv(x, J, y)
            assert ws.last_rank == c["rank"] and rel(x.download(), c["qr_undamped"]) <= 1e-9
            continue
        damp = np.array(c["damp"])
        L.DenseQRAllocatedSolver(ctx, m, n, True).ldiv(x, J, y, L.DeviceVector(ctx, n, damp))
        assert rel(x.download(), c["qr_damped"]) <= 1e-10
        L.DenseQRAllocatedSolver(ctx, m, n, False).ldiv(x, J, y)
        assert rel(x.download(), c["qr_undamped"]) <= 1e-10
        L.DenseCholeskyAllocatedSolver(ctx, m, n, True).ldiv(x, J, y, L.DeviceVector(ctx, n, damp))
        assert rel(x.download(), c["chol_damped"]) <= 1e-10
        ws = L.LSMRDampenedAllocatedSolver(ctx, m, n)
        ws.ldiv(x, J, y, L.DeviceVector(ctx, n, damp))
        assert (ws.last_iters, ws.last_istop) == (c["lsmr_damped"]["iters"], c["lsmr_damped"]["istop"])
        assert rel(x.download(), c["lsmr_damped"]["x"]) <= 2e-5
        ws = L.LSMRAllocatedSolver(ctx, m, n)
        ws.ldiv(x, J, y)
        assert (ws.last_iters, ws.last_istop) == (c["lsmr_undamped"]["iters"], c["lsmr_undamped"]["istop"])


def test_runs_match_golden(ctx):
    """Iteration / call counts of the GPU path equal the frozen oracle's on the well-conditioned runs; ssr agrees."""
    import lsob200 as L
    gold = json.load(open(os.path.join(G, "optimize_runs.json")))
    optc = {"dogleg": L.Dogleg, "lm": L.LevenbergMarquardt}
    checked = 0
    name, f, g, x0 = P.readme_rosenbrock()
    for opt in ("dogleg", "lm"):
        r = L.optimize_(L.LeastSquaresProblem(x=x0.copy(), y=np.zeros(2), f_=f, g_=g, J=np.zeros((2, 2), order="F")),
                        optc[opt](L.QR()))
        e = gold[f"readme_rosenbrock/{opt}/qr"]
        assert (r.iterations, r.f_calls, r.g_calls, r.mul_calls) == (e["iterations"], e["f_calls"], e["g_calls"], e["mul_calls"])
        assert rel(r.minimizer, e["minimizer"]) <= 1e-9
        checked += 1
    import json as _json
    allowed = _json.load(open(os.path.join(G, "iteration_allowlist.json")))
    mism = {}
    for i, (name, f, g, x0) in enumerate(P.minpack_cholesky()):
        for opt in ("dogleg", "lm"):
            e = gold[f"minpack_cholesky/{i:02d}_{name}_{x0.size}/{opt}"]
            n = x0.size
            r = L.optimize_(L.LeastSquaresProblem(x=x0.copy(), y=np.zeros(n), f_=f, g_=g, J=np.zeros((n, n), order="F")),
                            optc[opt](L.Cholesky()))
            assert r.converged == e["converged"] and r.ssr <= 1e-3
            if r.iterations != e["iterations"]:
                mism[f"{opt}/{name}_{n}"] = (r.iterations, e["iterations"])
            checked += 1
    unexpected = {k: v for k, v in mism.items() if k.split("/", 1)[1] not in allowed.get(f"minpack/{k.split('/')[0]}/cholesky/dense", {})}
    assert not unexpected, unexpected
