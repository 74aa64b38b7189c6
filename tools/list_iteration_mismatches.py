"""Run the GPU path and the oracle on every (problem, optimizer, solver) pair of tests/test_gpu_optimize.py /
test_gpu_golden.py and list the runs whose outer-iteration counts differ, with the diagnosis inputs the allow-list in
tests/golden/iteration_allowlist.json cites (cond of the final Jacobian, final ssr, both counts).
    python tools/list_iteration_mismatches.py > gpurun_out/iteration_mismatches.json"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import problems as P  # noqa: E402
from oracle import reference_port as O  # noqa: E402
import lsob200 as L  # noqa: E402

out = []
optc = {"dogleg": L.Dogleg, "lm": L.LevenbergMarquardt}
solc = {"qr": L.QR, "cholesky": L.Cholesky, "lsmr": L.LSMR}
for solver, dense in (("qr", True), ("lsmr", True), ("lsmr", False), ("cholesky", True)):
    probs = P.minpack_cholesky() if solver == "cholesky" else P.minpack_all()
    for opt in ("lm", "dogleg"):
        for name, f, g, x0 in probs:
            n = x0.size
            if dense:
                Jg, Jo, gg = np.zeros((n, n), order="F"), np.zeros((n, n), order="F"), g
            else:
                Jg, Jo, gg = P.dense_pattern_csc(n), P.dense_pattern_csc(n), P.sparse_adapter(g, n)
            rg = L.optimize_(L.LeastSquaresProblem(x=x0.copy(), y=np.zeros(n), f_=f, g_=gg, J=Jg), optc[opt](solc[solver]()))
            ro = O.optimize(f, gg, x0.copy(), Jo, n, optimizer=opt, solver=solver)
            if rg.iterations != ro.iterations:
                Jf = np.zeros((n, n), order="F")
                g(Jf, ro.minimizer)
                sv = np.linalg.svd(Jf, compute_uv=False)
                out.append({"set": f"minpack/{opt}/{solver}/{'dense' if dense else 'csc'}", "problem": f"{name}_{n}",
                            "gpu_iterations": rg.iterations, "oracle_iterations": ro.iterations, "gpu_ssr": rg.ssr,
                            "oracle_ssr": ro.ssr, "cond_J_final": float(sv[0] / max(sv[-1], 1e-300)),
                            "gpu_converged": bool(rg.converged), "oracle_converged": bool(ro.converged),
                            "minimizer_diff": float(np.linalg.norm(rg.minimizer - ro.minimizer))})
print(json.dumps(out, indent=1))
