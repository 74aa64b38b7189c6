"""Generates tests/golden/*.json from the ORACLE (oracle/reference_port.py).  Oracle-derived: no Julia is available
in this image, so these are NOT outputs of the reference itself; they freeze the oracle's behaviour
on the reference's own test problems so that a change in the oracle, LAPACK build or the GPU path is caught.

    python oracle/make_golden.py        # rewrites tests/golden/
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import problems as P  # noqa: E402
from oracle import reference_port as O  # noqa: E402
from oracle import synth_ref as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run(f, g, x0, m, opt, solver, sparse=False, **kw):
    n = x0.size
    if sparse:
        J, gg = P.dense_pattern_csc(n), P.sparse_adapter(g, n)
    else:
        J, gg = np.zeros((m, n), order="F"), g
    r = O.optimize(f, gg, x0.copy(), J, m, optimizer=opt, solver=solver, record=True, **kw)
    return {"iterations": r.iterations, "f_calls": r.f_calls, "g_calls": r.g_calls, "mul_calls": r.mul_calls,
            "ssr": r.ssr, "converged": bool(r.converged), "x_converged": bool(r.x_converged),
            "f_converged": bool(r.f_converged), "g_converged": bool(r.g_converged),
            "minimizer": [float(v) for v in r.minimizer],
            "first_deltas": [[float(v) for v in d] for d in r.deltas[:3]]}


def main():
    os.makedirs(OUT, exist_ok=True)
    gold = {"_note": "oracle-derived (oracle/reference_port.py via scipy/OpenBLAS); not produced by Julia"}
    name, f, g, x0 = P.readme_rosenbrock()
    for opt in ("dogleg", "lm"):
        gold[f"readme_rosenbrock/{opt}/qr"] = run(f, g, x0, 2, opt, "qr")
    for i, (name, f, g, x0) in enumerate(P.minpack_all()):
        for opt in ("dogleg", "lm"):
            for solver, sparse in (("qr", False), ("lsmr", False), ("lsmr", True)):
                gold[f"minpack/{i:02d}_{name}_{x0.size}/{opt}/{solver}{'_sparse' if sparse else ''}"] = \
                    run(f, g, x0, x0.size, opt, solver, sparse)
    for i, (name, f, g, x0) in enumerate(P.minpack_cholesky()):
        for opt in ("dogleg", "lm"):
            gold[f"minpack_cholesky/{i:02d}_{name}_{x0.size}/{opt}"] = run(f, g, x0, x0.size, opt, "cholesky")
    name, f, g, x0 = P.factor()
    for opt in ("dogleg", "lm"):
        gold[f"factor/{opt}/qr"] = run(f, g, x0, 9, opt, "qr")
    for name, f, g, x0, kw, xs in P.bounds_cases():
        for opt in ("dogleg", "lm"):
            gold[f"bounds/{name}/{opt}"] = run(f, g, x0, 2, opt, "qr", **kw)
    json.dump(gold, open(os.path.join(OUT, "optimize_runs.json"), "w"), indent=1)

    # per-solve vectors on seeded inputs (dense solves + LSMR), small enough to commit
    rng = np.random.default_rng(20240607)
    solves = {"_note": gold["_note"]}
    for (m, n) in [(9, 6), (60, 17), (300, 33)]:
        J = rng.standard_normal((m, n)) * np.exp2(rng.integers(-4, 5, n))
        y = rng.standard_normal(m)
        dtd = np.einsum("ij,ij->j", J, J)
        damp = np.clip(dtd, 1e-6 * dtd.mean(), 1e32 * dtd.mean()) / 10
        xl, nmul, it, istop = O.lsmr_ldiv(J, y, damp.copy())
        xu, nmulu, itu, istopu = O.lsmr_ldiv(J, y)
        solves[f"{m}x{n}"] = {"J": J.tolist(), "y": y.tolist(), "damp": damp.tolist(),
                              "qr_damped": O.qr_ldiv(J, y, damp)[0].tolist(), "qr_undamped": O.qr_ldiv(J, y)[0].tolist(),
                              "chol_damped": O.chol_ldiv(J, y, damp.copy()).tolist(),
                              "lsmr_damped": {"x": xl.tolist(), "iters": it, "istop": istop},
                              "lsmr_undamped": {"x": xu.tolist(), "iters": itu, "istop": istopu}}
    # rank-deficient: minimum-norm solution and rank
    B = rng.standard_normal((40, 5)) @ rng.standard_normal((5, 12))
    y = rng.standard_normal(40)
    x, rank = O.qr_ldiv(B, y)
    solves["rank_deficient_40x12"] = {"J": B.tolist(), "y": y.tolist(), "qr_undamped": x.tolist(), "rank": rank}
    json.dump(solves, open(os.path.join(OUT, "solves.json"), "w"))

    # generator known-answers (first entries) so the CPU replica itself is pinned
    gen = {"dense_matrix_5x3_seed99_row7": S.dense_matrix(5, 3, 99, row_offset=7).tolist(),
           "vector_6_seed5": S.vector(6, 5, 0.25, offset=3).tolist(),
           "csc_pattern_20x4x3_seed7": [a.tolist() for a in S.csc_pattern(20, 4, 3, 7)]}
    json.dump(gen, open(os.path.join(OUT, "generators.json"), "w"), indent=1)
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    main()
