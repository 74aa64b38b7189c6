"""ORACLE (TEST INFRASTRUCTURE): reference-held known answers for the QR path.

The only per-problem known answers the reference's own tests hold are the NIST StRD nonlinear-regression datasets
embedded in /root/reference/test/nonlinearfitting.jl (:6-1445): observations, the NIST start columns and the NIST
CERTIFIED parameter values (`solution`).  The reference's loop (:1457-1472) runs Dogleg(QR()) and
LevenbergMarquardt(QR()) from every start column with x_tol = 1e-50, f_tol = 1e-36, g_tol = 1e-50 and counts
norm(minimizer - solution) <= 1e-3.

This script reads the numeric tables out of that file IN THIS CONTAINER (the reference does not travel to the GPU box)
and freezes them as tests/golden/nist_strd.json.  Only numbers are taken (NIST's public data); the model functions are
restated, with analytic Jacobians, in tests/problems.py.

    python oracle/make_golden_nist.py      ->  tests/golden/nist_strd.json
"""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/test/nonlinearfitting.jl"


def _matrix(text):
    rows = []
    for line in re.split(r"[;\n]", text):
        line = line.split("#")[0].strip().strip(",")
        if not line:
            continue
        rows.append([float(t) for t in re.split(r"[\s,]+", line) if t])
    return rows


def _block(body, key):
    m = re.search(r"\b" + key + r"\s*=\s*\[", body)
    if not m:
        raise ValueError(key)
    end = body.index("]", m.end())
    return body[m.end():end], m.start()


def main():
    src = open(SRC).read()
    lines = src.split("\n")
    out = {"source": "test/nonlinearfitting.jl (NIST StRD nls datasets, certified values)", "problems": []}
    starts = [i for i, l in enumerate(lines) if re.match(r"^function \w+\(\)", l)]
    for k, i0 in enumerate(starts):
        i1 = starts[k + 1] if k + 1 < len(starts) else len(lines)
        body = "\n".join(lines[i0:i1])
        nm = re.search(r'name\s*=\s*"(\w+)"', body)
        if not nm:
            continue
        data, _ = _block(body, "data")
        par, _ = _block(body, "parameters")
        sol, pos = _block(body, "solution")
        sol = [v for row in _matrix(sol) for v in row]
        P = _matrix(par)
        assert all(len(r) == len(P[0]) for r in P) and len(P) == len(sol), nm.group(1)
        D = _matrix(data)
        assert all(len(r) == 2 for r in D), nm.group(1)
        out["problems"].append({
            "name": nm.group(1), "line": i0 + 1, "solution_line": i0 + 1 + body[:pos].count("\n"),
            "y": [r[0] for r in D], "x": [r[1] for r in D],
            "starts": [[P[i][j] for i in range(len(P))] for j in range(len(P[0]))],
            "certified": sol,
        })
    path = os.path.join(ROOT, "tests", "golden", "nist_strd.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", path, [(p["name"], len(p["y"]), len(p["starts"])) for p in out["problems"]])


if __name__ == "__main__":
    main()
