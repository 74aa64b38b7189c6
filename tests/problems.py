"""Test problems the reference's own test-suite runs through the hot path, restated with 0-based numpy:

  * the MINPACK `hybrj` set (test/nonlinearsolvers.jl:7-501; the list at :512-522), analytic f! / g!
  * the rank-deficient factor model (test/nonlinearleastsquares.jl:7-89)
  * README Rosenbrock (README.md:13-18,67-80; test/runtests.jl:20-41) and the bounds cases (test/bounds.jl)

Every problem is a tuple (name, f_, g_, x0) with f_(fvec, x) and g_(fjac, x) writing in place, as in Julia.
"""
import math

import numpy as np


def rosenbrock():
    def f(fvec, x):
        fvec[0] = 1 - x[0]
        fvec[1] = 10 * (x[1] - x[0] ** 2)

    def g(J, x):
        J[0, 0] = -1
        J[0, 1] = 0
        J[1, 0] = -20 * x[0]
        J[1, 1] = 10

    return "rosenbrock", f, g, np.array([-1.2, 1.0])


def powell_singular():
    def f(fvec, x):
        fvec[0] = x[0] + 10 * x[1]
        fvec[1] = math.sqrt(5) * (x[2] - x[3])
        fvec[2] = (x[1] - 2 * x[2]) ** 2
        fvec[3] = math.sqrt(10) * (x[0] - x[3]) ** 2

    def g(J, x):
        J[:] = 0
        J[0, 0] = 1
        J[0, 1] = 10
        J[1, 2] = math.sqrt(5)
        J[1, 3] = -J[1, 2]
        J[2, 1] = 2 * (x[1] - 2 * x[2])
        J[2, 2] = -2 * J[2, 1]
        J[3, 0] = 2 * math.sqrt(10) * (x[0] - x[3])
        J[3, 3] = -J[3, 0]

    return "powell_singular", f, g, np.array([3.0, -1.0, 0.0, 1.0])


def powell_badly_scaled():
    c1, c2 = 1e4, 1.0001

    def f(fvec, x):
        fvec[0] = c1 * x[0] * x[1] - 1
        fvec[1] = math.exp(-x[0]) + math.exp(-x[1]) - c2

    def g(J, x):
        J[0, 0] = c1 * x[1]
        J[0, 1] = c1 * x[0]
        J[1, 0] = -math.exp(-x[0])
        J[1, 1] = -math.exp(-x[1])

    return "powell_badly_scaled", f, g, np.array([0.0, 1.0])


def wood():
    c3, c4, c5, c6 = 2e2, 2.02e1, 1.98e1, 1.8e2

    def f(fvec, x):
        t1 = x[1] - x[0] ** 2
        t2 = x[3] - x[2] ** 2
        fvec[0] = -c3 * x[0] * t1 - (1 - x[0])
        fvec[1] = c3 * t1 + c4 * (x[1] - 1) + c5 * (x[3] - 1)
        fvec[2] = -c6 * x[2] * t2 - (1 - x[2])
        fvec[3] = c6 * t2 + c4 * (x[3] - 1) + c5 * (x[1] - 1)

    def g(J, x):
        J[:] = 0
        t1 = x[1] - 3 * x[0] ** 2
        t2 = x[3] - 3 * x[2] ** 2
        J[0, 0] = -c3 * t1 + 1
        J[0, 1] = -c3 * x[0]
        J[1, 0] = -2 * c3 * x[0]
        J[1, 1] = c3 + c4
        J[1, 3] = c5
        J[2, 2] = -c6 * t2 + 1
        J[2, 3] = -c6 * x[2]
        J[3, 1] = c5
        J[3, 2] = -2 * c6 * x[2]
        J[3, 3] = c6 + c4

    return "wood", f, g, np.array([-3.0, -1.0, -3.0, -1.0])


def helical_valley():
    tpi = 8 * math.atan(1)
    c7, c8 = 2.5e-1, 5e-1

    def f(fvec, x):
        if x[0] > 0:
            t1 = math.atan(x[1] / x[0]) / tpi
        elif x[0] < 0:
            t1 = math.atan(x[1] / x[0]) / tpi + c8
        else:
            t1 = c7 * np.sign(x[1])
        t2 = math.sqrt(x[0] ** 2 + x[1] ** 2)
        fvec[0] = 10 * (x[2] - 10 * t1)
        fvec[1] = 10 * (t2 - 1)
        fvec[2] = x[2]

    def g(J, x):
        t = x[0] ** 2 + x[1] ** 2
        t1 = tpi * t
        t2 = math.sqrt(t)
        J[0, 0] = 100 * x[1] / t1
        J[0, 1] = -100 * x[0] / t1
        J[0, 2] = 10
        J[1, 0] = 10 * x[0] / t2
        J[1, 1] = 10 * x[1] / t2
        J[1, 2] = 0
        J[2, 0] = 0
        J[2, 1] = 0
        J[2, 2] = 1

    return "helical_valley", f, g, np.array([-1.0, 0.0, 0.0])


def watson(n):
    c9 = 2.9e1

    def f(fvec, x):
        fvec[:] = 0
        for i in range(1, 30):
            ti = i / c9
            sum1, temp = 0.0, 1.0
            for j in range(2, n + 1):
                sum1 += (j - 1) * temp * x[j - 1]
                temp *= ti
            sum2, temp = 0.0, 1.0
            for j in range(1, n + 1):
                sum2 += temp * x[j - 1]
                temp *= ti
            temp1 = sum1 - sum2 ** 2 - 1
            temp2 = 2 * ti * sum2
            temp = 1 / ti
            for k in range(1, n + 1):
                fvec[k - 1] += temp * (k - 1 - temp2) * temp1
                temp *= ti
        temp = x[1] - x[0] ** 2 - 1
        fvec[0] += x[0] * (1 - 2 * temp)
        fvec[1] += temp

    def g(J, x):
        J[:] = 0
        for i in range(1, 30):
            ti = i / c9
            sum1, temp = 0.0, 1.0
            for j in range(2, n + 1):
                sum1 += (j - 1) * temp * x[j - 1]
                temp *= ti
            sum2, temp = 0.0, 1.0
            for j in range(1, n + 1):
                sum2 += temp * x[j - 1]
                temp *= ti
            temp1 = 2 * (sum1 - sum2 ** 2 - 1)
            temp2 = 2 * sum2
            temp = ti ** 2
            tk = 1.0
            for k in range(1, n + 1):
                tj = tk
                for j in range(k, n + 1):
                    J[k - 1, j - 1] += tj * (((k - 1) / ti - temp2) * ((j - 1) / ti - temp2) - temp1)
                    tj *= ti
                tk *= temp
        J[0, 0] += 6 * x[0] ** 2 - 2 * x[1] + 3
        J[0, 1] -= 2 * x[0]
        J[1, 1] += 1
        for k in range(n):
            for j in range(k, n):
                J[j, k] = J[k, j]

    return "watson", f, g, np.zeros(n)


def chebyquad(n):
    tk = 1 / n

    def f(fvec, x):
        fvec[:] = 0
        for j in range(n):
            temp1 = 1.0
            temp2 = 2 * x[j] - 1
            temp = 2 * temp2
            for i in range(n):
                fvec[i] += temp2
                ti = temp * temp2 - temp1
                temp1 = temp2
                temp2 = ti
        iev = -1.0
        for k in range(1, n + 1):
            fvec[k - 1] *= tk
            if iev > 0:
                fvec[k - 1] += 1 / (k ** 2 - 1)
            iev = -iev

    def g(J, x):
        for j in range(n):
            temp1 = 1.0
            temp2 = 2 * x[j] - 1
            temp = 2 * temp2
            temp3 = 0.0
            temp4 = 2.0
            for k in range(n):
                J[k, j] = tk * temp4
                ti = 4 * temp2 + temp * temp4 - temp3
                temp3 = temp4
                temp4 = ti
                ti = temp * temp2 - temp1
                temp1 = temp2
                temp2 = ti

    return "chebyquad", f, g, np.arange(1, n + 1) / (n + 1)


def brown_almost_linear(n):
    def f(fvec, x):
        sum1 = np.sum(x) - (n + 1)
        for k in range(n - 1):
            fvec[k] = x[k] + sum1
        fvec[n - 1] = np.prod(x) - 1

    def g(J, x):
        J[:] = 1
        J[np.arange(n), np.arange(n)] = 2
        prd = np.prod(x)
        for j in range(n):
            if x[j] == 0.0:
                J[n - 1, j] = 1.0
                for k in range(n):
                    if k != j:
                        J[n - 1, j] *= x[k]
            else:
                J[n - 1, j] = prd / x[j]

    return "brown_almost_linear", f, g, 0.5 * np.ones(n)


def discrete_boundary_value(n):
    h = 1 / (n + 1)

    def f(fvec, x):
        for k in range(1, n + 1):
            temp = (x[k - 1] + k * h + 1) ** 3
            temp1 = x[k - 2] if k != 1 else 0.0
            temp2 = x[k] if k != n else 0.0
            fvec[k - 1] = 2 * x[k - 1] - temp1 - temp2 + temp * h ** 2 / 2

    def g(J, x):
        for k in range(1, n + 1):
            temp = 3 * (x[k - 1] + k * h + 1) ** 2
            J[k - 1, :] = 0
            J[k - 1, k - 1] = 2 + temp * h ** 2 / 2
            if k != 1:
                J[k - 1, k - 2] = -1
            if k != n:
                J[k - 1, k] = -1

    x = np.arange(1, n + 1) * h
    return "discrete_boundary_value", f, g, x * (x - 1)


def discrete_integral_equation(n):
    h = 1 / (n + 1)

    def f(fvec, x):
        for k in range(1, n + 1):
            tk = k * h
            sum1 = 0.0
            for j in range(1, k + 1):
                tj = j * h
                sum1 += tj * (x[j - 1] + tj + 1) ** 3
            sum2 = 0.0
            for j in range(k + 1, n + 1):
                tj = j * h
                sum2 += (1 - tj) * (x[j - 1] + tj + 1) ** 3
            fvec[k - 1] = x[k - 1] + h * ((1 - tk) * sum1 + tk * sum2) / 2

    def g(J, x):
        for k in range(1, n + 1):
            tk = k * h
            for j in range(1, n + 1):
                tj = j * h
                J[k - 1, j - 1] = h * min(tj * (1 - tk), tk * (1 - tj)) * 3 * (x[j - 1] + tj + 1) ** 2 / 2
            J[k - 1, k - 1] += 1

    x = np.arange(1, n + 1) * h
    return "discrete_integral_equation", f, g, x * (x - 1)


def trigonometric(n):
    def f(fvec, x):
        for j in range(n):
            fvec[j] = math.cos(x[j])
        sum1 = np.sum(fvec)
        for k in range(1, n + 1):
            fvec[k - 1] = n + k - math.sin(x[k - 1]) - sum1 - k * fvec[k - 1]

    def g(J, x):
        for j in range(1, n + 1):
            temp = math.sin(x[j - 1])
            J[:, j - 1] = temp
            J[j - 1, j - 1] = (j + 1) * temp - math.cos(x[j - 1])

    return "trigonometric", f, g, np.ones(n) / n


def variably_dimensioned(n):
    def f(fvec, x):
        sum1 = 0.0
        for j in range(1, n + 1):
            sum1 += j * (x[j - 1] - 1)
        temp = sum1 * (1 + 2 * sum1 ** 2)
        for k in range(1, n + 1):
            fvec[k - 1] = x[k - 1] - 1 + k * temp

    def g(J, x):
        sum1 = 0.0
        for j in range(1, n + 1):
            sum1 += j * (x[j - 1] - 1)
        temp = 1 + 6 * sum1 ** 2
        for k in range(1, n + 1):
            for j in range(k, n + 1):
                J[k - 1, j - 1] = k * j * temp
                J[j - 1, k - 1] = J[k - 1, j - 1]
            J[k - 1, k - 1] += 1

    return "variably_dimensioned", f, g, np.arange(1, n + 1) / n


def broyden_tridiagonal(n):
    def f(fvec, x):
        for k in range(n):
            temp = (3 - 2 * x[k]) * x[k]
            temp1 = x[k - 1] if k != 0 else 0.0
            temp2 = x[k + 1] if k != n - 1 else 0.0
            fvec[k] = temp - temp1 - 2 * temp2 + 1

    def g(J, x):
        J[:] = 0
        for k in range(n):
            J[k, k] = 3 - 4 * x[k]
            if k != 0:
                J[k, k - 1] = -1
            if k != n - 1:
                J[k, k + 1] = -2

    return "broyden_tridiagonal", f, g, -np.ones(n)


def broyden_banded(n):
    ml, mu = 5, 1

    def f(fvec, x):
        for k in range(n):
            k1 = max(0, k - ml)
            k2 = min(k + mu, n - 1)
            temp = 0.0
            for j in range(k1, k2 + 1):
                if j != k:
                    temp += x[j] * (1 + x[j])
            fvec[k] = x[k] * (2 + 5 * x[k] ** 2) + 1 - temp

    def g(J, x):
        J[:] = 0
        for k in range(n):
            k1 = max(0, k - ml)
            k2 = min(k + mu, n - 1)
            for j in range(k1, k2 + 1):
                if j != k:
                    J[k, j] = -(1 + 2 * x[j])
            J[k, k] = 2 + 15 * x[k] ** 2

    return "broyden_banded", f, g, -np.ones(n)


def minpack_all():
    """test/nonlinearsolvers.jl:512-522"""
    return [rosenbrock(), powell_singular(), powell_badly_scaled(), wood(), helical_valley(), watson(6), watson(9),
            chebyquad(5), chebyquad(6), chebyquad(7), chebyquad(9), brown_almost_linear(10), brown_almost_linear(30),
            brown_almost_linear(40), discrete_boundary_value(10), discrete_integral_equation(1),
            discrete_integral_equation(10), trigonometric(10), variably_dimensioned(10), broyden_tridiagonal(10),
            broyden_banded(10)]


def minpack_cholesky():
    """test/nonlinearsolvers.jl:573-583"""
    return [rosenbrock(), powell_singular(), powell_badly_scaled(), wood(), helical_valley(), watson(6),
            chebyquad(5), chebyquad(6), chebyquad(7), chebyquad(9), brown_almost_linear(10),
            discrete_boundary_value(10), discrete_integral_equation(1), discrete_integral_equation(10),
            trigonometric(10), variably_dimensioned(10), broyden_tridiagonal(10), broyden_banded(10)]


def factor():
    """test/nonlinearleastsquares.jl:7-89 — 9 residuals, 6 parameters, J'J singular."""
    data = [3.0, 2.0, 5.0, 4.5, 3.2, 2.0, 5.0, 1.3, 1.5]

    def f(fvec, x):
        for a in range(3):
            for b in range(3):
                fvec[3 * a + b] = data[3 * a + b] - x[a] * x[3 + b]

    def g(J, x):
        J[:] = 0
        for a in range(3):
            for b in range(3):
                J[3 * a + b, a] = -x[3 + b]
                J[3 * a + b, 3 + b] = -x[a]

    return "factor", f, g, np.ones(6)


def factor_sparse_pattern():
    """Pattern of sparse(J) for the factor model and a g! that writes nonzeros(J) in CSC order
    (test/nonlinearleastsquares.jl:47-86 writes them in row order of a CSR-like walk; here the values are
    produced by evaluating the dense g! and reading them at the pattern, which is the same matrix)."""
    import scipy.sparse as sp
    name, f, g, x0 = factor()
    Jd = np.ones((9, 6))
    g(Jd, x0)
    pattern = sp.csc_matrix(Jd != 0, dtype=np.float64)
    pattern.sort_indices()
    rows, cols = pattern.nonzero()

    def g_sparse(J, x):
        Jd = np.zeros((9, 6))
        g(Jd, x)
        Jc = J.tocoo() if not hasattr(J, "indptr") else J
        for j in range(6):
            for k in range(J.indptr[j], J.indptr[j + 1]):
                J.data[k] = Jd[J.indices[k], j]

    return name, f, g_sparse, x0, pattern


def dense_pattern_csc(n):
    """sparse(Array(undef, n, n)) of test/nonlinearsolvers.jl:526-530: a CSC matrix storing every entry."""
    import scipy.sparse as sp
    A = sp.csc_matrix(np.ones((n, n)))
    A.sort_indices()
    return A


def sparse_adapter(g_dense, n):
    """g!(J::SparseMatrixCSC, x) for a fully-stored pattern: evaluate the dense g! and copy column-major."""
    scratch = np.zeros((n, n), order="F")

    def g_sparse(J, x):
        g_dense(scratch, x)
        J.data[:] = scratch.ravel(order="F")

    return g_sparse


# README / bounds ------------------------------------------------------------------------------------
def readme_rosenbrock():
    """README.md:13-18, test/runtests.jl:20-41 — x0 = zeros(2), analytic g!."""
    def f(out, x):
        out[0] = 1 - x[0]
        out[1] = 100 * (x[1] - x[0] ** 2)

    def g(J, x):
        J[0, 0] = -1
        J[0, 1] = 0
        J[1, 0] = -200 * x[0]
        J[1, 1] = 100

    return "readme_rosenbrock", f, g, np.zeros(2)


def bounds_cases():
    """test/bounds.jl:11-36 with analytic Jacobians. Each: (name, f, g, x0, kwargs, expected minimizer)."""
    _, f_r, g_r, _ = readme_rosenbrock()

    def flo(out, x):
        out[0] = x[0] - 0.5
        out[1] = x[1] ** 2 - 9

    def glo(J, x):
        J[0, 0] = 1; J[0, 1] = 0; J[1, 0] = 0; J[1, 1] = 2 * x[1]

    def fhi(out, x):
        out[0] = x[0] - 5
        out[1] = x[1] ** 2 - 4

    def ghi(J, x):
        J[0, 0] = 1; J[0, 1] = 0; J[1, 0] = 0; J[1, 1] = 2 * x[1]

    return [
        ("rosenbrock_lower_inactive", f_r, g_r, np.zeros(2), dict(lower=[0.0, 0.0]), np.array([1.0, 1.0])),
        ("lower_active", flo, glo, np.array([2.0, 1.0]), dict(lower=[1.0, -100.0], x_tol=1e-50, f_tol=1e-50),
         np.array([1.0, 3.0])),
        ("upper_active", fhi, ghi, np.array([0.0, 1.0]), dict(upper=[2.0, 100.0], x_tol=1e-50, f_tol=1e-50),
         np.array([2.0, 2.0])),
    ]


# NIST StRD nonlinear regression (test/nonlinearfitting.jl:6-1472) --------------------------------------
# Numbers (observations, NIST start columns, NIST certified values) come from tests/golden/nist_strd.json, frozen
# from the reference's test file by oracle/make_golden_nist.py.  The models below are the `f(x, beta)` lines of
# that file (cited per entry); the residual is ff!(fcur, x, f, data): fcur[i] = y[i] - f(x[i], beta)
# (test/nonlinearfitting.jl:1448-1452).  The reference differentiates f with ForwardDiff (exact derivatives up to
# rounding, types.jl:54-66); here the Jacobian d fcur / d beta = -d f / d beta is written out analytically.
def _nist_models():
    e = np.exp

    def misra1a(x, b):          # :28   b1*(1-exp(-b2*x))
        E = e(-b[1] * x)
        return b[0] * (1 - E), [1 - E, b[0] * x * E]

    def chwirut(x, b):          # :98, :327   exp(-b1*x)/(b2+b3*x)
        E, D = e(-b[0] * x), b[1] + b[2] * x
        return E / D, [-x * E / D, -E / D ** 2, -x * E / D ** 2]

    def lanczos3(x, b):         # :369   b1*exp(-b2*x) + b3*exp(-b4*x) + b5*exp(-b6*x)
        E1, E2, E3 = e(-b[1] * x), e(-b[3] * x), e(-b[5] * x)
        return b[0] * E1 + b[2] * E2 + b[4] * E3, [E1, -b[0] * x * E1, E2, -b[2] * x * E2, E3, -b[4] * x * E3]

    def gauss(x, b):            # :658, :928   b1*exp(-b2*x) + b3*exp(-(x-b4)^2/b5^2) + b6*exp(-(x-b7)^2/b8^2)
        E1 = e(-b[1] * x)
        d2, d3 = x - b[3], x - b[6]
        E2, E3 = e(-d2 ** 2 / b[4] ** 2), e(-d3 ** 2 / b[7] ** 2)
        return (b[0] * E1 + b[2] * E2 + b[5] * E3,
                [E1, -b[0] * x * E1, E2, b[2] * E2 * 2 * d2 / b[4] ** 2, b[2] * E2 * 2 * d2 ** 2 / b[4] ** 3,
                 E3, b[5] * E3 * 2 * d3 / b[7] ** 2, b[5] * E3 * 2 * d3 ** 2 / b[7] ** 3])

    def danwood(x, b):          # :958   b1*x^b2
        p = x ** b[1]
        return b[0] * p, [p, b[0] * p * np.log(x)]

    def misra1b(x, b):          # :989   b1*(1 - 1/(1+b2*x/2)^2)
        q = 1 + b[1] * x / 2
        return b[0] * (1 - 1 / q ** 2), [1 - 1 / q ** 2, b[0] * x / q ** 3]

    def mgh09(x, b):            # :1019  b1*(x^2+x*b2)/(x^2+x*b3+b4)
        N, D = x ** 2 + x * b[1], x ** 2 + x * b[2] + b[3]
        return b[0] * N / D, [N / D, b[0] * x / D, -b[0] * N * x / D ** 2, -b[0] * N / D ** 2]

    def thurber(x, b):          # :1081  (b1+b2x+b3x^2+b4x^3)/(1+b5x+b6x^2+b7x^3)
        N = b[0] + b[1] * x + b[2] * x ** 2 + b[3] * x ** 3
        D = 1 + b[4] * x + b[5] * x ** 2 + b[6] * x ** 3
        return N / D, [1 / D, x / D, x ** 2 / D, x ** 3 / D, -N * x / D ** 2, -N * x ** 2 / D ** 2, -N * x ** 3 / D ** 2]

    def rat42(x, b):            # :1140  b1/(1+exp(b2-b3*x))
        E = e(b[1] - b[2] * x)
        return b[0] / (1 + E), [1 / (1 + E), -b[0] * E / (1 + E) ** 2, b[0] * x * E / (1 + E) ** 2]

    def mgh10(x, b):            # :1173  b1*exp(b2/(x+b3))
        E = e(b[1] / (x + b[2]))
        return b[0] * E, [E, b[0] * E / (x + b[2]), -b[0] * E * b[1] / (x + b[2]) ** 2]

    def eckerle4(x, b):         # :1227  b1/b2*exp(-(x-b3)^2/(2*b2^2))
        d = x - b[2]
        E = e(-d ** 2 / (2 * b[1] ** 2))
        return b[0] / b[1] * E, [E / b[1], b[0] * E * (d ** 2 / b[1] ** 4 - 1 / b[1] ** 2), b[0] / b[1] * E * d / b[1] ** 2]

    def rat43(x, b):            # :1263  b1/(1+exp(b2-b3*x))^(1/b4)
        E = e(b[1] - b[2] * x)
        q = 1 + E
        P = q ** (-1 / b[3])
        return b[0] * P, [P, -b[0] * P * E / (b[3] * q), b[0] * P * x * E / (b[3] * q), b[0] * P * np.log(q) / b[3] ** 2]

    def bennett5(x, b):         # :1442  b1*(b2+x)^(-1/b3)
        q = b[1] + x
        P = q ** (-1 / b[2])
        return b[0] * P, [P, -b[0] * P / (b[2] * q), b[0] * P * np.log(q) / b[2] ** 2]

    return {"misra1a": misra1a, "Chwirut2": chwirut, "Chwirut1": chwirut, "Lanczos3": lanczos3, "Gauss1": gauss,
            "Gauss2": gauss, "DanWood": danwood, "Misra1b": misra1b, "MGH09": mgh09, "Thurber": thurber,
            "BoxBOD": misra1a, "Rat42": rat42, "MGH10": mgh10, "Eckerle4": eckerle4, "Rat43": rat43,
            "Bennet5": bennett5}        # BoxBOD (:1112) has misra1a's model; "Bennet5" is the file's own spelling (:1278)


def nist_strd():
    """[(name, f_, g_, [start columns], certified, m)] for the 16 datasets the reference's loop runs (:1455)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nist_strd.json")) as fh:
        blob = json.load(fh)
    models = _nist_models()
    out = []
    for p in blob["problems"]:
        model = models[p["name"]]
        xs, ys = np.array(p["x"]), np.array(p["y"])

        def f(fcur, beta, model=model, xs=xs, ys=ys):
            with np.errstate(all="ignore"):
                fcur[:] = ys - model(xs, beta)[0]

        def g(J, beta, model=model, xs=xs):
            with np.errstate(all="ignore"):
                cols = model(xs, beta)[1]
            for k, c in enumerate(cols):
                J[:, k] = -c

        out.append((p["name"], f, g, [np.array(s) for s in p["starts"]], np.array(p["certified"]), len(ys)))
    return out
