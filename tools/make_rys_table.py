"""Generate the Rys-quadrature lookup table used by the CUDA ERI kernels.

For n = 1..NMAX roots, the nodes u_r = t_r^2 and weights w_r of the n-point Gauss rule for
    int_0^1 exp(-x t^2) f(t^2) dt = sum_r w_r f(u_r)
are smooth functions of x.  On [0, XMAX) they are stored as piecewise Chebyshev expansions on
uniform intervals of width H (degree DEG); for x >= XMAX the upper limit is irrelevant and the rule
is the scaled half-range Hermite rule u_r = h_r^2 / x, w_r = W_r / sqrt(x).

Reference values are computed with mpmath at 80 digits: moments m_k = F_k(x) (Boys function) ->
Cholesky of the Hankel matrix -> three-term recurrence -> Jacobi matrix eigenproblem
(Golub-Welsch).  Output: dqc_b200/data/rys_table.npz
    coef_n  (nint, 2n, DEG+1) float64   rows 0..n-1 = roots u_r, rows n..2n-1 = weights w_r
    herm_n  (2, n)                      half-range Hermite nodes^2 and weights
    meta    [NMAX, H, DEG, XMAX]
"""
import sys
import time
import numpy as np
import mpmath as mp

mp.mp.dps = 80
NMAX = int(sys.argv[1]) if len(sys.argv) > 1 else 7
H = 1.0
DEG = 13
XMAX = 64.0  # checked below: asymptotic rule exact to < 1e-15 beyond this for all n <= NMAX


def boys_moments(x, kmax):
    x = mp.mpf(x)
    if x < mp.mpf("1e-30"):
        return [mp.mpf(1) / (2 * k + 1) for k in range(kmax + 1)]
    out = []
    for k in range(kmax + 1):
        a = mp.mpf(k) + mp.mpf(1) / 2
        out.append(mp.gammainc(a, 0, x) / (2 * x ** a))
    return out


def gauss_from_moments(m, n):
    """n-point Gauss rule of the measure with moments m[0..2n-1] (in the variable u = t^2)."""
    Hk = mp.matrix(n + 1, n + 1)
    for i in range(n + 1):
        for j in range(n + 1):
            if i + j <= 2 * n - 1:
                Hk[i, j] = m[i + j]
            else:
                Hk[i, j] = 0  # only used in row/col n, which we never read below
    # Cholesky of the leading n x n Hankel block plus one extra column: R^T R = H
    R = mp.matrix(n, n + 1)
    for i in range(n):
        for j in range(i, n + 1):
            if i == n - 1 and j == n and (i + j) > 2 * n - 1:
                continue
            s = Hk[i, j]
            for k in range(i):
                s -= R[k, i] * R[k, j]
            if j == i:
                R[i, i] = mp.sqrt(s)
            else:
                R[i, j] = s / R[i, i]
    alpha = [None] * n
    beta = [None] * n
    alpha[0] = R[0, 1] / R[0, 0]
    for j in range(1, n):
        alpha[j] = R[j, j + 1] / R[j, j] - R[j - 1, j] / R[j - 1, j - 1]
        beta[j] = R[j, j] / R[j - 1, j - 1]
    J = mp.matrix(n, n)
    for j in range(n):
        J[j, j] = alpha[j]
        if j > 0:
            J[j, j - 1] = J[j - 1, j] = beta[j]
    ev, V = mp.eigsy(J)
    idx = sorted(range(n), key=lambda i: ev[i])
    roots = [ev[i] for i in idx]
    wts = [m[0] * V[0, i] ** 2 for i in idx]
    return roots, wts


def rys_mp(n, x):
    return gauss_from_moments(boys_moments(x, 2 * n), n)


def hermite_half(n):
    """Gauss rule for int_0^inf exp(-t^2) f(t^2) dt: moments Gamma(k+1/2)/2."""
    m = [mp.gamma(mp.mpf(k) + mp.mpf(1) / 2) / 2 for k in range(2 * n + 1)]
    return gauss_from_moments(m, n)


def main():
    t0 = time.time()
    out = {}
    nint = int(round(XMAX / H))
    nodes = [mp.cos(mp.pi * (mp.mpf(j) + mp.mpf(1) / 2) / (DEG + 1)) for j in range(DEG + 1)]
    worst = 0.0
    for n in range(1, NMAX + 1):
        hr, hw = hermite_half(n)
        out["herm_%d" % n] = np.array([[float(v) for v in hr], [float(v) for v in hw]])
        coef = np.zeros((nint, 2 * n, DEG + 1))
        for it in range(nint):
            a, b = it * H, (it + 1) * H
            vals = np.zeros((DEG + 1, 2 * n))
            for j, xn in enumerate(nodes):
                x = (a + b) / 2 + (b - a) / 2 * xn
                r, w = rys_mp(n, x)
                vals[j, :n] = [float(v) for v in r]
                vals[j, n:] = [float(v) for v in w]
            # Chebyshev coefficients c_k = (2/N) sum_j f(x_j) cos(k pi (j+1/2)/N), c_0 halved
            N = DEG + 1
            jj = np.arange(N) + 0.5
            for k in range(N):
                ck = (2.0 / N) * (vals * np.cos(k * np.pi * jj / N)[:, None]).sum(axis=0)
                coef[it, :, k] = ck * (0.5 if k == 0 else 1.0)
        out["coef_%d" % n] = coef
        # verification at off-node points, fp64 Clenshaw vs mp
        rng = np.random.RandomState(n)
        for x in list(rng.uniform(0, XMAX, 40)) + [0.0, 1e-9, 0.5, XMAX - 1e-9]:
            it = min(int(x / H), nint - 1)
            tt = 2 * (x - it * H) / H - 1
            r, w = rys_mp(n, x)
            ref = np.array([float(v) for v in r] + [float(v) for v in w])
            T = np.polynomial.chebyshev.chebvander(tt, DEG)[0]
            got = coef[it] @ T
            scale = np.concatenate([np.abs(ref[:n]), np.full(n, abs(ref[n:]).max())])
            worst = max(worst, float(np.max(np.abs(got - ref) / scale)))
        # asymptotic check just beyond XMAX
        r, w = rys_mp(n, XMAX)
        ra = out["herm_%d" % n][0] / XMAX
        wa = out["herm_%d" % n][1] / np.sqrt(XMAX)
        ea = max(np.max(np.abs(ra - [float(v) for v in r]) / ra), np.max(np.abs(wa - [float(v) for v in w])) / wa.max())
        print("n=%d done %.0fs  interp err so far %.2e  asymptotic err at XMAX %.2e" % (n, time.time() - t0, worst, ea), flush=True)
    out["meta"] = np.array([NMAX, H, DEG, XMAX])
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dqc_b200", "data", "rys_table.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "max interpolation error", worst)


if __name__ == "__main__":
    main()
