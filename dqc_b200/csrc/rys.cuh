// Rys quadrature nodes/weights from the interpolation table (tools/make_rys_table.py).
//   int_0^1 exp(-x t^2) f(t^2) dt = sum_r w_r f(u_r)
// x < xmax: piecewise Chebyshev (Clenshaw); x >= xmax: scaled half-range Hermite rule.
#pragma once
#include "common.cuh"

// value of function `f` (0..n-1 nodes, n..2n-1 weights) of the n-point rule at x
__device__ __forceinline__ double rys_eval(int n, int f, double x) {
    if (x >= c_rys.xmax) {
        const double s = 1.0 / x;
        return f < n ? c_rys.herm[n - 1][0][f] * s : c_rys.herm[n - 1][1][f - n] * sqrt(s);
    }
    int it = (int)(x / c_rys.h);
    if (it > c_rys.nint - 1) it = c_rys.nint - 1;
    const double t = 2.0 * (x - it * c_rys.h) / c_rys.h - 1.0;
    const int nc = c_rys.deg + 1;
    const double *c = c_rys.coef[n - 1] + ((size_t)it * 2 * n + f) * nc;
    double b1 = 0.0, b2 = 0.0;
    for (int k = nc - 1; k >= 1; k--) {
        const double b0 = c[k] + 2.0 * t * b1 - b2;
        b2 = b1;
        b1 = b0;
    }
    return c[0] + t * b1 - b2;
}
