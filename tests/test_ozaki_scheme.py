"""CPU emulation of the arithmetic of the tcgen05 int8 kernels (csrc/vxc_i8.cuh, rho_i8.cuh, gemm_i8.cuh): operands cut
into S slices of 7 bits per row with a power-of-two scale, exact integer slice products with s + t < S, one
accumulator per anti-diagonal, exact int64 merge, one conversion.  Checks (i) the error bound the design relies on,
(ii) that the accumulators stay inside int32 / the merge inside int64 for the K limits the C side enforces
(QC_REQUIRE in b200qc_gemm_i8 / b200qc_vxc_sb_i8 / b200qc_rho_sb_i8) -- with the worst-case operands."""
import numpy as np
import pytest


def slice_rows(x, S):
    """x (R, K) fp64 -> (list of S int64 arrays q_s, scale (R,)) exactly as i8_quantise16 / the slicer kernels do."""
    m = np.abs(x).max(1)
    e = np.where(m > 0, np.frexp(np.where(m > 0, m, 1.0))[1], 0)
    scale = np.ldexp(1.0, e)
    y = x * np.ldexp(64.0, -e)[:, None]
    qs = []
    for _ in range(S):
        q = np.rint(y)
        qs.append(q.astype(np.int64))
        y = (y - q) * 128.0
    return qs, scale


def gemm_emulated(a, b, S):
    """a (M, K), b (N, K) -> (a b^T as the kernels compute it, max |accumulator|, max |merged int64|)."""
    qa, sa = slice_rows(a, S)
    qb, sb = slice_rows(b, S)
    for q in qa + qb:
        assert np.abs(q).max() <= 64                     # fits int8 with room to spare
    acc = [sum(qa[s] @ qb[d - s].T for s in range(d + 1)) for d in range(S)]
    t = acc[0].copy()
    for d in range(1, S):
        t = t * 128 + acc[d]                             # the int64 merge of i8_recombine16
    out = t.astype(np.float64) * np.ldexp(1.0, -12 - 7 * (S - 1)) * sa[:, None] * sb[None, :]
    return out, max(int(np.abs(x).max()) for x in acc), int(np.abs(t).max())


@pytest.mark.parametrize("S,tol", [(5, 6e-10), (6, 5e-12)])
def test_error_bound_of_the_sliced_product(S, tol):
    rng = np.random.default_rng(0)
    M, N, K = 96, 80, 512
    a = rng.standard_normal((M, K)) * np.exp(3 * rng.standard_normal((M, 1)))
    b = rng.standard_normal((N, K)) * np.exp(3 * rng.standard_normal((N, 1)))
    got, _, _ = gemm_emulated(a, b, S)
    want = a @ b.T
    # worst case 4 (S + 1) 2^(-7 S) K rowmax rowmax; typical errors are far smaller
    bound = 4 * (S + 1) * 2.0 ** (-7 * S) * K * np.abs(a).max(1)[:, None] * np.abs(b).max(1)[None, :]
    assert (np.abs(got - want) <= bound).all()
    assert (np.abs(got - want) / (np.abs(a) @ np.abs(b).T)).max() < tol


@pytest.mark.parametrize("S,K", [(5, 65536), (6, 32768), (5, 512), (6, 512)])
def test_integer_accumulators_cannot_overflow_at_the_enforced_limits(S, K):
    # worst case: every slice of every element at +-64 (all elements equal to the row maximum 2^e (1 - eps))
    q = np.full((1, K), 64, dtype=np.int64)
    acc = [(d + 1) * int((q @ q.T)[0, 0]) for d in range(S)]
    assert max(acc) < 2 ** 31                              # int32 TMEM accumulators
    t = acc[0]
    for d in range(1, S):
        t = t * 128 + acc[d]
    assert t < 2 ** 63                                     # exact int64 merge


def test_zero_rows_and_exact_powers_of_two():
    a = np.zeros((4, 64))
    a[1, 3] = 1.0
    a[2, :] = 0.5
    a[3, 7] = -2.0 ** -30
    b = np.eye(64)[:8] * 3.0
    got, _, _ = gemm_emulated(a, b, 6)
    assert np.array_equal(got, a @ b.T)                    # representable inputs come out exactly
