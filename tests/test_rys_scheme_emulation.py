"""CPU emulation of the arithmetic of the register-resident J/K engine (dqc_b200/csrc/jk_reg.cuh) in numpy: Rys roots
from the refined Chebyshev table (b200qc_rys_refine, host code of the library), the vertical recurrence and both
horizontal transfers of the 2-D tables, the component -> table index map, the prefactor, and the cart -> real-spherical
matrices (b200qc_c2s_matrix) -- against the McMurchie-Davidson oracle's (ij|kl).  Pins the formulas the CUDA templates
unroll (c00, c01, b10, b01, b00, AB / CD transfer directions, 2 pi^2.5 / (p q sqrt(p + q))) without a GPU."""
import ctypes
import os
import numpy as np
import pytest
from numpy.polynomial import chebyshev as C
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rys():
    from dqc_b200 import _lib
    lib = _lib.load(require_cuda=False)
    out = {}
    with np.load(os.path.join(ROOT, "dqc_b200", "data", "rys_table.npz")) as z:
        nmax, h, deg, xmax = z["meta"]
        out["h"], out["xmax"] = float(h) / 2, float(xmax)
        for n in range(1, 6):
            base = np.ascontiguousarray(z["coef_%d" % n], dtype=np.float64)
            fine = np.empty((base.shape[0] * 2, 2 * n, 10), dtype=np.float64)
            assert lib.b200qc_rys_refine(base.ctypes.data_as(ctypes.c_void_p), base.shape[0], 2 * n, int(deg), 2, 10,
                                         fine.ctypes.data_as(ctypes.c_void_p)) == 0
            out[n] = (fine, np.array(z["herm_%d" % n], dtype=np.float64))
    c2s = {}
    for l in range(3):
        m = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2))
        assert lib.b200qc_c2s_matrix(l, m.ctypes.data_as(ctypes.c_void_p)) == 0
        c2s[l] = m
    out["c2s"] = c2s
    return out


def rys_roots(rys, n, x):
    """u_r = t_r^2 and w_r exactly as jk_reg.cuh::rys_roots evaluates them."""
    fine, herm = rys[n]
    if x >= rys["xmax"]:
        return herm[0] / x, herm[1] / np.sqrt(x)
    h = rys["h"]
    it = min(int(x / h), fine.shape[0] - 1)
    t = 2.0 * (x - it * h) / h - 1.0
    vals = np.array([C.chebval(t, fine[it, f]) for f in range(2 * n)])
    return vals[:n], vals[n:]


def build_table(li, lj, lk, ll, w0, c00, c01, b10, b01, b00, ab, cd):
    """T[i][j][k][l] of one (root, dimension): jk_reg.cuh::build_table."""
    nij, nkl = li + lj + 1, lk + ll + 1
    W = np.zeros((nij, nkl))
    W[0, 0] = w0
    for n in range(nij - 1):
        W[n + 1, 0] = c00 * W[n, 0] + (n * b10 * W[n - 1, 0] if n > 0 else 0.0)
    for m in range(nkl - 1):
        for n in range(nij):
            v = c01 * W[n, m]
            if m > 0:
                v += m * b01 * W[n, m - 1]
            if n > 0:
                v += n * b00 * W[n - 1, m]
            W[n, m + 1] = v
    T = np.zeros((li + 1, lj + 1, lk + 1, ll + 1))
    for j in range(lj + 1):
        if j > 0:
            for n in range(nij - j):
                W[n, :] = W[n + 1, :] + ab * W[n, :]
        for i in range(li + 1):
            line = W[i, :].copy()
            for l in range(ll + 1):
                if l > 0:
                    for m in range(nkl - l):
                        line[m] = line[m + 1] + cd * line[m]
                T[i, j, :, l] = line[:lk + 1]
    return T


def cart_pows(l):
    return [(x, y, l - x - y) for x in range(l, -1, -1) for y in range(l - x, -1, -1)]


def shell(atm, bas, env, s):
    b = bas[s]
    pc = atm[b[0]][1]
    return dict(l=int(b[1]), r=np.array(env[pc:pc + 3]), a=np.array(env[b[5]:b[5] + b[2]]), c=np.array(env[b[6]:b[6] + b[2]]))


def quartet_block(rys, si, sj, sk, sl):
    li, lj, lk, ll = si["l"], sj["l"], sk["l"], sl["l"]
    nr = (li + lj + lk + ll) // 2 + 1
    comps = [cart_pows(l) for l in (li, lj, lk, ll)]
    acc = np.zeros([len(c) for c in comps])
    A, B, Cc, D = si["r"], sj["r"], sk["r"], sl["r"]
    ab, cd = A - B, Cc - D
    for ai, ci in zip(si["a"], si["c"]):
        for aj, cj in zip(sj["a"], sj["c"]):
            p = ai + aj
            P = (ai * A + aj * B) / p
            cb = ci * cj * np.exp(-ai * aj / p * ab @ ab) / p                  # JKPrim::c (without the s / p constants)
            for ak, ck in zip(sk["a"], sk["c"]):
                for al, cl in zip(sl["a"], sl["c"]):
                    q = ak + al
                    Q = (ak * Cc + al * D) / q
                    cq = ck * cl * np.exp(-ak * al / q * cd @ cd) / q
                    pq = p + q
                    d = P - Q
                    x = p * q / pq * (d @ d)
                    pref = cb * cq * 34.98683665524972497 / np.sqrt(pq)
                    a0, a1 = q / pq, p / pq
                    u, w = rys_roots(rys, nr, x)
                    for r in range(nr):
                        a0u, a1u = a0 * u[r], a1 * u[r]
                        b10, b01, b00 = (1 - a0u) * 0.5 / p, (1 - a1u) * 0.5 / q, 0.5 * u[r] / pq
                        T = [build_table(li, lj, lk, ll, (w[r] * pref if dim == 2 else 1.0), (P - A)[dim] - a0u * d[dim],
                                         (Q - Cc)[dim] + a1u * d[dim], b10, b01, b00, ab[dim], cd[dim]) for dim in range(3)]
                        for a, pa in enumerate(comps[0]):
                            for b, pb in enumerate(comps[1]):
                                for c, pc in enumerate(comps[2]):
                                    for e, pe in enumerate(comps[3]):
                                        acc[a, b, c, e] += (T[0][pa[0], pb[0], pc[0], pe[0]] * T[1][pa[1], pb[1], pc[1], pe[1]] *
                                                            T[2][pa[2], pb[2], pc[2], pe[2]])
    m = rys["c2s"]
    return np.einsum("abce,ia,jb,kc,le->ijkl", acc, m[li], m[lj], m[lk], m[ll])


@pytest.mark.parametrize("quartet", [(0, 0, 0, 0), (3, 0, 1, 0), (3, 4, 3, 1), (5, 3, 0, 2), (5, 4, 5, 3), (5, 5, 4, 3), (5, 5, 5, 5),
                                     (3, 6, 5, 9), (9, 6, 10, 3)])
def test_register_engine_arithmetic_matches_the_oracle(rys, quartet):
    """H2O / def2-SVP shells (O: 0-2 s, 3-4 p, 5 d; H: 6-7 / 9-10 s, 8 / 11 p): one contracted quartet per class from
    (ss|ss) to (dd|dd), three centres, contracted and uncontracted shells."""
    from oracle import cint
    w, _ = util.make_wrapper(*util.H2O, "def2-svp")
    atm, bas, env = w.atm_bas_env
    i, j, k, l = quartet
    ref = cint.int2e(atm, bas, env, (i, i + 1, j, j + 1, k, k + 1, l, l + 1))
    got = quartet_block(rys, *[shell(atm, bas, env, s) for s in quartet])
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())


def test_unique_quartet_digestion_matches_dense_contraction():
    """The digestion of jk_reg.cuh / jk.cuh in numpy: unique shell quartets (pairs i >= j, ket pair <= bra pair), factors
    1/2 per coincidence, J_ij += 2f B.D_kl, J_kl += 2f B.D_ij, K_ik += f B.D_jl, K_jk += f B.D_il, K_il += f B.D_jk,
    K_jl += f B.D_ik, then acc + acc^T -- on the oracle's (ij|kl) blocks of H2O / def2-SVP against the reference's dense
    einsums (hcgto.py:209,234)."""
    import torch
    from oracle import cint
    w, _ = util.make_wrapper(*util.H2O, "def2-svp")
    atm, bas, env = w.atm_bas_env
    eri = cint.int2e(atm, bas, env)
    nao = eri.shape[0]
    loc = cint.ao_loc_sph(bas)
    dm = util.seeded_dm(nao, 5, seed=3).numpy()
    nb = len(bas)
    pairs = [(i, j) for i in range(nb) for j in range(i + 1)]
    J, K = np.zeros((nao, nao)), np.zeros((nao, nao))
    sl = lambda s: slice(int(loc[s]), int(loc[s + 1]))
    for b, (i, j) in enumerate(pairs):
        for (k, l) in pairs[:b + 1]:
            f = 1.0
            if i == j:
                f *= 0.5
            if k == l:
                f *= 0.5
            if (i, j) == (k, l):
                f *= 0.5
            B = eri[sl(i), sl(j), sl(k), sl(l)]
            J[sl(i), sl(j)] += 2 * f * np.einsum("abcd,cd->ab", B, dm[sl(k), sl(l)])
            J[sl(k), sl(l)] += 2 * f * np.einsum("abcd,ab->cd", B, dm[sl(i), sl(j)])
            K[sl(i), sl(k)] += f * np.einsum("abcd,bd->ac", B, dm[sl(j), sl(l)])
            K[sl(j), sl(k)] += f * np.einsum("abcd,ad->bc", B, dm[sl(i), sl(l)])
            K[sl(i), sl(l)] += f * np.einsum("abcd,bc->ad", B, dm[sl(j), sl(k)])
            K[sl(j), sl(l)] += f * np.einsum("abcd,ac->bd", B, dm[sl(i), sl(k)])
    J, K = J + J.T, K + K.T
    jref = np.einsum("ij,ijkl->kl", dm, eri)
    kref = np.einsum("il,ijkl->jk", dm, eri)
    assert np.abs(J - jref).max() < 1e-12 and np.abs(K - kref).max() < 1e-12


def test_k1_tile_swizzle_is_a_conflict_free_bijection():
    """ao_pm_index of csrc/ao_eval.cuh (point-major K1 tile, 64 columns per row, no padding): for every point p the map
    column -> position is a bijection of the row; the 16 lanes of a half-warp writing one column for 16 consecutive points
    hit 16 different 8-byte slots of the 128-byte bank window; lane l of the write-out reads pair l ^ (p & 7), which holds
    columns 2 l, 2 l + 1 (swapped when bit 3 of p is set)."""
    def idx(p, col):
        return p * 64 + ((((col >> 1) ^ (p & 7)) << 1) | ((col & 1) ^ ((p >> 3) & 1)))
    for p in range(32):
        assert sorted(idx(p, c) - p * 64 for c in range(64)) == list(range(64))
        for lane in range(32):
            pair = lane ^ (p & 7)
            got = [c for c in range(64) if (idx(p, c) - p * 64) >> 1 == pair]
            assert sorted(got) == [2 * lane, 2 * lane + 1]
            first = [c for c in got if (idx(p, c) & 1) == 0][0]
            assert first == (2 * lane + 1 if p & 8 else 2 * lane)
    for col in range(64):
        for p0 in (0, 16):
            slots = {idx(p, col) % 16 for p in range(p0, p0 + 16)}
            assert len(slots) == 16
