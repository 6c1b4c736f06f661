"""Shared builders for the tests: molecules, wrappers and seeded density matrices."""
import numpy as np
import torch
from dqc_b200.api.loadbasis import loadbasis
from dqc_b200.utils.datastruct import AtomCGTOBasis, CGTOBasis
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper

dtype = torch.float64

H2O = ([8, 1, 1], [[0.0, 0.0, 0.2217], [0.0, 1.4309, -0.8867], [0.0, -1.4309, -0.8867]])
CH4ISH = ([6, 1, 1, 7], [[0.1, -0.2, 0.05], [1.9, 0.3, -0.4], [-0.7, 1.6, 0.9], [-1.1, -1.5, -1.7]])


def make_wrapper(atomzs, pos, basis):
    pos = torch.tensor(pos, dtype=dtype)
    ab = [AtomCGTOBasis(atomz=z, bases=loadbasis("%d:%s" % (z, basis)), pos=p) for z, p in zip(atomzs, pos)]
    return LibcintWrapper(ab), pos


def highl_wrapper():
    """Two off-axis centres carrying s..g shells with 1-3 primitives (exercises every l)."""
    rng = np.random.RandomState(7)
    pos = torch.tensor([[0.3, -0.2, 0.1], [-0.9, 0.8, 1.1]], dtype=dtype)
    abs_ = []
    for ia in range(2):
        shells = []
        for l in range(5):
            npr = 1 + (l + ia) % 3
            al = torch.tensor(np.sort(rng.uniform(0.3, 2.5, npr))[::-1].copy(), dtype=dtype)
            co = torch.tensor(rng.uniform(0.3, 1.0, npr), dtype=dtype)
            shells.append(CGTOBasis(angmom=l, alphas=al, coeffs=co))
        abs_.append(AtomCGTOBasis(atomz=1 + ia, bases=shells, pos=pos[ia]))
    return LibcintWrapper(abs_), pos


def random_points(n, seed=0, span=3.0):
    rng = np.random.RandomState(seed)
    return rng.uniform(-span, span, (n, 3))


def seeded_dm(nao, nocc, seed=0):
    """D = 2 C C^T, C = first nocc columns of qr(randn), symmetric PSD."""
    g = torch.Generator().manual_seed(seed)
    q, _ = torch.linalg.qr(torch.randn(nao, nao, dtype=dtype, generator=g))
    c = q[:, :nocc]
    return 2 * c @ c.T
