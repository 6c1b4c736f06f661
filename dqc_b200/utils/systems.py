"""Deterministic geometries of the BASELINE.json workloads (coordinates in Bohr).

None of these come from the reference (it ships no geometries beyond the diatomics in its tests);
they are generated from closed formulas so that every run and every rank builds the same molecule.
"""
import math
import numpy as np

ANG = 1.0 / 0.52917721092  # Angstrom -> Bohr


def h2o():
    """H2O of dqc/test/test_properties.py:20-27 (already in Bohr)."""
    return [8, 1, 1], np.array([[0.0, 0.0, 0.2217], [0.0, 1.4309, -0.8867], [0.0, -1.4309, -0.8867]])


def benzene():
    """D6h benzene, R_CC = 1.39 A, R_CH = 1.09 A, in the xy plane."""
    zs, pos = [], []
    for k in range(6):
        a = math.pi / 3 * k
        zs.append(6)
        pos.append([1.39 * math.cos(a), 1.39 * math.sin(a), 0.0])
    for k in range(6):
        a = math.pi / 3 * k
        zs.append(1)
        pos.append([2.48 * math.cos(a), 2.48 * math.sin(a), 0.0])
    return zs, np.array(pos) * ANG


def c60(r_single=1.45, r_double=1.40):
    """Ih C60 as a truncated icosahedron: a carbon sits on every icosahedron edge at distance
    r_single from each end vertex, so pentagon edges are r_single and the remaining (hexagon-hexagon)
    edges are r_double; the icosahedron edge is r_double + 2 r_single."""
    phi = (1 + math.sqrt(5)) / 2
    verts = []
    for s1 in (1, -1):
        for s2 in (1, -1):
            verts += [[0, s1, s2 * phi], [s1, s2 * phi, 0], [s2 * phi, 0, s1]]
    v = np.array(verts, dtype=np.float64)          # edge length 2
    a = r_double + 2 * r_single
    v *= a / 2.0
    pos = []
    for i in range(12):
        d = np.linalg.norm(v - v[i], axis=1)
        for j in np.where(np.abs(d - a) < 1e-8)[0]:
            pos.append(v[i] + (v[j] - v[i]) * (r_single / a))
    pos = np.array(pos)
    assert pos.shape == (60, 3)
    return [6] * 60, pos * ANG


def taxol_like(seed=20260117):
    """113 atoms with Taxol's composition C47 H51 N O14.

    PROVENANCE: there is no Taxol structure file in this offline image and the reference ships
    none, so this is NOT the experimental geometry: heavy atoms are placed by a seeded random walk
    on a diamond-like lattice of 1.50 A bonds (a compact organic-like skeleton, every heavy atom
    >= 1.45 A from the others), hydrogens 1.09 A from a heavy atom and >= 1.6 A from one another.
    It reproduces what the benchmark depends on -- element counts (nao = 1123 in def2-SVP, the grid
    size) and an organic-molecule spatial density -- and is labelled "taxol-like" wherever used."""
    rng = np.random.RandomState(seed)
    # diamond lattice neighbours (two sublattices)
    d = 1.50 / math.sqrt(3)
    nb = [np.array(x) * d for x in ([1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1])]
    heavy = [np.zeros(3)]
    parity = [0]
    frontier = [0]
    nheavy = 47 + 1 + 14
    while len(heavy) < nheavy:
        i = frontier[rng.randint(len(frontier))]
        k = rng.randint(4)
        p = heavy[i] + (nb[k] if parity[i] == 0 else -nb[k])
        if min(np.linalg.norm(p - q) for q in heavy) < 1.45:
            continue
        # keep it compact: reject growth far from the centroid
        if np.linalg.norm(p - np.mean(heavy, axis=0)) > 6.5:
            continue
        heavy.append(p)
        parity.append(1 - parity[i])
        frontier.append(len(heavy) - 1)
    heavy = np.array(heavy)
    order = rng.permutation(nheavy)
    zs = np.empty(nheavy, dtype=int)
    zs[order[:47]] = 6
    zs[order[47:48]] = 7
    zs[order[48:]] = 8
    hpos = []
    tries = 0
    while len(hpos) < 51:
        tries += 1
        i = rng.randint(nheavy)
        v = rng.normal(size=3)
        p = heavy[i] + 1.09 * v / np.linalg.norm(v)
        dh = np.linalg.norm(heavy - p, axis=1)
        dh[i] = 9.0
        if dh.min() < 1.55:
            continue
        if hpos and min(np.linalg.norm(p - q) for q in hpos) < 1.6:
            continue
        hpos.append(p)
        assert tries < 200000
    pos = np.concatenate([heavy, np.array(hpos)])
    return [int(z) for z in zs] + [1] * 51, pos * ANG


def carbon_cluster(natom, spacing=2.7):
    """Simple-cubic carbon cluster for the nbasis/grid sweep: the natom lattice
    sites closest to the origin, spacing in Bohr."""
    n = int(math.ceil(natom ** (1.0 / 3))) + 2
    g = np.arange(-n, n + 1)
    pts = np.array([[x, y, z] for x in g for y in g for z in g], dtype=np.float64)
    key = np.round((pts ** 2).sum(1), 9)
    idx = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0], key))[:natom]
    return [6] * natom, pts[idx] * spacing
