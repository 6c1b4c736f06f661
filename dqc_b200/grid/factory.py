"""Predefined molecular grids.  Restates dqc/grid/factory.py:132-321: "sg2" = 75 x 302 and
"sg3" = 99 x 590 (uniform integrator + DE2 map with per-Z alpha, <r> radii, Becke partition,
Dasgupta pruning); integer levels 0-9 = PySCF's (nr, nang) per period with Chebyshev-2 + Treutler
M4 (xi table), Bragg radii, Treutler adjustment and NWChem pruning.  One atomic grid per distinct Z
is built and shared (:206-223).  Atomic grids are built on the host; the Becke weights on the GPU."""
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Union
import torch
from dqc_b200.grid.base_grid import BaseGrid
from dqc_b200.grid.radial_grid import RadialGrid, LogM3Transformation, TreutlerM4Transformation, \
    DE2Transformation
from dqc_b200.grid.lebedev_grid import LebedevGrid, TruncatedLebedevGrid
from dqc_b200.grid.multiatoms_grid import BeckeGrid
from dqc_b200.grid.truncation_rules import DasguptaTrunc, NWChemTrunc, NoTrunc
from dqc_b200.utils.periodictable import atom_bragg_radii, atom_expected_radii, get_period

__all__ = ["get_grid", "get_predefined_grid"]

# DE2 alpha per Z (Dasgupta & Herbert 2017, tables for SG-2 / SG-3); default 1.0
_SG2_ALPHA = defaultdict(lambda: 1.0, {1: 2.6, 3: 3.2, 4: 2.4, 5: 2.4, 6: 2.2, 7: 2.2, 8: 2.2, 9: 2.2,
                                       11: 3.2, 12: 2.4, 13: 2.5, 14: 2.3, 15: 2.5, 16: 2.5, 17: 2.5})
_SG3_ALPHA = defaultdict(lambda: 1.0, {1: 2.7, 3: 3.0, 4: 2.4, 5: 2.4, 6: 2.4, 7: 2.4, 8: 2.6, 9: 2.1,
                                       11: 3.2, 12: 2.6, 13: 2.6, 14: 2.8, 15: 2.4, 16: 2.4, 17: 2.6})
# Treutler & Ahlrichs, JCP 102, 346 (1995), Table I
_TREUTLER_XI = defaultdict(lambda: 1.0, dict(zip(range(1, 37), [
    0.8, 0.9, 1.8, 1.4, 1.3, 1.1, 0.9, 0.9, 0.9, 0.9, 1.4, 1.3, 1.3, 1.2, 1.1, 1.0, 1.0, 1.0,
    1.5, 1.4, 1.3, 1.2, 1.2, 1.2, 1.2, 1.2, 1.2, 1.1, 1.1, 1.1, 1.1, 1.0, 0.9, 0.9, 0.9, 0.9])))
# Lebedev: number of points -> algebraic order
_NANG2PREC = dict(zip(
    [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194, 230, 266, 302, 350, 434, 590, 770, 974, 1202, 1454,
     1730, 2030, 2354, 2702, 3074, 3470, 3890, 4334, 4802, 5294, 5810],
    [3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31, 35, 41, 47, 53, 59, 65, 71, 77, 83, 89, 95,
     101, 107, 113, 119, 125, 131]))

_NR_LEVELS = ((10, 15, 20, 30, 35, 40, 50), (30, 40, 50, 60, 65, 70, 75), (40, 60, 65, 75, 80, 85, 90),
              (50, 75, 80, 90, 95, 100, 105), (60, 90, 95, 105, 110, 115, 120),
              (70, 105, 110, 120, 125, 130, 135), (80, 120, 125, 135, 140, 145, 150),
              (90, 135, 140, 150, 155, 160, 165), (100, 150, 155, 165, 170, 175, 180),
              (200, 200, 200, 200, 200, 200, 200))
_NANG_LEVELS = ((50, 86, 110, 110, 110, 110, 110), (110, 194, 194, 194, 194, 194, 194),
                (194, 302, 302, 302, 302, 302, 302), (302, 302, 434, 434, 434, 434, 434),
                (434, 590, 590, 590, 590, 590, 590), (590, 770, 770, 770, 770, 770, 770),
                (770, 974, 974, 974, 974, 974, 974), (974, 1202, 1202, 1202, 1202, 1202, 1202),
                (1202, 1202, 1202, 1202, 1202, 1202, 1202), (1454, 1454, 1454, 1454, 1454, 1454, 1454))


def _opt(name, key, table):
    if key not in table:
        raise ValueError("Unknown %s: %s. The available options are: %s" % (name, key, list(table.keys())))
    return table[key]


def _val(v, atz):
    return v if isinstance(v, int) else v(atz)


def get_grid(atomzs, atompos: torch.Tensor, *, lattice=None,
             nr: Union[int, Callable[[int], int]] = 99, nang: Union[int, Callable[[int], int]] = 590,
             radgrid_generator: str = "uniform", radgrid_transform: str = "sg2-dasgupta",
             atom_radii: str = "expected", multiatoms_scheme: str = "becke",
             truncate: Optional[str] = "dasgupta", dtype: torch.dtype = torch.double,
             device: torch.device = torch.device("cpu")) -> BaseGrid:
    if lattice is not None:
        raise NotImplementedError("periodic grids are outside the B200 Fock-build path")
    assert atompos.ndim == 2 and atompos.shape[-2] == len(atomzs)
    zs = [int(a) for a in atomzs]
    radii = _opt("atom radii", atom_radii, {"expected": atom_expected_radii, "bragg": atom_bragg_radii})
    atomradii = torch.tensor([radii[z] for z in zs], dtype=dtype, device=device)
    tf_of_z = _opt("radial grid transformation", radgrid_transform, {
        "sg2-dasgupta": lambda z: DE2Transformation(alpha=_SG2_ALPHA[z], rmin=1e-7, rmax=15 * radii[z]),
        "sg3-dasgupta": lambda z: DE2Transformation(alpha=_SG3_ALPHA[z], rmin=1e-7, rmax=15 * radii[z]),
        "logm3": lambda z: LogM3Transformation(ra=radii[z]),
        "treutlerm4": lambda z: TreutlerM4Transformation(xi=_TREUTLER_XI[z], alpha=0.6),
    })
    if isinstance(nang, int):
        prec = _opt("number of angular points", nang, _NANG2PREC)
    else:
        prec = lambda z: _opt("number of angular points", nang(z), _NANG2PREC)
    trunc = _opt("truncation rule", truncate if truncate is not None else "no", {
        "dasgupta": lambda: DasguptaTrunc(nr),
        "nwchem": lambda: NWChemTrunc(radii, prec, list(_NANG2PREC.values()), dtype=dtype,
                                      device=torch.device("cpu")),
        "no": lambda: NoTrunc(),
    })()

    host = torch.device("cpu")
    if torch.device(device).type == "cuda":
        # On a GPU only the 1-D radial rules (tens of numbers per element) are formed on the host; the radial x Lebedev
        # products with their pruning, the translation to the nuclei and the Becke weights are CUDA kernels
        # (b200qc_grid_assemble, b200qc_becke_weights) -- bit-identical points to the host construction below.
        return _device_grid(zs, atompos.to(device), nr, radgrid_generator, tf_of_z, prec, trunc, atomradii, multiatoms_scheme,
                            dtype)
    # host path (CPU tensors: the oracle's grids in the tests): atomic grids built once per distinct Z
    per_z: Dict[int, BaseGrid] = {}
    sph: List[BaseGrid] = []
    for z in zs:
        if z not in per_z:
            rad = RadialGrid(_val(nr, z), grid_integrator=radgrid_generator, grid_transform=tf_of_z(z),
                             dtype=dtype, device=host)
            if trunc.to_truncate(z):
                per_z[z] = TruncatedLebedevGrid([rad[sl] for sl in trunc.rad_slices(z, rad)],
                                                trunc.precs(z, rad))
            else:
                per_z[z] = LebedevGrid(rad, prec=_val(prec, z))
        sph.append(per_z[z])

    atompos = atompos.to(device)
    if multiatoms_scheme == "becke":
        return BeckeGrid(sph, atompos, atomradii=atomradii)
    if multiatoms_scheme == "treutler":
        return BeckeGrid(sph, atompos, atomradii=atomradii, ratom_adjust="treutler")
    raise ValueError("Unknown multiatoms scheme: %s" % multiatoms_scheme)


def get_predefined_grid(grid_inp: Union[int, str], atomzs, atompos: torch.Tensor, *, lattice=None,
                        dtype: torch.dtype = torch.double,
                        device: torch.device = torch.device("cpu")) -> BaseGrid:
    if isinstance(grid_inp, str):
        if grid_inp not in ("sg2", "sg3"):
            raise ValueError(f"Unknown grid name: {grid_inp}")
        nr, nang = (75, 302) if grid_inp == "sg2" else (99, 590)
        return get_grid(atomzs, atompos, lattice=lattice, nr=nr, nang=nang, radgrid_generator="uniform",
                        radgrid_transform=grid_inp + "-dasgupta", atom_radii="expected",
                        multiatoms_scheme="becke", truncate="dasgupta", dtype=dtype, device=device)
    if isinstance(grid_inp, int):
        nrs, nangs = _NR_LEVELS[grid_inp], _NANG_LEVELS[grid_inp]
        return get_grid(atomzs, atompos, lattice=lattice,
                        nr=lambda z: nrs[get_period(z) - 1], nang=lambda z: nangs[get_period(z) - 1],
                        radgrid_generator="chebyshev2", radgrid_transform="treutlerm4", atom_radii="bragg",
                        multiatoms_scheme="treutler", truncate="nwchem", dtype=dtype, device=device)
    raise TypeError("Unknown type of grid_inp: %s" % type(grid_inp))


def _device_grid(zs, atompos, nr, radgrid_generator, tf_of_z, prec, trunc, atomradii, multiatoms_scheme, dtype) -> BaseGrid:
    import numpy as np
    from dqc_b200 import _lib
    from dqc_b200.grid.lebedev_grid import load_lebedev
    if multiatoms_scheme not in ("becke", "treutler"):
        raise ValueError("Unknown multiatoms scheme: %s" % multiatoms_scheme)
    host = torch.device("cpu")
    types = sorted(set(zs))
    ang_rows, ang_off = [], {}
    node_r, node_dv, node_ang, node_pt, type_node_off, type_npts = [], [], [], [], [0], {}

    def table(p: int) -> int:
        if p not in ang_off:
            t = np.asarray(load_lebedev(p), dtype=np.float64)                 # (phi, theta, w)
            phi, theta = torch.tensor(t[:, 0]), torch.tensor(t[:, 1])         # same sin / cos as LebedevGrid (torch, host)
            rows = torch.stack([torch.sin(theta), torch.cos(theta), torch.sin(phi), torch.cos(phi), torch.tensor(t[:, 2])], dim=1)
            ang_off[p] = (sum(r.shape[0] for r in ang_rows), rows.shape[0])
            ang_rows.append(rows.numpy())
        return p
    for z in types:
        rad = RadialGrid(_val(nr, z), grid_integrator=radgrid_generator, grid_transform=tf_of_z(z), dtype=dtype, device=host)
        r_all, dv_all = rad.get_rgrid().reshape(-1).numpy(), rad.get_dvolume().reshape(-1).numpy()
        if trunc.to_truncate(z):
            pieces = list(zip(trunc.rad_slices(z, rad), trunc.precs(z, rad)))
        else:
            pieces = [(slice(0, len(r_all)), _val(prec, z))]
        npt = 0
        for sl, p in pieces:
            o, n = ang_off[table(int(p))]
            for i in range(*sl.indices(len(r_all))):
                node_r.append(r_all[i]); node_dv.append(dv_all[i]); node_ang.append(o); node_pt.append(npt)
                npt += n
        type_node_off.append(len(node_r))
        type_npts[z] = npt
    tid = {z: i for i, z in enumerate(types)}
    xyz, dvol_atoms, owner = _lib.grid_assemble(atompos.to(torch.float64), [tid[z] for z in zs], [type_npts[z] for z in zs],
                                                type_node_off, node_r, node_dv, node_ang, node_pt, np.concatenate(ang_rows))
    return BeckeGrid(None, atompos, atomradii=atomradii, ratom_adjust="becke" if multiatoms_scheme == "becke" else "treutler",
                     prebuilt=(xyz, dvol_atoms, owner, [type_npts[z] for z in zs]))
