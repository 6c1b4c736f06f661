"""Element data used by the grid and the molecule parser (reference: dqc/utils/periodictable.py).
Radii are published constants: Bragg-Slater radii (JCP 41, 3199 (1964), the PySCF list, Angstrom ->
Bohr) and the <r> 'expected' radii used by the DE2 grids (DOI 10.1007/s00214-012-1169-z)."""
import torch

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge "
            "As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe").split()
periodic_table_atomz = {s: z for z, s in enumerate(_SYMBOLS) if z > 0}
_BOHR = 0.52917721092

_bragg_angstrom = [
    2.00,
    0.35, 1.40,
    1.45, 1.05, 0.85, 0.70, 0.65, 0.60, 0.50, 1.50,
    1.80, 1.50, 1.25, 1.10, 1.00, 1.00, 1.00, 1.80,
    2.20, 1.80,
    1.60, 1.40, 1.35, 1.40, 1.40, 1.40, 1.35, 1.35, 1.35, 1.35,
    1.30, 1.25, 1.15, 1.15, 1.15, 1.90,
    2.35, 2.00,
    1.80, 1.55, 1.45, 1.45, 1.35, 1.30, 1.35, 1.40, 1.60, 1.55,
    1.55, 1.45, 1.45, 1.40, 1.40, 2.10]
atom_bragg_radii = [r / _BOHR for r in _bragg_angstrom]

atom_expected_radii = [
    1.0,
    1.0, 0.927272, 3.873661, 2.849396, 2.204757, 1.714495, 1.409631, 1.232198, 1.084786, 0.965273,
    4.208762, 3.252938, 3.433889, 2.752216, 2.322712, 2.060717, 1.842024, 1.662954,
    5.243652, 4.218469, 3.959716, 3.778855, 3.626288, 3.675012, 3.381917, 3.258487, 3.153572,
    3.059109, 3.330979, 2.897648, 3.424103, 2.866859, 2.512233, 2.299617, 2.111601, 1.951590,
    5.631401, 4.632850, 4.299870, 4.091705, 3.985219, 3.841740, 3.684647, 3.735235, 3.702057,
    1.533028, 3.655961, 3.237216, 3.777242, 3.248093, 2.901067, 2.691328, 2.501704, 2.337950]


def get_atomz(elmt):
    if isinstance(elmt, str):
        return periodic_table_atomz[elmt]
    if isinstance(elmt, torch.Tensor):
        assert elmt.numel() == 1
    return elmt


def get_period(atz: int) -> int:
    for period, zmax in enumerate((2, 10, 18, 36, 54, 86, 118), start=1):
        if atz <= zmax:
            return period
    raise RuntimeError("Unimplemented atomz: %d" % atz)
