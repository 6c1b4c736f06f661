"""Angular pruning rules (reference: dqc/grid/truncation_rules.py:6-217).  The SG-2/SG-3 tables
are the published ones of Dasgupta & Herbert, J. Comput. Chem. 38, 869 (2017); the NWChem rule is
the one PySCF uses (radii fractions per period)."""
from typing import Callable, List, Union
import torch
from dqc_b200.grid.radial_grid import RadialGrid

# (nr) -> Z -> (radial break points, Lebedev precision per segment)
_STD = {
    75: {
        1: ([0, 35, 47, 63, 70, 75], [3, 17, 29, 15, 7]),
        3: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 15, 11]),
        4: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 15, 11]),
        5: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 19, 7]),
        6: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 19, 7]),
        7: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 15, 7]),
        8: ([0, 30, 44, 62, 70, 75], [3, 17, 29, 19, 11]),
        9: ([0, 26, 42, 61, 69, 75], [3, 17, 29, 17, 11]),
        11: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 15, 11]),
        12: ([0, 35, 47, 64, 71, 75], [3, 17, 29, 15, 11]),
        13: ([0, 32, 47, 64, 71, 75], [3, 17, 29, 19, 11]),
        14: ([0, 32, 47, 64, 71, 75], [3, 17, 29, 19, 11]),
        15: ([0, 30, 44, 61, 68, 75], [3, 17, 29, 19, 9]),
        16: ([0, 30, 44, 61, 68, 75], [3, 17, 29, 19, 9]),
        17: ([0, 26, 42, 61, 69, 75], [3, 17, 29, 17, 11]),
    },
    99: {
        1: ([0, 45, 61, 82, 92, 99], [3, 17, 41, 23, 11]),
        3: ([0, 46, 62, 84, 93, 99], [3, 17, 41, 19, 11]),
        4: ([0, 42, 48, 62, 84, 87, 93, 99], [3, 15, 17, 41, 23, 19, 11]),
        5: ([0, 42, 48, 62, 84, 93, 99], [3, 15, 17, 41, 23, 11]),
        6: ([0, 46, 62, 84, 85, 87, 93, 99], [3, 19, 41, 29, 23, 19, 15]),
        7: ([0, 40, 58, 82, 93, 99], [3, 17, 41, 19, 11]),
        8: ([0, 40, 54, 56, 58, 82, 83, 84, 92, 99], [3, 17, 23, 29, 41, 29, 23, 19, 11]),
        9: ([0, 35, 52, 56, 81, 83, 91, 99], [3, 17, 23, 41, 23, 17, 11]),
        11: ([0, 46, 62, 84, 93, 99], [3, 17, 41, 19, 11]),
        12: ([0, 48, 63, 83, 90, 99], [3, 17, 41, 19, 11]),
        13: ([0, 42, 48, 62, 84, 87, 93, 99], [3, 15, 17, 41, 23, 19, 11]),
        14: ([0, 42, 48, 62, 84, 93, 99], [3, 15, 17, 41, 23, 11]),
        15: ([0, 35, 36, 54, 58, 83, 85, 93, 99], [3, 15, 17, 23, 41, 23, 19, 11]),
        16: ([0, 35, 36, 54, 58, 83, 85, 93, 99], [3, 15, 17, 23, 41, 23, 19, 11]),
        17: ([0, 35, 52, 56, 81, 83, 91, 99], [3, 17, 23, 41, 23, 17, 11]),
    },
}


def _val(v: Union[int, Callable[[int], int]], atz: int) -> int:
    return v if isinstance(v, int) else v(atz)


class NoTrunc(object):
    def to_truncate(self, atz: int) -> bool:
        return False


class DasguptaTrunc(object):
    def __init__(self, nr: Union[int, Callable[[int], int]]):
        self._nr = nr

    def to_truncate(self, atz: int) -> bool:
        return atz in _STD[_val(self._nr, atz)]

    def rad_slices(self, atz: int, radgrid: RadialGrid) -> List[slice]:
        idx = _STD[_val(self._nr, atz)][atz][0]
        return [slice(a, b, None) for a, b in zip(idx[:-1], idx[1:])]

    def precs(self, atz: int, radgrid: RadialGrid) -> List[int]:
        return _STD[_val(self._nr, atz)][atz][1]


class NWChemTrunc(object):
    def __init__(self, radii_list: List[float], prec, precs_list: List[int], dtype, device):
        self._radii = radii_list
        self._alphas = torch.tensor([[0.25, 0.5, 1.0, 4.5], [0.1667, 0.5, 0.9, 3.5], [0.1, 0.4, 0.8, 2.5]],
                                    dtype=dtype, device=device)
        self._prec = prec
        self._plist = precs_list

    def _get_precs(self, atz: int) -> List[int]:
        p = _val(self._prec, atz)
        if p == 13:
            ids = [5, 6, 6, 6, 5]
        elif p > 13:
            k = self._plist.index(p)
            ids = [5, 7, k - 1, k, k - 1]
        else:
            raise RuntimeError("This shouldn't be displayed. Please report to Github")
        return [self._plist[i] for i in ids]

    def to_truncate(self, atz: int) -> bool:
        return _val(self._prec, atz) >= 13

    def rad_slices(self, atz: int, radgrid: RadialGrid) -> List[slice]:
        row = 0 if atz <= 2 else (1 if atz <= 10 else 2)
        bounds = self._alphas[row] * self._radii[atz]
        place = torch.sum(radgrid.get_rgrid().reshape(-1, 1) > bounds, dim=-1)
        _, counts = torch.unique_consecutive(place, return_counts=True)
        out, start = [], 0
        for i in range(len(self._get_precs(atz))):
            c = int(counts[i])
            out.append(slice(start, start + c, None))
            start += c
        return out

    def precs(self, atz: int, radgrid: RadialGrid) -> List[int]:
        return self._get_precs(atz)
