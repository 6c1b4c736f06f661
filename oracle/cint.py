"""ctypes front-end of oracle/cint_oracle.c (McMurchie-Davidson CPU restatement of the libcint /
libcgto entry points the reference calls, see that file's header).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.realpath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"])
        _LIB = ctypes.CDLL(so)
        _LIB.orc_init()
        _LIB.orc_c2s.restype = ctypes.POINTER(ctypes.c_double)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(atm, bas, env):
    return (np.ascontiguousarray(atm, dtype=np.int32), np.ascontiguousarray(bas, dtype=np.int32),
            np.ascontiguousarray(env, dtype=np.float64))


def ao_loc_sph(bas):
    return np.concatenate([[0], np.cumsum(2 * np.asarray(bas)[:, 1] + 1)]).astype(np.int32)


def c2s(l):
    nc = (l + 1) * (l + 2) // 2
    return np.ctypeslib.as_array(lib().orc_c2s(l), shape=(2 * l + 1, nc)).copy()


def _call(fn, out, slices, atm, bas, env, *lead):
    atm, bas, env = _prep(atm, bas, env)
    ao_loc = ao_loc_sph(bas)
    sl = (ctypes.c_int * len(slices))(*slices)
    fn(*lead, _p(out), sl, _p(ao_loc), _p(atm), ctypes.c_int(len(atm)), _p(bas), ctypes.c_int(len(bas)), _p(env))
    return out


def _n(bas, s0, s1):
    loc = ao_loc_sph(bas)
    return int(loc[s1] - loc[s0])


_KINDS = {"ovlp": 0, "kin": 1, "nuc": 2, "rinv": 3}


def int1e(kind, atm, bas, env, shls=None):
    nb = len(bas)
    s = shls or (0, nb, 0, nb)
    out = np.zeros((_n(bas, s[0], s[1]), _n(bas, s[2], s[3])))
    return _call(lib().orc_int1e, out, s, atm, bas, env, ctypes.c_int(_KINDS[kind]))


def int2c2e(atm, bas, env, shls=None):
    nb = len(bas)
    s = shls or (0, nb, 0, nb)
    out = np.zeros((_n(bas, s[0], s[1]), _n(bas, s[2], s[3])))
    return _call(lib().orc_int2c2e, out, s, atm, bas, env)


def int3c2e(atm, bas, env, shls):
    out = np.zeros(tuple(_n(bas, shls[2 * q], shls[2 * q + 1]) for q in range(3)))
    return _call(lib().orc_int3c2e, out, shls, atm, bas, env)


def int2e(atm, bas, env, shls=None):
    nb = len(bas)
    s = shls or (0, nb) * 4
    out = np.zeros(tuple(_n(bas, s[2 * q], s[2 * q + 1]) for q in range(4)))
    return _call(lib().orc_int2e, out, s, atm, bas, env)


def eval_gto(atm, bas, env, coords, deriv=0, shls=None):
    """deriv=0 -> (ngrid, nao); deriv=1 -> (3, ngrid, nao); deriv=2 -> (ngrid, nao) Laplacian (the to_transpose=True
    layouts of eval_gto / eval_gradgto / eval_laplgto)."""
    atm, bas, env = _prep(atm, bas, env)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    nb = len(bas)
    s = shls or (0, nb)
    ng = coords.shape[0]
    nao = _n(bas, s[0], s[1])
    out = np.zeros((3, ng, nao) if deriv == 1 else (ng, nao))
    ao_loc = ao_loc_sph(bas)
    sl = (ctypes.c_int * 2)(*s)
    lib().orc_eval_gto(ctypes.c_int(deriv), ctypes.c_int(ng), _p(coords), _p(out), sl, _p(ao_loc), _p(atm),
                       ctypes.c_int(len(atm)), _p(bas), ctypes.c_int(nb), _p(env))
    return out
