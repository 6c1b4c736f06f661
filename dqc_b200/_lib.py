"""ctypes binding of libb200qc.so -- the only route from the PyTorch host code to the CUDA kernels.

Mirrors the reference's own idiom (ctypes CDLL handles + raw pointers, dqc/hamilton/intor/utils.py:15-29,
molintor.py:629-638), except that every pointer is a device pointer of a torch CUDA tensor and the
current torch stream is passed explicitly.  There is NO CPU fallback: if the library cannot be
loaded or no CUDA device is present, every entry point raises.
"""
import ctypes
import os
from typing import Optional
import numpy as np
import torch

_HERE = os.path.dirname(os.path.realpath(__file__))
_SO_PATH = os.path.join(_HERE, "libb200qc.so")
_lib = None
_rys_loaded = set()      # CUDA device indices that hold the Rys table (constant memory is per device)

GRID_ALIGN = 128  # leading dimension of the grid axis (K2/K4 CTA tile)
AO_ALIGN = 64     # leading dimension of the AO axis

LMAX = 4     # highest angular momentum of the kernels (B200QC_LMAX)
FUNC_IDS = {"lda_x": 1, "lda_c_pw": 2, "lda_c_pw_mod": 3, "lda_c_vwn": 4, "lda_c_vwn_rpa": 5,
            "gga_x_pbe": 101, "gga_c_pbe": 102, "gga_x_b88": 103, "gga_c_lyp": 104, "mgga_x_scan": 201}
FUNC_FAMILY = {name: (4 if fid >= 200 else 2 if fid >= 100 else 1) for name, fid in FUNC_IDS.items()}

_SIGS = {
    # name: (restype, argtypes)
    "b200qc_last_error": (ctypes.c_char_p, []),
    "b200qc_version": (ctypes.c_int, []),
    "b200qc_launch_count": (ctypes.c_int64, []),
    "b200qc_profile": (ctypes.c_int, [ctypes.c_int]),
    "b200qc_profile_nkernels": (ctypes.c_int, []),
    "b200qc_profile_name": (ctypes.c_char_p, [ctypes.c_int]),
    "b200qc_profile_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_peak_fp64_dmma": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_peak_fp64_fma": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_peak_i8_mma": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_basis_upload": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_grid_assemble": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_basis_free": (ctypes.c_int, [ctypes.c_void_p]),
    "b200qc_basis_set_cartesian": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "b200qc_c2s_matrix": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "b200qc_rys_refine": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p]),
    "b200qc_rys_upload": (ctypes.c_int, [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                         ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_eval_gto": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_void_p]),
    "b200qc_becke_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                            ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_rho": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_xc_unpol": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_xc_mgga_unpol": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_rho_sb_mgga": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p]),
    "b200qc_vxc_sb_mgga": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_xc_pol": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_vxc_worksize": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int64]),
    "b200qc_vxc_mat": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "b200qc_ao_screen": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_eval_gto_sb": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_rho_sb": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_vxc_sb": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_vxc_i8_prepare": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_vxc_sb_i8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_i8_debug_variant": (ctypes.c_int, [ctypes.c_int]),
    "b200qc_i8_mode": (ctypes.c_int, [ctypes.c_int]),
    "b200qc_rho_i8_prepare": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "b200qc_rho_sb_i8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "b200qc_int1e": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_int2c2e": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_int3c2e": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_int2e": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_int3c2e_packed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                             ctypes.c_void_p]),
    "b200qc_eri_store": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]),
    "b200qc_gemv": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_jkplan_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_jkplan_nquartets": (ctypes.c_int64, [ctypes.c_void_p]),
    "b200qc_jkplan_nquartets_reg": (ctypes.c_int64, [ctypes.c_void_p]),
    "b200qc_jkplan_flops": (ctypes.c_double, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "b200qc_jkplan_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "b200qc_jkplan_free": (ctypes.c_int, [ctypes.c_void_p]),
    "b200qc_jk_direct": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_dfj_worksize": (ctypes.c_int64, [ctypes.c_int64, ctypes.c_int64]),
    "b200qc_dfj": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_dfj_rowmask": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                          ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_dfj_pass1_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_int64, ctypes.c_void_p]),
    "b200qc_dfj_pass2_masked": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_dfj_pass1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_dfj_pass2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_pack_tril": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_i8_slice": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]),
    "b200qc_gemm_i8": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_void_p,
                                      ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int, ctypes.c_void_p]),
    "b200qc_i8_slice_dual": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}


class B200QCError(RuntimeError):
    pass


def exported_symbols():
    """Names declared in include/b200qc.h (used by the CPU test that checks the .so exports them)."""
    return sorted(_SIGS.keys())


def load(require_cuda: bool = True):
    """Load libb200qc.so (once).  Raises -- never falls back -- when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO_PATH):
            raise B200QCError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback for the Fock-build path." % _SO_PATH)
        lib = ctypes.CDLL(_SO_PATH)
        missing = [name for name in _SIGS if not hasattr(lib, name)]
        if missing:
            raise B200QCError("libb200qc.so does not export %s (header/library mismatch; rebuild)" % missing)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    if require_cuda and not torch.cuda.is_available():
        raise B200QCError("no CUDA device: the B200 Fock-build path has no CPU fallback")
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load(False).b200qc_last_error()
        raise B200QCError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "device-contiguous tensor required"
    return ctypes.c_void_p(t.data_ptr())


def _np(a: np.ndarray) -> ctypes.c_void_p:
    return a.ctypes.data_as(ctypes.c_void_p)


def launch_count() -> int:
    return int(load(False).b200qc_launch_count())


def profile_enable(on: bool = True):
    """Start (or stop) per-kernel CUDA-event timing inside the library."""
    _check(load().b200qc_profile(1 if on else 0), "profile")


def profile_read():
    """{kernel name: (launch count, total device ms)} since the last read; synchronises."""
    lib = load()
    n = lib.b200qc_profile_nkernels()
    ms = np.zeros(n, dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int64)
    _check(lib.b200qc_profile_read(_np(ms), _np(cnt)), "profile_read")
    return {lib.b200qc_profile_name(i).decode(): (int(cnt[i]), float(ms[i])) for i in range(n) if cnt[i] > 0}


def peak_fp64_dmma(iters: int = 20000) -> float:
    """Measured fp64 tensor-pipe peak (TFLOP/s) of the current device."""
    lib = load()
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    out = ctypes.c_double(0.0)
    _check(lib.b200qc_peak_fp64_dmma(iters, _ptr(scratch), ctypes.byref(out), _stream()), "peak_fp64_dmma")
    return float(out.value)


def peak_fp64_fma(iters: int = 20000) -> float:
    """Measured plain fp64 pipe (DFMA) peak (TFLOP/s) of the current device."""
    lib = load()
    scratch = torch.zeros(8, dtype=torch.float64, device="cuda")
    out = ctypes.c_double(0.0)
    _check(lib.b200qc_peak_fp64_fma(iters, _ptr(scratch), ctypes.byref(out), _stream()), "peak_fp64_fma")
    return float(out.value)


def peak_i8_mma(iters: int = 40000) -> float:
    """Measured tcgen05 int8 tensor-pipe peak (TOP/s, 2 ops per multiply-add) of the current device."""
    lib = load()
    out = ctypes.c_double(0.0)
    _check(lib.b200qc_peak_i8_mma(iters, ctypes.byref(out), _stream()), "peak_i8_mma")
    return float(out.value)


def round_up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


def ensure_rys_table():
    """Upload the Rys interpolation table (dqc_b200/data/rys_table.npz) once per process and device."""
    dev = torch.cuda.current_device()
    if dev in _rys_loaded:
        return
    lib = load()
    with np.load(os.path.join(_HERE, "data", "rys_table.npz")) as z:
        nmax, h, deg, xmax = z["meta"]
        nmax, deg = int(nmax), int(deg)
        coefs = [np.ascontiguousarray(z["coef_%d" % n], dtype=np.float64) for n in range(1, nmax + 1)]
        herms = [np.ascontiguousarray(z["herm_%d" % n], dtype=np.float64) for n in range(1, nmax + 1)]
    cptr = (ctypes.c_void_p * nmax)(*[c.ctypes.data for c in coefs])
    hptr = (ctypes.c_void_p * nmax)(*[c.ctypes.data for c in herms])
    _check(lib.b200qc_rys_upload(nmax, float(h), deg, float(xmax), cptr, hptr), "rys_upload")
    _rys_loaded.add(dev)


class DeviceBasis(object):
    """Handle of a basis uploaded to the GPU (b200qc_basis_upload)."""

    def __init__(self, atm, bas, env, ao_loc, spherical=True, device=None):
        """spherical=False: raw cartesian output x^a y^b z^c sum_p c_p exp(-a_p r^2) of the integral kernels (ao_loc
        counting (l + 1)(l + 2) / 2 per shell) -- the building block of the derivative integrals, not an AO basis."""
        lib = load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.atm = np.ascontiguousarray(atm, dtype=np.int32)
        self.bas = np.ascontiguousarray(bas, dtype=np.int32)
        self.env = np.ascontiguousarray(env, dtype=np.float64)
        self.ao_loc = np.ascontiguousarray(ao_loc, dtype=np.int32)
        h = ctypes.c_void_p(0)
        with torch.cuda.device(self.device):
            _check(lib.b200qc_basis_upload(_np(self.atm), len(self.atm), _np(self.bas), len(self.bas),
                                           _np(self.env), len(self.env), _np(self.ao_loc), ctypes.byref(h)),
                   "basis_upload")
        self.handle = h
        if not spherical:
            _check(lib.b200qc_basis_set_cartesian(self.handle, 1), "basis_set_cartesian")

    def __del__(self):
        try:
            if self.handle:
                load(False).b200qc_basis_free(self.handle)
                self.handle = None
        except Exception:
            pass

    def nao(self, sh0, sh1):
        return int(self.ao_loc[sh1] - self.ao_loc[sh0])


# ------------------------------------------------------------------------------------------
# thin tensor-level wrappers (one per C entry point)

def eval_gto(basis: DeviceBasis, sh0: int, sh1: int, coords: torch.Tensor, deriv: int,
             ngrid_ld: Optional[int] = None, ao_ld: Optional[int] = None) -> torch.Tensor:
    """Returns the padded AO buffer (ncomp, ngrid_ld, ao_ld), zero in the padding."""
    lib = load()
    ngrid = coords.shape[0]
    nao = basis.nao(sh0, sh1)
    ngrid_ld = round_up(max(ngrid, 1), GRID_ALIGN) if ngrid_ld is None else ngrid_ld
    ao_ld = round_up(nao, AO_ALIGN) if ao_ld is None else ao_ld
    ncomp = {0: 1, 1: 4, 2: 5}[int(deriv)]
    ao = torch.zeros((ncomp, ngrid_ld, ao_ld), dtype=torch.float64, device=coords.device)
    coords = coords.contiguous()
    _check(lib.b200qc_eval_gto(basis.handle, sh0, sh1, deriv, _ptr(coords), ngrid, _ptr(ao), ngrid_ld, ao_ld,
                               _stream()), "eval_gto")
    return ao


def c2s_matrix(l: int) -> np.ndarray:
    """(2l + 1, (l + 1)(l + 2) / 2) cartesian -> real-spherical matrix the kernels apply (angular normalisation included)."""
    out = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2), dtype=np.float64)
    _check(load(require_cuda=False).b200qc_c2s_matrix(int(l), _np(out)), "c2s_matrix")
    return out


def grid_assemble(atompos: torch.Tensor, atom_type, atom_npts, type_node_off, node_r, node_dv, node_ang_off, node_pt_off,
                  ang: np.ndarray):
    """Molecular grid on the device from per-type radial rules and Lebedev tables (b200qc_grid_assemble):
    returns xyz (ngrid, 3), dvol (ngrid,) (radial x angular weights), owner (ngrid,) int32."""
    dev = atompos.device
    up = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev)
    off = np.concatenate([[0], np.cumsum(np.asarray(atom_npts, dtype=np.int64))])
    ngrid = int(off[-1])
    xyz = torch.empty((ngrid, 3), dtype=torch.float64, device=dev)
    dvol = torch.empty(ngrid, dtype=torch.float64, device=dev)
    owner = torch.empty(ngrid, dtype=torch.int32, device=dev)
    bufs = [up(atom_type, torch.int32), up(off, torch.int64), up(type_node_off, torch.int32), up(node_r, torch.float64),
            up(node_dv, torch.float64), up(node_ang_off, torch.int32), up(node_pt_off, torch.int32), up(ang, torch.float64)]
    _check(load().b200qc_grid_assemble(int(atompos.shape[0]), _ptr(atompos.contiguous()), *[_ptr(b) for b in bufs], ngrid,
                                       _ptr(xyz), _ptr(dvol), _ptr(owner), _stream()), "grid_assemble")
    return xyz, dvol, owner


def becke_weights(xyz, owner, atompos, aij=None):
    lib = load()
    xyz = xyz.contiguous()
    owner = owner.to(torch.int32).contiguous()
    w = torch.empty(xyz.shape[0], dtype=torch.float64, device=xyz.device)
    _check(lib.b200qc_becke_weights(_ptr(xyz), _ptr(owner), xyz.shape[0], _ptr(atompos.contiguous()),
                                    atompos.shape[0], _ptr(aij), _ptr(w), _stream()), "becke_weights")
    return w


def rho(ao: torch.Tensor, dm_pad: torch.Tensor, with_grad: bool):
    """ao (ncomp, ngrid_ld, ao_ld); dm_pad (ao_ld, ao_ld).  Returns rho (ngrid_ld), grad (3, ngrid_ld)|None."""
    lib = load()
    _, ngl, aol = ao.shape
    r = torch.empty(ngl, dtype=torch.float64, device=ao.device)
    g = torch.empty((3, ngl), dtype=torch.float64, device=ao.device) if with_grad else None
    _check(lib.b200qc_rho(_ptr(ao), ngl, aol, _ptr(dm_pad), _ptr(r), _ptr(g), _stream()), "rho")
    return r, g


def _terms(terms):
    ids = np.array([FUNC_IDS[n] for _, n in terms], dtype=np.int32)
    coefs = np.array([float(c) for c, _ in terms], dtype=np.float64)
    return ids, coefs


def xc_unpol(terms, rho_t, grad_t, want_e=True, want_v=True):
    """terms: [(coef, name)].  rho (n,), grad (3, n)|None -> edens, vrho, vgrad (3, n)|None."""
    lib = load()
    ids, coefs = _terms(terms)
    n = rho_t.shape[-1]
    gga = grad_t is not None
    e = torch.empty_like(rho_t) if want_e else None
    vr = torch.empty_like(rho_t) if want_v else None
    vg = torch.empty_like(grad_t) if (want_v and gga) else None
    _check(lib.b200qc_xc_unpol(len(ids), _np(ids), _np(coefs), n, n, _ptr(rho_t), _ptr(grad_t), _ptr(e), _ptr(vr),
                               _ptr(vg), _stream()), "xc_unpol")
    return e, vr, vg


def xc_mgga_unpol(terms, rho_t, grad_t, lapl_t, kin_t, want_e=True, want_v=True):
    """Meta-GGA sums: rho (n,), grad (3, n), lapl (n,), kin (n,) -> edens, vrho, vgrad (3, n), vlapl, vkin."""
    lib = load()
    ids, coefs = _terms(terms)
    n = rho_t.shape[-1]
    e = torch.empty_like(rho_t) if want_e else None
    vr = torch.empty_like(rho_t) if want_v else None
    vg = torch.empty_like(grad_t) if want_v else None
    vl = torch.empty_like(rho_t) if want_v else None
    vk = torch.empty_like(rho_t) if want_v else None
    _check(lib.b200qc_xc_mgga_unpol(len(ids), _np(ids), _np(coefs), n, n, _ptr(rho_t), _ptr(grad_t), _ptr(lapl_t),
                                    _ptr(kin_t), _ptr(e), _ptr(vr), _ptr(vg), _ptr(vl), _ptr(vk), _stream()), "xc_mgga_unpol")
    return e, vr, vg, vl, vk


def xc_pol(terms, rho_t, grad_t, want_e=True, want_v=True):
    """rho (2, n), grad (2, 3, n)|None -> edens (n,), vrho (2, n), vgrad (2, 3, n)|None."""
    lib = load()
    ids, coefs = _terms(terms)
    n = rho_t.shape[-1]
    gga = grad_t is not None
    e = torch.empty(n, dtype=rho_t.dtype, device=rho_t.device) if want_e else None
    vr = torch.empty_like(rho_t) if want_v else None
    vg = torch.empty_like(grad_t) if (want_v and gga) else None
    _check(lib.b200qc_xc_pol(len(ids), _np(ids), _np(coefs), n, n, _ptr(rho_t), _ptr(grad_t), _ptr(e), _ptr(vr),
                             _ptr(vg), _stream()), "xc_pol")
    return e, vr, vg


_work_cache = {}


def _workspace(numel: int, device) -> torch.Tensor:
    key = (str(device),)
    buf = _work_cache.get(key)
    if buf is None or buf.numel() < numel:
        _work_cache[key] = None
        buf = torch.empty(numel, dtype=torch.float64, device=device)
        _work_cache[key] = buf
    return buf


def vxc_mat(ao, weights, vrho, vgrad):
    """Returns the padded (ao_ld, ao_ld) matrix sum_g w phi^T (vrho phi + 2 vgrad . dphi)."""
    lib = load()
    _, ngl, aol = ao.shape
    work = _workspace(int(lib.b200qc_vxc_worksize(ngl, aol)), ao.device)
    mat = torch.empty((aol, aol), dtype=torch.float64, device=ao.device)
    _check(lib.b200qc_vxc_mat(_ptr(ao), ngl, aol, _ptr(weights), _ptr(vrho), _ptr(vgrad), _ptr(mat), _ptr(work),
                              _stream()), "vxc_mat")
    return mat


def int1e(basis: DeviceBasis, kind: str, shls, rinv_orig=None):
    lib = load()
    ensure_rys_table()
    kinds = {"ovlp": 0, "kin": 1, "nuc": 2, "rinv": 3}
    sl = np.array(shls, dtype=np.int32)
    out = torch.zeros((basis.nao(sl[0], sl[1]), basis.nao(sl[2], sl[3])), dtype=torch.float64, device=basis.device)
    ro = np.zeros(3) if rinv_orig is None else np.ascontiguousarray(rinv_orig, dtype=np.float64)
    _check(lib.b200qc_int1e(basis.handle, kinds[kind], _np(sl), _np(ro), _ptr(out), _stream()), "int1e")
    return out


def int2c2e(basis: DeviceBasis, shls):
    lib = load()
    ensure_rys_table()
    sl = np.array(shls, dtype=np.int32)
    out = torch.zeros((basis.nao(sl[0], sl[1]), basis.nao(sl[2], sl[3])), dtype=torch.float64, device=basis.device)
    _check(lib.b200qc_int2c2e(basis.handle, _np(sl), _ptr(out), _stream()), "int2c2e")
    return out


def int3c2e(basis: DeviceBasis, shls):
    lib = load()
    ensure_rys_table()
    sl = np.array(shls, dtype=np.int32)
    shape = tuple(basis.nao(sl[2 * q], sl[2 * q + 1]) for q in range(3))
    out = torch.zeros(shape, dtype=torch.float64, device=basis.device)
    _check(lib.b200qc_int3c2e(basis.handle, _np(sl), _ptr(out), _stream()), "int3c2e")
    return out


def int2e(basis: DeviceBasis, shls):
    lib = load()
    ensure_rys_table()
    sl = np.array(shls, dtype=np.int32)
    shape = tuple(basis.nao(sl[2 * q], sl[2 * q + 1]) for q in range(4))
    out = torch.zeros(shape, dtype=torch.float64, device=basis.device)
    _check(lib.b200qc_int2e(basis.handle, _np(sl), _ptr(out), _stream()), "int2e")
    return out


def int3c2e_packed(basis: DeviceBasis, shls, ld: Optional[int] = None):
    """(ij|P) for AO pairs i >= j: (npair, ld) with ld = naux rounded up to even, zero padded."""
    lib = load()
    ensure_rys_table()
    sl = np.array(shls, dtype=np.int32)
    nao, naux = basis.nao(sl[0], sl[1]), basis.nao(sl[4], sl[5])
    ld = round_up(naux, 2) if ld is None else ld
    out = torch.zeros((nao * (nao + 1) // 2, ld), dtype=torch.float64, device=basis.device)
    _check(lib.b200qc_int3c2e_packed(basis.handle, _np(sl), _ptr(out), ld, _stream()), "int3c2e_packed")
    return out


class StoredERI(object):
    """Both dense layouts of (ij|kl) resident in HBM (small molecules); J / K are HBM-bound GEMVs."""

    def __init__(self, basis: DeviceBasis, sh0: int, sh1: int):
        lib = load()
        ensure_rys_table()
        n = basis.nao(sh0, sh1)
        self.nao = n
        self.eri_j = torch.empty(n ** 4, dtype=torch.float64, device=basis.device)
        self.eri_k = torch.empty(n ** 4, dtype=torch.float64, device=basis.device)
        with torch.cuda.device(basis.device):
            _check(lib.b200qc_eri_store(basis.handle, sh0, sh1, _ptr(self.eri_j), _ptr(self.eri_k), _stream()),
                   "eri_store")

    @staticmethod
    def nbytes(nao: int) -> int:
        return 2 * 8 * nao ** 4

    def _gemv(self, A, dm):
        lib = load()
        n2 = self.nao * self.nao
        ld = n2 + (n2 & 1)
        if ld != n2:
            raise B200QCError("stored-ERI GEMV needs an even nao^2")  # never hit: handled in run()
        x = dm.reshape(-1).contiguous()
        y = torch.empty(n2, dtype=torch.float64, device=dm.device)
        _check(lib.b200qc_gemv(_ptr(A), n2, n2, n2, _ptr(x), _ptr(y), _stream()), "gemv")
        return y.reshape(self.nao, self.nao)

    def run(self, dm: torch.Tensor, with_j=True, with_k=True, rank=0, world=1):
        """dm (nset, nao, nao) -> vj, vk (nset, nao, nao) (same contract as JKPlan.run; world must be 1)."""
        assert world == 1
        vj = torch.stack([self._gemv(self.eri_j, d) for d in dm]) if with_j else None
        vk = torch.stack([self._gemv(self.eri_k, d) for d in dm]) if with_k else None
        return vj, vk


class JKPlan(object):
    """Schwarz-screened direct J/K plan of shells [sh0, sh1) (b200qc_jkplan_*)."""

    def __init__(self, basis: DeviceBasis, sh0: int, sh1: int, thresh: float = 1e-13):
        lib = load()
        ensure_rys_table()
        self.basis = basis  # keeps the device basis alive
        self.nao = basis.nao(sh0, sh1)
        h = ctypes.c_void_p(0)
        with torch.cuda.device(basis.device):
            _check(lib.b200qc_jkplan_create(basis.handle, sh0, sh1, float(thresh), ctypes.byref(h), _stream()),
                   "jkplan_create")
        self.handle = h
        self.nquartets = int(lib.b200qc_jkplan_nquartets(h))
        self.nquartets_reg = int(lib.b200qc_jkplan_nquartets_reg(h))   # on the register-resident engine (l <= 1 classes)

    def flops(self, with_j=True, with_k=True) -> float:
        """fp64 operations of one build of one density (integrals once per quartet + digestion)."""
        return float(load().b200qc_jkplan_flops(self.handle, int(with_j), int(with_k)))

    def run(self, dm: torch.Tensor, with_j=True, with_k=True, rank=0, world=1):
        """dm (nset, nao, nao) symmetric -> vj, vk (nset, nao, nao); partial sums when world > 1."""
        lib = load()
        dm = dm.contiguous()
        assert dm.ndim == 3 and dm.shape[1] == self.nao
        vj = torch.empty_like(dm) if with_j else None
        vk = torch.empty_like(dm) if with_k else None
        _check(lib.b200qc_jkplan_run(self.handle, _ptr(dm), dm.shape[0], _ptr(vj), _ptr(vk), rank, world,
                                     _stream()), "jkplan_run")
        return vj, vk

    def __del__(self):
        try:
            if self.handle:
                load(False).b200qc_jkplan_free(self.handle)
                self.handle = None
        except Exception:
            pass


def jk_direct(basis: DeviceBasis, sh0, sh1, dm, with_j=True, with_k=True):
    """dm (nset, nao, nao) symmetric AO basis -> vj, vk (nset, nao, nao) (either may be None)."""
    lib = load()
    ensure_rys_table()
    dm = dm.contiguous()
    vj = torch.zeros_like(dm) if with_j else None
    vk = torch.zeros_like(dm) if with_k else None
    _check(lib.b200qc_jk_direct(basis.handle, sh0, sh1, _ptr(dm), dm.shape[0], _ptr(vj), _ptr(vk), _stream()),
           "jk_direct")
    return vj, vk


def pack_tril(full, ld=None):
    lib = load()
    nao, _, naux = full.shape
    ld = round_up(naux, 2) if ld is None else ld
    out = torch.empty((nao * (nao + 1) // 2, ld), dtype=torch.float64, device=full.device)
    _check(lib.b200qc_pack_tril(_ptr(full.contiguous()), nao, naux, ld, _ptr(out), _stream()), "pack_tril")
    return out


def dfj(j3c_packed, nao, naux, inv_j2c, dm):
    """One-GPU density-fitted J: (nao, nao) from the packed (npair, ld) tensor."""
    lib = load()
    ld = j3c_packed.shape[1]
    work = _workspace(int(lib.b200qc_dfj_worksize(nao, ld)), dm.device)
    vj = torch.empty((nao, nao), dtype=torch.float64, device=dm.device)
    _check(lib.b200qc_dfj(_ptr(j3c_packed), nao, naux, ld, _ptr(inv_j2c.contiguous()), _ptr(dm.contiguous()),
                          _ptr(vj), _ptr(work), _stream()), "dfj")
    return vj


def dfj_rowmask(j3c_packed, nao, naux, thresh):
    """uint8 (npair,): 1 for pair rows whose every |(ij|P)| is below thresh (skipped by the masked passes)."""
    lib = load()
    mask = torch.empty(j3c_packed.shape[0], dtype=torch.uint8, device=j3c_packed.device)
    _check(lib.b200qc_dfj_rowmask(_ptr(j3c_packed), nao, naux, j3c_packed.shape[1], float(thresh), _ptr(mask),
                                  _stream()), "dfj_rowmask")
    return mask


def dfj_pass1(j3c_packed, nao, naux, dm, rows=None):
    """temp_P = sum_ij D_ij (ij|P); rows: optional int32 list of the pair rows to read (ascending)."""
    lib = load()
    ld = j3c_packed.shape[1]
    work = _workspace(int(lib.b200qc_dfj_worksize(nao, ld)), dm.device)
    temp = torch.empty(naux, dtype=torch.float64, device=dm.device)
    assert rows is None or rows.dtype == torch.int32
    _check(lib.b200qc_dfj_pass1_rows(_ptr(j3c_packed), nao, naux, ld, _ptr(dm.contiguous()), _ptr(temp), _ptr(work),
                                     _ptr(rows), 0 if rows is None else rows.numel(), _stream()), "dfj_pass1")
    return temp


def dfj_pass2(j3c_packed, nao, naux, coef, mask=None):
    lib = load()
    ld = j3c_packed.shape[1]
    cpad = torch.zeros(ld, dtype=torch.float64, device=coef.device)
    cpad[:naux] = coef
    vj = torch.empty((nao, nao), dtype=torch.float64, device=coef.device)
    _check(lib.b200qc_dfj_pass2_masked(_ptr(j3c_packed), nao, naux, ld, _ptr(cpad), _ptr(vj), _ptr(mask), _stream()),
           "dfj_pass2")
    return vj


# ------------------------------------------------------------------------------------------
# block-sparse grid path (csrc/xc_sb.cuh)

SB_DTYPE = np.dtype([("ao_off", np.int64), ("d_off", np.int64), ("nsp", np.int32), ("idx_off", np.int32),
                     ("shell_off", np.int32), ("nshell", np.int32), ("dsb_idx_off", np.int32), ("pad", np.int32)])


# ------------------------------------------------------------------------------------------
# fp64-accurate GEMM on tcgen05 (sliced int8, csrc/gemm_i8.cuh)

class I8Operand(object):
    """One fp64 operand of ``gemm_i8`` cut into int8 planes in the tiled UMMA order (b200qc_i8_slice).

    Logical shape (nbatch, R, K): element (b, r, k) is read from ``src`` (a CUDA fp64 tensor, any layout) at
    element offset ``b * sb + r * sr + k * sk`` -- or, with ``pair_ld``, from the packed (ij|P) tensor
    (b = i, k = j, r = P).  ``role`` is "A" (128-row tiles) or "B" (64-row tiles)."""

    def __init__(self, role: str, nbatch: int, R: int, K: int, nslice: int = 6, K_last: Optional[int] = None,
                 device=None):
        assert role in ("A", "B")
        self.role, self.nbatch, self.R, self.K, self.S = role, int(nbatch), int(R), int(K), int(nslice)
        self.K_last = self.K if K_last is None else int(K_last)
        self.W = 128 if role == "A" else 64
        self.Rpad, self.Kpad = round_up(self.R, self.W), round_up(self.K, 32)
        self.rtiles, self.nk, self.nk_last = self.Rpad // self.W, self.Kpad // 32, round_up(self.K_last, 32) // 32
        self.bstride = self.Rpad * self.Kpad * self.S           # bytes per batch
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.planes = torch.empty(self.nbatch * self.bstride, dtype=torch.int8, device=dev)
        self.scales = torch.empty(self.nbatch * self.Rpad, dtype=torch.float64, device=dev)

    def fill(self, src: torch.Tensor, sb: int, sr: int, sk: int, pair_ld: int = 0, offset: int = 0):
        assert src.is_cuda and src.dtype == torch.float64
        ptr = ctypes.c_void_p(src.data_ptr() + 8 * int(offset))
        _check(load().b200qc_i8_slice(ptr, self.nbatch, int(sb), int(sr), int(sk), 1 if pair_ld else 0, int(pair_ld),
                                      self.R, self.K, self.K_last, self.Rpad, self.Kpad, self.W, self.S,
                                      _ptr(self.planes), _ptr(self.scales), _stream()), "i8_slice")
        return self


def i8_slice_dual(src: torch.Tensor, nbatch: int, sb: int, sr: int, rowmax: torch.Tensor, rm_ld: int, R: int, K: int,
                  K_last: int, nslice: int = 6):
    """One pass over a K-contiguous fp64 operand with known row maxima -> (A-form, B-form) I8Operands."""
    a = I8Operand("A", nbatch, R, K, nslice, K_last=K_last, device=src.device)
    b = I8Operand("B", nbatch, R, K, nslice, K_last=K_last, device=src.device)
    _check(load().b200qc_i8_slice_dual(_ptr(src), int(nbatch), int(sb), int(sr), _ptr(rowmax), int(rm_ld), int(R),
                                       int(K), int(K_last), a.Kpad, int(nslice), _ptr(a.planes), _ptr(a.scales),
                                       _ptr(b.planes), _ptr(b.scales), _stream()), "i8_slice_dual")
    return a, b


def gemm_i8(a: I8Operand, b: I8Operand, out: torch.Tensor, c_bstride: int, ldc: int, M: int, N: int, mode: int = 0,
            alpha: float = 1.0, nbatch: Optional[int] = None, a_shared: bool = False, b_shared: bool = False,
            rowmax: Optional[torch.Tensor] = None, rm_bstride: int = 0, rm_div: int = 1):
    """out[b][m][n] (=, +=) alpha * sum_k A[b][m][k] B[b][n][k] on the tcgen05 int8 engine (b200qc_gemm_i8)."""
    assert a.role == "A" and b.role == "B" and a.S == b.S and a.nk == b.nk and a.nk_last == b.nk_last
    nb = int(nbatch if nbatch is not None else max(a.nbatch, b.nbatch))
    assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
    _check(load().b200qc_gemm_i8(_ptr(a.planes), _ptr(a.scales), 0 if a_shared else a.bstride, 0 if a_shared else a.Rpad,
                                 _ptr(b.planes), _ptr(b.scales), 0 if b_shared else b.bstride, 0 if b_shared else b.Rpad,
                                 nb, a.rtiles, b.rtiles, a.nk, a.nk_last, a.S, int(M), int(N), float(alpha),
                                 _ptr(out), int(c_bstride), int(ldc), int(mode), _ptr(rowmax), int(rm_bstride),
                                 int(rm_div), _stream()), "gemm_i8")
    return out


I8_KCHUNK = {5: 65536, 6: 32768}   # longest K per batch with exact integer accumulation


def gemm_f64emu(x: torch.Tensor, y: torch.Tensor, nslice: int = 6, kchunk: Optional[int] = None) -> torch.Tensor:
    """x (M, K) @ y (N, K)^T in fp64 accuracy on the int8 tensor cores (split over K chunks when K is long)."""
    assert x.ndim == 2 and y.ndim == 2 and x.shape[1] == y.shape[1]
    x, y = x.contiguous(), y.contiguous()
    M, K = x.shape
    N = y.shape[0]
    kc = min(I8_KCHUNK[nslice] if kchunk is None else int(kchunk), round_up(K, 32))
    nchunk = (K + kc - 1) // kc
    k_last = K - (nchunk - 1) * kc
    a = I8Operand("A", nchunk, M, kc, nslice, K_last=k_last, device=x.device).fill(x, kc, K, 1)
    b = I8Operand("B", nchunk, N, kc, nslice, K_last=k_last, device=x.device).fill(y, kc, K, 1)
    out = torch.zeros(M, N, dtype=torch.float64, device=x.device)
    return gemm_i8(a, b, out, 0, N, M, N, mode=0 if nchunk == 1 else 1)


def ao_screen(basis: DeviceBasis, sh0: int, sh1: int, coords: torch.Tensor, sbp: int, eps: float, deriv: int):
    """(nsb, nshell) uint8: 1 where the shell's envelope can exceed eps somewhere on the superblock."""
    lib = load()
    ngrid = coords.shape[0]
    nsb = (ngrid + sbp - 1) // sbp
    flags = torch.zeros((nsb, sh1 - sh0), dtype=torch.uint8, device=coords.device)
    _check(lib.b200qc_ao_screen(basis.handle, sh0, sh1, _ptr(coords.contiguous()), ngrid, sbp, float(eps), deriv,
                                _ptr(flags), _stream()), "ao_screen")
    return flags


class GridBlocks(object):
    """Superblock decomposition of a grid slice: screening flags -> compact AO storage + descriptors.

    Built once per geometry (HamiltonCGTO.setup_grid); ``rho`` and ``vxc_mat`` are the per-iteration
    kernels (K2 / K4) on it."""

    def __init__(self, basis: DeviceBasis, sh0: int, sh1: int, coords: torch.Tensor, weights: torch.Tensor,
                 deriv: int, sbp: int = 1024, eps: float = 1e-12, flags: Optional[np.ndarray] = None,
                 i8_slices: int = 0, i8_variant: int = 0, rho_i8_slices: int = 0):
        lib = load()
        dev = coords.device
        self.basis, self.sh0, self.sh1 = basis, sh0, sh1
        self.ngrid = int(coords.shape[0])
        self.sbp, self.deriv, self.eps = int(sbp), int(deriv), float(eps)
        self.ncomp = {0: 1, 1: 4, 2: 5}[self.deriv]      # deriv 2: phi, grad phi, lapl phi (meta-GGA)
        if self.deriv == 2:
            i8_slices = rho_i8_slices = 0                 # the meta-GGA contractions run on the fp64 DMMA engine
        self.nsb = (self.ngrid + self.sbp - 1) // self.sbp
        self.ngl = self.nsb * self.sbp
        self.nao = basis.nao(sh0, sh1)
        coords = coords.contiguous()
        if flags is None:
            flags = ao_screen(basis, sh0, sh1, coords, self.sbp, eps, deriv).cpu().numpy()
        kept = flags.astype(bool)
        nshell = sh1 - sh0
        loc = basis.ao_loc[sh0:sh1 + 1].astype(np.int64) - int(basis.ao_loc[sh0])
        sizes = (loc[1:] - loc[:-1]).astype(np.int64)
        ksz = kept * sizes[None, :]
        nsig = ksz.sum(1)
        nsp = np.maximum(64, (nsig + 63) // 64 * 64).astype(np.int64)
        nsh = kept.sum(1).astype(np.int64)
        excl = lambda a: np.concatenate([[0], np.cumsum(a)[:-1]]).astype(np.int64)
        sb_i, sh_i = np.nonzero(kept)
        col = (np.cumsum(ksz, axis=1) - ksz)[kept]                       # compact first column of each kept shell
        idx_off, shell_off = excl(nsp), excl(nsh)
        ao_off, d_off, vb_off = excl(self.ncomp * self.sbp * nsp), excl(nsp * nsp), excl(self.sbp * nsp)
        idx = np.full(int(nsp.sum()), self.nao, dtype=np.int32)
        rep = sizes[sh_i]
        if rep.sum() > 0:
            start = np.repeat(idx_off[sb_i] + col, rep)
            within = np.arange(int(rep.sum()), dtype=np.int64) - np.repeat(excl(rep), rep)
            idx[start + within] = (np.repeat(loc[sh_i], rep) + within).astype(np.int32)
        desc = np.zeros(self.nsb, dtype=SB_DTYPE)
        desc["ao_off"], desc["d_off"], desc["nsp"] = ao_off, d_off, nsp
        desc["idx_off"], desc["shell_off"], desc["nshell"] = idx_off, shell_off, nsh
        # superblocks with the kept-AO list of their predecessor (consecutive radial shells of an atom: a quarter of
        # them at C60) use its gathered, sliced D_sb: rep[sb] = first superblock of the run
        same = np.zeros(self.nsb, dtype=bool)
        if self.nsb > 1:
            same[1:] = (kept[1:] == kept[:-1]).all(1)
        rep = np.arange(self.nsb)
        for sb in range(1, self.nsb):
            if same[sb]:
                rep[sb] = rep[sb - 1]
        self.sb_rep = rep
        desc["dsb_idx_off"] = idx_off[rep]
        self.nsp = nsp
        self.max_nsp = int(nsp.max()) if self.nsb else 64
        self.kept_fraction = float(nsig.sum()) / max(1, self.nao * self.nsb)
        self.flops_per_pass = float(2.0 * self.sbp * (nsp.astype(np.float64) ** 2).sum())   # K2 == K4 GEMM flops
        self.ao_bytes = float(self.ncomp * self.sbp * nsp.sum() * 8)
        tt = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a)).to(dt).to(dev)
        self.d_desc = torch.as_tensor(desc.view(np.uint8)).to(dev)
        self.d_idx = tt(idx, torch.int32)
        self.d_shell_ids = tt(sh_i + sh0, torch.int32)
        self.d_shell_col = tt(col, torch.int32)
        self.d_vb_off = tt(vb_off, torch.int64)
        self.ao = torch.zeros(int(self.ncomp * self.sbp * nsp.sum()), dtype=torch.float64, device=dev)
        self._dsb = self._vb = None      # fp64 scratch of the fp64-DMMA / unfused kernels, allocated on first use
        self.w = torch.zeros(self.ngl, dtype=torch.float64, device=dev)
        self.w[:self.ngrid] = weights
        if self.nsb:
            _check(lib.b200qc_eval_gto_sb(basis.handle, deriv, _ptr(coords), self.ngrid, self.sbp, self.nsb,
                                          _ptr(self.d_desc), _ptr(self.d_shell_ids), _ptr(self.d_shell_col),
                                          _ptr(self.ao), _stream()), "eval_gto_sb")
        # optional tcgen05 int8 (Ozaki) form of the Vxc GEMM: the AO values are sliced once here
        self.i8_slices, self.i8_variant = int(i8_slices), int(i8_variant)
        from dqc_b200.utils.config import config as _cfg
        _check(lib.b200qc_i8_mode(int(_cfg.I8_MODE)), "i8_mode")
        lib.b200qc_i8_debug_variant(self.i8_variant)
        if self.i8_slices and self.nsb:
            S = self.i8_slices
            # N tile of the GEMM: 96 with 5 slices (5 x 96 = 480 TMEM columns), else 64
            self.i8_bn = bn = 96 if (S == 5 and _cfg.VXC_I8_BN == 96) else 64
            a_bytes = S * self.sbp * ((nsp + 127) // 128 * 128)      # A operand: whole 128-column M tiles
            b_bytes = S * self.sbp * ((nsp + bn - 1) // bn * bn)      # B operand: whole N tiles (zero padding)
            ntile = ((nsp + 127) // 128) * ((nsp + bn - 1) // bn)
            self.d_tile_off = tt(excl(ntile), torch.int32)
            self.ntiles = int(ntile.sum())
            nptile = ((nsp + 127) // 128) ** 2                        # cluster mode: (M tile, pair of N tiles) units
            self.d_ptile_off = tt(excl(nptile), torch.int32)
            self.nptiles = int(nptile.sum())
            self.d_a_off = tt(excl(a_bytes), torch.int64)
            self.d_b_off = tt(excl(b_bytes), torch.int64)
            self.aplanes = torch.zeros(int(a_bytes.sum()), dtype=torch.int8, device=dev)
            self.bplanes = torch.zeros(int(b_bytes.sum()), dtype=torch.int8, device=dev)
            self.ascale = torch.empty(int(nsp.sum()), dtype=torch.float64, device=dev)
            self.bscale = torch.empty(int(nsp.sum()), dtype=torch.float64, device=dev)
            # fused operand preparation (vb cut into int8 planes as it is formed): static column maxima per superblock
            self.colmax = torch.empty(int(nsp.sum()) * (self.sbp // 32) * self.ncomp, dtype=torch.float32, device=dev) \
                if (_cfg.VXC_FUSED_VB and self.sbp <= 1536) else None
            # ... and the per-call flags of the 64-column blocks whose bound came out loose (cut again with exact exponents)
            self.fixflag = torch.zeros(self.nsb * (self.max_nsp // 64), dtype=torch.int32, device=dev) \
                if self.colmax is not None else None
            _check(lib.b200qc_vxc_i8_prepare(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, S, self.ncomp,
                                             _ptr(self.ao), _ptr(self.d_a_off), _ptr(self.aplanes), _ptr(self.ascale),
                                             _ptr(self.colmax), _stream()),
                   "vxc_i8_prepare")

        # optional tcgen05 int8 form of the density GEMM: AO rows sliced once here
        self.rho_i8_slices = int(rho_i8_slices)
        if self.rho_i8_slices and self.nsb:
            S = self.rho_i8_slices
            # row tile of the sliced density: 128 = point-stationary kernel (phi planes in 64-row tiles, read from HBM
            # once), 64 / 96 = the round-1 kernel (phi planes in 128-row tiles, streamed once per N tile)
            self.rho_bn = rbn = _cfg.RHO_I8_BN if (_cfg.RHO_I8_BN in (64, 128) or (S == 5 and _cfg.RHO_I8_BN == 96)) else 128
            rb_bytes = S * nsp * ((nsp + rbn - 1) // rbn * rbn)      # sliced density: whole row tiles (zero padding)
            self.d_ra_off = tt(excl(S * self.sbp * nsp), torch.int64)
            rb_off = excl(np.where(self.sb_rep == np.arange(self.nsb), rb_bytes, 0))   # planes of the distinct lists only
            self.d_rb_off = tt(rb_off[self.sb_rep], torch.int64)
            rb_bytes = np.where(self.sb_rep == np.arange(self.nsb), rb_bytes, 0)
            self.r_aplanes = torch.empty(int((S * self.sbp * nsp).sum()), dtype=torch.int8, device=dev)
            self.r_bplanes = torch.zeros(int(rb_bytes.sum()), dtype=torch.int8, device=dev)
            self.r_rscale = torch.empty(self.nsb * self.sbp, dtype=torch.float64, device=dev)
            self.r_cscale = torch.empty(int(nsp.sum()), dtype=torch.float64, device=dev)
            _check(lib.b200qc_rho_i8_prepare(_ptr(self.d_desc), self.nsb, self.sbp, S, 64 if rbn == 128 else 128, _ptr(self.ao),
                                             _ptr(self.d_ra_off), _ptr(self.r_aplanes), _ptr(self.r_rscale), _stream()),
                   "rho_i8_prepare")

    @property
    def dsb(self) -> torch.Tensor:
        if self._dsb is None:
            self._dsb = torch.empty(int((self.nsp * self.nsp).sum()), dtype=torch.float64, device=self.ao.device)
        return self._dsb

    @property
    def vb(self) -> torch.Tensor:
        if self._vb is None:
            self._vb = torch.empty(int(self.sbp * self.nsp.sum()), dtype=torch.float64, device=self.ao.device)
        return self._vb

    def rho(self, dm: torch.Tensor, with_grad: bool):
        """dm (nao, nao) symmetric AO-basis density -> rho (ngl,), grad (3, ngl) | None (zero in the padding)."""
        lib = load()
        assert dm.shape == (self.nao, self.nao) and (not with_grad or self.deriv)
        r = torch.empty(self.ngl, dtype=torch.float64, device=dm.device)
        g = torch.empty((3, self.ngl), dtype=torch.float64, device=dm.device) if with_grad else None
        if self.rho_i8_slices and self.nsb:
            _check(lib.b200qc_rho_sb_i8(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, self.rho_i8_slices,
                                        _ptr(self.d_idx), _ptr(self.ao), _ptr(dm.contiguous()), self.nao,
                                        _ptr(self.r_aplanes), _ptr(self.d_ra_off), _ptr(self.r_rscale),
                                        _ptr(self.r_bplanes), _ptr(self.d_rb_off), _ptr(self.r_cscale), self.rho_bn,
                                        _ptr(r), _ptr(g), _stream()), "rho_sb_i8")
            return r, g
        _check(lib.b200qc_rho_sb(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, _ptr(self.d_idx), _ptr(self.ao),
                                 _ptr(dm.contiguous()), self.nao, _ptr(self.dsb), _ptr(r), _ptr(g), _stream()), "rho_sb")
        return r, g

    def rho_mgga(self, dm: torch.Tensor):
        """dm (nao, nao) -> rho (ngl,), grad (3, ngl), lapl rho (ngl,), tau (ngl,)   (hcgto.py:399-438)."""
        assert self.deriv == 2 and dm.shape == (self.nao, self.nao)
        dev = dm.device
        r, lp, kn = (torch.empty(self.ngl, dtype=torch.float64, device=dev) for _ in range(3))
        g = torch.empty((3, self.ngl), dtype=torch.float64, device=dev)
        if self.nsb:
            _check(load().b200qc_rho_sb_mgga(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, _ptr(self.d_idx),
                                             _ptr(self.ao), _ptr(dm.contiguous()), self.nao, _ptr(self.dsb), _ptr(r), _ptr(g),
                                             _ptr(lp), _ptr(kn), _stream()), "rho_sb_mgga")
        return r, g, lp, kn

    def vxc_mat_mgga(self, vrho, vgrad, vlapl, vkin) -> torch.Tensor:
        """(nao, nao) = sum_g w [phi^T (vrho phi + 2 vgrad . grad phi + 2 vlapl lapl phi) + sum_d dphi_d^T (2 vlapl + vkin / 2) dphi_d]."""
        assert self.deriv == 2 and vrho.shape[0] == self.ngl
        mat = torch.empty((self.nao, self.nao), dtype=torch.float64, device=vrho.device)
        _check(load().b200qc_vxc_sb_mgga(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, _ptr(self.d_idx), _ptr(self.ao),
                                         _ptr(self.w), _ptr(vrho.contiguous()), _ptr(vgrad.contiguous()),
                                         _ptr(vlapl.contiguous()), _ptr(vkin.contiguous()), self.nao, _ptr(self.d_vb_off),
                                         _ptr(self.vb), _ptr(mat), _stream()), "vxc_sb_mgga")
        return mat

    def vxc_mat(self, vrho: torch.Tensor, vgrad: Optional[torch.Tensor]) -> torch.Tensor:
        """vrho (ngl,), vgrad (3, ngl) | None -> (nao, nao) = sum_g w phi^T (vrho phi + 2 vgrad . grad phi)."""
        lib = load()
        assert vrho.shape[0] == self.ngl and (vgrad is None or self.deriv)
        mat = torch.empty((self.nao, self.nao), dtype=torch.float64, device=vrho.device)
        if self.i8_slices and self.nsb:
            # the fused slicer bounds vb with the maxima of ALL stored components: an LDA potential on a GGA grid
            # (vgrad None, 4 components stored) takes the two-pass form
            fused = self.colmax is not None and (vgrad is not None or self.ncomp == 1)
            _check(lib.b200qc_vxc_sb_i8(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, self.i8_slices,
                                        _ptr(self.d_idx), _ptr(self.ao), _ptr(self.w), _ptr(vrho.contiguous()),
                                        _ptr(vgrad), self.nao, _ptr(self.d_vb_off), _ptr(None if fused else self.vb),
                                        _ptr(self.colmax if fused else None), _ptr(self.fixflag if fused else None),
                                        _ptr(self.aplanes),
                                        _ptr(self.d_a_off), _ptr(self.ascale), _ptr(self.bplanes), _ptr(self.d_b_off),
                                        _ptr(self.bscale), self.i8_bn, _ptr(self.d_tile_off), self.ntiles, _ptr(self.d_ptile_off),
                                        self.nptiles, _ptr(mat), _stream()),
                   "vxc_sb_i8")
            return mat
        _check(lib.b200qc_vxc_sb(_ptr(self.d_desc), self.nsb, self.sbp, self.max_nsp, _ptr(self.d_idx), _ptr(self.ao),
                                 _ptr(self.w), _ptr(vrho.contiguous()), _ptr(vgrad), self.nao, _ptr(self.d_vb_off),
                                 _ptr(self.vb), _ptr(mat), _stream()), "vxc_sb")
        return mat

    def dense_ao(self) -> torch.Tensor:
        """(ncomp, ngrid, nao) dense AO values rebuilt from the compact storage (API parity / tests only)."""
        out = torch.zeros((self.ncomp, self.ngrid, self.nao + 1), dtype=torch.float64, device=self.ao.device)
        idx = self.d_idx.long()
        desc = self.d_desc.cpu().numpy().view(SB_DTYPE)
        for sb in range(self.nsb):
            n, a0, i0 = int(desc["nsp"][sb]), int(desc["ao_off"][sb]), int(desc["idx_off"][sb])
            blk = self.ao[a0:a0 + self.ncomp * self.sbp * n].reshape(self.ncomp, self.sbp, n)
            r0 = sb * self.sbp
            r1 = min(r0 + self.sbp, self.ngrid)
            out[:, r0:r1, idx[i0:i0 + n]] = blk[:, :r1 - r0, :]
        return out[:, :, :self.nao]
