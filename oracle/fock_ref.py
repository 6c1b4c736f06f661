"""ORACLE (test infrastructure only): the reference's Fock-build path restated on CPU.

Per-iteration torch ops are the reference's own, line for line in meaning:
  orthogonaliser X = U s^-1/2 (eigenvalues > 1e-6)            dqc/hamilton/orbconverter.py:67-116
  J (dense ERI), K (dense ERI)                                 dqc/hamilton/hcgto.py:204-241
  density on grid (16 MiB chunks), Vxc integration             dqc/hamilton/hcgto.py:371-495
  density-fitted J (both THRESHOLD_MEMORY branches equivalent) dqc/df/dfmol.py:24-79
  energies                                                     dqc/hamilton/hcgto.py:302-328
with the native pieces (libcint, libcgto, libxc) replaced by oracle/cint_oracle.c and
oracle/xc_ref.py.  This module is also what bench.py times as the CPU baseline
("reference-equivalent PyTorch-CPU path, libcint/libxc replaced by the oracle").
"""
import numpy as np
import torch
from oracle import cint, xc_ref, becke_ref

CHUNK_MEMORY = 16 * 1024 ** 2  # dqc/utils/config.py:10


def chunk_ranges(nrow, ncol, itemsize=8):
    """Row ranges of the (nrow, ncol) AO tensor as dqc/utils/mem.py:6-38 chunkify(dim=0) yields them."""
    maxnumel = CHUNK_MEMORY // itemsize
    csize = max(maxnumel // ncol, 1)
    return [(i, min(i + csize, nrow)) for i in range(0, nrow, csize)]


class RefHamilton:
    """CPU mirror of HamiltonCGTO (+DFMol) fed by oracle integrals."""

    def __init__(self, wrapper, auxwrapper=None, orthozer=True):
        atm, bas, env = wrapper.atm_bas_env
        self.wrapper = wrapper
        self.atm, self.bas, self.env = atm, bas, env
        s0, s1 = wrapper.shell_idxs
        self.shl = (s0, s1)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
        sl2 = (s0, s1, s0, s1)
        self._t, self._sl2 = t, sl2
        self.S = t(cint.int1e("ovlp", atm, bas, env, sl2))
        ev, evec = torch.linalg.eigh(self.S)
        if orthozer:
            keep = ev > 1e-6
            self.X = evec[:, keep] * ev[keep] ** -0.5
        else:
            self.X = torch.eye(self.S.shape[0], dtype=torch.float64)
        self.orthozer = orthozer
        self.nao = self.X.shape[1]
        self.olp_mat = self.S if not orthozer else self.conv2(self.S)
        self._kinnucl = None
        self.el_mat = None
        self.j3c = None
        self.aux = auxwrapper

    # T and V are built on first use: the grid-only checks at the C60 / taxol sizes never need them
    @property
    def T(self):
        if not hasattr(self, "_T"):
            self._T = self._t(cint.int1e("kin", self.atm, self.bas, self.env, self._sl2))
        return self._T

    @property
    def V(self):
        if not hasattr(self, "_V"):
            self._V = self._t(cint.int1e("nuc", self.atm, self.bas, self.env, self._sl2))
        return self._V

    @property
    def kinnucl_mat(self):
        if self._kinnucl is None:
            self._kinnucl = self.conv2(self.T + self.V)
        return self._kinnucl

    # ---- build ----
    def build_eri(self):
        s0, s1 = self.shl
        eri = torch.as_tensor(cint.int2e(self.atm, self.bas, self.env, (s0, s1) * 4))
        X = self.X
        if self.orthozer:  # convert4, orbconverter.py:99-107 (done index by index: same numbers)
            eri = torch.einsum("ijkl,im->mjkl", eri, X)
            eri = torch.einsum("mjkl,jn->mnkl", eri, X)
            eri = torch.einsum("mnkl,kp->mnpl", eri, X)
            eri = torch.einsum("mnpl,lq->mnpq", eri, X)
        self.el_mat = eri
        return self

    def build_df(self):
        """aux wrapper must come from LibcintWrapper.concatenate(basis, aux) (dfmol.py:30-33)."""
        basisw, auxw = self.wrapper, self.aux
        atm, bas, env = auxw.atm_bas_env
        a0, a1 = auxw.shell_idxs
        b0, b1 = basisw.shell_idxs
        self.j2c = torch.as_tensor(cint.int2c2e(atm, bas, env, (a0, a1, a0, a1)))
        self.j3c = torch.as_tensor(cint.int3c2e(atm, bas, env, (b0, b1, b0, b1, a0, a1)))
        self.inv_j2c = torch.inverse(self.j2c)
        return self

    def setup_grid(self, rgrid, dvolume, xcstr):
        self.xcstr = xcstr
        self.xcfamily = 1 if xcstr is None else xc_ref.family(xcstr)
        self.rgrid = np.asarray(rgrid)
        self.dvolume = torch.as_tensor(np.asarray(dvolume))
        self.basis = torch.as_tensor(cint.eval_gto(self.atm, self.bas, self.env, self.rgrid, 0, self.shl))
        self.basis_dvolume = self.basis * self.dvolume.unsqueeze(-1)
        if self.xcfamily >= 2:
            self.grad_basis = torch.as_tensor(cint.eval_gto(self.atm, self.bas, self.env, self.rgrid, 1, self.shl))
        if self.xcfamily == 4:      # hcgto.py:183-186
            self.lapl_basis = torch.as_tensor(cint.eval_gto(self.atm, self.bas, self.env, self.rgrid, 2, self.shl))

    # ---- orbital converter ----
    def conv2(self, m):
        return self.X.T @ m @ self.X

    def unconv_dm(self, dm):
        return self.X @ dm @ self.X.T

    # ---- operators (orthogonalised basis in, orthogonalised basis out) ----
    def get_elrep(self, dm):
        if self.j3c is None:
            mat = torch.einsum("...ij,ijkl->...kl", dm, self.el_mat)
            return (mat + mat.transpose(-2, -1)) * 0.5
        dmao = self.unconv_dm(dm)
        temp = torch.einsum("...ij,ijl->...l", dmao, self.j3c)
        coef = torch.einsum("...l,lk->...k", temp, self.inv_j2c)
        mat = torch.einsum("...k,ijk->...ij", coef, self.j3c)
        mat = (mat + mat.transpose(-2, -1)) * 0.5
        return self.conv2(mat)

    def get_exchange(self, dm):
        if self.j3c is not None and self.el_mat is None:
            return self.get_exchange_df(dm)
        mat = -0.5 * torch.einsum("...il,ijkl->...ijk", dm, self.el_mat).sum(dim=-3)
        return (mat + mat.transpose(-2, -1)) * 0.5

    def get_exchange_df(self, dm):
        """Density-fitted exact exchange -- NOT in the reference (it raises, hcgto.py:229-230).  The extension is
        the one SURVEY 8a defines, with the reference's own conventions for the factor and the symmetrisation
        (hcgto.py:234-241): K'_ij = -1/2 sum_PQ sum_kl (ik|P) (P|Q)^-1 (Q|jl) D_kl, written with the plain
        inverse the reference uses for J (dfmol.py:48)."""
        dmao = self.unconv_dm((dm + dm.transpose(-2, -1)) * 0.5)
        half = torch.einsum("ikp,kl->ilp", self.j3c, dmao)                 # (i, l, P)
        half = torch.einsum("ilp,pq->ilq", half, self.inv_j2c)
        mat = -0.5 * torch.einsum("ilq,jlq->ij", half, self.j3c)
        mat = (mat + mat.transpose(-2, -1)) * 0.5
        return self.conv2(mat)

    def dm2densinfo(self, dm):
        dmdmt = self.unconv_dm((dm + dm.transpose(-2, -1)) * 0.5)
        ng, nb = self.basis.shape
        dens = torch.empty(ng, dtype=torch.float64)
        gdens = torch.empty(3, ng, dtype=torch.float64) if self.xcfamily >= 2 else None
        for i0, i1 in chunk_ranges(ng, nb):
            b = self.basis[i0:i1]
            dmao = torch.matmul(b, dmdmt)
            dens[i0:i1] = torch.einsum("ri,ri->r", dmao, b)
            if gdens is not None:
                for d in range(3):
                    gdens[d, i0:i1] = torch.einsum("ri,ri->r", dmao, self.grad_basis[d, i0:i1]) * 2
        return dens, gdens

    def dm2densinfo_mgga(self, dm):
        """hcgto.py:399-438 with xcfamily == 4: (dens, gdens, lapldens, kindens)."""
        dmdmt = self.unconv_dm((dm + dm.transpose(-2, -1)) * 0.5)
        ng, nb = self.basis.shape
        dens, lapldens, kindens = (torch.empty(ng, dtype=torch.float64) for _ in range(3))
        gdens = torch.empty(3, ng, dtype=torch.float64)
        for i0, i1 in chunk_ranges(ng, nb):
            b = self.basis[i0:i1]
            dmao = torch.matmul(b, dmdmt)
            dens[i0:i1] = torch.einsum("ri,ri->r", dmao, b)
            gg = torch.zeros(i1 - i0, dtype=torch.float64)
            for d in range(3):
                gb = self.grad_basis[d, i0:i1]
                gdens[d, i0:i1] = torch.einsum("ri,ri->r", dmao, gb) * 2
                gg += torch.einsum("ri,ri->r", torch.matmul(gb, dmdmt), gb)
            lapl_basis = torch.einsum("ri,ri->r", dmao, self.lapl_basis[i0:i1])
            lapldens[i0:i1] = (lapl_basis + gg) * 2
            kindens[i0:i1] = gg * 0.5
        return dens, gdens, lapldens, kindens

    def vxc_from_potinfo_mgga(self, vrho, vgrad, vlapl, vkin):
        """hcgto.py:445-495 with xcfamily == 4."""
        ng, nb = self.basis.shape
        mat = torch.zeros(nb, nb, dtype=torch.float64)
        for i0, i1 in chunk_ranges(ng, nb):
            vb = vrho[i0:i1].unsqueeze(-1) * self.basis[i0:i1]
            vg = vgrad[:, i0:i1] * 2
            for d in range(3):
                vb += vg[d].unsqueeze(-1) * self.grad_basis[d, i0:i1]
            lapl, kin = vlapl[i0:i1], vkin[i0:i1]
            vb += 2 * lapl.unsqueeze(-1) * self.lapl_basis[i0:i1]
            mat += torch.matmul(self.basis_dvolume[i0:i1].T, vb)
            lapl_kin_dvol = (2 * lapl + 0.5 * kin) * self.dvolume[i0:i1]
            for d in range(3):
                gb = self.grad_basis[d, i0:i1]
                mat += torch.einsum("r,rb,rc->bc", lapl_kin_dvol, gb, gb)
        mat = self.conv2(mat)
        return (mat + mat.T) * 0.5

    def vxc_from_potinfo(self, vrho, vgrad):
        ng, nb = self.basis.shape
        mat = torch.zeros(nb, nb, dtype=torch.float64)
        for i0, i1 in chunk_ranges(ng, nb):
            vb = vrho[i0:i1].unsqueeze(-1) * self.basis[i0:i1]
            if vgrad is not None:
                vg = vgrad[:, i0:i1] * 2
                for d in range(3):
                    vb += vg[d].unsqueeze(-1) * self.grad_basis[d, i0:i1]
            mat += torch.matmul(self.basis_dvolume[i0:i1].T, vb)
        mat = self.conv2(mat)
        return (mat + mat.T) * 0.5

    def get_vxc(self, dm):
        """dm: tensor (unpolarised) or (dm_u, dm_d) tuple."""
        if self.xcfamily == 4:
            if isinstance(dm, tuple):
                du, dd = self.dm2densinfo_mgga(dm[0]), self.dm2densinfo_mgga(dm[1])
                _, vu, vd = xc_ref.eval_pol_mgga(self.xcstr, du[0], dd[0], du[1], dd[1], du[2], dd[2], du[3], dd[3])
                return self.vxc_from_potinfo_mgga(*vu), self.vxc_from_potinfo_mgga(*vd)
            out = xc_ref.eval_unpol_mgga(self.xcstr, *self.dm2densinfo_mgga(dm))
            return self.vxc_from_potinfo_mgga(*out[1:])
        if isinstance(dm, tuple):
            (ru, gu), (rd, gd) = self.dm2densinfo(dm[0]), self.dm2densinfo(dm[1])
            _, (vu, vd), vg = xc_ref.eval_pol(self.xcstr, ru, rd, gu, gd)
            vgu, vgd = (None, None) if vg is None else vg
            return self.vxc_from_potinfo(vu, vgu), self.vxc_from_potinfo(vd, vgd)
        rho, grad = self.dm2densinfo(dm)
        _, vrho, vgrad = xc_ref.eval_unpol(self.xcstr, rho, grad)
        return self.vxc_from_potinfo(vrho, vgrad)

    def get_e_xc(self, dm):
        if self.xcfamily == 4:
            if isinstance(dm, tuple):
                du, dd = self.dm2densinfo_mgga(dm[0]), self.dm2densinfo_mgga(dm[1])
                e = xc_ref.eval_pol_mgga(self.xcstr, du[0], dd[0], du[1], dd[1], du[2], dd[2], du[3], dd[3])[0]
            else:
                e = xc_ref.eval_unpol_mgga(self.xcstr, *self.dm2densinfo_mgga(dm))[0]
            return torch.sum(self.dvolume * e)
        if isinstance(dm, tuple):
            (ru, gu), (rd, gd) = self.dm2densinfo(dm[0]), self.dm2densinfo(dm[1])
            e = xc_ref.eval_pol(self.xcstr, ru, rd, gu, gd)[0]
        else:
            rho, grad = self.dm2densinfo(dm)
            e = xc_ref.eval_unpol(self.xcstr, rho, grad)[0]
        return torch.sum(self.dvolume * e)

    def get_e_hcore(self, dm):
        return torch.einsum("ij,ji->", self.kinnucl_mat, dm)

    def get_e_elrep(self, dm):
        return 0.5 * torch.einsum("ij,ji->", self.get_elrep(dm), dm)

    def get_e_exchange(self, dm):
        if isinstance(dm, tuple):
            return sum(0.5 * torch.einsum("ij,ji->", self.get_exchange(2 * d), d) for d in dm)
        return 0.5 * torch.einsum("ij,ji->", self.get_exchange(dm), dm)

    def ao_orb2dm(self, orb, w):
        return torch.matmul(orb * w.unsqueeze(-2), orb.transpose(-2, -1))

    def aodm2dens(self, dm, xyz):
        dmao = self.unconv_dm(dm)
        b = torch.as_tensor(cint.eval_gto(self.atm, self.bas, self.env, np.asarray(xyz), 0, self.shl))
        return torch.einsum("ri,ij,rj->r", b, dmao, b)


def nuclei_energy(atomzs, atompos):
    z = np.asarray(atomzs, dtype=np.float64)
    p = np.asarray(atompos, dtype=np.float64)
    e = 0.0
    for i in range(len(z)):
        for j in range(i):
            e += z[i] * z[j] / np.linalg.norm(p[i] - p[j])
    return e


def becke_dvolume(rgrid, owner, dvol_atoms, atompos, radii, ratom_adjust="becke"):
    return dvol_atoms * becke_ref.becke_weights(rgrid, owner, atompos, radii, ratom_adjust)
