"""Coulomb-metric density fitting on the GPU.  Same role and surface as dqc/df/dfmol.py:13-101
(``build``, ``get_elrep``, ``j2c``, ``j3c``): J_ij = sum_P (ij|P) c_P with c = (sum_ij D_ij (ij|P)) .
inv(j2c), explicit inverse like the reference (:48).

B200 layout: (ij|P) is computed by the Rys kernel straight into its resident form -- packed over
AO pairs i >= j, (npair, ld) -- and never exists as the reference's (nao, nao, naux) tensor
(C60/def2-SVP: 12.7 GB packed vs 25 GB).  With several GPUs the aux shells are split into
contiguous slices, one per rank: pass 1 gives the rank's slice of temp, an all-gather (naux
doubles) completes it, every rank applies the replicated inv(j2c), pass 2 gives a partial J that is
summed by the caller's all-reduce."""
from typing import List, Optional
import numpy as np
import torch
from dqc_b200 import _lib
from dqc_b200.df.base_df import BaseDF
from dqc_b200.hamilton.intor.lcintwrap import LibcintWrapper
from dqc_b200.hamilton.intor import molintor as intor
from dqc_b200.utils.datastruct import DensityFitInfo
from dqc_b200.utils.linop import LinearOperator
from dqc_b200.utils.dist import ParallelContext, get_context
from dqc_b200.utils.config import config
from dqc_b200.utils.misc import logger

__all__ = ["DFMol", "dfk_chunking"]


def dfk_chunking(nao: int, nl: int, npad: int, kchunk_max: int, nsm: int = 148):
    """Split-K plan of DF-K stage 2, K_ik = sum_(P,o) Y_i(P,o) Y_k(P,o): the contraction index runs over the rank's
    ``nl`` aux functions times ``npad`` (padded) occupied orbitals and is cut into chunks of whole aux functions,
    ``pc`` per chunk -- short enough for exact integer accumulation (``pc * npad <= kchunk_max``) and numerous
    enough for ~4 waves of lower-triangle output tiles over the SMs.  Returns (pc, nchunk, kc, ktot, k_last)."""
    mt = (nao + 127) // 128
    ntl = sum(min((nao + 63) // 64, 2 * t + 2) for t in range(mt))       # 128 x 64 tiles touching the lower triangle
    want = max(1, (4 * nsm + ntl - 1) // ntl)
    pc = max(1, min(kchunk_max // npad, (nl + want - 1) // want))
    nchunk = (nl + pc - 1) // pc
    kc, ktot = pc * npad, nl * npad
    return pc, nchunk, kc, ktot, ktot - (nchunk - 1) * kc


class DFMol(BaseDF):
    def __init__(self, dfinfo: DensityFitInfo, wrapper: LibcintWrapper, orthozer=None,
                 ctx: Optional[ParallelContext] = None):
        self.dfinfo = dfinfo
        self.wrapper = wrapper
        self._is_built = False
        self._orthozer = orthozer
        self._ctx = ctx if ctx is not None else get_context()
        self._k_planes = None       # whitened (ij|P) as int8 operand planes (build_exchange)

    def build(self) -> BaseDF:
        self._is_built = True
        method = self.dfinfo.method
        auxw = LibcintWrapper(self.dfinfo.auxbases, spherical=self.wrapper.spherical)
        basisw, auxbw = LibcintWrapper.concatenate(self.wrapper, auxw)
        if method == "coulomb":
            logger.log("Calculating the 2e2c integrals")
            j2c = intor.coul2c(auxbw)
        elif method == "overlap":
            raise NotImplementedError("Density fitting with overlap minimization is not implemented")
        else:
            raise RuntimeError("Unknown density fitting method: %s" % method)
        self._basisw, self._auxbw = basisw, auxbw
        self._j2c = j2c
        self._inv_j2c = torch.inverse(j2c).contiguous()
        self._nao = basisw.nao()
        self._naux = auxbw.nao()
        # contiguous aux-shell slice of every rank, balanced by function count
        a0, a1 = auxbw.shell_idxs
        loc = auxbw.full_shell_to_aoloc
        world, rank = self._ctx.world, self._ctx.rank
        bounds = [a0]
        for r in range(1, world):
            target = loc[a0] + (loc[a1] - loc[a0]) * r / world
            bounds.append(int(a0 + np.searchsorted(loc[a0:a1 + 1], target)))
        bounds.append(a1)
        bounds = [min(max(b, a0), a1) for b in bounds]
        for r in range(1, len(bounds)):
            bounds[r] = max(bounds[r], bounds[r - 1])
        self._aux_bounds = bounds
        self._aux_sizes = [int(loc[bounds[r + 1]] - loc[bounds[r]]) for r in range(world)]
        self._aux_off = int(loc[bounds[rank]] - loc[a0])
        self._naux_local = self._aux_sizes[rank]
        logger.log("Calculating the 2e3c integrals")
        if self._naux_local > 0:
            self._j3c_packed = intor.coul3c_packed(basisw, auxbw, aux_slice=(bounds[rank], bounds[rank + 1]))
        else:
            self._j3c_packed = None
        # pair rows of (ij|P) that are negligible on this rank's columns are not read by the two DF-J passes
        self._row_mask, self._row_list, self.row_kept_fraction = None, None, 1.0
        if self._j3c_packed is not None and self._j3c_packed.is_cuda and config.DFJ_ROW_SKIP > 0.0:
            self._row_mask = _lib.dfj_rowmask(self._j3c_packed, self._nao, self._naux_local, config.DFJ_ROW_SKIP)
            self.row_kept_fraction = 1.0 - float(self._row_mask.double().mean())
            self._row_list = torch.nonzero(self._row_mask == 0).flatten().to(torch.int32).contiguous()
        logger.log("Density fitting done")
        return self

    # ---- AO-basis pieces (used by the fused Fock build) ----
    def elrep_ao_partial(self, dmao: torch.Tensor) -> torch.Tensor:
        """dmao (nao, nao) in the AO basis -> this rank's partial J (nao, nao); the caller sums over ranks."""
        nao, nl = self._nao, self._naux_local
        if nl > 0:
            temp_l = _lib.dfj_pass1(self._j3c_packed, nao, nl, dmao, self._row_list)
        else:
            temp_l = torch.zeros(0, dtype=dmao.dtype, device=dmao.device)
        temp = self._ctx.allgather_cat(temp_l, self._aux_sizes)
        coef = torch.matmul(temp, self._inv_j2c)                      # temp @ inv_j2c (dfmol.py:72)
        if nl == 0:
            return torch.zeros(nao, nao, dtype=dmao.dtype, device=dmao.device)
        return _lib.dfj_pass2(self._j3c_packed, nao, nl, coef[self._aux_off:self._aux_off + nl], self._row_mask)

    # ---- density-fitted exact exchange (extension: the reference raises, hcgto.py:229-230) ----
    def build_exchange(self, nslice: Optional[int] = None) -> "DFMol":
        """One-off set-up of DF-K: B_ij,P = sum_Q (ij|Q) L^-T_QP with (P|Q) = L L^T, kept as the
        int8 operand planes of stage 1 (batch = AO i, rows = this rank's aux functions P, K = AO j)."""
        if self._k_planes is not None:
            return self
        from dqc_b200.utils.config import config
        S = int(nslice if nslice is not None else config.DFK_I8_SLICES)
        nao, naux, nl, p0 = self._nao, self._naux, self._naux_local, self._aux_off
        dev = self._j2c.device
        logger.log("Whitening the 2e3c integrals for the exchange build")
        chol = torch.linalg.cholesky(self._j2c)
        linv = torch.linalg.solve_triangular(chol, torch.eye(naux, dtype=chol.dtype, device=dev), upper=False)
        wmat = linv[p0:p0 + nl, :].t().contiguous()                       # (naux, nl): L^-T columns of this rank
        npair = nao * (nao + 1) // 2
        b3 = torch.zeros(npair, max(nl, 1), dtype=torch.float64, device=dev)
        q0 = 0
        for r in range(self._ctx.world):
            nq = self._aux_sizes[r]
            if nq > 0 and nl > 0:
                if r == self._ctx.rank:
                    j3c_r = self._j3c_packed
                else:   # the other ranks' aux slices are recomputed here (one-off), not communicated
                    j3c_r = intor.coul3c_packed(self._basisw, self._auxbw,
                                                aux_slice=(self._aux_bounds[r], self._aux_bounds[r + 1]))
                for r0 in range(0, npair, 32768):
                    r1 = min(npair, r0 + 32768)
                    b3[r0:r1] += torch.matmul(j3c_r[r0:r1, :nq], wmat[q0:q0 + nq])
                del j3c_r
            q0 += nq
        self._k_S = S
        self._k_planes = _lib.I8Operand("A", nao, max(nl, 1), nao, S, device=dev).fill(b3, 0, 0, 0, pair_ld=b3.shape[1])
        del b3
        return self

    def exchange_ao_partial(self, cw: torch.Tensor) -> torch.Tensor:
        """cw (nao, nocc): AO-basis orbitals times sqrt(occupation), D = cw cw^T  ->  this rank's partial
        K_ij = sum_P sum_o Y_iPo Y_jPo,  Y_iPo = sum_j B_ij,P cw_jo  (two tcgen05 int8 GEMMs, csrc/gemm_i8.cuh)."""
        self.build_exchange()
        nao, nl, S = self._nao, self._naux_local, self._k_S
        dev = cw.device
        kmat = torch.zeros(nao, nao, dtype=torch.float64, device=dev)
        nocc = int(cw.shape[1])
        if nl == 0 or nocc == 0:
            return kmat
        cw = cw.contiguous()
        npad = _lib.round_up(nocc, 64)
        # K chunks of stage 2 = whole aux functions: pc of them (pc * npad columns) per chunk
        pc, nchunk, kc, ktot, k_last = dfk_chunking(nao, nl, npad, _lib.I8_KCHUNK[S])
        # stage 1: Y[i][P][o] = sum_j B[i][P][j] cw[j][o]      (batch i; M = P, N = o, K = j); the epilogue also
        # leaves max |Y| per (i, chunk of P) -- the row scales of the stage-2 operands
        cop = _lib.I8Operand("B", 1, nocc, nao, S, device=dev).fill(cw, 0, 1, nocc)
        y = torch.empty(nao, nl, npad, dtype=torch.float64, device=dev)
        ymax = torch.zeros(nao, nchunk, dtype=torch.float64, device=dev)
        _lib.gemm_i8(self._k_planes, cop, y, nl * npad, npad, nl, npad, mode=0, nbatch=nao, b_shared=True,
                     rowmax=ymax, rm_bstride=nchunk, rm_div=pc)
        # stage 2: K[i][k] = sum_(P,o) Y[i][(P,o)] Y[k][(P,o)]   (split-K chunks as batches, lower-triangle tiles)
        ya, yb = _lib.i8_slice_dual(y, nchunk, kc, ktot, ymax, nchunk, nao, kc, k_last, S)
        _lib.gemm_i8(ya, yb, kmat, 0, nao, nao, nao, mode=2)
        low = torch.tril(kmat)
        return low + torch.tril(kmat, -1).t()

    def get_elrep(self, dm: torch.Tensor) -> LinearOperator:
        # dm: (*BD, nao2, nao2) in the orthogonalised basis
        if self._orthozer is not None:
            dm = self._orthozer.unconvert_dm(dm)
        bshape = dm.shape[:-2]
        dm2 = dm.reshape(-1, *dm.shape[-2:])
        mats = torch.stack([self.elrep_ao_partial(dm2[b].contiguous()) for b in range(dm2.shape[0])])
        self._ctx.allreduce_(mats)
        mat = mats.reshape(*bshape, *mats.shape[-2:])
        mat = (mat + mat.transpose(-2, -1)) * 0.5
        if self._orthozer is not None:
            mat = self._orthozer.convert2(mat)
        return LinearOperator.m(mat, is_hermitian=True)

    @property
    def j2c(self) -> torch.Tensor:
        return self._j2c

    @property
    def j3c(self) -> torch.Tensor:
        """Dense (nao, nao, naux) view for API parity -- materialised on demand (single rank only)."""
        if self._ctx.world != 1:
            raise RuntimeError("the dense j3c view is only available on a single GPU")
        nao, naux = self._nao, self._naux
        ii, jj = torch.tril_indices(nao, nao, device=self._j3c_packed.device)
        out = torch.empty(nao, nao, naux, dtype=self._j3c_packed.dtype, device=self._j3c_packed.device)
        out[ii, jj] = self._j3c_packed[:, :naux]
        out[jj, ii] = self._j3c_packed[:, :naux]
        return out

    def getparamnames(self, methodname: str, prefix: str = "") -> List[str]:
        if methodname == "get_elrep":
            params = [prefix + "_inv_j2c", prefix + "_j3c_packed"]
            if self._orthozer is not None:
                params += self._orthozer.getparamnames("unconvert_dm", prefix=prefix + "_orthozer.")
            return params
        raise KeyError("getparamnames has no %s method" % methodname)
