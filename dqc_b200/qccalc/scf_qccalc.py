"""SCF driver: the CALLER of the Fock-build path.  Mirrors the reference's surface
(dqc/qccalc/scf_qccalc.py:18-316, hf.py:14-316, ks.py:15-238): ``HF(system).run().energy()``,
``KS(system, xc).run()``, ``aodm()``, ``dm2energy(dm)``, engines with ``dm2scp / scp2dm / scp2scp /
dm2energy``; the self-consistent parameter is the Fock matrix, the initial guess ``dm0="1e"`` is
zero density -> core Fock -> density (:88-91).

The reference hands ``scp2scp`` to ``xitorch.optimize.equilibrium`` (Broyden's good method,
alpha=-0.5, maxiter=50, :48-53, :109-113); xitorch is absent here, so the fixed point is found by
the two solvers below (device-resident, no host round trip except the convergence scalar).  The
fixed point -- hence the energy -- is the same; only the path differs."""
from abc import abstractmethod
from typing import Any, Dict, Optional, Union
import warnings
import torch
from dqc_b200.utils.config import config
from dqc_b200.utils.datastruct import SpinParam
from dqc_b200.utils.linop import EditableModule

__all__ = ["SCF_QCCalc", "BaseSCFEngine", "equilibrium", "ConvergenceWarning"]


class ConvergenceWarning(UserWarning):
    """The fixed-point iteration stopped at ``maxiter`` with max|fcn(y) - y| >= f_tol (xitorch warns likewise)."""


def _not_converged(method: str, maxiter: int, err: float, f_tol: float, info: Optional[dict]):
    if info is not None:
        info.update(converged=False, niter=maxiter, residual=err)
    warnings.warn("equilibrium(method=%r) did not converge in %d iterations: max|f(y) - y| = %.3e >= f_tol = %.1e; "
                  "the last iterate is returned" % (method, maxiter, err, f_tol), ConvergenceWarning, stacklevel=3)


def equilibrium(fcn, y0: torch.Tensor, method: str = "diis", maxiter: int = 50, f_tol: float = 1e-9,
                alpha: float = -0.5, history: int = 8, verbose: bool = False, info: Optional[dict] = None,
                **unused) -> torch.Tensor:
    """Solve y = fcn(y).  method "diis": Anderson/Pulay mixing of the last `history` iterates on the
    residual fcn(y) - y; "broyden1": Broyden's good method with J0 = alpha^-1 I like the reference's
    default; "simple": y <- fcn(y).  Stops when max|fcn(y) - y| < f_tol; when ``maxiter`` is exhausted first a
    ``ConvergenceWarning`` is emitted and the last iterate returned.  ``info`` (a dict, optional) receives
    ``converged``, ``niter`` and the last ``residual``."""
    y = y0
    shape = y0.shape
    err = float("inf")
    if method == "simple":
        for it in range(maxiter):
            fy = fcn(y)
            err = float((fy - y).abs().max())
            if err < f_tol:
                if info is not None:
                    info.update(converged=True, niter=it + 1, residual=err)
                return fy
            y = fy
        _not_converged(method, maxiter, err, f_tol, info)
        return y
    if method == "broyden1":
        x = y0.reshape(-1)
        f = (fcn(x.reshape(shape)).reshape(-1) - x)
        us, vs = [], []          # B^-1 = alpha I + sum_k u_k v_k^T  (Sherman-Morrison updates)

        def binv(v):
            out = alpha * v
            for u, w in zip(us, vs):
                out = out + u * torch.dot(w, v)
            return out

        def binv_t(v):
            out = alpha * v
            for u, w in zip(us, vs):
                out = out + w * torch.dot(u, v)
            return out
        for it in range(maxiter + 1):
            err = float(f.abs().max())
            if err < f_tol:
                if info is not None:
                    info.update(converged=True, niter=it, residual=err)
                return x.reshape(shape)
            if it == maxiter:
                break
            dx = -binv(f)
            xn = x + dx
            fn = fcn(xn.reshape(shape)).reshape(-1) - xn
            df = fn - f
            bdf = binv(df)
            denom = torch.dot(dx, bdf)
            us.append((dx - bdf) / denom)
            vs.append(binv_t(dx))
            x, f = xn, fn
            if verbose:
                print("broyden1 iter %3d  max|f| = %.3e" % (it, float(f.abs().max())))
        _not_converged(method, maxiter, err, f_tol, info)
        return x.reshape(shape)
    if method != "diis":
        raise RuntimeError("Unknown equilibrium method: %s (available: diis, broyden1, simple)" % method)
    ys, rs = [], []
    # Convergence is tested WITHOUT stalling the device queue: the residual norm of iteration k travels to pinned host
    # memory behind an event and is looked at while iteration k + lag is being enqueued (lag = config.SCF_CHECK_LAG, 0 =
    # test every iteration synchronously like the reference's solver).  A converged iterate found that way is returned as
    # it was; the `lag` builds enqueued after it are discarded.  (torch.linalg.eigh inside scp2dm still checks its own
    # status on the host; the DIIS solve below does not.)
    lag = int(config.SCF_CHECK_LAG) if y0.is_cuda else 0
    pending = []            # (iteration, pinned residual, event, iterate)

    def poll(force: bool):
        nonlocal err
        while pending and (force or len(pending) > lag or pending[0][2].query()):
            it_k, host, ev, fy_k = pending.pop(0)
            ev.synchronize()
            err = float(host)
            if verbose:
                print("diis iter %3d  max|f| = %.3e" % (it_k, err))
            if err < f_tol:
                if info is not None:
                    info.update(converged=True, niter=it_k + 1, residual=err)
                return fy_k
        return None
    for it in range(maxiter):
        fy = fcn(y)
        r = (fy - y).reshape(-1)
        if lag > 0:
            host = torch.empty((), dtype=r.dtype, pin_memory=True)
            host.copy_(r.abs().max(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((it, host, ev, fy))
            done = poll(False)
            if done is not None:
                return done
        else:
            err = float(r.abs().max())
            if verbose:
                print("diis iter %3d  max|f| = %.3e" % (it, err))
            if err < f_tol:
                if info is not None:
                    info.update(converged=True, niter=it + 1, residual=err)
                return fy
        ys.append(fy.reshape(-1))
        rs.append(r)
        ys, rs = ys[-history:], rs[-history:]
        n = len(rs)
        if n == 1:
            y = fy
            continue
        R = torch.stack(rs)
        B = torch.zeros(n + 1, n + 1, dtype=R.dtype, device=R.device)
        B[:n, :n] = R @ R.T
        B[n, :n] = B[:n, n] = -1.0
        rhs = torch.zeros(n + 1, dtype=R.dtype, device=R.device)
        rhs[n] = -1.0
        # solve_ex: no status round trip; a singular DIIS matrix (linearly dependent residuals) gives info != 0 and the
        # plain iterate is used for this step (selected on the device)
        c, status = torch.linalg.solve_ex(B, rhs)
        c = torch.nan_to_num(c[:n], nan=0.0, posinf=0.0, neginf=0.0)
        ymix = (c.unsqueeze(-1) * torch.stack(ys)).sum(0)
        y = torch.where(status == 0, ymix, fy.reshape(-1)).reshape(shape)
    done = poll(True)
    if done is not None:
        return done
    _not_converged(method, maxiter, err, f_tol, info)
    return y


class BaseSCFEngine(EditableModule):
    @abstractmethod
    def get_system(self):
        pass

    @abstractmethod
    def dm2scp(self, dm):
        pass

    @abstractmethod
    def scp2dm(self, scp):
        pass

    @abstractmethod
    def scp2scp(self, scp):
        pass

    @abstractmethod
    def dm2energy(self, dm):
        pass

    @abstractmethod
    def set_eigen_options(self, eigen_options: Dict[str, Any]) -> None:
        pass


class SCF_QCCalc(object):
    def __init__(self, engine: BaseSCFEngine, variational: bool = False):
        if variational:
            raise NotImplementedError("variational (direct-minimisation) SCF is outside the Fock-build path "
                                      "(SURVEY section 2, component 14); use variational=False")
        self._engine = engine
        self._polarized = engine.polarized
        self._shape = engine.shape
        self.dtype = engine.dtype
        self.device = engine.device
        self._has_run = False
        self._variational = variational

    def get_system(self):
        return self._engine.get_system()

    def run(self, dm0="1e", eigen_options: Optional[Dict[str, Any]] = None,
            fwd_options: Optional[Dict[str, Any]] = None, bck_options: Optional[Dict[str, Any]] = None):
        fwd = {"method": "diis", "alpha": -0.5, "maxiter": 50, "verbose": config.VERBOSE > 0}
        fwd.update(fwd_options or {})
        self._engine.set_eigen_options(eigen_options or {"method": "exacteig"})
        if dm0 is None:
            dm = self._get_zero_dm()
        elif isinstance(dm0, str):
            if dm0 != "1e":
                raise RuntimeError("Unknown dm0: %s" % dm0)
            dm = self._engine.scp2dm(self._engine.dm2scp(self._get_zero_dm()))
        else:
            dm = SpinParam.apply_fcn(lambda d: d.detach().to(self.device), dm0)
        if isinstance(dm, torch.Tensor) and self._polarized:
            dm = SpinParam(u=dm * 0.5, d=dm * 0.5)
        elif isinstance(dm, SpinParam) and not self._polarized:
            dm = dm.u + dm.d
        scp0 = self._engine.dm2scp(dm)
        self.scf_info: Dict[str, Any] = {}
        scp = equilibrium(self._engine.scp2scp, scp0, info=self.scf_info, **fwd)
        self._dm = self._engine.scp2dm(scp)
        self._has_run = True
        return self

    @property
    def converged(self) -> bool:
        """False when the last ``run`` stopped at ``maxiter`` (a ConvergenceWarning was emitted then)."""
        assert self._has_run, "run() must be called first"
        return bool(self.scf_info.get("converged", False))

    def energy(self) -> torch.Tensor:
        assert self._has_run, "run() must be called first"
        return self.dm2energy(self._dm)

    def aodm(self):
        assert self._has_run, "run() must be called first"
        return self._dm

    def dm2energy(self, dm) -> torch.Tensor:
        assert (isinstance(dm, torch.Tensor) and not self._polarized) or \
            (isinstance(dm, SpinParam) and self._polarized), \
            "The dm must be a Tensor for unpolarized case and SpinParam of Tensor for polarized case"
        return self._engine.dm2energy(dm)

    def _get_zero_dm(self):
        z = torch.zeros(self._shape, dtype=self.dtype, device=self.device)
        return SpinParam(u=z, d=z.clone()) if self._polarized else z
